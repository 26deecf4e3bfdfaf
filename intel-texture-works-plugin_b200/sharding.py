"""Row-band sharding of the encode over several GPUs (one process per GPU, torch.distributed).

4x4 blocks are independent (no halo, no exchange during compute), so the only communication is
putting the packed output back together: ONE all-gather per job.  The band rule is
the reference's own thread split (win32Threads.cpp:217-230): linesPerThread = ceil(h/N), band
edges rounded down to multiples of 4 rows -- so N ranks produce exactly the bytes N reference
threads would.

The encoder is passed in as a callable so that the host logic can be tested on CPU (gloo) with the
oracle standing in for the GPU kernels; on the GPU box it is `ItwBcn.encode_device`.
"""
import torch
import torch.distributed as dist


def band_rows(height, parts, index):
    """Rows [y0, y1) of band `index` of `parts` (win32Threads.cpp:217-230). May be empty."""
    lines = (height + parts - 1) // parts
    y0 = (lines * index) // 4 * 4
    y1 = min((lines * (index + 1)) // 4 * 4, height)
    return y0, max(y0, y1)


def band_bytes(width, y0, y1, bytes_per_block):
    return (width // 4) * ((y1 - y0) // 4) * bytes_per_block


def encode_levels_sharded(levels, bytes_per_block, encode_band, group=None, device="cpu"):
    """Encode a list of surfaces (e.g. a mip chain), every one row-sharded over the ranks of `group`.

    levels       list of (width, height) of each surface (multiples of 4)
    encode_band  callable(level_index, y0, y1) -> 1-D uint8 tensor on `device` holding the packed
                 blocks of rows [y0, y1) of that level (never called for empty bands)
    returns      list of 1-D uint8 tensors (one per level, the complete packed output), identical on
                 every rank.  Exactly one collective (all_gather_into_tensor) is issued.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # per-rank payload = concatenation over levels of that rank's band; sizes are known everywhere
    sizes = [[band_bytes(w, *band_rows(h, world, r), bytes_per_block) for (w, h) in levels] for r in range(world)]
    slot = max(sum(s) for s in sizes)
    slot = (slot + 15) // 16 * 16
    mine = torch.zeros(slot, dtype=torch.uint8, device=device)
    off = 0
    for li, (w, h) in enumerate(levels):
        y0, y1 = band_rows(h, world, rank)
        n = sizes[rank][li]
        if n:
            part = encode_band(li, y0, y1)
            assert part.numel() == n, (part.numel(), n)
            mine[off:off + n] = part
        off += n
    if world > 1:
        gathered = torch.empty(world * slot, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    # local reorder: level-major, rank-minor
    out = []
    offs = [0] * world
    for li in range(len(levels)):
        parts = []
        for r in range(world):
            n = sizes[r][li]
            if n:
                parts.append(gathered[r * slot + offs[r]: r * slot + offs[r] + n])
            offs[r] += n
        out.append(torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=device))
    return out


def shardable_levels(height, levels, world):
    """Number of leading mip levels whose row bands are exact 2:1 images of the level-0 bands: level l
    qualifies while (height >> l) / world is a positive multiple of 4 rows (then band_rows() gives every rank
    rows [r*h_l/world, (r+1)*h_l/world) and a rank can filter its own band without a halo)."""
    n = 0
    for l in range(levels):
        hl = height >> l
        if hl <= 0 or hl % world or (hl // world) % 4:
            break
        n += 1
    return n


def encode_mip_chain_sharded(lib, fmt, band0, width, height, levels, settings=None, group=None):
    """Config C4: encode the full mip chain of a width x height RGBA8 texture whose level 0 is ROW-SHARDED over
    the ranks (rank r holds rows band_rows(height, world, r) in the CUDA uint8 tensor `band0`).

    Mips are made on the GPU (itw_generate_mips_device).  The first `shardable_levels` levels are filtered and
    encoded band-locally; the last of them is all-gathered as raw texels (a few KiB) so that every rank can
    finish the tiny remaining levels redundantly; the packed blocks are reassembled with ONE all-gather
    (encode_levels_sharded).  Returns the list of complete packed levels (identical on every rank)."""
    import ctypes
    from .binding import FORMATS, RgbaSurface
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bpb = FORMATS[fmt][1]
    dev = band0.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    nshard = shardable_levels(height, levels, world)
    assert nshard >= 1, "level 0 itself must split into bands of a multiple of 4 rows"
    y0, y1 = band_rows(height, world, rank)
    bh = y1 - y0
    # band-local levels 0..nshard-1
    local = (RgbaSurface * nshard)()
    scratch_a = torch.empty(max(lib.lib.itw_mip_scratch_bytes(width, bh, nshard, 1), 16), dtype=torch.uint8, device=dev)
    top = RgbaSurface(band0.data_ptr(), width, bh, width * 4)
    if lib.lib.itw_generate_mips_device(ctypes.byref(top), nshard, local, ctypes.c_void_p(scratch_a.data_ptr()), ctypes.c_void_p(stream)) != 0:
        lib.check()
    # replicated tail: gather the raw texels of level nshard-1, then filter the remaining levels everywhere
    tail = None
    keep = [scratch_a]
    if nshard < levels:
        last = local[nshard - 1]
        nbytes = last.width * last.height * 4
        mine = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ctypes_copy = torch.cuda.current_stream(dev)
        # the band of the last shardable level is tightly packed in scratch (or is band0 itself when nshard == 1)
        src = band0 if nshard == 1 else scratch_a
        off = 0 if nshard == 1 else last.ptr - scratch_a.data_ptr()
        mine.copy_(src[off:off + nbytes])
        full = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(full, mine, group=group)
        else:
            full = mine
        lw, lh = width >> (nshard - 1), height >> (nshard - 1)
        rest = levels - nshard + 1
        tail = (RgbaSurface * rest)()
        scratch_b = torch.empty(max(lib.lib.itw_mip_scratch_bytes(lw, lh, rest, 1), 16), dtype=torch.uint8, device=dev)
        ftop = RgbaSurface(full.data_ptr(), lw, lh, lw * 4)
        if lib.lib.itw_generate_mips_device(ctypes.byref(ftop), rest, tail, ctypes.c_void_p(scratch_b.data_ptr()), ctypes.c_void_p(stream)) != 0:
            lib.check()
        keep += [full, scratch_b]

    def encode_band(li, r0, r1):
        w_l = max(width >> li, 1)
        pw = (w_l + 3) // 4 * 4
        out = torch.empty((pw // 4) * ((r1 - r0) // 4) * bpb, dtype=torch.uint8, device=dev)
        if li < nshard:                                   # my own band of a shardable level
            s = local[li]
            ptr, stride = s.ptr, s.stride
        else:                                             # a band of a replicated (padded) small level
            s = tail[li - nshard + 1]
            ptr, stride = s.ptr + r0 * s.stride, s.stride
        lib.encode_device(fmt, ptr, pw, r1 - r0, stride, out.data_ptr(), settings, stream)
        return out

    dims = [((max(width >> l, 1) + 3) // 4 * 4, (max(height >> l, 1) + 3) // 4 * 4) for l in range(levels)]
    out = encode_levels_sharded(dims, bpb, encode_band, group=group, device=dev)
    torch.cuda.current_stream(dev).synchronize()
    del keep
    return out


def tile_owner(tile_index, world):
    """Independent tiles (config C5) are dealt round-robin; no collective at all."""
    return tile_index % world
