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


def tile_owner(tile_index, world):
    """Independent tiles (config C5) are dealt round-robin; no collective at all."""
    return tile_index % world
