"""Row-band sharding of the encode over several GPUs (one process per GPU).

The work itself is in the C++ library (csrc/itw_shard.inc): `itw_shard_plan_make` (band rule and slot /
chain layout, pure arithmetic), `itw_shard_init` (NCCL communicator) and `itw_encode_mip_chain_sharded`
(band mips -> encode -> ONE ncclAllGather -> chain order -> replicated tail).  This module is the thin
Python face of those entry points for torch.distributed hosts (bench.py, tests):

  * `shard_init(lib)`          bootstrap the library's NCCL communicator from an initialised
                               torch.distributed group (the id travels by a broadcast);
  * `encode_mip_chain_sharded` config C4: returns the complete packed chain on every rank;
  * `run_plan_on_cpu`          the same plan executed with a CPU encoder and any torch.distributed backend
                               (gloo) -- TEST helper that exercises the plan arithmetic without a GPU.

4x4 blocks are independent (no halo, no exchange during compute), so the only communication is putting
the packed output back together.  The band rule is the reference's own thread split
(win32Threads.cpp:217-230): linesPerThread = ceil(h/N), band edges rounded down to multiples of 4 rows --
so N ranks produce exactly the bytes N reference threads would.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from .binding import FORMATS, RgbaSurface, ShardPlan


def band_rows(height, parts, index):
    """Rows [y0, y1) of band `index` of `parts` (win32Threads.cpp:217-230). May be empty."""
    lines = (height + parts - 1) // parts
    y0 = (lines * index) // 4 * 4
    y1 = min((lines * (index + 1)) // 4 * 4, height)
    return y0, max(y0, y1)


def shardable_levels(height, levels, world):
    """Number of leading mip levels whose row bands are exact 2:1 images of the level-0 bands: level l
    qualifies while (height >> l) / world is a positive multiple of 4 rows (itw_shard_plan.band_levels)."""
    n = 0
    for l in range(levels):
        hl = height >> l
        if hl <= 0 or hl % world or (hl // world) % 4:
            break
        n += 1
    return n


def make_plan(lib, fmt, width, height, levels, world, rank):
    """itw_shard_plan_make as a ShardPlan structure (no CUDA call)."""
    plan = ShardPlan()
    f = lib.lib.itw_shard_plan_make
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ShardPlan)]
    if f(FORMATS[fmt][0], width, height, levels, world, rank, ctypes.byref(plan)) != 0:
        lib.check()
        raise RuntimeError("itw_shard_plan_make failed")
    return plan


def shard_init(lib, group=None):
    """Create the library's NCCL communicator over the ranks of an initialised torch.distributed group:
    rank 0 draws the id (ncclGetUniqueId), a broadcast carries it, every rank calls itw_shard_init."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ident = (ctypes.c_uint8 * 128)()
    if rank == 0 and lib.lib.itw_shard_unique_id(ident) != 0:
        lib.check()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    ident = (ctypes.c_uint8 * 128)(*t.cpu().tolist())
    lib.lib.itw_shard_init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    if lib.lib.itw_shard_init(rank, world, ident) != 0:
        lib.check()
        raise RuntimeError("itw_shard_init failed")


def shard_finalize(lib):
    lib.lib.itw_shard_finalize.restype = None
    lib.lib.itw_shard_finalize()


def encode_mip_chain_sharded(lib, fmt, band0, width, height, levels, settings=None, chain=None, stream=None):
    """Config C4: the packed mip chain of a width x height RGBA8 texture whose level 0 is ROW-SHARDED over the
    ranks (rank r holds rows band_rows(height, world, r) in the CUDA uint8 tensor `band0`).  Everything is
    enqueued on `stream` (default: torch's current stream); returns (chain tensor, plan): the complete chain,
    level l at plan.level_offset[l], identical on every rank.  Needs shard_init() first when world > 1."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    plan = make_plan(lib, fmt, width, height, levels, world, rank)
    if chain is None:
        chain = torch.empty(plan.chain_bytes, dtype=torch.uint8, device=band0.device)
    if stream is None:
        stream = torch.cuda.current_stream(band0.device).cuda_stream
    surf = RgbaSurface(band0.data_ptr(), width, plan.band_y1 - plan.band_y0, width * 4)
    f = lib.lib.itw_encode_mip_chain_sharded
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int, ctypes.POINTER(RgbaSurface), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_void_p]
    sp = ctypes.cast(ctypes.byref(settings), ctypes.c_void_p) if settings is not None else None
    if f(FORMATS[fmt][0], ctypes.byref(surf), width, height, levels, sp, ctypes.c_void_p(chain.data_ptr()), ctypes.c_void_p(stream)) != 0:
        lib.check()
        raise RuntimeError("itw_encode_mip_chain_sharded failed")
    return chain, plan


def split_chain(chain, plan):
    """The packed levels of a chain buffer as a list of views."""
    return [chain[plan.level_offset[l]: plan.level_offset[l] + plan.level_bytes[l]] for l in range(plan.levels)]


def run_plan_on_cpu(lib, fmt, level0, levels, encode, mip_chain, group=None):
    """TEST helper: execute the library's shard plan with a CPU encoder over any torch.distributed backend.
    level0 = the whole level-0 image (numpy, every rank has it; only its own band is read before the gather),
    encode(img) -> packed blocks (numpy uint8), mip_chain(img) -> list of padded levels.  Returns the chain."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    h, w = level0.shape[:2]
    plan = make_plan(lib, fmt, w, h, levels, world, rank)
    assert (plan.band_y0, plan.band_y1) == band_rows(h, world, rank)
    nshard = plan.band_levels
    local = mip_chain(np.ascontiguousarray(level0[plan.band_y0:plan.band_y1]))[:nshard]      # band-local filtering
    slot = np.zeros(plan.slot_bytes, np.uint8)
    for l in range(nshard):
        part = encode(np.ascontiguousarray(local[l]))
        assert part.size == plan.band_bytes[l], (l, part.size, plan.band_bytes[l])
        slot[plan.send_offset[l]: plan.send_offset[l] + part.size] = part
    if plan.texel_bytes:
        raw = np.ascontiguousarray(local[nshard - 1]).reshape(-1)
        assert raw.size == plan.texel_bytes
        slot[plan.texel_offset: plan.texel_offset + raw.size] = raw
    mine = torch.from_numpy(slot)
    if world > 1:
        gathered = torch.empty(world * plan.slot_bytes, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, mine, group=group)                              # the ONE collective
    else:
        gathered = mine
    g = gathered.numpy()
    chain = np.zeros(plan.chain_bytes, np.uint8)
    for l in range(nshard):
        for r in range(world):
            src = r * plan.slot_bytes + plan.send_offset[l]
            dst = plan.level_offset[l] + r * plan.band_bytes[l]
            chain[dst: dst + plan.band_bytes[l]] = g[src: src + plan.band_bytes[l]]
    if plan.texel_bytes:
        lw = max(w >> (nshard - 1), 1)
        pw = (lw + 3) // 4 * 4
        rows = np.concatenate([g[r * plan.slot_bytes + plan.texel_offset: r * plan.slot_bytes + plan.texel_offset + plan.texel_bytes]
                               for r in range(world)]).reshape(-1, pw, 4)
        tail = mip_chain(np.ascontiguousarray(rows[:, :lw]))                                  # valid columns only, like the library
        for l in range(nshard, levels):
            part = encode(np.ascontiguousarray(tail[l - nshard + 1]))
            assert part.size == plan.level_bytes[l]
            chain[plan.level_offset[l]: plan.level_offset[l] + part.size] = part
    return chain, plan


def tile_owner(tile_index, world):
    """Independent tiles (config C5) are dealt round-robin; no collective at all (itw_encode_batch does the same
    over the devices of itw_set_devices inside one process)."""
    return tile_index % world
