"""Multi-GPU host logic.

CPU (gloo, world size 2): the library's own shard plan (itw_shard_plan_make, pure arithmetic in csrc/itw_shard.inc)
executed with the oracle standing in for the kernels -- the sharded result after the single all-gather must equal the
single-process encode byte for byte.  GPU (-m gpu, 2+ GPUs): the real thing, itw_encode_mip_chain_sharded over NCCL."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import itw_testlib as T

sharding = importlib.import_module("intel-texture-works-plugin_b200.sharding")


def test_band_rule_matches_reference_thread_split():
    # win32Threads.cpp:217-230 with 4096 rows and 8 threads: 8 bands of 512 rows
    assert [sharding.band_rows(4096, 8, i) for i in range(8)] == [(512 * i, 512 * i + 512) for i in range(8)]
    # more parts than block rows: some bands are empty, none overlap, all rows covered
    for h, n in ((8, 3), (4, 8), (20, 3), (64, 7), (8192, 8)):
        bands = [sharding.band_rows(h, n, i) for i in range(n)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(b[0] % 4 == 0 and b[1] % 4 == 0 for b in bands)
        assert all(bands[i][1] == bands[i + 1][0] for i in range(n - 1))


def test_shardable_levels_rule():
    assert [sharding.shardable_levels(8192, 14, n) for n in (1, 2, 4, 8)] == [12, 11, 10, 9]
    assert sharding.shardable_levels(256, 9, 2) == 6 and sharding.shardable_levels(4, 3, 1) == 1


def test_plan_layout_c4():
    """itw_shard_plan_make for config C4 (8192^2 BC3, 14 levels, 8 ranks): the numbers of SURVEY.md 8e."""
    lib = T.product()
    plans = [sharding.make_plan(lib, "BC3", 8192, 8192, 14, 8, r) for r in range(8)]
    p = plans[3]
    assert p.band_levels == 9 and (p.band_y0, p.band_y1) == (3072, 4096)
    assert p.chain_bytes == sum(((max(8192 >> l, 1) + 3) // 4) ** 2 * 16 for l in range(14))       # 64 MiB + 21.3 MiB
    assert p.band_bytes[0] == 8 << 20 and p.level_bytes[0] == 64 << 20
    assert p.texel_bytes == 32 * 4 * 4                                                              # level 8 is 32x32: 4 rows per rank
    assert p.slot_bytes % 16 == 0 and p.slot_bytes >= sum(p.band_bytes[l] for l in range(9)) + p.texel_bytes
    assert all(q.slot_bytes == p.slot_bytes and q.chain_bytes == p.chain_bytes for q in plans)
    # the rule of sharding.shardable_levels is the library's
    for h, levels, n in ((8192, 14, 1), (8192, 14, 2), (256, 9, 2), (64, 7, 4), (48, 6, 3)):
        assert sharding.make_plan(lib, "BC1", 64, h, levels, n, 0).band_levels == sharding.shardable_levels(h, levels, n)
    # errors are reported, not swallowed
    with pytest.raises(RuntimeError, match="multiple of 4 rows"):
        sharding.make_plan(lib, "BC3", 64, 40, 3, 4, 0)
    with pytest.raises(RuntimeError, match="RGBA8"):
        sharding.make_plan(lib, "BC6H", 64, 64, 3, 2, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fmt, prof, shape, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle, lib = T.oracle(), T.product()               # the product library is only asked for the PLAN (no CUDA call)
        settings = oracle.profile(prof) if prof else None
        h, w = shape
        base = T.synth.mixed_rgba8(h, w)
        levels = max(h, w).bit_length()

        def encode(img):
            return oracle.encode(fmt, img, settings)

        chain_of = lambda img: T.oracle_mip_chain_rgba8(img, 0)      # the product's RGBA8 contract (box / linear, DirectXTex non-WIC)
        got, plan = sharding.run_plan_on_cpu(lib, fmt, base, levels, encode, chain_of)
        want = np.concatenate([encode(np.ascontiguousarray(l)) for l in chain_of(base)])
        ret[rank] = bool(np.array_equal(got, want)) and plan.band_levels < levels
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt,prof,shape", [("BC3", None, (64, 32)), ("BC7", "veryfast", (64, 32)), ("BC1", None, (256, 4)), ("BC3", None, (32, 20))],
                         ids=["BC3-64x32", "BC7-64x32", "BC1-narrow", "BC3-npot-width"])
def test_row_sharded_mip_chain_equals_single_process(fmt, prof, shape):
    """World size 2 over gloo; the narrow / non-power-of-two widths cover the padded-width tail (a level whose stored
    width is pad4(width) but whose valid width is smaller)."""
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fmt, prof, shape, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _gpu_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        lib = T.product()
        lib.set_device(rank)
        sharding.shard_init(lib)
        ok = True
        for fmt, prof, (h, w) in (("BC3", None, (256, 256)), ("BC1", None, (512, 8)), ("BC7", "veryfast", (128, 64)), ("BC3", None, (64, 20))):
            settings = lib.profile(prof) if prof else None
            levels = max(h, w).bit_length()
            base = T.synth.mixed_rgba8(h, w, seed=3 + h)
            y0, y1 = sharding.band_rows(h, world, rank)
            band = torch.from_numpy(np.ascontiguousarray(base[y0:y1]).reshape(-1)).cuda()
            chain, plan = sharding.encode_mip_chain_sharded(lib, fmt, band, w, h, levels, settings)
            torch.cuda.synchronize()
            want = np.concatenate([lib.encode(fmt, np.ascontiguousarray(l), settings) for l in T.oracle_mip_chain_rgba8(base, 0)])
            ok = ok and np.array_equal(chain.cpu().numpy(), want)
        sharding.shard_finalize(lib)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_row_sharded_mip_chain_on_two_gpus_nccl():
    """Config C4 in small: full mip chains, level 0 row-sharded over 2 GPUs, ONE ncclAllGather issued by the library."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_gpu_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2)), dict(ret)


@pytest.mark.gpu
def test_sharded_entry_on_one_gpu_equals_the_plain_chain():
    """Without a communicator itw_encode_mip_chain_sharded is the single-GPU chain encode (bands = the whole image)."""
    lib = T.product()
    for fmt, (h, w) in (("BC3", (256, 128)), ("BC1", (64, 20)), ("BC5", (128, 128))):
        base = T.synth.mixed_rgba8(h, w, seed=h)
        levels = max(h, w).bit_length()
        band = torch.from_numpy(base.reshape(-1)).cuda()
        chain, plan = sharding.encode_mip_chain_sharded(lib, fmt, band, w, h, levels)
        torch.cuda.synchronize()
        want = np.concatenate([lib.encode(fmt, np.ascontiguousarray(l)) for l in T.oracle_mip_chain_rgba8(base, 0)])
        assert np.array_equal(chain.cpu().numpy(), want), fmt
