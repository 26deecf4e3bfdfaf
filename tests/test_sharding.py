"""Multi-GPU host logic on CPU: world_size-2 gloo processes, the oracle standing in for the kernels.
The sharded result after the single all-gather must equal the single-process encode byte for byte."""
import importlib
import os
import socket
import sys

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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fmt, prof, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = T.oracle()
        settings = oracle.profile(prof) if prof else None
        base = T.synth.mixed_rgba8(64, 32)
        chain = T.synth.mip_chain(base)                      # 7 levels, the small ones padded to 4x4
        bpb = T.binding.FORMATS[fmt][1]

        def encode_band(li, y0, y1):
            img = np.ascontiguousarray(chain[li][y0:y1])
            return torch.from_numpy(oracle.encode(fmt, img, settings))

        got = sharding.encode_levels_sharded([(l.shape[1], l.shape[0]) for l in chain], bpb, encode_band)
        want = [oracle.encode(fmt, np.ascontiguousarray(l), settings) for l in chain]
        ok = all(np.array_equal(g.numpy(), w) for g, w in zip(got, want))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt,prof", [("BC3", None), ("BC7", "veryfast")])
def test_row_sharded_mip_chain_equals_single_process(fmt, prof):
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fmt, prof, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _gpu_worker(rank, world, port, ret):
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        lib = T.product()
        lib.set_device(rank)
        fmt, bpb = "BC3", 16
        chain = T.synth.mip_chain(T.synth.mixed_rgba8(256, 256))
        dev = [torch.from_numpy(np.ascontiguousarray(l).reshape(-1)).cuda() for l in chain]

        def encode_band(li, y0, y1):
            h, w = chain[li].shape[:2]
            out = torch.empty((w // 4) * ((y1 - y0) // 4) * bpb, dtype=torch.uint8, device="cuda")
            lib.encode_device(fmt, dev[li].data_ptr() + y0 * w * 4, w, y1 - y0, w * 4, out.data_ptr(), None,
                              torch.cuda.current_stream().cuda_stream)
            return out

        got = sharding.encode_levels_sharded([(l.shape[1], l.shape[0]) for l in chain], bpb, encode_band, device="cuda")
        torch.cuda.synchronize()
        want = [lib.encode(fmt, np.ascontiguousarray(l)) for l in chain]
        ok = all(np.array_equal(g.cpu().numpy(), w) for g, w in zip(got, want))
        # config C4 proper: level 0 row-sharded, mips made on the GPUs, one small texel gather + one block gather
        base = T.synth.mixed_rgba8(256, 256, seed=3)
        y0, y1 = sharding.band_rows(256, world, rank)
        band = torch.from_numpy(np.ascontiguousarray(base[y0:y1]).reshape(-1)).cuda()
        got2 = sharding.encode_mip_chain_sharded(lib, fmt, band, 256, 256, 9)
        ref_chain = T.synth.mip_chain(base)
        want2 = [lib.encode(fmt, np.ascontiguousarray(l)) for l in ref_chain]
        ok = ok and len(got2) == 9 and all(np.array_equal(g.cpu().numpy(), w) for g, w in zip(got2, want2))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shardable_levels_rule():
    assert [sharding.shardable_levels(8192, 14, n) for n in (1, 2, 4, 8)] == [12, 11, 10, 9]
    assert sharding.shardable_levels(256, 9, 2) == 6 and sharding.shardable_levels(4, 3, 1) == 1


@pytest.mark.gpu
def test_row_sharded_mip_chain_on_two_gpus_nccl():
    """Config C4 in small: BC3 + full mip chain, row-sharded over 2 GPUs, one NCCL all-gather."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_gpu_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2)), dict(ret)
