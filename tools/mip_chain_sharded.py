#!/usr/bin/env python3
"""Config C4: BC3 + full mip chain of an 8192x8192 RGBA8 texture, level 0 row-sharded over the ranks
(torchrun --nproc-per-node N).  Times the whole sharded job (GPU mips, encode, gathers) with CUDA events,
max over ranks, and checks rank 0's result against a single-GPU encode of the same texture."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("intel-texture-works-plugin_b200")
sharding = importlib.import_module("intel-texture-works-plugin_b200.sharding")

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--format", default="BC3")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = pkg.ItwBcn()
lib.set_device(local)
n, levels = a.size, a.size.bit_length()
base = pkg.synth.mixed_rgba8(n, n)
y0, y1 = sharding.band_rows(n, world, rank)
band = torch.from_numpy(np.ascontiguousarray(base[y0:y1]).reshape(-1)).cuda()


def job():
    return sharding.encode_mip_chain_sharded(lib, a.format, band, n, n, levels)


out = job()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    out = job()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    texels = sum(max(n >> l, 1) ** 2 for l in range(levels))
    d = pkg.DdsDesc(n, n, levels, 1, 77 if a.format == "BC3" else 71, 0)
    whole = lib.dds_encode_texture(d, [base])
    import ctypes
    ok = True
    for l in range(levels):
        off = lib.lib.itw_dds_image_offset(ctypes.byref(d), 0, l)
        ok = ok and np.array_equal(out[l].cpu().numpy(), whole[off:off + out[l].numel()])
    print(json.dumps({"config": f"C4 {a.format} + {levels}-level mip chain {n}x{n}, row-sharded over {world} GPU(s)",
                      "ms_per_job": round(float(ms.item()), 4), "Mtexels_per_s": round(texels / float(ms.item()) / 1e3, 1),
                      "equals_single_gpu_result": bool(ok)}))
if world > 1:
    dist.destroy_process_group()
