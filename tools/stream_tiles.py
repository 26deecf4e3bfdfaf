#!/usr/bin/env python3
"""Config C5 in small: N independent tiles streamed through one GPU, pinned host memory -> packed blocks in
pinned host memory.  Compares the pipelined itw_encode_batch with a loop of CompressBlocks* calls.
usage: python tools/stream_tiles.py [--format BC7] [--profile basic] [--tiles 64] [--size 1024]"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pkg = importlib.import_module("intel-texture-works-plugin_b200")
binding = importlib.import_module("intel-texture-works-plugin_b200.binding")

ap = argparse.ArgumentParser()
ap.add_argument("--format", default="BC7")
ap.add_argument("--profile", default="basic")
ap.add_argument("--tiles", type=int, default=64)
ap.add_argument("--size", type=int, default=1024)
a = ap.parse_args()
lib = pkg.ItwBcn()
fmt, n, size = a.format, a.tiles, a.size
kind = binding.FORMATS[fmt][3]
settings = lib.profile(a.profile) if kind else None
_, bpb, texel, _ = binding.FORMATS[fmt]
ndistinct = min(n, 8)
srcs = []
for t in range(ndistinct):
    img = pkg.synth.random_rgba16f(size, size, seed=t) if fmt == "BC6H" else (
        pkg.synth.random_rgba8(size, size, seed=0xB2000005 + t) if t % 2 == 0 else pkg.synth.gradient_rgba8(size, size, phase=t))
    srcs.append(torch.from_numpy(img.view(np.uint8).reshape(-1)).pin_memory())
out_bytes = (size // 4) ** 2 * bpb
outs = [torch.zeros(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
surf = [(srcs[i % ndistinct].data_ptr(), size, size, size * texel) for i in range(n)]
dst = [o.data_ptr() for o in outs]

lib.encode_batch(fmt, surf[:4], dst[:4], settings)          # warm-up (allocations)
t0 = time.perf_counter()
lib.encode_batch(fmt, surf, dst, settings)
t_batch = time.perf_counter() - t0
ref = [o.clone() for o in outs]
t0 = time.perf_counter()
for s, d in zip(surf, dst):
    lib.encode_raw(fmt, s[0], s[1], s[2], s[3], d, settings)
t_loop = time.perf_counter() - t0
assert all(torch.equal(a_, b_) for a_, b_ in zip(ref, outs)), "batch and loop disagree"
mt = n * size * size / 1e6
print(f"{fmt} {a.profile if kind else ''} {n} tiles of {size}^2: batch {mt / t_batch:.1f} Mtexels/s ({t_batch * 1e3:.1f} ms), "
      f"loop {mt / t_loop:.1f} Mtexels/s ({t_loop * 1e3:.1f} ms)")
