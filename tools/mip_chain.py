#!/usr/bin/env python3
"""Config C4 on one GPU, device-resident: 8192x8192 RGBA8 -> 14-level mip chain (GPU box filter + pad) -> BC3.
Prints the time of the pre-pass and of the encode, and the pre-pass's HBM traffic rate."""
import argparse
import ctypes
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("intel-texture-works-plugin_b200")
binding = importlib.import_module("intel-texture-works-plugin_b200.binding")

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--format", default="BC3")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--srgb", action="store_true", help="the sRGB-correct chain of the *_SRGB encodings (itw_generate_mips_device_srgb)")
a = ap.parse_args()
lib = pkg.ItwBcn()
n = a.size
levels = n.bit_length()
img = torch.from_numpy(pkg.synth.mixed_rgba8(n, n).reshape(-1)).cuda()
scratch = torch.empty(lib.lib.itw_mip_scratch_bytes(n, n, levels, 1), dtype=torch.uint8, device="cuda")
outs = (binding.RgbaSurface * levels)()
top = binding.RgbaSurface(img.data_ptr(), n, n, n * 4)
bpb = binding.FORMATS[a.format][1]
dst = [torch.empty(((max(n >> l, 1) + 3) // 4) ** 2 * bpb, dtype=torch.uint8, device="cuda") for l in range(levels)]
stream = torch.cuda.current_stream()
sp = ctypes.c_void_p(stream.cuda_stream)


def mips():
    f = lib.lib.itw_generate_mips_device_srgb if a.srgb else lib.lib.itw_generate_mips_device
    assert f(ctypes.byref(top), levels, outs, ctypes.c_void_p(scratch.data_ptr()), sp) == 0


def encode():
    for l in range(levels):
        lib.encode_device(a.format, outs[l].ptr, outs[l].width, outs[l].height, outs[l].stride, dst[l].data_ptr(), None, stream.cuda_stream)


def timed(f):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        f()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


t_mip, t_enc = timed(mips), timed(encode)
texels = sum(max(n >> l, 1) ** 2 for l in range(levels))
mip_bytes = sum(max(n >> l, 1) ** 2 * 4 for l in range(0, levels - 1)) + sum(o.width * o.height * 4 for o in list(outs)[1:])
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"config": f"C4 {a.format} + {levels}-level {'sRGB-correct ' if a.srgb else ''}mip chain, {n}x{n} RGBA8, 1 GPU, device-resident",
                  "mip_prepass_ms": round(t_mip, 4), "mip_prepass_GBps": round(mip_bytes / t_mip / 1e6, 1),
                  "mip_prepass_frac_of_hbm_peak": round(mip_bytes / t_mip / 1e6 / peak, 3),
                  "encode_ms": round(t_enc, 4), "Mtexels_per_s_encode": round(texels / t_enc / 1e3, 1),
                  "Mtexels_per_s_total": round(texels / (t_enc + t_mip) / 1e3, 1)}))
