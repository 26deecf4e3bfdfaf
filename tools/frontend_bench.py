"""Device-resident timing of the pixel-format front end (include/itw_bcn.h section 6) and of the image-level entry.

    python tools/frontend_bench.py [--size 4096] [--reps 20]

One JSON line per case: kernel time (CUDA events inside the library), algorithmic HBM rate (source bytes read +
surface bytes written) against MEASURED_PEAKS.json."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("intel-texture-works-plugin_b200")
B = pkg.binding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3, help="untimed launches before the timed ones")
    a = ap.parse_args()
    import torch
    api = pkg.ItwBcn()
    n = a.size
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        peak = None
    rng = np.random.default_rng(1)
    cases = [("BC7", 8, 4, B.FRONT_HAS_ALPHA), ("BC7", 8, 3, 0), ("BC7", 16, 4, B.FRONT_HAS_ALPHA), ("BC7", 32, 3, B.FRONT_GAMMA),
             ("BC5", 8, 3, B.FRONT_FLIP_Y | B.FRONT_NORMALIZE), ("BC6H", 32, 3, 0), ("BC6H", 16, 4, B.FRONT_HAS_ALPHA)]
    for fmt, depth, planes, flags in cases:
        texel = B.FORMATS[fmt][2]
        if depth == 8:
            host = rng.integers(0, 256, (n, n, planes), dtype=np.uint8)
        elif depth == 16:
            host = rng.integers(0, 32769, (n, n, planes), dtype=np.uint16).view(np.int16)
        else:
            host = rng.random((n, n, planes), dtype=np.float32)
        srcs = [torch.from_numpy(host).cuda() for _ in range(2)]
        outs = [torch.empty(n * n * texel, dtype=torch.uint8, device="cuda") for _ in range(3)]
        ms = []
        for i in range(a.reps + a.warm):
            src = B.PixelSource(srcs[i % 2].data_ptr(), n, n, planes, depth, 0)
            api.convert_pixels_raw(fmt, src, flags, outs[i % 3].data_ptr(), n, n, n * texel)
            if i >= a.warm:
                ms.append(api.last_kernel_ms())
        t = float(np.median(ms))
        traffic = n * n * (planes * depth // 8 + texel)
        line = {"op": "convert_pixels", "to": fmt, "depth": depth, "planes": planes, "flags": flags, "size": n, "kernel_ms": round(t, 4),
                "algorithmic_gbps": round(traffic / t / 1e6, 1)}
        if peak:
            line["frac_of_measured_hbm"] = round(traffic / t / 1e6 / peak, 3)
        print(json.dumps(line), flush=True)
    # image-level entry: planes in HBM -> blocks in HBM (front end + encoder, two launches)
    for fmt, prof, depth, planes, flags in (("BC1", None, 8, 3, 0), ("BC7", "veryfast", 8, 4, 1)):
        host = rng.integers(0, 256, (n, n, planes), dtype=np.uint8)
        d_src = torch.from_numpy(host).cuda()
        blocks = torch.empty((n // 4) * (n // 4) * B.FORMATS[fmt][1], dtype=torch.uint8, device="cuda")
        src = B.PixelSource(d_src.data_ptr(), n, n, planes, depth, 0)
        settings = api.profile(prof) if prof else None
        ms = []
        for i in range(a.warm + max(a.reps, 1)):
            api.encode_pixels_raw(fmt, src, flags, blocks.data_ptr(), settings)
            if i >= a.warm:
                ms.append(api.last_kernel_ms())
        print(json.dumps({"op": "encode_pixels", "format": fmt, "profile": prof, "size": n, "kernel_ms": round(float(np.median(ms)), 4),
                          "mtexel_s": round(n * n / float(np.median(ms)) / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
