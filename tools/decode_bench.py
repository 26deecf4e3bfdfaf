"""Device-resident decode timing (include/itw_bcn.h section 5): blocks in HBM -> RGBA8 / RGBA16F surface in HBM.

    python tools/decode_bench.py [--size 4096] [--reps 20]

Prints one JSON line per format: kernel time (CUDA events inside the library, itw_last_kernel_ms), Gtexel/s and the
algorithmic HBM traffic rate (block bytes read + texel bytes written) against MEASURED_PEAKS.json."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("intel-texture-works-plugin_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3, help="untimed launches before the timed ones")
    a = ap.parse_args()
    import torch
    api = pkg.ItwBcn()
    n = a.size
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass
    rng = np.random.default_rng(1)
    img8 = torch.from_numpy(rng.integers(0, 256, (n, n, 4), dtype=np.uint8)).cuda()
    img16 = torch.from_numpy(rng.integers(0, 0x7BFF, (n, n, 4), dtype=np.uint16).view(np.int16)).cuda()
    for fmt, prof in (("BC1", None), ("BC3", None), ("BC4", None), ("BC5", None), ("BC7", "veryfast"), ("BC6H", "bc6h_veryfast")):
        _, bpb, texel, _ = pkg.binding.FORMATS[fmt]
        src = img16 if fmt == "BC6H" else img8
        nblk = (n // 4) * (n // 4)
        blocks = torch.empty(nblk * bpb, dtype=torch.uint8, device="cuda")
        api.encode_raw(fmt, src.data_ptr(), n, n, n * texel, blocks.data_ptr(), api.profile(prof) if prof else None)
        # rotate over several destinations so that consecutive launches do not hit lines already in L2
        outs = [torch.empty(n * n * texel, dtype=torch.uint8, device="cuda") for _ in range(4)]
        ms = []
        for i in range(a.reps + a.warm):
            api.decode_raw(fmt, blocks.data_ptr(), outs[i % 4].data_ptr(), n, n, n * texel)
            if i >= a.warm:
                ms.append(api.last_kernel_ms())
        t = float(np.median(ms))
        traffic = nblk * bpb + n * n * texel
        line = {"op": "decode", "format": fmt, "size": n, "kernel_ms": round(t, 4), "gtexel_s": round(n * n / t / 1e6, 1),
                "algorithmic_gbps": round(traffic / t / 1e6, 1)}
        if peak:
            line["frac_of_measured_hbm"] = round(traffic / t / 1e6 / peak, 3)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
