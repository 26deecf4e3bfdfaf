#!/usr/bin/env python3
"""Generate tests/golden/frontend_digests.json from the reference's OWN front-end / mip-generator bodies
(oracle/_ref/libitw_ref_frontend.so, built by oracle/build_ref_frontend.py from IntelPlugin.h, IntelPlugin.cpp,
DirectXTexMipmaps.cpp and Filters.h).

Run here (where /root/reference exists):   python tests/golden/make_golden_frontend.py
The inputs are the seeded generators of tests/test_frontend.py / tests/test_mips_f16.py."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import itw_testlib as T  # noqa: E402
import test_frontend as F  # noqa: E402
import test_mips_f16 as M  # noqa: E402


def main():
    lib = T.ref_frontend()
    assert lib is not None, "needs /root/reference (or a prebuilt oracle/_ref/libitw_ref_frontend.so)"
    convert, mips = {}, {}
    for fmt, depth, planes in F.CASES:
        px = F.source(depth, planes, 13, 7, seed=depth + planes)
        for flags in F.flag_sets(fmt, depth, planes):
            convert[f"{fmt}:{depth}:{planes}:{flags}"] = hashlib.sha256(F.ref_convert(lib, fmt, px, flags).tobytes()).hexdigest()
    for h, w in M.SIZES:
        img = M.random_f16(h, w, seed=h * 131 + w)
        chain = M.chain_with(lib.ref_mip_chain_f16, img, M.full_levels(w, h))
        mips[f"{h}x{w}"] = hashlib.sha256(b"".join(l.tobytes() for l in chain)).hexdigest()
    import ctypes
    import numpy as np
    import test_mips as M8
    lib.ref_mip_chain_rgba8.restype = ctypes.c_int
    lib.ref_mip_chain_rgba8.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    mips8 = {}
    for h, w in M8.SIZES:
        img = M8.random_rgba8(h, w)
        for srgb in (0, 1):
            chain = M8.chain_with(lib.ref_mip_chain_rgba8, img, srgb)
            mips8[f"{h}x{w}:{'srgb' if srgb else 'unorm'}"] = hashlib.sha256(b"".join(l.tobytes() for l in chain)).hexdigest()
    out = {"generator": "tests/golden/make_golden_frontend.py", "produced_by": "reference function bodies (oracle/build_ref_frontend.py)",
           "convert": convert, "mip_chain_f16": mips, "mip_chain_rgba8": mips8}
    path = os.path.join(HERE, "frontend_digests.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(convert), "conversion digests,", len(mips), "+", len(mips8), "mip-chain digests")


if __name__ == "__main__":
    main()
