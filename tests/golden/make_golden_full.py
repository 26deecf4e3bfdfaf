#!/usr/bin/env python3
"""Generate tests/golden/full_digests.json: SHA-256 of the reference's output on the WHOLE BASELINE configs
(C1, C2a, C2b, C3, C4, first tiles of C5), produced by the REFERENCE-SOURCE build oracle/_ref
(kernel.ispc + ispc_texcomp.cpp compiled scalar by oracle/build_ref.py), multi-threaded with the row-band split of
CompressImageMT (bands are independent surfaces, so the bytes do not depend on the thread count).

Run here (where /root/reference exists):   python tests/golden/make_golden_full.py      (a few minutes on 8 cores)

Each config stores the digest of the whole packed output and of every chunk of 64 block rows (so that a mismatch can be
localised).  C4's mip levels come from synth.mip_chain -- the integer box filter that is the product's stated RGBA8 mip
contract (csrc/mips.cuh) -- and every level is encoded by the reference's BC3 encoder."""
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import itw_testlib as T  # noqa: E402


def encode_mt(api, fmt, img, settings, threads):
    h, w = img.shape[:2]
    bpb = T.binding.FORMATS[fmt][1]
    out = np.zeros((h // 4) * (w // 4) * bpb, np.uint8)
    lines = (h + threads - 1) // threads
    jobs = []
    for t in range(threads):
        y0, y1 = (lines * t) // 4 * 4, min((lines * (t + 1)) // 4 * 4, h)
        if y1 > y0:
            jobs.append((y0, y1))

    def work(j):
        y0, y1 = j
        api.encode_raw(fmt, img.ctypes.data + y0 * img.strides[0], w, y1 - y0, img.strides[0], out.ctypes.data + (y0 // 4) * (w // 4) * bpb, settings)
    ts = [threading.Thread(target=work, args=(j,)) for j in jobs]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def digest_entry(fmt, img, out):
    h, w = img.shape[:2]
    bpb = T.binding.FORMATS[fmt][1]
    row = (w // 4) * bpb
    chunks = [hashlib.sha256(out[r * row:(r + 64) * row].tobytes()).hexdigest() for r in range(0, h // 4, 64)]
    return {"sha256": hashlib.sha256(out.tobytes()).hexdigest(), "bytes": int(out.size), "input_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
            "chunk_block_rows": 64, "chunks": chunks}


def configs():
    """name -> (format, profile, list of surfaces); shared with tests/test_full_configs.py"""
    S = T.synth
    yield "C1", "BC1", None, [S.gradient_rgba8(512, 512)]
    yield "C2a", "BC7", "slow", [S.random_rgba8(4096, 4096)]
    yield "C2b", "BC7", "alpha_slow", [S.random_rgba8(4096, 4096)]
    yield "C3", "BC6H", "bc6h_slow", [S.random_rgba16f(4096, 4096)]
    yield "C4", "BC3", None, S.mip_chain(S.mixed_rgba8(8192, 8192))
    yield "C5", "BC7", "basic", [S.c5_tile(t) for t in range(8)]


def main():
    ref = T.ref()
    assert ref is not None, "needs /root/reference (or a prebuilt oracle/_ref/libitw_ref.so)"
    threads = len(os.sched_getaffinity(0))
    out = {"generator": "tests/golden/make_golden_full.py", "produced_by": "oracle/_ref (reference-source build)", "configs": {}}
    for name, fmt, prof, surfaces in configs():
        t0 = time.time()
        settings = ref.profile(prof) if prof else None
        entries = [digest_entry(fmt, np.ascontiguousarray(s), encode_mt(ref, fmt, np.ascontiguousarray(s), settings, threads)) for s in surfaces]
        out["configs"][name] = {"format": fmt, "profile": prof, "surfaces": entries}
        print(name, fmt, prof, len(entries), "surfaces", f"{time.time() - t0:.1f}s", flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_digests.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
