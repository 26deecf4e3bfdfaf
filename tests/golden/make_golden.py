#!/usr/bin/env python3
"""Generate tests/golden/digests.json from the REFERENCE-SOURCE build (oracle/_ref), i.e. from the
reference's own kernel.ispc + ispc_texcomp.cpp compiled scalar by oracle/build_ref.py.

Run here (where /root/reference exists):   python tests/golden/make_golden.py
BC4/BC5 are not part of the ISPC reference: their digests come from DirectXTex's own encoder bodies
(oracle/_ref/libitw_ref_frontend.so, cut from BC4BC5.cpp / BC.h by oracle/build_ref_frontend.py), fed with
texel floats byte * (1/255)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import itw_testlib as T  # noqa: E402


def main():
    ref = T.ref()
    assert ref is not None, "needs /root/reference (or a prebuilt oracle/_ref/libitw_ref.so)"
    digests, inputs, source = {}, {}, {}
    import test_bc45_vs_directxtex as DX
    dxlib = DX.ref_lib()
    for fmt, prof in T.ALL_CASES:
        for name, img in T.corpus_for(fmt).items():
            if fmt in ("BC4", "BC5"):
                data, source[fmt] = DX.ref_encode(dxlib, fmt, img), "DirectXTex encoder bodies (oracle/build_ref_frontend.py)"
            else:
                data, source[fmt] = T.run(ref, fmt, img, prof), "reference-source"
            digests[f"{fmt}:{prof}:{name}"] = hashlib.sha256(data.tobytes()).hexdigest()
    for fmt in ("BC7", "BC6H"):
        for name, img in T.corpus_for(fmt).items():
            inputs[f"{fmt}:{name}"] = hashlib.sha256(img.tobytes()).hexdigest()
    out = {"generator": "tests/golden/make_golden.py", "produced_by": source, "corpus": "tests/itw_testlib.py corpus8/corpus16 (size 64)",
           "inputs": inputs, "digests": digests}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "digests.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(digests), "digests")


if __name__ == "__main__":
    main()
