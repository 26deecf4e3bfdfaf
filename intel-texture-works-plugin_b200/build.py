"""Build libitw_bcn.so (the product: sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

The float flags are part of the contract, not tuning: the kernels must reproduce the canonical
strict-IEEE execution of the reference encoder, so FMA contraction is off and
division / square root are the IEEE-rounded versions.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libitw_bcn.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-shared", "-cudart", "static",
]


def sources():
    return [os.path.join(CSRC, "itw_bcn.cu")]


def deps():
    out = [os.path.join(HERE, "..", "include", "itw_bcn.h"), os.path.abspath(__file__)]
    for n in os.listdir(CSRC):
        out.append(os.path.join(CSRC, n))
    return out


def build(force=False, verbose=True, extra=()):
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(d) <= t for d in deps()):
            return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = OUT + ".tmp"                     # written aside and renamed: a snapshot of the tree never sees a half-written library
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + sources() + ["-o", tmp]
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, extra=[a for a in sys.argv[1:] if a != "--force"])
    print(OUT)
