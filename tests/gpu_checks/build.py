"""Builds tests/gpu_checks/libdivsweep.so (TEST-ONLY CUDA code, sm_100a) -- called by __graft_entry__.build()."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libdivsweep.so")


def build(verbose=True):
    src = os.path.join(HERE, "div_sweep.cu")
    if os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(src):
        return OUT
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", src, "-o", OUT]
    if verbose:
        print("[gpu_checks]", " ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
