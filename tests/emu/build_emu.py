"""Build tests/emu/libitw_emu.so: the product's kernel logic compiled for the CPU (TEST-ONLY)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
OUT = os.path.join(HERE, "libitw_emu.so")
CSRC = os.path.join(ROOT, "intel-texture-works-plugin_b200", "csrc")


def build(verbose=True):
    deps = [os.path.join(HERE, "emu.cpp")] + [os.path.join(CSRC, n) for n in os.listdir(CSRC)]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fwrapv",
           "-mfpmath=sse", "-msse2", "-Wno-unknown-pragmas", os.path.join(HERE, "emu.cpp"), "-o", OUT]
    if verbose:
        print("[build_emu]", " ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
