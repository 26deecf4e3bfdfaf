"""Proof by exhaustion, on the GPU, that the FMA-corrected quotient of itw_device.cuh (div_by_rcp) equals the IEEE
quotient for ARBITRARY normal operands: all 2^23 x 2^23 significand pairs (tests/gpu_checks/div_sweep.cu).  bc6h.cuh
relies on it for the projection / squared-length division of the index search, whose operands are not confined to a
small domain (the small domains are proved on the CPU by tests/test_exact_division.py)."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_fma_corrected_quotient_equals_ieee_division_for_every_significand_pair():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "gpu_checks"))
    try:
        import build as gpu_checks_build
        path = gpu_checks_build.build(verbose=False)
    finally:
        sys.path.pop(0)
    lib = ctypes.CDLL(path)
    lib.div_sweep.restype = ctypes.c_int
    lib.div_sweep.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_uint)]
    bad = ctypes.c_ulonglong(0)
    examples = (ctypes.c_uint * 32)()
    count = int(os.environ.get("ITW_DIV_SWEEP_COUNT", 1 << 23))          # all divisor significands by default
    assert lib.div_sweep(0, count, ctypes.byref(bad), examples) == 0
    pairs = [(hex(examples[2 * i]), hex(examples[2 * i + 1])) for i in range(min(int(bad.value), 16))]
    assert bad.value == 0, (bad.value, pairs)
