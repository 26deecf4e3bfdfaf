"""The product's kernel logic, executed lane by lane on the CPU (tests/emu), against the oracle.
This covers the warp program (phases, lane<->task maps, winner selection, bit packing) without a
GPU; the GPU's own float instructions are covered by the -m gpu tests."""
import numpy as np
import pytest

import itw_testlib as T


@pytest.mark.parametrize("fmt,prof", T.ALL_CASES, ids=[f"{f}-{p}" for f, p in T.ALL_CASES])
def test_emulated_kernels_bit_exact(fmt, prof):
    bpb = T.binding.FORMATS[fmt][1]
    for name, img in T.corpus_for(fmt, 32).items():
        got = T.run(T.emu(), fmt, img, prof)
        want = T.run(T.oracle(), fmt, img, prof)
        assert T.differing_blocks(got, want, bpb) == 0, f"{fmt}/{prof}/{name}"


@pytest.mark.parametrize("fmt,prof", [("BC7", "alpha_basic"), ("BC6H", "bc6h_basic"), ("BC3", None)])
def test_ragged_batches_and_strides(fmt, prof):
    """Block counts that are not multiples of the per-warp batch, non-square surfaces, padded strides."""
    rng = np.random.default_rng(5)
    dt = np.uint16 if fmt == "BC6H" else np.uint8
    hi = 0x7C00 if fmt == "BC6H" else 256
    for h, w, pad in ((4, 4, 0), (4, 12, 3), (12, 20, 5), (8, 4, 1)):
        buf = rng.integers(0, hi, (h, w + pad, 4)).astype(dt)
        img = buf[:, :w]                       # row stride larger than the row
        got = T.run(T.emu(), fmt, img, prof)
        want = T.run(T.oracle(), fmt, np.ascontiguousarray(img), prof)
        assert np.array_equal(got, want), (h, w, pad)
