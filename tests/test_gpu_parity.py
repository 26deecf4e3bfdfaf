"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle (bit-exact)."""
import numpy as np
import pytest

import itw_testlib as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt,prof", T.ALL_CASES, ids=[f"{f}-{p}" for f, p in T.ALL_CASES])
def test_corpus_bit_exact(fmt, prof):
    lib, oracle = T.product(), T.oracle()
    bpb = T.binding.FORMATS[fmt][1]
    for name, img in T.corpus_for(fmt).items():
        got = T.run(lib, fmt, img, prof)
        want = T.run(oracle, fmt, img, prof)
        assert T.differing_blocks(got, want, bpb) == 0, f"{fmt}/{prof}/{name}"
