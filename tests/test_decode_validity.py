"""Decode validity: every block the oracle emits must parse with an INDEPENDENT decoder
(tests/bcn_decode.py, written from the format definition), decode close to the input on compressible images, and
the corpus must actually reach every BC7 and BC6H mode -- otherwise "bit-exact on the corpus" would say little."""
import collections
import os

import numpy as np
import pytest

import bcn_decode as D
import itw_testlib as T

LAYOUTS = D.bc6_layouts(open(os.path.join(T.ROOT, "oracle", "itw_oracle.cpp")).read())
# PSNR floors (dB) on the two compressible LDR images; measured values are 2-10 dB above these (BC7 ultrafast, mode 6 only, is the lowest)
FLOORS = {"BC1": (35.0, 28.0), "BC3": (36.0, 28.0), "BC4": (48.0, 42.0), "BC5": (48.0, 42.0), "BC7": (38.0, 27.0)}


@pytest.mark.parametrize("fmt,prof", [c for c in T.ALL_CASES if c[0] != "BC6H"], ids=lambda v: str(v))
def test_ldr_streams_decode_and_are_close(fmt, prof):
    o = T.oracle()
    corpus = T.corpus8()
    ch = {"BC1": 3, "BC3": 4, "BC4": 1, "BC5": 2}.get(fmt, 4 if (prof or "").startswith("alpha") else 3)
    for name in ("gradient", "smooth", "flat", "twocolour", "random"):
        img = corpus[name]
        dec, _ = D.decode_image(fmt, T.run(o, fmt, img, prof), 64, 64)        # raises on an illegal stream
        p = D.psnr(dec[..., :ch], img[..., :ch])
        if name == "gradient":
            assert p >= FLOORS[fmt][0], (name, p)
        if name == "smooth":
            assert p >= FLOORS[fmt][1], (name, p)
        if name in ("flat", "twocolour") and fmt in ("BC4", "BC5", "BC7"):
            assert p >= 50.0, (name, p)


@pytest.mark.parametrize("prof", T.binding.BC6H_PROFILES)
def test_hdr_streams_decode_and_are_close(prof):
    o = T.oracle()
    corpus = T.corpus16()
    for name in ("smooth", "flat", "lowvar", "narrow", "zeros", "maxhalf", "random"):
        img = corpus[name]
        dec, _ = D.decode_image("BC6H", T.run(o, "BC6H", img, prof), 64, 64, LAYOUTS)
        p = D.psnr(dec, img[..., :3].astype(np.int64), peak=31743.0)            # on half bit patterns (~log scale)
        if name in ("smooth", "flat", "narrow"):
            assert p >= 60.0, (name, p)
        if name == "lowvar":
            assert p >= 50.0, (name, p)
        if name in ("zeros", "maxhalf"):
            assert p >= 90.0, (name, p)


def test_corpus_reaches_every_mode():
    o = T.oracle()
    bc7, bc6 = collections.Counter(), collections.Counter()
    for prof in ("slow", "alpha_slow"):
        for img in T.corpus8().values():
            bc7.update(D.decode_image("BC7", T.run(o, "BC7", img, prof), 64, 64)[1])
    for prof in ("bc6h_slow", "bc6h_fast"):
        for img in T.corpus16().values():
            bc6.update(D.decode_image("BC6H", T.run(o, "BC6H", img, prof), 64, 64, LAYOUTS)[1])
    assert set(bc7) == set(range(8)), dict(bc7)
    assert set(bc6) == set(range(14)), dict(bc6)
