"""Property tests ("Property (hypothesis)"): invariants that hold for every input, driven by hypothesis
over sizes, strides, band splits, formats and profiles -- not only over fixed seeds.

  * output size = (w/4)(h/4) * bytes-per-block, and every block parses (legal mode) in the independent decoder;
  * row-band split invariance: encoding any set of bands separately == encoding the whole image (win32Threads.cpp:217-230);
  * stride / sub-surface invariance: a padded row stride or a bumped base pointer does not change the blocks;
  * block permutation equivariance: the encoder sees 4x4 blocks only;
  * determinism;
  * DDS header write -> read round trip; shard plan tiles the chain exactly.
CPU: the oracle and the kernels' logic on the CPU (tests/emu).  GPU (-m gpu): the product through the C-ABI."""
import ctypes

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import itw_testlib as T

B = T.binding
CASES = [("BC1", None), ("BC3", None), ("BC4", None), ("BC5", None), ("BC7", "veryfast"), ("BC7", "alpha_fast"), ("BC6H", "bc6h_fast")]
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large, HealthCheck.function_scoped_fixture])


def image(fmt, bh, bw, seed, pad=0, kind=0):
    rng = np.random.default_rng(seed)
    h, w = 4 * bh, 4 * bw
    if fmt == "BC6H":
        buf = rng.integers(0, 0x7C00, (h, w + pad, 4)).astype(np.uint16)
    elif kind == 1:                                            # smooth: exercises near-ties
        y, x = np.mgrid[0:h, 0:w + pad]
        buf = np.stack([(x * 3 + seed) & 255, (y * 5 + x) & 255, ((x + y) * 2) & 255, 255 - ((x * y) & 31)], -1).astype(np.uint8)
    elif kind == 2:                                            # few levels: flat blocks, coincident endpoints
        buf = (rng.integers(0, 3, (h // 4, (w + pad + 3) // 4, 4)) * 120).astype(np.uint8).repeat(4, 0).repeat(4, 1)[:, :w + pad]
    else:
        buf = rng.integers(0, 256, (h, w + pad, 4), dtype=np.uint8)
    return buf[:, :w]


def check_invariants(api, fmt, prof, bh, bw, seed, pad, kind, cuts):
    bpb = B.FORMATS[fmt][1]
    img = image(fmt, bh, bw, seed, pad, kind)                  # row stride = (w + pad) texels
    tight = np.ascontiguousarray(img)
    whole = T.run(api, fmt, tight, prof)
    assert whole.size == bh * bw * bpb
    assert np.array_equal(T.run(api, fmt, img, prof), whole), "stride invariance"
    assert np.array_equal(T.run(api, fmt, tight, prof), whole), "determinism"
    # any split into bands of whole block rows
    edges = sorted({0, bh, *[c % (bh + 1) for c in cuts]})
    parts = [T.run(api, fmt, np.ascontiguousarray(img[4 * a:4 * b]), prof) for a, b in zip(edges, edges[1:]) if b > a]
    assert np.array_equal(np.concatenate(parts), whole), "band split invariance"
    # permute the blocks of the image: the output blocks permute the same way
    perm = np.random.default_rng(seed + 1).permutation(bh * bw)
    blocks = tight.reshape(bh, 4, bw, 4, 4).transpose(0, 2, 1, 3, 4).reshape(bh * bw, 4, 4, 4)
    shuffled = blocks[perm].reshape(bh, bw, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(4 * bh, 4 * bw, 4)
    got = T.run(api, fmt, np.ascontiguousarray(shuffled), prof).reshape(-1, bpb)
    assert np.array_equal(got, whole.reshape(-1, bpb)[perm]), "block permutation equivariance"
    return whole, tight


@settings(max_examples=25, **COMMON)
@given(case=st.sampled_from(CASES), bh=st.integers(1, 6), bw=st.integers(1, 9), seed=st.integers(0, 2 ** 31), pad=st.integers(0, 5),
       kind=st.integers(0, 2), cuts=st.lists(st.integers(0, 64), max_size=4))
def test_oracle_invariants(case, bh, bw, seed, pad, kind, cuts):
    fmt, prof = case
    whole, tight = check_invariants(T.oracle(), fmt, prof, bh, bw, seed, pad, kind, cuts)
    # every block the encoder emits is a legal block of its format (the independent decoder raises on reserved modes)
    import bcn_decode as D
    import os
    layouts = D.bc6_layouts(open(os.path.join(T.ROOT, "oracle", "itw_oracle.cpp")).read()) if fmt == "BC6H" else None
    D.decode_image(fmt, whole.tobytes(), 4 * bw, 4 * bh, layouts)


@settings(max_examples=20, **COMMON)
@given(case=st.sampled_from(CASES), bh=st.integers(1, 5), bw=st.integers(1, 9), seed=st.integers(0, 2 ** 31), pad=st.integers(0, 5),
       kind=st.integers(0, 2), cuts=st.lists(st.integers(0, 64), max_size=3))
def test_emulated_kernels_invariants_and_oracle_parity(case, bh, bw, seed, pad, kind, cuts):
    fmt, prof = case
    whole, tight = check_invariants(T.emu(), fmt, prof, bh, bw, seed, pad, kind, cuts)
    assert np.array_equal(whole, T.run(T.oracle(), fmt, tight, prof)), "kernel logic == oracle"


@settings(max_examples=200, **COMMON)
@given(w=st.integers(1, 5000), h=st.integers(1, 5000), mips=st.integers(1, 13), fmt=st.sampled_from([71, 72, 77, 78, 80, 83, 95, 96, 98, 99]),
       cube=st.booleans(), items=st.integers(1, 3))
def test_dds_header_round_trip(w, h, mips, fmt, cube, items):
    lib = T.product().lib
    array = 6 * items if cube else items
    d = B.DdsDesc(w, h, min(mips, max(w, h).bit_length()), array, fmt, 1 if cube else 0)
    n = lib.itw_dds_header_bytes(ctypes.byref(d))
    assert n in (128, 148)
    buf = (ctypes.c_uint8 * n)()
    assert lib.itw_dds_write_header(ctypes.byref(d), buf, n) == n
    back = B.DdsDesc()
    assert lib.itw_dds_read_header(buf, n, ctypes.byref(back)) == n
    assert (back.width, back.height, back.mip_levels, back.array_size, back.is_cubemap) == (d.width, d.height, d.mip_levels, d.array_size, d.is_cubemap)
    assert back.dxgi_format == d.dxgi_format
    # the payload is the sum of the tightly packed images, item-major / mip-minor
    total = sum(lib.itw_dds_image_bytes(ctypes.byref(d), m) for m in range(d.mip_levels)) * array
    assert lib.itw_dds_file_bytes(ctypes.byref(d)) == n + total
    assert lib.itw_dds_image_offset(ctypes.byref(d), array - 1, d.mip_levels - 1) + lib.itw_dds_image_bytes(ctypes.byref(d), d.mip_levels - 1) == n + total


@settings(max_examples=200, **COMMON)
@given(wb=st.integers(1, 600), hq=st.integers(1, 200), nranks=st.sampled_from([1, 2, 3, 4, 8]), fmt=st.sampled_from(["BC1", "BC3", "BC7"]), data=st.data())
def test_shard_plan_tiles_the_chain(wb, hq, nranks, fmt, data):
    """itw_shard_plan_make: bands cover level 0 exactly, the per-rank bands of every band-local level tile that level of the chain
    without gap or overlap, slots are 16-byte multiples and hold everything a rank sends."""
    import importlib
    sharding = importlib.import_module("intel-texture-works-plugin_b200.sharding")
    lib = T.product()
    w, h = 4 * wb, 4 * hq * nranks
    levels = data.draw(st.integers(1, max(w, h).bit_length()))
    plans = [sharding.make_plan(lib, fmt, w, h, levels, nranks, r) for r in range(nranks)]
    bpb = B.FORMATS[fmt][1]
    assert [(p.band_y0, p.band_y1) for p in plans] == [sharding.band_rows(h, nranks, r) for r in range(nranks)]
    assert plans[0].band_y0 == 0 and plans[-1].band_y1 == h and all(a.band_y1 == b.band_y0 for a, b in zip(plans, plans[1:]))
    p = plans[0]
    off = 0
    for l in range(levels):
        pw, ph = (max(w >> l, 1) + 3) // 4, (max(h >> l, 1) + 3) // 4
        assert p.level_offset[l] == off and p.level_bytes[l] == pw * ph * bpb
        off += p.level_bytes[l]
        if l < p.band_levels:
            assert p.band_bytes[l] * nranks == p.level_bytes[l]
    assert p.chain_bytes == off and 1 <= p.band_levels <= levels
    sent = sum(p.band_bytes[l] for l in range(p.band_levels)) + p.texel_bytes
    assert p.slot_bytes % 16 == 0 and sent <= p.slot_bytes < sent + 48
    assert (p.texel_bytes == 0) == (p.band_levels == levels)


@pytest.mark.gpu
@settings(max_examples=30, **COMMON)
@given(case=st.sampled_from(CASES), bh=st.integers(1, 40), bw=st.integers(1, 70), seed=st.integers(0, 2 ** 31), pad=st.integers(0, 5),
       kind=st.integers(0, 2), cuts=st.lists(st.integers(0, 64), max_size=4))
def test_gpu_invariants_and_oracle_parity(case, bh, bw, seed, pad, kind, cuts):
    fmt, prof = case
    whole, tight = check_invariants(T.product(), fmt, prof, bh, bw, seed, pad, kind, cuts)
    if bh * bw <= 400:
        assert np.array_equal(whole, T.run(T.oracle(), fmt, tight, prof)), "GPU == oracle"
