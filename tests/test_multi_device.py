"""One process, several GPUs (itw_set_devices), the reference's coarse seam (CompressImage*, win32Threads.h) and the
deferred mode for the reference's unchanged slice loop (IntelPlugin.cpp:851-879)."""
import ctypes
import time

import numpy as np
import pytest

import itw_testlib as T

pytestmark = pytest.mark.gpu
B = T.binding


def _devices():
    import torch
    return torch.cuda.device_count()


def test_fanout_over_devices_gives_the_single_device_bytes():
    """Host -> host calls cut into one band per device: identical bytes.  With one GPU the pool degenerates to the default
    device; with 2+ the bands really run on different GPUs."""
    lib = T.product()
    n = _devices()
    cases = [("BC7", "veryfast", T.synth.random_rgba8(1024, 512)), ("BC6H", "bc6h_veryfast", T.synth.random_rgba16f(768, 256)),
             ("BC1", None, T.synth.random_rgba8(2048, 1024)), ("BC3", None, T.synth.random_rgba8(516, 1024))]
    want = [lib.encode(f, im, lib.profile(p) if p else None) for f, p, im in cases]
    try:
        lib.set_devices(list(range(n)))
        for (f, p, im), w in zip(cases, want):
            assert np.array_equal(lib.encode(f, im, lib.profile(p) if p else None), w), f
        # tile stream over the devices
        tiles = [T.synth.c5_tile(t, 256) for t in range(11)]
        s = lib.profile("veryfast")
        single = [lib.encode("BC7", t, s) for t in tiles]
        outs = [np.zeros_like(x) for x in single]
        lib.encode_batch("BC7", [(t.ctypes.data, 256, 256, t.strides[0]) for t in tiles], [o.ctypes.data for o in outs], s)
        assert all(np.array_equal(a, b) for a, b in zip(outs, single))
        # an error in one band is reported
        bad = lib.profile("slow")
        bad.fastSkipTreshold_mode1 = 65
        with pytest.raises(RuntimeError, match="fastSkipTreshold"):
            lib.encode("BC7", cases[0][2], bad)
    finally:
        lib.set_devices([])
    assert np.array_equal(lib.encode("BC1", cases[2][2]), want[2])


def test_multi_gpu_e2e_scales():
    """2+ GPUs: a BC7 basic 4096 x 4096 host surface through ONE call; the fan-out must beat one GPU clearly."""
    import torch
    n = _devices()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    lib = T.product()
    img = torch.from_numpy(T.synth.random_rgba8(4096, 4096).reshape(-1)).pin_memory()
    out = torch.empty(1024 * 1024 * 16, dtype=torch.uint8).pin_memory()
    s = lib.profile("basic")

    def run():
        lib.encode_raw("BC7", img.data_ptr(), 4096, 4096, 4096 * 4, out.data_ptr(), s)
    run()
    one = min(_timed(run) for _ in range(3))
    ref = out.numpy().copy()
    try:
        lib.set_devices(list(range(n)))
        run()
        many = min(_timed(run) for _ in range(3))
    finally:
        lib.set_devices([])
    assert np.array_equal(out.numpy(), ref)
    assert many < one / (0.5 * n) + 1e-3, (one, many, n)         # at least half of linear scaling (measured: 0.8-1.0)


def test_zero_height_band_is_a_no_op():
    """CompressImageMT hands zero-height bands to the threads beyond height/4 (win32Threads.cpp:217-230); the reference's
    loops simply do not run.  No error, nothing written."""
    lib = T.product()
    img = T.synth.random_rgba8(8, 16)
    out = np.full(64, 0xAB, np.uint8)
    lib.encode_raw("BC1", img.ctypes.data, 16, 0, img.strides[0], out.ctypes.data, None)
    assert lib.last_error() == "" and (out == 0xAB).all()


PROFILE_FUNCS = ([("BC1", None, "CompressImageBC1"), ("BC3", None, "CompressImageBC3")]
                 + [("BC7", p, "CompressImageBC7_" + p) for p in B.BC7_PROFILES]
                 + [("BC6H", p, "CompressImage" + p.replace("bc6h_", "BC6H_")) for p in B.BC6H_PROFILES])


def test_compress_image_wrappers_equal_profile_plus_compress_blocks():
    """win32Threads.cpp:289-330: CompressImageBC7_<profile> = GetProfile_<profile> + CompressBlocksBC7, through MT and ST."""
    lib = T.product()
    L = lib.lib
    img8, img16 = T.synth.random_rgba8(32, 64, seed=5), T.synth.random_rgba16f(32, 64, seed=6)
    L.CompressImageMT.restype = ctypes.c_bool
    L.CompressImageST.restype = ctypes.c_bool
    for fmt, prof, name in PROFILE_FUNCS:
        img = img16 if fmt == "BC6H" else img8
        want = lib.encode(fmt, img, lib.profile(prof) if prof else None)
        surf = B.RgbaSurface(img.ctypes.data, 64, 32, img.strides[0])
        fn = getattr(L, name)
        fn.restype = None
        for entry in (L.CompressImageMT, L.CompressImageST):
            out = np.zeros_like(want)
            assert entry(ctypes.byref(surf), ctypes.c_void_p(out.ctypes.data), fn, B.FORMATS[fmt][0]) is True
            lib.check()
            assert np.array_equal(out, want), name
    assert [L.GetBytesPerBlock(f) for f in (71, 72, 77, 78, 95, 96, 98, 99, 80)] == [8, 8, 16, 16, 16, 16, 16, 16, 8]
    assert L.GetProcessorCount() == max(_devices(), 1)


def _slice_loop(lib, name, fmt, img, out, deferred):
    """The plug-in's loop, unchanged (IntelPlugin.cpp:851-879): slices of 256 K texels, one CompressImageMT each."""
    L = lib.lib
    h, w = img.shape[:2]
    texel, bpb = B.FORMATS[fmt][2], B.FORMATS[fmt][1]
    fn = getattr(L, name)
    slices = max((w * h) // 0x40000, 1)
    row_pitch = (w // 4) * bpb
    if deferred:
        lib.begin_deferred()
    for i in range(slices):
        ylo, yhi = (i * h // slices) & ~3, min(((i + 1) * h // slices) & ~3, h)
        if yhi > ylo:
            surf = B.RgbaSurface(img.ctypes.data + img.strides[0] * ylo, w, yhi - ylo, img.strides[0])
            L.CompressImageMT(ctypes.byref(surf), ctypes.c_void_p(out.ctypes.data + row_pitch * (ylo >> 2)), fn, B.FORMATS[fmt][0])
    if deferred:
        lib.flush()
    lib.check()
    return slices


@pytest.mark.parametrize("fmt,name,limit", [("BC7", "CompressImageBC7_basic", 2.0), ("BC1", "CompressImageBC1", 4.0)])
def test_reference_slice_loop_replayed(fmt, name, limit):
    """A 4096^2 image through the reference's own call pattern (64 slices): bytes equal the whole-image encode, and the time
    stays within `limit` x the single-call end-to-end time (deferred mode; the plain synchronous loop is reported too)."""
    import torch
    lib = T.product()
    img_t = torch.from_numpy(T.synth.random_rgba8(4096, 4096).reshape(4096, 4096, 4)).pin_memory()
    img = img_t.numpy()
    prof = name.split("_", 1)[1] if fmt == "BC7" else None
    settings = lib.profile(prof) if prof else None
    bpb = B.FORMATS[fmt][1]
    whole_t = torch.empty(1024 * 1024 * bpb, dtype=torch.uint8).pin_memory()
    whole = whole_t.numpy()

    def single():
        lib.encode_raw(fmt, img.ctypes.data, 4096, 4096, 4096 * 4, whole.ctypes.data, settings)
    single()
    t_single = min(_timed(single) for _ in range(3))
    out_t = torch.zeros(1024 * 1024 * bpb, dtype=torch.uint8).pin_memory()
    out = out_t.numpy()
    times = {}
    for deferred in (False, True):
        out[:] = 0
        _slice_loop(lib, name, fmt, img, out, deferred)            # warm
        times[deferred] = min(_timed(lambda: _slice_loop(lib, name, fmt, img, out, deferred)) for _ in range(3))
        assert np.array_equal(out, whole), ("deferred" if deferred else "synchronous")
    print(f"\n{fmt}: single call {t_single * 1e3:.2f} ms, 64-slice loop {times[False] * 1e3:.2f} ms synchronous, "
          f"{times[True] * 1e3:.2f} ms deferred")
    assert times[True] <= limit * t_single, (times, t_single)


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def test_deferred_mode_reports_the_first_failure_at_flush():
    lib = T.product()
    img = T.synth.random_rgba8(64, 64)
    out = np.zeros(16 * 16 * 16, np.uint8)
    bad = lib.profile("slow")
    bad.fastSkipTreshold_mode3 = 99
    lib.begin_deferred()
    lib.lib.CompressBlocksBC7(ctypes.byref(B.RgbaSurface(img.ctypes.data, 64, 64, img.strides[0])), ctypes.c_void_p(out.ctypes.data), ctypes.byref(bad))
    with pytest.raises(RuntimeError, match="fastSkipTreshold"):
        lib.flush()
    assert np.array_equal(lib.encode("BC7", img, lib.profile("veryfast")), T.run(T.oracle(), "BC7", img, "veryfast"))


def test_texture_save_path_over_devices_equals_single_device():
    """itw_dds_encode_texture (level 0 in host memory -> mips, encode, DDS) fanned over the selected GPUs: the same file."""
    lib = T.product()
    n = _devices()
    cases = [(1024, 1024, 77, None, 1, 0), (2048, 1024, 72, None, 1, 0), (1024, 1024, 99, "veryfast", 1, 0), (1024, 1024, 71, None, 6, 1), (1536, 1024, 78, None, 1, 0)]
    for (w, h, fmt, prof, items, cube) in cases:
        tops = [T.synth.mixed_rgba8(h, w, seed=s) for s in range(items)]
        d = B.DdsDesc(w, h, max(w, h).bit_length(), items, fmt, cube)
        s = lib.profile(prof) if prof else None
        want = lib.dds_encode_texture(d, tops, s)
        try:
            lib.set_devices(list(range(n)))
            got = lib.dds_encode_texture(d, tops, s)
        finally:
            lib.set_devices([])
        assert np.array_equal(got, want), (w, h, fmt)
