"""On-GPU pre-pass: mip chain by the integer 2x2 box filter + pad-to-4 by edge replication.
Contract = the numpy restatement in synth.box_mip / synth.pad_to_4 (IntelPlugin.cpp:893-928 for the padding)."""
import ctypes

import numpy as np
import pytest

import itw_testlib as T

D = T.binding.DdsDesc


def reference_chain(img, levels):
    out, cur = [T.synth.pad_to_4(img)], img
    for _ in range(1, levels):
        cur = T.synth.box_mip(cur)
        out.append(T.synth.pad_to_4(cur))
    return out


def full_levels(w, h):
    n, m = 1, max(w, h)
    while m > 1:
        m >>= 1
        n += 1
    return n


@pytest.mark.parametrize("h,w", [(64, 64), (32, 128), (5, 12), (7, 7), (1, 9), (100, 3), (4, 4)])
def test_emulated_mip_texel_matches_numpy(h, w):
    emu = T.emu().lib
    rng = np.random.default_rng(h * 1000 + w)
    cur = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    want = reference_chain(cur, full_levels(w, h))
    for l in range(1, len(want)):
        dh, dw = max(1, cur.shape[0] >> 1), max(1, cur.shape[1] >> 1)
        ph, pw = want[l].shape[:2]
        got = np.zeros((ph, pw, 4), np.uint8)
        src = np.ascontiguousarray(cur)
        emu.emu_mip_level(src.ctypes.data_as(ctypes.c_void_p), src.shape[1], src.shape[0], src.strides[0],
                          got.ctypes.data_as(ctypes.c_void_p), dw, dh, pw, ph)
        assert np.array_equal(got, want[l]), (l, dh, dw)
        cur = got[:dh, :dw]


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(256, 256), (64, 200), (37, 61), (4, 4), (1, 16)])
def test_gpu_mip_chain_matches_numpy(h, w):
    import torch
    lib = T.product()
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 4), dtype=np.uint8)
    levels = full_levels(w, h)
    want = reference_chain(img, levels)
    d_img = torch.from_numpy(img.reshape(-1)).cuda()
    pad0 = (w % 4 != 0) or (h % 4 != 0)
    nbytes = lib.lib.itw_mip_scratch_bytes(w, h, levels, 0 if pad0 else 1)
    assert nbytes == sum(x.size for x in want[(0 if pad0 else 1):])
    scratch = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device="cuda")
    outs = (T.binding.RgbaSurface * levels)()
    top = T.binding.RgbaSurface(d_img.data_ptr(), w, h, w * 4)
    assert lib.lib.itw_generate_mips_device(ctypes.byref(top), levels, outs, ctypes.c_void_p(scratch.data_ptr()),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    torch.cuda.synchronize()
    base = scratch.data_ptr()
    host = scratch.cpu().numpy()
    for l in range(levels):
        ph, pw = want[l].shape[:2]
        assert (outs[l].width, outs[l].height, outs[l].stride) == (pw, ph, pw * 4) or (l == 0 and not pad0)
        if l == 0 and not pad0:
            continue
        off = outs[l].ptr - base
        got = host[off:off + ph * pw * 4].reshape(ph, pw, 4)
        assert np.array_equal(got, want[l]), l


@pytest.mark.gpu
def test_whole_texture_save_path():
    """itw_dds_encode_texture: level 0 in, .dds out (mips made and encoded on the GPU) == numpy mips + per-level encodes."""
    lib = T.product()
    for (w, h, fmt, name, prof, items, cube) in ((128, 64, 77, "BC3", None, 1, 0), (60, 36, 98, "BC7", "veryfast", 1, 0), (32, 32, 71, "BC1", None, 6, 1)):
        tops = [T.synth.mixed_rgba8(h, w, seed=s) for s in range(items)]
        levels = full_levels(w, h)
        d = D(w, h, levels, items, fmt, cube)
        s = lib.profile(prof) if prof else None
        blob = lib.dds_encode_texture(d, tops, s)
        for item in range(items):
            chain = reference_chain(tops[item], levels)
            for mip in range(levels):
                off = lib.lib.itw_dds_image_offset(ctypes.byref(d), item, mip)
                want = lib.encode(name, np.ascontiguousarray(chain[mip]), s)
                assert np.array_equal(blob[off:off + want.size], want), (name, item, mip)
