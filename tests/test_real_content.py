"""Real-content check: the reference's own sample images through the reference-source build, the oracle and
the emulated kernels.  The images live under /root/reference/Sample Images and are NOT copied into the repository, so
this test runs only where the reference is mounted (it is skipped on the GPU box)."""
import os
import struct

import numpy as np
import pytest

import itw_testlib as T

SAMPLES = "/root/reference/Sample Images"
pytestmark = pytest.mark.skipif(not os.path.isdir(SAMPLES), reason="reference sample images not present")


def load_rgba8(name, crop=None):
    from PIL import Image
    img = np.array(Image.open(os.path.join(SAMPLES, name)).convert("RGBA"))
    if crop:
        y, x, h, w = crop
        img = img[y:y + h, x:x + w]
    h, w = (img.shape[0] // 4) * 4, (img.shape[1] // 4) * 4
    return np.ascontiguousarray(img[:h, :w])


def load_radiance_hdr(name, crop):
    """Minimal Radiance RGBE reader (new-style RLE scanlines) -> RGBA16F half bits."""
    data = open(os.path.join(SAMPLES, name), "rb").read()
    pos = data.index(b"\n\n") + 2
    end = data.index(b"\n", pos)
    tokens = data[pos:end].split()
    assert tokens[0] == b"-Y" and tokens[2] == b"+X", tokens
    h, w = int(tokens[1]), int(tokens[3])
    pos = end + 1
    rows = []
    y0, x0, ch, cw = crop
    for y in range(min(h, y0 + ch)):
        assert data[pos] == 2 and data[pos + 1] == 2 and ((data[pos + 2] << 8) | data[pos + 3]) == w
        pos += 4
        line = np.zeros((4, w), np.uint8)
        for c in range(4):
            x = 0
            while x < w:
                n = data[pos]
                pos += 1
                if n > 128:
                    line[c, x:x + n - 128] = data[pos]
                    pos += 1
                    x += n - 128
                else:
                    line[c, x:x + n] = np.frombuffer(data[pos:pos + n], np.uint8)
                    pos += n
                    x += n
        if y >= y0:
            rows.append(line[:, x0:x0 + cw].copy())
    rgbe = np.stack(rows).transpose(0, 2, 1).astype(np.float32)              # H x W x 4
    scale = np.where(rgbe[..., 3] > 0, np.exp2(rgbe[..., 3] - 136.0), 0.0).astype(np.float32)
    rgb = rgbe[..., :3] * scale[..., None]
    out = np.zeros(rgb.shape[:2] + (4,), np.float16)
    out[..., :3] = np.clip(rgb, 0, 65504).astype(np.float16)
    out[..., 3] = 1.0
    return np.ascontiguousarray(out.view(np.uint16))


LDR = [("baboon.png", None), ("gradients.png", None), ("colors-260K.png", (128, 128, 128, 128)), ("colors-16M.png", (1024, 2048, 64, 128)),
       ("juggling-balls.jpg", (200, 300, 128, 128))]


@pytest.mark.parametrize("name,crop", LDR)
def test_ldr_samples_reference_build_oracle_and_emulated_kernels_agree(name, crop):
    ref, o, e = T.ref(), T.oracle(), T.emu()
    img = load_rgba8(name, crop)
    for fmt, prof in (("BC1", None), ("BC3", None), ("BC7", "slow"), ("BC7", "alpha_basic"), ("BC7", "veryfast")):
        want = T.run(ref, fmt, img, prof)
        assert np.array_equal(T.run(o, fmt, img, prof), want), (name, fmt, prof, "oracle")
        assert np.array_equal(T.run(e, fmt, img, prof), want), (name, fmt, prof, "emulated kernel")
    for fmt in ("BC4", "BC5"):
        assert np.array_equal(T.run(e, fmt, img, None), T.run(o, fmt, img, None)), (name, fmt)


def test_hdr_sample_reference_build_oracle_and_emulated_kernel_agree():
    ref, o, e = T.ref(), T.oracle(), T.emu()
    img = load_radiance_hdr("HDR.hdr", (64, 64, 128, 128))
    assert len(np.unique(img[..., :3])) > 300                            # a real HDR crop, not a flat region
    for prof in ("bc6h_slow", "bc6h_basic", "bc6h_veryfast"):
        want = T.run(ref, "BC6H", img, prof)
        assert np.array_equal(T.run(o, "BC6H", img, prof), want), prof
        assert np.array_equal(T.run(e, "BC6H", img, prof), want), prof
