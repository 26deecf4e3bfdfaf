"""Synthetic inputs of the BASELINE.json configs, integer-only so that every
implementation of the generator agrees.  numpy on the host; the bench moves them to the device."""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(x):
    """Vectorised splitmix64 finaliser of uint64 counters (pixel i of seed S uses counter S+i)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _counters(seed, h, w):
    with np.errstate(over="ignore"):
        return np.uint64(seed & MASK64) + np.arange(h * w, dtype=np.uint64)


def random_rgba8(h, w, seed=0xB2000002):
    """C2: every channel uniform 0..255 (low 32 bits of the hash)."""
    z = splitmix64(_counters(seed, h, w))
    return (z & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.uint8).reshape(h, w, 4).copy()


def gradient_rgba8(h, w, phase=0):
    """C1: R=(255x)/(w-1), G=(255y)/(h-1), B=(255(x+y))/(w+h-2), A=255 (optionally phase-shifted)."""
    x = (np.arange(w, dtype=np.int64)[None, :] + phase) % w
    y = (np.arange(h, dtype=np.int64)[:, None] + phase) % h
    img = np.empty((h, w, 4), np.uint8)
    img[..., 0] = (255 * x) // max(w - 1, 1)
    img[..., 1] = (255 * y) // max(h - 1, 1)
    img[..., 2] = (255 * (x + y)) // max(w + h - 2, 1)
    img[..., 3] = 255
    return img


def random_rgba16f(h, w, seed=0xB2000003):
    """C3: half bit patterns uniform over the non-negative finite halves [0, 0x7C00), A = 1.0."""
    z = splitmix64(_counters(seed, h, w))
    img = np.empty((h, w, 4), np.uint16)
    for c in range(3):
        img[..., c] = (((z >> np.uint64(16 * c)) & np.uint64(0xFFFF)) % np.uint64(0x7C00)).astype(np.uint16).reshape(h, w)
    img[..., 3] = 0x3C00
    return img


def smooth_rgba16f(h, w):
    """Secondary HDR input: half(2^((x+y)/1024 - 4) * (1 + c/8))."""
    x = np.arange(w, dtype=np.float32)[None, :]
    y = np.arange(h, dtype=np.float32)[:, None]
    v = np.exp2((x + y) / 1024.0 - 4.0)
    img = np.stack([v * (1.0 + c / 8.0) for c in range(4)], -1).astype(np.float16)
    return img.view(np.uint16)


def mixed_rgba8(h, w, seed=0xB2000004):
    """C4 level 0: R=(x^y)&255, G=((3x+5y)>>6)&255, B=noise, A=((x+y)>>6)&255."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    img = np.empty((h, w, 4), np.uint8)
    img[..., 0] = (x ^ y) & 255
    img[..., 1] = ((3 * x + 5 * y) >> 6) & 255
    img[..., 2] = (splitmix64(_counters(seed, h, w)) & np.uint64(255)).astype(np.uint8).reshape(h, w)
    img[..., 3] = ((x + y) >> 6) & 255
    return img


def c5_tile(t, size=1024):
    """C5 tile t: even t = the C2 generator with seed 0xB2000005 + t, odd t = the C1 gradient with a
    per-tile phase."""
    if t % 2 == 0:
        return random_rgba8(size, size, seed=0xB2000005 + t)
    return gradient_rgba8(size, size, phase=37 * t)


def box_mip(img):
    """Next mip level of a SQUARE POWER-OF-TWO texture (config C4): the integer form (a+b+c+d+2)>>2 of DirectXTex's box filter, which
    is what the library runs on exact 2:1 levels (include/itw_bcn.h section 4; other shapes follow DirectXTex's linear filter / its
    one-texel-high quirk, see tests/itw_testlib.oracle_mip_chain_rgba8).  max(1, w>>1) x max(1, h>>1);
    a 1-texel-wide/high level degenerates to the 2-tap average; an odd trailing row/column is dropped (floor)."""
    h, w = img.shape[:2]
    a = img.astype(np.uint16)
    if h > 1:
        a = a[: h // 2 * 2]
    if w > 1:
        a = a[:, : w // 2 * 2]
    if h > 1 and w > 1:
        s = a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2]
        return ((s + 2) >> 2).astype(np.uint8)
    if h > 1:
        return ((a[0::2] + a[1::2] + 1) >> 1).astype(np.uint8)
    if w > 1:
        return ((a[:, 0::2] + a[:, 1::2] + 1) >> 1).astype(np.uint8)
    return img.copy()


def pad_to_4(img):
    """Edge-replicate to multiples of 4 (IntelPlugin.cpp:893-928)."""
    h, w = img.shape[:2]
    ph, pw = (-h) % 4, (-w) % 4
    if ph == 0 and pw == 0:
        return img
    return np.pad(img, ((0, ph), (0, pw), (0, 0)), mode="edge")


def mip_chain(img):
    """All levels down to 1x1, each padded to a multiple of 4 (config C4)."""
    out = [pad_to_4(img)]
    cur = img
    while cur.shape[0] > 1 or cur.shape[1] > 1:
        cur = box_mip(cur)
        out.append(pad_to_4(cur))
    return out
