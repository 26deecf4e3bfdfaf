"""Independent BCn DECODERS (numpy / pure Python), test infrastructure only.

Written from the format definitions (the D3D BC1-BC7 specification; DirectXTex's decoders
DirectXTex/BC.cpp:322, :897, BC4BC5.cpp:42-95, BC6HBC7.cpp:1077, :1937 implement the same formats), NOT from the
encoder: they are used to check that what the oracle / kernels emit are legal bit streams that decode back to
something close to the input ("decode validity").  Small images only (pure Python loops)."""
import numpy as np

W2 = [0, 21, 43, 64]
W3 = [0, 9, 18, 27, 37, 46, 55, 64]
W4 = [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]
WEIGHTS = {2: W2, 3: W3, 4: W4}

# two-subset shapes as bit masks (bit k = texel k in subset 1), three-subset shapes as 2 bits/texel, anchors:
# the BC7 / BC6H partition tables of the format definition
SHAPE2 = [0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8, 0xFF00, 0xFFF0, 0xF000,
          0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110, 0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C,
          0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696, 0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660,
          0x0272, 0x04E4, 0x4E40, 0x2720, 0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22]
SHAPE3 = [0xAA685050, 0x6A5A5040, 0x5A5A4200, 0x5450A0A8, 0xA5A50000, 0xA0A05050, 0x5555A0A0, 0x5A5A5050, 0xAA550000, 0xAA555500, 0xAAAA5500,
          0x90909090, 0x94949494, 0xA4A4A4A4, 0xA9A59450, 0x2A0A4250, 0xA5945040, 0x0A425054, 0xA5A5A500, 0x55A0A0A0, 0xA8A85454, 0x6A6A4040,
          0xA4A45000, 0x1A1A0500, 0x0050A4A4, 0xAAA59090, 0x14696914, 0x69691400, 0xA08585A0, 0xAA821414, 0x50A4A450, 0x6A5A0200, 0xA9A58000,
          0x5090A0A8, 0xA8A09050, 0x24242424, 0x00AA5500, 0x24924924, 0x24499224, 0x50A50A50, 0x500AA550, 0xAAAA4444, 0x66660000, 0xA5A0A5A0,
          0x50A050A0, 0x69286928, 0x44AAAA44, 0x66666600, 0xAA444444, 0x54A854A8, 0x95809580, 0x96969600, 0xA85454A8, 0x80959580, 0xAA141414,
          0x96960000, 0xAAAA1414, 0xA05050A0, 0xA0A5A5A0, 0x96000000, 0x40804080, 0xA9A8A9A8, 0xAAAAAA44, 0x2A4A5254]
ANCHOR2 = [15] * 16 + [15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2, 15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6,
                       6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15]
ANCHOR3A = [3, 3, 15, 15, 8, 3, 15, 15, 8, 8, 6, 6, 6, 5, 3, 3, 3, 3, 8, 15, 3, 3, 6, 10, 5, 8, 8, 6, 8, 5, 15, 15,
            8, 15, 3, 5, 6, 10, 8, 15, 15, 3, 15, 5, 15, 15, 15, 15, 3, 15, 5, 5, 5, 8, 5, 10, 5, 10, 8, 13, 15, 12, 3, 3]
ANCHOR3B = [15, 8, 8, 3, 15, 15, 3, 8, 15, 15, 15, 15, 15, 15, 15, 8, 15, 8, 15, 3, 15, 8, 15, 8, 3, 15, 6, 10, 15, 15, 10, 8,
            15, 3, 15, 10, 10, 8, 9, 10, 6, 15, 8, 15, 3, 6, 6, 8, 15, 3, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 3, 15, 15, 8]


class Bits:
    def __init__(self, block):
        self.v = int.from_bytes(bytes(block), "little")
        self.pos = 0

    def get(self, n):
        r = (self.v >> self.pos) & ((1 << n) - 1)
        self.pos += n
        return r


def subset_of(ns, shape, k):
    if ns == 1:
        return 0
    if ns == 2:
        return (SHAPE2[shape] >> k) & 1
    return (SHAPE3[shape] >> (2 * k)) & 3


def anchors_of(ns, shape):
    if ns == 1:
        return [0]
    if ns == 2:
        return [0, ANCHOR2[shape]]
    return [0, ANCHOR3A[shape], ANCHOR3B[shape]]


# ---- BC1 / BC3 / BC4 / BC5 ----
def _unorm8(x):
    return int(np.float32(np.float32(x) * np.float32(255.0)) + np.float32(0.5))


def decode_bc1_block(blk, force4=False):
    """DirectXTex DecodeBC1 (BC.cpp:322-370) in float32, stored as x * 255 + 0.5 truncated; NOT 565 bit replication."""
    c0, c1 = int.from_bytes(bytes(blk[0:2]), "little"), int.from_bytes(bytes(blk[2:4]), "little")
    idx = int.from_bytes(bytes(blk[4:8]), "little")
    four = c0 > c1 or force4
    pal = np.zeros((4, 4), np.int32)
    pal[:, 3] = 255
    if not four:
        pal[3, 3] = 0
    for c, (shift, mask, scale) in enumerate(((11, 31, np.float32(1.0) / np.float32(31.0)), (5, 63, np.float32(1.0) / np.float32(63.0)),
                                              (0, 31, np.float32(1.0) / np.float32(31.0)))):
        f0 = np.float32((c0 >> shift) & mask) * scale
        f1 = np.float32((c1 >> shift) & mask) * scale
        ln = np.float32(f1 - f0)
        third, two_thirds = np.float32(1.0) / np.float32(3.0), np.float32(2.0) / np.float32(3.0)
        f2 = np.float32(np.float32(ln * third) + f0) if four else np.float32(np.float32(ln * np.float32(0.5)) + f0)
        f3 = np.float32(np.float32(ln * two_thirds) + f0)
        pal[0, c], pal[1, c], pal[2, c] = _unorm8(f0), _unorm8(f1), _unorm8(f2)
        pal[3, c] = _unorm8(f3) if four else 0
    out = np.zeros((16, 4), np.int32)
    for k in range(16):
        out[k] = pal[(idx >> (2 * k)) & 3]
    return out


def decode_alpha_block(blk):
    a0, a1 = int(blk[0]), int(blk[1])
    idx = int.from_bytes(bytes(blk[2:8]), "little")
    if a0 > a1:
        pal = [a0, a1] + [((7 - i) * a0 + i * a1 + 3) // 7 for i in range(1, 7)]
    else:
        pal = [a0, a1] + [((5 - i) * a0 + i * a1 + 2) // 5 for i in range(1, 5)] + [0, 255]
    return np.array([pal[(idx >> (3 * k)) & 7] for k in range(16)], np.int32)


# ---- BC7 ----
# mode: (subsets, partition bits, rotation bits, index-selection bits, colour bits, alpha bits, endpoint p-bits, shared p-bits, index bits, 2nd index bits)
BC7_MODES = [(3, 4, 0, 0, 4, 0, 1, 0, 3, 0), (2, 6, 0, 0, 6, 0, 0, 1, 3, 0), (3, 6, 0, 0, 5, 0, 0, 0, 2, 0), (2, 6, 0, 0, 7, 0, 1, 0, 2, 0),
             (1, 0, 2, 1, 5, 6, 0, 0, 2, 3), (1, 0, 2, 0, 7, 8, 0, 0, 2, 2), (1, 0, 0, 0, 7, 7, 1, 0, 4, 0), (2, 6, 0, 0, 5, 5, 1, 0, 2, 0)]


def decode_bc7_block(blk):
    """-> (16x4 int array, mode); raises ValueError for the reserved mode."""
    b = Bits(blk)
    mode = 0
    while mode < 8 and b.get(1) == 0:
        mode += 1
    if mode == 8:
        raise ValueError("reserved BC7 mode")
    ns, pbits, rbits, isb, cb, ab, epb, spb, ib, ib2 = BC7_MODES[mode]
    shape = b.get(pbits)
    rot = b.get(rbits)
    isel = b.get(isb)
    ne = 2 * ns
    ep = np.zeros((ne, 4), np.int32)
    for c in range(3):
        for e in range(ne):
            ep[e, c] = b.get(cb)
    for e in range(ne):
        ep[e, 3] = b.get(ab) if ab else 255
    cbits, abits = cb, ab
    if epb:
        for e in range(ne):
            p = b.get(1)
            ep[e, :3] = (ep[e, :3] << 1) | p
            if ab:
                ep[e, 3] = (ep[e, 3] << 1) | p
        cbits += 1
        abits += 1 if ab else 0
    if spb:
        for s in range(ns):
            p = b.get(1)
            for e in (2 * s, 2 * s + 1):
                ep[e, :3] = (ep[e, :3] << 1) | p
        cbits += 1
    ep[:, :3] = (ep[:, :3] << (8 - cbits)) | (ep[:, :3] >> (2 * cbits - 8))
    if ab:
        ep[:, 3] = (ep[:, 3] << (8 - abits)) | (ep[:, 3] >> (2 * abits - 8))
    anchors = anchors_of(ns, shape)

    def read_indices(bits, multi):
        out = []
        for k in range(16):
            s = subset_of(ns, shape, k) if multi else 0
            is_anchor = (k == anchors[s]) if multi else (k == 0)
            out.append(b.get(bits - 1 if is_anchor else bits))
        return out
    i1 = read_indices(ib, True)
    i2 = read_indices(ib2, False) if ib2 else None
    assert b.pos == 128, (mode, b.pos)
    out = np.zeros((16, 4), np.int32)
    for k in range(16):
        s = subset_of(ns, shape, k)
        e0, e1 = ep[2 * s], ep[2 * s + 1]
        if i2 is None:
            w = WEIGHTS[ib][i1[k]]
            out[k] = ((64 - w) * e0 + w * e1 + 32) >> 6
        else:
            ci, ai, cbw, abw = (i1[k], i2[k], ib, ib2) if not isel else (i2[k], i1[k], ib2, ib)
            wc, wa = WEIGHTS[cbw][ci], WEIGHTS[abw][ai]
            out[k, :3] = ((64 - wc) * e0[:3] + wc * e1[:3] + 32) >> 6
            out[k, 3] = ((64 - wa) * e0[3] + wa * e1[3] + 32) >> 6
        if rot:
            out[k, rot - 1], out[k, 3] = out[k, 3], out[k, rot - 1]
    return out, mode


# ---- BC6H (unsigned) ----
# 5-bit (or 2-bit) mode field -> (our mode number, endpoint bits, delta bits r,g,b, transformed, regions); layouts are read from
# the oracle's table so that this decoder stays consistent with what tests/test_tables.py checks
BC6_PREFIX = {0: 0, 1: 1, 2: 2, 6: 3, 10: 4, 14: 5, 18: 6, 22: 7, 26: 8, 30: 9, 3: 10, 7: 11, 11: 12, 15: 13}
BC6_EPB = [10, 7, 11, 11, 11, 9, 8, 8, 8, 6, 10, 11, 12, 16]
BC6_DELTA = [(5, 5, 5), (6, 6, 6), (5, 4, 4), (4, 5, 4), (4, 4, 5), (5, 5, 5), (6, 5, 5), (5, 6, 5), (5, 5, 6), None, None, (9, 9, 9), (8, 8, 8), (4, 4, 4)]


def bc6_layouts(oracle_src):
    import re
    rows = re.findall(r'/\*\s*(\d+)\*/\s*"([^"]+)"', oracle_src)
    out = {}
    for n, text in rows:
        steps = []
        for tok in text.split():
            name, rng = tok.split(".")
            a, _, bb = rng.partition("-")
            a = int(a)
            bb = int(bb) if bb else a
            seq = list(range(a, bb + 1)) if bb >= a else list(range(a, bb - 1, -1))
            steps += [(name, bit) for bit in seq]
        out[int(n)] = steps
    return out


def _unq(v, bits):
    if bits >= 15:
        return v
    if v == 0:
        return 0
    if v == (1 << bits) - 1:
        return 0xFFFF
    return ((v << 15) + 0x4000) >> (bits - 1)


def decode_bc6h_block(blk, layouts):
    """-> (16x3 array of half bit patterns, mode)."""
    v = int.from_bytes(bytes(blk), "little")
    m2 = v & 3
    field = m2 if m2 < 2 else (v & 31)
    if field not in BC6_PREFIX:
        raise ValueError("reserved BC6H mode")
    mode = BC6_PREFIX[field]
    steps = layouts[mode]
    comp = {}
    for pos, (name, bit) in enumerate(steps):
        if name == "m":
            continue
        comp[name] = comp.get(name, 0) | (((v >> pos) & 1) << bit)
    regions = 2 if mode < 10 else 1
    epb = BC6_EPB[mode]
    e = np.zeros((4, 3), np.int64)
    for i in range(2 * regions):
        for c, ch in enumerate("rgb"):
            e[i, c] = comp.get(f"{ch}{i}", 0)
    if BC6_DELTA[mode] is not None:
        for i in range(1, 2 * regions):
            for c in range(3):
                d, nb = int(e[i, c]), BC6_DELTA[mode][c]
                if d & (1 << (nb - 1)):
                    d -= 1 << nb
                e[i, c] = (int(e[0, c]) + d) & ((1 << epb) - 1)
    pos = len(steps)
    shape = 0
    if regions == 2:
        shape = (v >> pos) & 31
        pos += 5
    ib = 3 if regions == 2 else 4
    anchors = anchors_of(regions, shape)
    out = np.zeros((16, 3), np.int64)
    for k in range(16):
        s = subset_of(regions, shape, k)
        nb = ib - 1 if k == anchors[s] else ib
        q = (v >> pos) & ((1 << nb) - 1)
        pos += nb
        w = WEIGHTS[ib][q]
        for c in range(3):
            a, b = _unq(int(e[2 * s, c]), epb), _unq(int(e[2 * s + 1, c]), epb)
            out[k, c] = ((((64 - w) * a + w * b + 32) >> 6) * 31) >> 6
    assert pos == 128, (mode, pos)
    return out, mode


# ---- whole images ----
def decode_image(fmt, data, w, h, layouts=None):
    """-> (H x W x C int array, list of per-block modes)."""
    bpb = 8 if fmt in ("BC1", "BC4") else 16
    blocks = np.frombuffer(bytes(data), np.uint8).reshape(-1, bpb)
    ch = 3 if fmt == "BC6H" else 4
    img = np.zeros((h, w, ch), np.int64)
    modes = []
    for i, blk in enumerate(blocks):
        by, bx = divmod(i, w // 4)
        if fmt == "BC1":
            px = decode_bc1_block(blk)
        elif fmt == "BC3":
            px = decode_bc1_block(blk[8:], force4=True)
            px[:, 3] = decode_alpha_block(blk[:8])
        elif fmt == "BC4":
            px = np.zeros((16, 4), np.int32)
            px[:, 0] = decode_alpha_block(blk)
            px[:, 3] = 255
        elif fmt == "BC5":
            px = np.zeros((16, 4), np.int32)
            px[:, 0] = decode_alpha_block(blk[:8])
            px[:, 1] = decode_alpha_block(blk[8:])
            px[:, 3] = 255
        elif fmt == "BC7":
            px, mode = decode_bc7_block(blk)
            modes.append(mode)
        else:
            px, mode = decode_bc6h_block(blk, layouts)
            modes.append(mode)
        img[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4] = px.reshape(4, 4, -1)
    return img, modes


def psnr(a, b, peak=255.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
