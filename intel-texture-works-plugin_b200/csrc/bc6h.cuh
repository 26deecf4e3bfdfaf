// bc6h.cuh -- BC6H (UF16) encoder (reference: kernel.ispc:2039-3139, cited as K:line).
//
// Mapping (same scheme as bc7.cuh): one warp owns a batch of kBc6Slots blocks; lanes are spread
// over (block, shape) for the 32 PCA split bounds, over (block, mode, ranked shape) for the
// two-region candidates, and over (block, mode) for the refinement chains.  BC6H texels are
// non-integer floats up to 65535 and the per-texel error is truncated through an x86 float->int
// conversion that overflows to INT_MIN (K:1178, quirk Q3), so NOTHING here may be re-associated:
// every float sum stays inside one lane in the reference's texel order.
//
// The reference recomputes the 32 split bounds AND the PCA segments of each ranked shape for every
// two-region mode it tries (K:2257-2273, :2174-2193); both depend only on the texels, so they are
// computed once per block here (same values) and the six modes share them.  A chain ends with quantised endpoints and
// indices; only the block's winner is packed into 128 bits, one bit field per lane (pick / pack / store phases).
#pragma once
#include "bc67_core.cuh"

#ifndef ITW_BC6_ASSIGN_UNROLL
#define ITW_BC6_ASSIGN_UNROLL 16          // tuned on B200, tools/tune_unroll.sh (profiles/r2_tune_unroll.txt)
#endif

namespace itw {

struct Bc6Params {
    int slow_mode, fast_mode, refine_1p, refine_2p, fast_skip;
};

constexpr int kBc6Slots = 3;
constexpr int kBc6MaxTwo = 6;    // two-region modes tried per block
constexpr int kBc6MaxOne = 4;    // one-region modes tried per block

// One mode the block will be encoded with: K's mode index (2..4 / 6..8 carry the wide channel),
// endpoint precision and the quantised clamp window of K:2302-2330
struct Bc6Entry {
    int mode, epb;
    int qbounds[8];
};
struct Bc6Step;
// Result of one (block, role) chain: quantised endpoints (per channel: A0 | B0 << 16 | A1 << 32 | B1 << 48), indices, shape.
// Only the winning role of a block is packed into 128 bits, and its ~40 bit fields are packed by as many lanes (pick / pack /
// store phases): walking the header layout field by field in every candidate's lane cost 8 % of the kernel.
struct Bc6Role {
    unsigned long long ch[3];
    u32 idx0, idx1;
    int shape, flips;                 // flips: texels whose index is mirrored (set by the pick phase for the winner)
};
struct Bc6Warp {
    const Bc6Step* layout;                        // the 14 header layouts: CTA-shared copy on the device (bc6h_kernel), host table in the emulation
    float px[kBc6Slots][64];
    float lo[kBc6Slots][3], hi[kBc6Slots][3];
    float max_span[kBc6Slots];
    int max_span_idx[kBc6Slots];
    int keys[kBc6Slots][32], order[kBc6Slots][32];
    Bc6Entry two[kBc6Slots][kBc6MaxTwo], one[kBc6Slots][kBc6MaxOne];
    union {
        struct {
            Bc6Entry tmp[kBc6Slots][10];          // the ten base modes, filled in parallel by the setup phase
            int tmp_fits[kBc6Slots][10];          // 1 if the mode's span test passed under the profile's margin
        } setup;                                  // dead once the select phase has filled two[] / one[]
        Bc6Role role[kBc6Slots][kBc6MaxTwo + kBc6MaxOne];   // chain phase -> store phase: what the winner is packed from
    };
    int ntwo[kBc6Slots], none[kBc6Slots];
    float fit[kBc6Slots][32][16];                 // unquantised segments of ranked shape n: [subset][A rgb., B rgb.]
    float cand_err[kBc6Slots][kBc6MaxTwo][32];
    int win_pos[kBc6Slots][kBc6MaxTwo];
    float res_err[kBc6Slots][kBc6MaxTwo + kBc6MaxOne];
    int win_role[kBc6Slots];                      // winning role per block, -1 = none (all candidates infinite)
    u32 code[kBc6Slots][4];                       // the block being assembled: lanes OR their fields in (pack phase)
    int nvalid;
};

// ---- format data; K:2080-2125 ----
ITW_HD int bc6_prefix(int mode)
{
    // 5-bit mode prefixes {0,1,2,6,10,14,18,22,26,30,3,7,11,15} packed 5 bits each
    const unsigned long long lo = 0ull | (1ull << 5) | (2ull << 10) | (6ull << 15) | (10ull << 20) | (14ull << 25) |
                                  (18ull << 30) | (22ull << 35) | (26ull << 40) | (30ull << 45) | (3ull << 50) | (7ull << 55);
    if (mode < 12) return (int)((lo >> (5 * mode)) & 31ull);
    return (mode == 12) ? 11 : 15;
}
ITW_HD int bc6_epb(int mode)      // endpoint bits of K's base modes 0,1,2,5,6,9,10..13
{
    switch (mode) {
        case 0: return 10; case 1: return 7; case 2: return 11; case 5: return 9; case 6: return 8; case 9: return 6;
        case 10: return 10; case 11: return 11; case 12: return 12; default: return 16;
    }
}
// K:2090-2111: a float table read back through an int (truncation, quirk Q4)
ITW_HD float bc6_span(int mode)
{
    const float f = 65535.0f;
    float v;
    switch (mode) {
        case 0: v = 0.9f * f / 64.0f; break;
        case 1: v = 0.9f * f / 4.0f; break;
        case 2: v = 0.8f * f / 256.0f; break;
        case 5: v = 0.9f * f / 32.0f; break;
        case 6: v = 0.9f * f / 16.0f; break;
        case 9: case 10: v = f; break;
        case 11: v = 0.95f * f / 8.0f; break;
        case 12: v = 0.95f * f / 32.0f; break;
        default: v = 6.0f; break;
    }
    return (float)cvt_x86(v);
}

// ---- endpoint quantisation; K:2130-2169 ----
ITW_HD int bc6_dequant(int v, int bits)
{
    if (bits >= 15) return v;
    if (v == 0) return 0;
    if (v == (1 << bits) - 1) return 0xFFFF;
    return (int)(((u32)v * 2u + 1u) << (15 - bits));
}
ITW_HD int bc6_quant1(float e, int bits)
{
    const int top = (1 << bits) - 1;
    // +-0 / 65535 * top + 0.5 = 0.5 -> 0.  The never-written fourth component is always zero here and a zero
    // numerator sends the IEEE division to its slow path, so it is answered directly (same result).
    if (e == 0.0f) return 0;
    return clampi(cvt_x86(div_by_rcp(e, 65535.0f, 1.0f / 65535.0f) * (float)top + 0.5f), 0, top);   // e / 65535, see bc6_assign
}
// 8*pairs values: quantise, clamp RGB into the entry's window, decode in place
ITW_HD void bc6_quant_dequant(const Bc6Entry& E, int* q, float* ep, int pairs)
{
    for (int i = 0; i < 2 * pairs; i++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int v = bc6_quant1(ep[4 * i + c], E.epb);
            if (c < 3) v = clampi(v, E.qbounds[c], E.qbounds[4 + c]);
            q[4 * i + c] = v;
            ep[4 * i + c] = (float)bc6_dequant(v, E.epb);
        }
    }
}

// Decide whether `mode` fits the block's range and, if so, fill the entry; K:2332-2365, :2302-2330
ITW_HD_NOINLINE bool bc6_make_entry(Bc6Entry& E, const Bc6Warp& W, int slot, int mode, float margin)
{
    const float span = bc6_span(mode);
    const bool fits = !(W.max_span[slot] * margin > span);     // the entry is filled either way (mode 1 is also used untested, K:3098)
    const bool wide = !(mode >= 10 || mode <= 1 || mode == 5 || mode == 9);
    const int widx = W.max_span_idx[slot];
    E.epb = bc6_epb(mode);
    E.mode = wide ? mode + widx : mode;
    float bounds[8];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float sp = span;
        if (wide) sp *= (c == widx) ? 2.0f : 1.0f;
        float middle = (W.lo[slot][c] + W.hi[slot][c]) / 2.0f;
        bounds[c] = middle - sp / 2.0f;
        bounds[4 + c] = middle + sp / 2.0f;
    }
    bounds[3] = bounds[7] = 0.0f;                               // never-written slots read as zero (F6, Q2)
#pragma unroll
    for (int i = 0; i < 8; i++) E.qbounds[i] = bc6_quant1(bounds[i], E.epb);
    return fits;
}

// ---- header layouts -------------------------------------------------------------------------
// The 5+72 (two-region) or 5+60 (one-region) header bits, LSB first.  Each step copies `n`
// consecutive bits starting at bit `b` of field f (ascending, or descending when n is negative):
// f = 0 mode prefix; f = 1 + 3*e + c is component c of endpoint e.  Endpoints 1..3 are stored as
// differences from endpoint 0 (wrapped by the extraction) except in modes 9 and 10.  Same layout
// as K:2392-2980, held as data (it is the D3D BC6H format definition).
struct Bc6Step { int8_t f, b, n; };
#define S_(f, b, n) {f, b, n}
#define R0 1
#define G0 2
#define B0 3
#define R1 4
#define G1 5
#define B1 6
#define R2 7
#define G2 8
#define B2 9
#define R3 10
#define G3 11
#define B3 12
#define ITW_BC6_LAYOUT_INIT                                                                                         \
    /* 0*/ {S_(0,0,2), S_(G2,4,1), S_(B2,4,1), S_(B3,4,1), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,5), S_(G3,4,1), S_(G2,0,4), S_(G1,0,5), S_(B3,0,1), S_(G3,0,4), S_(B1,0,5), S_(B3,1,1), S_(B2,0,4), S_(R2,0,5), S_(B3,2,1), S_(R3,0,5), S_(B3,3,1)}, \
    /* 1*/ {S_(0,0,2), S_(G2,5,1), S_(G3,4,1), S_(G3,5,1), S_(R0,0,7), S_(B3,0,1), S_(B3,1,1), S_(B2,4,1), S_(G0,0,7), S_(B2,5,1), S_(B3,2,1), S_(G2,4,1), S_(B0,0,7), S_(B3,3,1), S_(B3,5,1), S_(B3,4,1), S_(R1,0,6), S_(G2,0,4), S_(G1,0,6), S_(G3,0,4), S_(B1,0,6), S_(B2,0,4), S_(R2,0,6), S_(R3,0,6)}, \
    /* 2*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,5), S_(R0,10,1), S_(G2,0,4), S_(G1,0,4), S_(G0,10,1), S_(B3,0,1), S_(G3,0,4), S_(B1,0,4), S_(B0,10,1), S_(B3,1,1), S_(B2,0,4), S_(R2,0,5), S_(B3,2,1), S_(R3,0,5), S_(B3,3,1)}, \
    /* 3*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,4), S_(R0,10,1), S_(G3,4,1), S_(G2,0,4), S_(G1,0,5), S_(G0,10,1), S_(G3,0,4), S_(B1,0,4), S_(B0,10,1), S_(B3,1,1), S_(B2,0,4), S_(R2,0,4), S_(B3,0,1), S_(B3,2,1), S_(R3,0,4), S_(G2,4,1), S_(B3,3,1)}, \
    /* 4*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,4), S_(R0,10,1), S_(B2,4,1), S_(G2,0,4), S_(G1,0,4), S_(G0,10,1), S_(B3,0,1), S_(G3,0,4), S_(B1,0,5), S_(B0,10,1), S_(B2,0,4), S_(R2,0,4), S_(B3,1,1), S_(B3,2,1), S_(R3,0,4), S_(B3,4,1), S_(B3,3,1)}, \
    /* 5*/ {S_(0,0,5), S_(R0,0,9), S_(B2,4,1), S_(G0,0,9), S_(G2,4,1), S_(B0,0,9), S_(B3,4,1), S_(R1,0,5), S_(G3,4,1), S_(G2,0,4), S_(G1,0,5), S_(B3,0,1), S_(G3,0,4), S_(B1,0,5), S_(B3,1,1), S_(B2,0,4), S_(R2,0,5), S_(B3,2,1), S_(R3,0,5), S_(B3,3,1)}, \
    /* 6*/ {S_(0,0,5), S_(R0,0,8), S_(G3,4,1), S_(B2,4,1), S_(G0,0,8), S_(B3,2,1), S_(G2,4,1), S_(B0,0,8), S_(B3,3,1), S_(B3,4,1), S_(R1,0,6), S_(G2,0,4), S_(G1,0,5), S_(B3,0,1), S_(G3,0,4), S_(B1,0,5), S_(B3,1,1), S_(B2,0,4), S_(R2,0,6), S_(R3,0,6)}, \
    /* 7*/ {S_(0,0,5), S_(R0,0,8), S_(B3,0,1), S_(B2,4,1), S_(G0,0,8), S_(G2,5,1), S_(G2,4,1), S_(B0,0,8), S_(G3,5,1), S_(B3,4,1), S_(R1,0,5), S_(G3,4,1), S_(G2,0,4), S_(G1,0,6), S_(G3,0,4), S_(B1,0,5), S_(B3,1,1), S_(B2,0,4), S_(R2,0,5), S_(B3,2,1), S_(R3,0,5), S_(B3,3,1)}, \
    /* 8*/ {S_(0,0,5), S_(R0,0,8), S_(B3,1,1), S_(B2,4,1), S_(G0,0,8), S_(B2,5,1), S_(G2,4,1), S_(B0,0,8), S_(B3,5,1), S_(B3,4,1), S_(R1,0,5), S_(G3,4,1), S_(G2,0,4), S_(G1,0,5), S_(B3,0,1), S_(G3,0,4), S_(B1,0,6), S_(B2,0,4), S_(R2,0,5), S_(B3,2,1), S_(R3,0,5), S_(B3,3,1)}, \
    /* 9*/ {S_(0,0,5), S_(R0,0,6), S_(G3,4,1), S_(B3,0,1), S_(B3,1,1), S_(B2,4,1), S_(G0,0,6), S_(G2,5,1), S_(B2,5,1), S_(B3,2,1), S_(G2,4,1), S_(B0,0,6), S_(G3,5,1), S_(B3,3,1), S_(B3,5,1), S_(B3,4,1), S_(R1,0,6), S_(G2,0,4), S_(G1,0,6), S_(G3,0,4), S_(B1,0,6), S_(B2,0,4), S_(R2,0,6), S_(R3,0,6)}, \
    /*10*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,10), S_(G1,0,10), S_(B1,0,10)}, \
    /*11*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,9), S_(R0,10,1), S_(G1,0,9), S_(G0,10,1), S_(B1,0,9), S_(B0,10,1)}, \
    /*12*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,8), S_(R0,11,-2), S_(G1,0,8), S_(G0,11,-2), S_(B1,0,8), S_(B0,11,-2)}, \
    /*13*/ {S_(0,0,5), S_(R0,0,10), S_(G0,0,10), S_(B0,0,10), S_(R1,0,4), S_(R0,15,-6), S_(G1,0,4), S_(G0,15,-6), S_(B1,0,4), S_(B0,15,-6)},
constexpr int kBc6MaxSteps = 24;
#if defined(__CUDACC__)
static __device__ const Bc6Step d_bc6_layout[14][kBc6MaxSteps] = {ITW_BC6_LAYOUT_INIT};
#endif
static const Bc6Step h_bc6_layout[14][kBc6MaxSteps] = {ITW_BC6_LAYOUT_INIT};
#undef S_
#undef R0
#undef G0
#undef B0
#undef R1
#undef G1
#undef B1
#undef R2
#undef G2
#undef B2
#undef R3
#undef G3
#undef B3

// Register-resident working set of the refinement chain.  Everything below is passed BY VALUE between the
// (non-inlined) routines: nvcc keeps small structs in registers across calls, while arrays handed over by pointer
// live in local memory, and the chain's dependent loads from it were 80 % of this kernel's long-scoreboard stalls
// (profiles/r1_bc6h_v4_ncu.txt).
struct Bc6Seg { float a[3], b[3]; };                      // one subset's endpoints A, B (r, g, b)
struct Bc6Quant {
    unsigned long long ch[3];                             // per channel: quantised A0, B0, A1, B1, 16 bits each
    u32 dec[2][3];                                        // per subset and channel: decoded A | decoded B << 16
};
struct Bc6Search { float err; u32 idx0, idx1; };
ITW_HD u32 bc6_q(const Bc6Quant& Q, int endpoint, int c) { return (u32)(Q.ch[c] >> (16 * endpoint)) & 0xFFFFu; }

// Header field `i` of `mode`: its bits (LSB first, as they go into the block), their count and their position; count 0 = no
// such step.  The position is the sum of the widths of the steps before it.
ITW_HD void bc6_header_field(u32& bits, int& count, int& pos, int i, const unsigned long long (&ch)[3], int mode, const Bc6Step* layout)
{
    const Bc6Step* steps = layout + mode * kBc6MaxSteps;
    pos = 0;
    for (int j = 0; j < i; j++) { const int w = steps[j].n; pos += (w < 0) ? -w : w; }
    const int f = steps[i].f, b = steps[i].b, n = steps[i].n;
    count = (n < 0) ? -n : n;
    bits = 0u;
    if (n == 0) return;
    int value;
    if (f == 0) value = bc6_prefix(mode);
    else {
        const bool delta = !(mode == 9 || mode == 10);
        const int e = (f - 1) / 3, c = (f - 1) % 3;
        const unsigned long long sel = (c == 0) ? ch[0] : ((c == 1) ? ch[1] : ch[2]);
        value = (int)((u32)(sel >> (16 * e)) & 0xFFFFu);
        if (delta && e > 0) value -= (int)((u32)sel & 0xFFFFu);
    }
    if (n > 0) bits = ((u32)value >> b) & ((1u << n) - 1u);
    else
        for (int k = 0; k < -n; k++) bits |= (((u32)value >> (b - k)) & 1u) << k;      // descending source bits
}

// ---- index search, three channels, decoded endpoints are integers 0..65535; K:1133-1193 ----
// D[j][c] = decoded endpoints A | B << 16 of subset j.  Per subset and channel the loop keeps A and B - A as floats (both
// exact).  A palette entry is ((64-w) a + w b + 32) / 64 cast to int in the reference, every intermediate an exact integer
// below 2^23; here it is floor((64 a + 32 + w (b - a)) / 64) in float -- the same integer, exact at every step, without a
// conversion per texel.  Everything that involves the (non-integer) texels stays in float, in the reference's order, with
// the x86 conversion rule.  Sums the reference starts from 0.0f start from their first term: 0 + x differs from x only in
// the sign of a zero, which the quotient below (0 / d -> index 1 either way) and the squared errors (never negative) hide.
ITW_HD_NOINLINE Bc6Search bc6_assign(const float* px, int bits, u32 pattern, u32 d00, u32 d01, u32 d02, u32 d10, u32 d11, u32 d12)
{
    const int levels = 1 << bits;
    const float flevels = (float)levels;
    float fa[2][3], fd[2][3];
    float div[2], rdiv[2];
    const u32 D[2][3] = {{d00, d01, d02}, {d10, d11, d12}};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        float d2 = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int a = (int)(D[j][c] & 0xFFFFu), b = (int)(D[j][c] >> 16);
            fa[j][c] = (float)a;
            fd[j][c] = (float)(b - a);                       // K:1155; the difference of two exact integers is exact
            d2 += sq(fd[j][c]);
        }
        div[j] = d2;
        rdiv[j] = 1.0f / d2;                                 // inf when the endpoints coincide: 0 * inf = NaN below, as 0 / 0
    }
    float total = 0.0f;
    u32 out0 = 0u, out1 = 0u;
    ITW_UNROLL(ITW_BC6_ASSIGN_UNROLL)
    for (int k = 0; k < 16; k++) {
        const bool second = ((pattern >> (2 * k)) & 3u) != 0u;
        float a[3], d[3], t[3], proj = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            a[c] = second ? fa[1][c] : fa[0][c];
            d[c] = second ? fd[1][c] : fd[0][c];
            t[c] = px[16 * c + k];
            const float term = (t[c] - a[c]) * d[c];
            proj = (c == 0) ? term : proj + term;
        }
        // the FMA-corrected quotient equals the IEEE quotient for every pair of normal floats (proved by exhaustion over
        // all 2^46 significand pairs, tests/test_gpu_division.py); proj is 0 or >= 2^-22 in magnitude and div >= 1
        proj = div_by_rcp(proj, second ? div[1] : div[0], second ? rdiv[1] : rdiv[0]);
        const int q1 = clampi(cvt_x86(fma_rn(proj, flevels, 0.5f)), 1, levels - 1);   // proj*levels is exact
        const float w0 = (float)bc7_weight(bits, q1 - 1), w1 = (float)bc7_weight(bits, q1);
        float err0 = 0.0f, err1 = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float base = fma_rn(a[c], 64.0f, 32.0f);                             // exact
            const float d0 = floorf(fma_rn(w0, d[c], base) * 0.015625f);
            const float d1 = floorf(fma_rn(w1, d[c], base) * 0.015625f);
            const float s0 = sq(d0 - t[c]), s1 = sq(d1 - t[c]);
            err0 = (c == 0) ? s0 : err0 + s0;
            err1 = (c == 0) ? s1 : err1 + s1;
        }
        const bool first = err0 < err1;
        const int best_err = cvt_x86(first ? err0 : err1);                             // K:1178-1183 (quirk Q3)
        const u32 bq = (u32)(first ? q1 - 1 : q1) << (4 * (k & 7));
        if (k < 8) out0 += bq; else out1 += bq;
        total += (float)best_err;
    }
    return Bc6Search{total, out0, out1};
}
ITW_HD Bc6Search bc6_assign(const float* px, int bits, u32 pattern, const Bc6Quant& Q)
{
    return bc6_assign(px, bits, pattern, Q.dec[0][0], Q.dec[0][1], Q.dec[0][2], Q.dec[1][0], Q.dec[1][1], Q.dec[1][2]);
}
// Quantise the endpoints of `pairs` subsets for entry E (K:2139-2169).  The never-written fourth component of the
// reference's arrays is not carried: nothing downstream reads it (K:2392-2980 packs r, g, b only).
ITW_HD Bc6Quant bc6_quantise(const Bc6Entry& E, const Bc6Seg (&seg)[2], int pairs)
{
    Bc6Quant Q;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        Q.ch[c] = 0ull;
        const int lo = E.qbounds[c], hi = E.qbounds[4 + c];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            u32 qa = 0u, qb = 0u;
            if (j < pairs) {
                qa = (u32)clampi(bc6_quant1(seg[j].a[c], E.epb), lo, hi);
                qb = (u32)clampi(bc6_quant1(seg[j].b[c], E.epb), lo, hi);
            } else {                                         // one-region entries: the second pair quantises zeros
                qa = qb = (u32)clampi(0, lo, hi);
            }
            Q.ch[c] |= ((unsigned long long)qa | ((unsigned long long)qb << 16)) << (32 * j);
            Q.dec[j][c] = (u32)bc6_dequant((int)qa, E.epb) | ((u32)bc6_dequant((int)qb, E.epb) << 16);
        }
    }
    return Q;
}
// One candidate of the two-region search: stored fit -> quantise for E -> index search; K:2188-2191
ITW_HD_NOINLINE float bc6_eval_two_region(const float* px, const Bc6Entry& E, int shape, const float* fit)
{
    Bc6Seg seg[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int c = 0; c < 3; c++) { seg[j].a[c] = fit[8 * j + c]; seg[j].b[c] = fit[8 * j + 4 + c]; }
    const Bc6Quant Q = bc6_quantise(E, seg, 2);
    return bc6_assign(px, 3, shape_pattern(shape), Q).err;
}

// PCA segment of the masked texels (3 channels, not clamped), by value; K:857-905 via bc67_core's fit_segment
ITW_HD_NOINLINE Bc6Seg bc6_fit(const float* px, int mask)
{
    float ep[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ep[i] = 0.0f;
    fit_segment_inl(ep, px, mask, 3, false);
    return Bc6Seg{{ep[0], ep[1], ep[2]}, {ep[4], ep[5], ep[6]}};
}
// Least-squares endpoints of the masked texels from their indices, by value; K:1198-1262 via bc67_core
ITW_HD_NOINLINE Bc6Seg bc6_solve(const float* px, int bits, u32 idx0, u32 idx1, int mask)
{
    float ep[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ep[i] = 0.0f;
    solve_endpoints_inl(ep, px, bits, idx0, idx1, mask, 3);
    return Bc6Seg{{ep[0], ep[1], ep[2]}, {ep[4], ep[5], ep[6]}};
}

// ---- the generic chain: initial candidate + refinement + encode of one (block, role) ----
// roles 0..ntwo-1: two-region entries (K:2195-2255); then the one-region entries (K:2275-2300)
ITW_HD_NOINLINE void bc6_chain(Bc6Warp& W, const Bc6Params& P, int slot, int r)
{
    const float* px = W.px[slot];
    const bool two = r < W.ntwo[slot];
    const Bc6Entry& E = two ? W.two[slot][r] : W.one[slot][r - W.ntwo[slot]];
    const int pairs = two ? 2 : 1, bits = two ? 3 : 4;
    int shape = 0;
    Bc6Seg seg[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int c = 0; c < 3; c++) seg[j].a[c] = seg[j].b[c] = 0.0f;
    if (two) {
        const int pos = W.win_pos[slot][r];
        if (pos < 0) { W.res_err[slot][r] = inf_f(); return; }
        shape = W.order[slot][pos] & 31;
        const float* fit = W.fit[slot][pos];
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int c = 0; c < 3; c++) { seg[j].a[c] = fit[8 * j + c]; seg[j].b[c] = fit[8 * j + 4 + c]; }
    } else {
        seg[0] = bc6_fit(px, 0xFFFF);
    }
    const u32 pattern = two ? shape_pattern(shape) : 0u;
    const int mask0 = two ? shape_mask(shape, 0) : 0xFFFF, mask1 = two ? shape_mask(shape, 1) : 0;
    Bc6Quant best_q = bc6_quantise(E, seg, pairs);
    Bc6Search best = bc6_assign(px, bits, pattern, best_q);

    const int refine = two ? P.refine_2p : P.refine_1p;
    for (int it = 0; it < refine; it++) {
        seg[0] = bc6_solve(px, bits, best.idx0, best.idx1, mask0);
        if (two) seg[1] = bc6_solve(px, bits, best.idx0, best.idx1, mask1);
        const Bc6Quant q = bc6_quantise(E, seg, pairs);
        const Bc6Search found = bc6_assign(px, bits, pattern, q);
        // two-region keeps the best iterate (K:2242); one-region keeps the last (K:2288-2293)
        if (!two || found.err < best.err) { best_q = q; best = found; }
    }
    W.res_err[slot][r] = best.err;
    Bc6Role& R = W.role[slot][r];
#pragma unroll
    for (int c = 0; c < 3; c++) R.ch[c] = best_q.ch[c];
    R.idx0 = best.idx0; R.idx1 = best.idx1; R.shape = shape;
}
// Orientation of the winning role, in place: the anchor index of every subset must have a clear top bit -- swap that subset's
// endpoints and mirror its indices (K:1694-1733, :2982-3031); the swap is a 32-bit rotate of the packed pair.
ITW_HD void bc6_orient_role(Bc6Role& R, bool two)
{
    const int bits = two ? 3 : 4, half = (1 << bits) / 2;
    int flips = 0;
    if (two) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int k0 = shape_anchor(R.shape, j);
            const int v = (int)(((k0 < 8 ? R.idx0 : R.idx1) >> (4 * (k0 & 7))) & 15u);
            if (v >= half) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const u32 pair = (u32)(R.ch[c] >> (32 * j));
                    const u32 swapped = (pair >> 16) | (pair << 16);
                    R.ch[c] = (R.ch[c] & ~(0xFFFFFFFFull << (32 * j))) | ((unsigned long long)swapped << (32 * j));
                }
                flips |= shape_mask(R.shape, j);
            }
        }
    } else if ((int)(R.idx0 & 15u) >= half) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const u32 pair = (u32)R.ch[c];
            R.ch[c] = (R.ch[c] & 0xFFFFFFFF00000000ull) | (unsigned long long)((pair >> 16) | (pair << 16));
        }
        flips = 0xFFFF;                                       // every index mirrored
    }
    R.flips = flips;
}

// =============================================================================================
// warp program
// =============================================================================================
// half bits -> the reference's working scale v = (h/31)*64; K:134-151, :3048
ITW_HD void bc6_phase_load(int lane, Bc6Warp& W, const SurfaceView& s, long long first_block, int nvalid)
{
    const int bw = s.width >> 2;
    for (int t = lane; t < nvalid * 16; t += 32) {
        const int slot = t >> 4, k = t & 15;
        const long long id = first_block + slot;
        const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);
        const uint8_t* p = s.ptr + (size_t)(by * 4 + (k >> 2)) * (size_t)s.stride + (size_t)(bx * 4 + (k & 3)) * 8;
        float* px = W.px[slot];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int h = (int)p[2 * c] | ((int)p[2 * c + 1] << 8);
            px[16 * c + k] = ((float)h / 31.0f) * 64.0f;
        }
        px[48 + k] = 0.0f;
    }
    if (lane == 0) W.nvalid = nvalid;
}
// per-block range; K:3036-3067
ITW_HD void bc6_phase_range(int lane, Bc6Warp& W)
{
    for (int t = lane; t < W.nvalid * 3; t += 32) {
        const int slot = t / 3, c = t - slot * 3;
        const float* px = W.px[slot];
        float lo = 65535.0f, hi = 0.0f;
        for (int k = 0; k < 16; k++) { lo = min_sse(lo, px[16 * c + k]); hi = max_sse(hi, px[16 * c + k]); }
        W.lo[slot][c] = lo;
        W.hi[slot][c] = hi;
    }
}
ITW_HD void bc6_phase_span(int lane, Bc6Warp& W)
{
    for (int slot = lane; slot < W.nvalid; slot += 32) {
        float max_span = 0.0f;
        int max_idx = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float sp = W.hi[slot][c] - W.lo[slot][c];
            if (sp > max_span) { max_idx = c; max_span = sp; }
        }
        W.max_span[slot] = max_span;
        W.max_span_idx[slot] = max_idx;
    }
}
// The ten base modes in the reference's order {0,1,2,5,6,9 | 10,11,12,13} and the margin each is tested with.
ITW_HD int bc6_base_mode(int i) { return (i < 3) ? i : ((i == 3) ? 5 : ((i == 4) ? 6 : ((i == 5) ? 9 : i + 4))); }
ITW_HD float bc6_margin(const Bc6Params& P, int i)
{
    if (P.slow_mode) return 0.0f;                               // K:3075-3084
    const float m12 = 1.0f / 1.2f;                              // K:3090-3104
    switch (i) {
        case 0: return m12; case 1: return 1.0f; case 2: return 1.0f; case 3: return m12; case 4: return m12; case 5: return 0.0f;
        case 6: return 0.0f; default: return 1.0f;
    }
}
// lane <-> (block, base mode): span test + clamp window; K:2332-2365
ITW_HD void bc6_phase_entries(int lane, Bc6Warp& W, const Bc6Params& P)
{
    for (int t = lane; t < W.nvalid * 10; t += 32) {
        const int slot = t / 10, i = t - slot * 10;
        W.setup.tmp_fits[slot][i] = bc6_make_entry(W.setup.tmp[slot][i], W, slot, bc6_base_mode(i), bc6_margin(P, i)) ? 1 : 0;
    }
}
// the list of modes each block is encoded with; K:3069-3107
ITW_HD void bc6_phase_select(int lane, Bc6Warp& W, const Bc6Params& P)
{
    for (int slot = lane; slot < W.nvalid; slot += 32) {
        int ntwo = 0, none = 0;
        if (P.slow_mode) {                                       // every mode whose (margin 0) test passes, in order
            for (int i = 0; i < 6; i++)
                if (W.setup.tmp_fits[slot][i]) W.two[slot][ntwo++] = W.setup.tmp[slot][i];
            for (int i = 6; i < 10; i++)
                if (W.setup.tmp_fits[slot][i]) W.one[slot][none++] = W.setup.tmp[slot][i];
        } else {
            if (P.fast_skip > 0) {                               // the LAST passing test of 9, [1], 6, 5, 0, 2 wins; K:3090-3095
                int pick = 5;                                    // mode 9 (margin 0) always passes
                if (P.fast_mode && W.setup.tmp_fits[slot][1]) pick = 1;
                if (W.setup.tmp_fits[slot][4]) pick = 4;
                if (W.setup.tmp_fits[slot][3]) pick = 3;
                if (W.setup.tmp_fits[slot][0]) pick = 0;
                if (W.setup.tmp_fits[slot][2]) pick = 2;
                W.two[slot][0] = W.setup.tmp[slot][pick];
                ntwo = 1;
                if (!P.fast_mode) { W.two[slot][1] = W.setup.tmp[slot][1]; ntwo = 2; }     // mode 1 with margin 0; K:3098
            }
            int pick = 6;                                        // 10, then 11, 12, 13 if they fit; K:3101-3105
            for (int i = 7; i < 10; i++)
                if (W.setup.tmp_fits[slot][i]) pick = i;
            W.one[slot][0] = W.setup.tmp[slot][pick];
            none = 1;
        }
        W.ntwo[slot] = ntwo;
        W.none[slot] = none;
    }
}
ITW_HD void bc6_phase_keys(int lane, Bc6Warp& W)
{
    for (int t = lane; t < W.nvalid * 32; t += 32) {
        const int slot = t >> 5, shape = t & 31;
        float full[15];
        masked_moments(full, W.px[slot], 0xFFFF, 3);
        W.keys[slot][shape] = split_bound_key(W.px[slot], shape, full, 3);
    }
}
ITW_HD void bc6_phase_rank(int lane, Bc6Warp& W)
{
    for (int t = lane; t < W.nvalid * 32; t += 32) {
        const int slot = t >> 5, i = t & 31;
        W.order[slot][rank_of(W.keys[slot], 32, i)] = W.keys[slot][i];
    }
}
// PCA segments of the first fast_skip ranked shapes, once per block; K:2181-2186
ITW_HD void bc6_phase_fits(int lane, Bc6Warp& W, const Bc6Params& P)
{
    const int count = P.fast_skip;
    for (int t = lane; t < W.nvalid * count; t += 32) {
        const int slot = t / count, n = t - slot * count;
        const int shape = W.order[slot][n] & 31;
        float* ep = W.fit[slot][n];
#pragma unroll
        for (int i = 0; i < 16; i++) ep[i] = 0.0f;               // never-written slots read as zero (F6)
        for (int j = 0; j < 2; j++) {
            const Bc6Seg seg = bc6_fit(W.px[slot], shape_mask(shape, j));
#pragma unroll
            for (int c = 0; c < 3; c++) { ep[8 * j + c] = seg.a[c]; ep[8 * j + 4 + c] = seg.b[c]; }
        }
    }
}
ITW_HD void bc6_phase_candidates(int lane, Bc6Warp& W, const Bc6Params& P)
{
    const int count = P.fast_skip, per = kBc6MaxTwo * count;
    for (int t = lane; t < W.nvalid * per; t += 32) {
        const int slot = t / per, r = t - slot * per;
        const int e = r / count, n = r - e * count;
        if (e >= W.ntwo[slot]) continue;
        W.cand_err[slot][e][n] = bc6_eval_two_region(W.px[slot], W.two[slot][e], W.order[slot][n] & 31, W.fit[slot][n]);
    }
}
ITW_HD void bc6_phase_winners(int lane, Bc6Warp& W, const Bc6Params& P)
{
    for (int t = lane; t < W.nvalid * kBc6MaxTwo; t += 32) {
        const int slot = t / kBc6MaxTwo, e = t - slot * kBc6MaxTwo;
        int best = -1;
        float best_err = inf_f();
        if (e < W.ntwo[slot])
            for (int n = 0; n < P.fast_skip; n++) {
                float v = W.cand_err[slot][e][n];
                if (v < best_err) { best_err = v; best = n; }
            }
        W.win_pos[slot][e] = best;
    }
}
ITW_HD void bc6_phase_chains(int lane, Bc6Warp& W, const Bc6Params& P)
{
    const int per = kBc6MaxTwo + kBc6MaxOne;
    for (int t = lane; t < W.nvalid * per; t += 32) {
        const int slot = t / per, r = t - slot * per;
        if (r < W.ntwo[slot] + W.none[slot]) bc6_chain(W, P, slot, r);
    }
}
// the block's winner: first strict minimum in role order (K:2257, :2296), oriented in place
ITW_HD void bc6_phase_pick(int lane, Bc6Warp& W)
{
    for (int t = lane; t < W.nvalid; t += 32) {
        const int nroles = W.ntwo[t] + W.none[t];
        float best_err = inf_f();
        int best = -1;
        for (int r = 0; r < nroles; r++) {
            const float e = W.res_err[t][r];
            if (e < best_err) { best_err = e; best = r; }
        }
        W.win_role[t] = best;
#pragma unroll
        for (int i = 0; i < 4; i++) W.code[t][i] = 0u;
        if (best >= 0) bc6_orient_role(W.role[t][best], best < W.ntwo[t]);
    }
}
// The 128 bits of the winners, one bit field per lane: up to 24 header steps, the shape id, sixteen indices (K:2392-3031).
// Lanes OR their fields into W.code (several lanes per word: shared_or).
constexpr int kBc6PackItems = kBc6MaxSteps + 1 + 16;
ITW_HD void bc6_phase_pack(int lane, Bc6Warp& W)
{
    for (int t = lane; t < W.nvalid * kBc6PackItems; t += 32) {
        const int slot = t / kBc6PackItems, i = t - slot * kBc6PackItems;
        const int r = W.win_role[slot];
        if (r < 0) continue;
        const bool two = r < W.ntwo[slot];
        const Bc6Entry& E = two ? W.two[slot][r] : W.one[slot][r - W.ntwo[slot]];
        const Bc6Role& R = W.role[slot][r];
        u32 bits = 0u;
        int count = 0, pos = 0;
        if (i < kBc6MaxSteps) {
            const unsigned long long ch[3] = {R.ch[0], R.ch[1], R.ch[2]};
            bc6_header_field(bits, count, pos, i, ch, E.mode, W.layout);
        } else if (i == kBc6MaxSteps) {
            if (two) { bits = (u32)R.shape; count = 5; pos = 77; }                      // 5 + 72 header bits before it
        } else {
            const int k = i - kBc6MaxSteps - 1, width = two ? 3 : 4, top = (1 << width) - 1;
            const int anchor1 = two ? shape_anchor(R.shape, 1) : -1;
            int q = (int)(((k < 8 ? R.idx0 : R.idx1) >> (4 * (k & 7))) & 15u);
            if ((R.flips >> k) & 1) q = top - q;
            const bool narrow = (k == 0) || (k == anchor1);
            bits = (u32)q;
            count = narrow ? width - 1 : width;
            pos = (two ? 82 : 65) + width * k - (k > 0 ? 1 : 0) - ((anchor1 >= 0 && k > anchor1) ? 1 : 0);
        }
        if (count == 0) continue;
        bits &= (1u << count) - 1u;
        const unsigned long long wide = (unsigned long long)bits << (pos & 31);
        const int word = pos >> 5;
        if ((u32)wide) shared_or(&W.code[slot][word], (u32)wide);
        if ((u32)(wide >> 32) && word < 3) shared_or(&W.code[slot][word + 1], (u32)(wide >> 32));
    }
}
ITW_HD void bc6_phase_store(int lane, Bc6Warp& W, uint8_t* dst, long long first_block)
{
    for (int t = lane; t < W.nvalid * 4; t += 32)
        reinterpret_cast<u32*>(dst + (size_t)(first_block + (t >> 2)) * 16)[t & 3] = W.code[t >> 2][t & 3];
}

#define ITW_BC6_PROGRAM(PHASE)                                             \
    PHASE(bc6_phase_load(lane, W, surf, first_block, nvalid));             \
    ITW_BC6_PROGRAM_AFTER_LOAD(PHASE)
#define ITW_BC6_PROGRAM_AFTER_LOAD(PHASE)                                  \
    PHASE(bc6_phase_range(lane, W));                                       \
    PHASE(bc6_phase_span(lane, W));                                        \
    PHASE(bc6_phase_entries(lane, W, P));                                  \
    PHASE(bc6_phase_select(lane, W, P));                                   \
    if (P.fast_skip > 0) {                                                 \
        PHASE(bc6_phase_keys(lane, W));                                    \
        PHASE(bc6_phase_rank(lane, W));                                    \
        PHASE(bc6_phase_fits(lane, W, P));                                 \
        PHASE(bc6_phase_candidates(lane, W, P));                           \
    }                                                                      \
    PHASE(bc6_phase_winners(lane, W, P));                                  \
    PHASE(bc6_phase_chains(lane, W, P));                                   \
    PHASE(bc6_phase_pick(lane, W));                                        \
    PHASE(bc6_phase_pack(lane, W));                                        \
    PHASE(bc6_phase_store(lane, W, dst, first_block));

#if defined(__CUDACC__)
// Lock-step phases over one 16-warp CTA per SM, for instruction-cache locality, and (kTma) the round's 48
// blocks fetched by the TMA engine into a double-buffered shared-memory tile -- see bc7.cuh.
constexpr int kBc6WarpsPerCta = 16;
constexpr size_t kBc6SmemBytes = sizeof(Bc6Warp) * kBc6WarpsPerCta;
constexpr int kBc6TileBlocks = kBc6WarpsPerCta * kBc6Slots;              // 48
constexpr int kBc6TileRowBytes = kBc6TileBlocks * 32;                     // 1536

template <bool kTma>
__global__ void __launch_bounds__(kBc6WarpsPerCta * 32, 1)
bc6h_kernel(SurfaceView gsurf, uint8_t* __restrict__ dst, Bc6Params P, long long nblocks)
{
    extern __shared__ __align__(16) unsigned char bc6_smem[];
    __shared__ __align__(128) unsigned char stage[kTma ? 2 : 1][kTma ? 4 * kBc6TileRowBytes : 16];
    __shared__ __align__(8) unsigned long long full[2];
    // header layouts in shared memory: the pack phase walks them, and from global memory every step
    // was a dependent L1/L2 round trip (long-scoreboard stalls, profiles/r1_final_bc6h_slow_ncu.txt)
    __shared__ Bc6Step layout[14 * kBc6MaxSteps];
    for (int i = threadIdx.x; i < 14 * kBc6MaxSteps; i += blockDim.x) layout[i] = d_bc6_layout[i / kBc6MaxSteps][i % kBc6MaxSteps];
    Bc6Warp& W = reinterpret_cast<Bc6Warp*>(bc6_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) W.layout = layout;
    __syncthreads();
    const long long nbatches = (nblocks + kBc6Slots - 1) / kBc6Slots;
    const long long nwarps = (long long)gridDim.x * kBc6WarpsPerCta;
    const long long rounds = (nbatches + nwarps - 1) / nwarps;

    auto tile_first = [&](long long round) { return ((long long)blockIdx.x * kBc6WarpsPerCta + round * nwarps) * kBc6Slots; };
    auto prefetch = [&](long long round) {
        const long long fb = tile_first(round);
        if (round >= rounds || fb >= nblocks) return;
        const long long left = nblocks - fb;
        tma_prefetch_tile(stage[round & 1], kBc6TileRowBytes, &full[round & 1], gsurf, fb,
                          (int)(left < kBc6TileBlocks ? left : kBc6TileBlocks), 32);
    };
    if (kTma) {
        if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
        __syncthreads();
        if (threadIdx.x == 0) prefetch(0);
    }
    for (long long round = 0; round < rounds; round++) {
        const long long batch = (long long)blockIdx.x * kBc6WarpsPerCta + warp + round * nwarps;
        long long first_block = batch * kBc6Slots;
        const long long left = nblocks - first_block;
        const int nvalid = (int)(left <= 0 ? 0 : (left < kBc6Slots ? left : kBc6Slots));
        SurfaceView surf = gsurf;
        const long long out_block = first_block;
        if (kTma) {
            if (threadIdx.x == 0) prefetch(round + 1);
            if (tile_first(round) < nblocks) mbar_wait(&full[round & 1], (unsigned)((round >> 1) & 1));
            surf = SurfaceView{stage[round & 1], kBc6TileBlocks * 4, 4, kBc6TileRowBytes};
            first_block = (long long)warp * kBc6Slots;
        }
#define ITW_PHASE_DEVICE(call) call; __syncthreads()
        {
            ITW_PHASE_DEVICE(bc6_phase_load(lane, W, surf, first_block, nvalid));
            first_block = out_block;
            ITW_BC6_PROGRAM_AFTER_LOAD(ITW_PHASE_DEVICE)
        }
#undef ITW_PHASE_DEVICE
    }
}
#endif

}  // namespace itw
