// bc1_pair.cuh -- EXPERIMENT, not part of the product (DESIGN.md section 4 reports the measurements): BC1 / BC3 encoder,
// TWO blocks per thread on Blackwell's packed float lanes
// (reference: kernel.ispc:231-614, cited as K:line; the one-block form of the same arithmetic is bc1_bc3.cuh).
//
// Why.  One 4x4 block costs ~1.5 k scalar instructions against 72 bytes of traffic, and three quarters of them are
// binary32 adds and multiplies whose ORDER is fixed by the reference (K:377-417 accumulates inexact sums texel by texel).
// The order cannot change, but sm_100a can issue one FADD2 / FMUL2 / FFMA2 for two independent lanes.  A thread therefore
// owns two blocks -- lane x and lane y of every float register pair -- and walks the reference's algorithm ONCE: the same
// roundings per lane, half the issue slots.  Integer steps (565 packing, index words) run per lane.
//
// Also different from bc1_bc3.cuh, each argued where it happens:
//   * byte -> float through PRMT + one packed subtraction (bytes_to_f2) and float -> 2-bit index through a round-towards-
//     zero add of 2^23 -- neither touches the quarter-rate conversion pipe;
//   * products of exact small numbers are fused into their accumulation (fma2): covariance terms, refinement sums;
//   * "0 + x" first additions are dropped (x when x is not -0, and the sign of a zero is shown unobservable).
//
// Staging: a persistent CTA fetches tiles of 2 x blockDim consecutive blocks with the TMA engine (cp.async.bulk, mbarrier)
// into a 4-row shared-memory tile; the next tile streams in while the current one is being encoded from registers.
#pragma once
#include "../../../intel-texture-works-plugin_b200/csrc/bc1_bc3.cuh"

namespace itw {

// ---- two independent float lanes per instruction (sm_100a: FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2) ----------
// Blackwell issues ONE instruction for two IEEE binary32 operations on a 64-bit register pair.  Each lane is an ordinary
// round-to-nearest-even add / multiply (or fused multiply-add), so a pair op gives exactly the two results the scalar
// ops would: the float model of this file is unchanged, the issue slots are halved.  The encoders use the lanes for two
// independent blocks (BC1/BC3) or two independent quantities of one block.  fma2 FUSES: it is used only where the product
// is proved exact (then fused == multiply-then-add), like fma_rn above.  Host (tests/emu): plain scalar operations.
#if defined(__CUDACC__)
typedef float2 f2;
#else
struct f2 { float x, y; };
#endif
ITW_HD f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
ITW_HD f2 splat2(float a) { return mk2(a, a); }
ITW_HD f2 add2(f2 a, f2 b)
{
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return mk2(a.x + b.x, a.y + b.y);
#endif
}
ITW_HD f2 mul2(f2 a, f2 b)
{
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return mk2(a.x * b.x, a.y * b.y);
#endif
}
ITW_HD f2 fma2(f2 a, f2 b, f2 c)
{
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// RN(RN(a*b) + c) per lane -- multiply, then add, two roundings (the reference's unfused expression).
// ptxas of CUDA 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 whenever the product has no other use, even with
// --fmad false and although both instructions carry an explicit .rn (scalar mul.rn / add.rn are never contracted; checked with
// cuobjdump, see tools/ubench/check_f32x2.cu).  A fused multiply-add rounds once and changes results.  The addition is therefore
// issued as an FFMA2 whose multiplier `one` is a KERNEL ARGUMENT that only the host knows to be 1.0: p * 1 is exact, so
// fma(p, one, c) rounds exactly like p + c, and a product that feeds the multiplicand of an FMA cannot be contracted into it.
// Same instruction count as the packed multiply + packed add it replaces.
ITW_HD f2 madd2(f2 a, f2 b, f2 c, f2 one) { return fma2(mul2(a, b), one, c); }
// a + b rounded TOWARDS ZERO per lane (the float->int "magic number" conversion below)
ITW_HD f2 add2_rz(f2 a, f2 b)
{
#if defined(__CUDA_ARCH__)
    return __fadd2_rz(a, b);
#else
    // a, b >= 0 wherever this is used: towards zero == the double sum truncated to 24 bits == nextafter correction of RN
    f2 r;
    const float in[2][2] = {{a.x, b.x}, {a.y, b.y}};
    float out[2];
    for (int i = 0; i < 2; i++) {
        const double exact = (double)in[i][0] + (double)in[i][1];       // exact: both operands are floats of similar magnitude
        float s = (float)exact;
        if ((double)s > exact) s = nextafterf(s, 0.0f);
        out[i] = s;
    }
    r.x = out[0]; r.y = out[1];
    return r;
#endif
}
ITW_HD f2 min2(f2 a, f2 b) { return mk2(fminf(a.x, b.x), fminf(a.y, b.y)); }      // finite / NaN-first operands only (see callers)
ITW_HD f2 max2(f2 a, f2 b) { return mk2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
// byte `c` of each of two packed words as floats, without the conversion pipe: PRMT drops the byte into the mantissa of
// 2^23 (0x4B000000 | byte == 8388608 + byte exactly), one packed subtraction removes the 2^23 (exact).
ITW_HD f2 bytes_to_f2(u32 wa, u32 wb, int c)
{
#if defined(__CUDA_ARCH__)
    const u32 sel = 0x7440u | (u32)c;                       // result bytes: [byte c of w, 00, 00, 4B]
    const float ma = __uint_as_float(__byte_perm(wa, 0x4B000000u, sel)), mb = __uint_as_float(__byte_perm(wb, 0x4B000000u, sel));
    return __fadd2_rn(make_float2(ma, mb), make_float2(-8388608.0f, -8388608.0f));
#else
    return mk2((float)((wa >> (8 * c)) & 255u), (float)((wb >> (8 * c)) & 255u));
#endif
}



// v -> clamp((int)v, 0, top) for both lanes WITHOUT F2I, top = 3 or 7.  (int)v truncates; clamping first to [0, top + 0.5]
// cannot change the clamped result, NaN (p0 == p1, K:321-336: INT_MIN -> 0) becomes 0 through fmaxf.  Then
// RZ(v + 2^23) has floor(v) = trunc(v) in its low mantissa bits (0 <= v < 8).  Returns the float 2^23 + q per lane.
ITW_HD f2 trunc_clamp_magic(f2 v, float top_plus_half)
{
    const f2 c = min2(max2(v, splat2(0.0f)), splat2(top_plus_half));
    return add2_rz(c, splat2(8388608.0f));
}

struct PairWords { u32 a0, a1, b0, b1; };      // (w0, w1) of lane x and of lane y

// rgb565 of clamped floats (callers clamp to [0,255]); K:245-248
ITW_HD int pack565_lane(float r, float g, float b) { return pack565(r, g, b); }

// Linear 2-bit indices of both lanes along p0 -> p1; K:308-344.  kRefine: also accumulate sum_k (3 - q_k) * px_k per
// channel (K:432-440) while the indices are at hand.
// The 2 x 48 texel values live in SHARED memory on the device -- px[(c*16 + k) * ps], one 8-byte column per thread -- instead
// of 96 registers: that is what lets 14 warps share an SM instead of 8 (the load/store pipe is idle in this kernel, every
// pass reads its operands with LDS.64 at constant offsets).  The emulation passes a local array with ps = 1.
#define ITW_PX(c, k) px[((c) * 16 + (k)) * ps]
// Between two passes over the texel values: the compiler must not keep the 96 values of one pass in registers for the next
// (it would, the columns being read-only after the conversion) -- each pass re-reads them from shared memory.
#if defined(__CUDA_ARCH__)
#define ITW_PASS_FENCE() asm volatile("" ::: "memory")
#else
#define ITW_PASS_FENCE() ((void)0)
#endif
template <bool kRefine>
ITW_HD void bc1_indices_pair(const f2* px, const int ps, int p0x, int p1x, int p0y, int p1y, u32& bits_x, u32& bits_y, f2 (&atb1)[3], const f2 one)
{
    float ax[3], bx[3], ay[3], by[3];
    unpack565(ax, p0x); unpack565(bx, p1x);
    unpack565(ay, p0y); unpack565(by, p1y);
    f2 na[3], dir[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        na[c] = mk2(-ax[c], -ay[c]);                                   // exact
        dir[c] = add2(mk2(bx[c], by[c]), na[c]);                       // b - a, exact small integers
    }
    f2 n2 = mul2(dir[0], dir[0]);                                      // 0 + x dropped: squares are never -0
    n2 = madd2(dir[1], dir[1], n2, one);
    n2 = madd2(dir[2], dir[2], n2, one);
    const f2 inv3 = mul2(mk2(1.0f / n2.x, 1.0f / n2.y), splat2(3.0f)); // inf when p0 == p1 -> NaN below, as in K
#pragma unroll
    for (int c = 0; c < 3; c++) dir[c] = mul2(dir[c], inv3);
    f2 bias = splat2(0.5f);
#pragma unroll
    for (int c = 0; c < 3; c++) bias = madd2(na[c], dir[c], bias, one);    // bias -= a*dir: -(a*dir) == (-a)*dir
    u32 bx_bits = 0u, by_bits = 0u;
    if (kRefine) { atb1[0] = atb1[1] = atb1[2] = splat2(0.0f); }
#pragma unroll
    for (int k = 15; k >= 0; k--) {                                   // texels are independent: descending order makes
        const f2 pr = ITW_PX(0, k), pg = ITW_PX(1, k), pb = ITW_PX(2, k);
        f2 d = mul2(pr, dir[0]);                                      // bits = bits*4 + q one multiply-add per lane
        d = madd2(pg, dir[1], d, one);                                // (0 + first term dropped: d only feeds d + bias,
        d = madd2(pb, dir[2], d, one);                                //  where the sign of a zero vanishes)
        const f2 t = trunc_clamp_magic(add2(d, bias), 3.5f);          // 2^23 + q
        bx_bits = bx_bits * 4u + (float_bits(t.x) & 3u);
        by_bits = by_bits * 4u + (float_bits(t.y) & 3u);
        if (kRefine) {
            // x = 3 - q = (2^23 + 3) - t exactly; x*px (<= 765) and the running sums (<= 12240) are exact integers, so
            // the fused multiply-add is the reference's multiply-then-add, and the order of an exact sum is free
            const f2 x = add2(splat2(8388611.0f), mk2(-t.x, -t.y));
            atb1[0] = fma2(x, pr, atb1[0]); atb1[1] = fma2(x, pg, atb1[1]); atb1[2] = fma2(x, pb, atb1[2]);
        }
    }
    bits_x = bx_bits; bits_y = by_bits;
}

// One lane of the refinement solve; K:419-480 (sums from bc1_indices_pair)
ITW_HD void bc1_refine_lane(u32 bits, const float (&mean)[3], const float (&atb1)[3], int& p0, int& p1)
{
    float ea[3], eb[3];
    if ((bits ^ (bits * 4u)) < 4u) {
#pragma unroll
        for (int c = 0; c < 3; c++) ea[c] = eb[c] = mean[c];
    } else {
        const u32 lo_b = bits & 0x55555555u, hi_b = (bits >> 1) & 0x55555555u;
        const float sq1 = (float)(popcount32(lo_b) + 2 * popcount32(hi_b));
        const float sqq = (float)(popcount32(lo_b) + 4 * popcount32(hi_b) + 4 * popcount32(lo_b & hi_b));
        const float cxx = 16.0f * 9.0f - 6.0f * sq1 + sqq;
        const float cyy = sqq;
        const float cxy = 3.0f * sq1 - sqq;
        const float scale = 3.0f * (1.0f / (cxx * cyy - cxy * cxy));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float total = mean[c] * 16.0f;
            const float atb2 = 3.0f * total - atb1[c];
            const float a = (atb1[c] * cyy - atb2 * cxy) * scale;
            const float b = (atb2 * cxx - atb1[c] * cxy) * scale;
            ea[c] = clamp_sse(a, 0.0f, 255.0f);
            eb[c] = clamp_sse(b, 0.0f, 255.0f);
        }
    }
    p0 = pack565(ea[0], ea[1], ea[2]);
    p1 = pack565(eb[0], eb[1], eb[2]);
    if (p0 < p1) { const int t = p0; p0 = p1; p1 = t; }
}

// The colour half of both lanes; K:494-533.  out = (w0, w1) per lane.
ITW_HD PairWords bc1_colour_pair(const f2* px, const int ps, const f2 one)
{
    // mean: sixteen exact integers, any order; K:377-385.  The first "0 + x" is dropped (x >= 0).
    f2 mean[3], nmean[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        f2 acc = ITW_PX(c, 0);
#pragma unroll
        for (int k = 1; k < 16; k++) acc = add2(acc, ITW_PX(c, k));
        mean[c] = mul2(acc, splat2(0.0625f));                          // /16, exact
        nmean[c] = mul2(acc, splat2(-0.0625f));
    }
    ITW_PASS_FENCE();
    // centred covariance in texel order; K:386-417.  px - mean is exact (multiples of 1/16 below 256), the products are
    // exact (24 significant bits at most), so fusing each product into its running sum rounds exactly like multiply-then-add
    f2 crr = splat2(0.0f), crg = crr, crb = crr, cgg = crr, cgb = crr, cbb = crr;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const f2 r = add2(ITW_PX(0, k), nmean[0]), g = add2(ITW_PX(1, k), nmean[1]), b = add2(ITW_PX(2, k), nmean[2]);
        crr = fma2(r, r, crr); crg = fma2(r, g, crg); crb = fma2(r, b, crb);
        cgg = fma2(g, g, cgg); cgb = fma2(g, b, cgb); cbb = fma2(b, b, cbb);
    }
    ITW_PASS_FENCE();
    const f2 eps = splat2(0.001f);
    crr = add2(crr, eps); cgg = add2(cgg, eps); cbb = add2(cbb, eps);

    // 4 power iterations from (1,1,1), renormalised after iterations 1 and 3; K:184-205
    f2 v0 = splat2(1.0f), v1 = v0, v2 = v0;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const f2 a0 = madd2(crb, v2, madd2(crg, v1, mul2(crr, v0), one), one);      // (crr*v0 + crg*v1) + crb*v2
        const f2 a1 = madd2(cgb, v2, madd2(cgg, v1, mul2(crg, v0), one), one);
        const f2 a2 = madd2(cbb, v2, madd2(cgb, v1, mul2(crb, v0), one), one);
        v0 = a0; v1 = a1; v2 = a2;
        if (it & 1) {
            f2 n2 = mul2(a0, a0);                                      // 0 + x dropped (a square is never -0)
            n2 = madd2(a1, a1, n2, one);
            n2 = madd2(a2, a2, n2, one);
            const f2 rn = mk2(1.0f / sqrtf(n2.x), 1.0f / sqrtf(n2.y));
            v0 = mul2(v0, rn); v1 = mul2(v1, rn); v2 = mul2(v2, rn);
        }
    }
    // extreme projections -> endpoints; K:274-306 (min starts at 65536, max at 0).  "0 + first term" dropped: d only
    // reaches min/max and then mean + d*inv*axis, where the sign of a zero cannot be seen (mean >= 0)
    f2 dmin = splat2(65536.0f), dmax = splat2(0.0f);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        f2 d = mul2(add2(ITW_PX(0, k), nmean[0]), v0);
        d = madd2(add2(ITW_PX(1, k), nmean[1]), v1, d, one);
        d = madd2(add2(ITW_PX(2, k), nmean[2]), v2, d, one);
        dmin = min2(dmin, d);                                          // d is finite: fminf == the reference's (a<b)?a:b
        dmax = max2(dmax, d);
    }
    ITW_PASS_FENCE();
    if (dmax.x - dmin.x < 1.0f) { dmin.x -= 0.5f; dmax.x += 0.5f; }
    if (dmax.y - dmin.y < 1.0f) { dmin.y -= 0.5f; dmax.y += 0.5f; }
    f2 n2 = mul2(v0, v0);
    n2 = madd2(v1, v1, n2, one);
    n2 = madd2(v2, v2, n2, one);
    const f2 inv = mk2(1.0f / n2.x, 1.0f / n2.y);
    const f2 tmin = mul2(dmin, inv), tmax = mul2(dmax, inv);
    const f2 ax[3] = {v0, v1, v2};
    f2 lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // clamp_sse(v, 0, 255) == fminf(fmaxf(v, 0), 255) for every v that can occur (a NaN gives 0 in both forms; -0
        // and +0 truncate alike)
        lo[c] = min2(max2(madd2(tmin, ax[c], mean[c], one), splat2(0.0f)), splat2(255.0f));
        hi[c] = min2(max2(madd2(tmax, ax[c], mean[c], one), splat2(0.0f)), splat2(255.0f));
    }
    int p0x = pack565(lo[0].x, lo[1].x, lo[2].x), p1x = pack565(hi[0].x, hi[1].x, hi[2].x);
    int p0y = pack565(lo[0].y, lo[1].y, lo[2].y), p1y = pack565(hi[0].y, hi[1].y, hi[2].y);
    if (p0x < p1x) { const int t = p0x; p0x = p1x; p1x = t; }
    if (p0y < p1y) { const int t = p0y; p0y = p1y; p1y = t; }
    u32 bits_x, bits_y;
    f2 atb1[3];
    bc1_indices_pair<true>(px, ps, p0x, p1x, p0y, p1y, bits_x, bits_y, atb1, one);

    ITW_PASS_FENCE();
    // one least-squares refinement pass; K:419-480, :524-530
    {
        const float mx[3] = {mean[0].x, mean[1].x, mean[2].x}, my[3] = {mean[0].y, mean[1].y, mean[2].y};
        const float sx[3] = {atb1[0].x, atb1[1].x, atb1[2].x}, sy[3] = {atb1[0].y, atb1[1].y, atb1[2].y};
        bc1_refine_lane(bits_x, mx, sx, p0x, p1x);
        bc1_refine_lane(bits_y, my, sy, p0y, p1y);
    }
    bc1_indices_pair<false>(px, ps, p0x, p1x, p0y, p1y, bits_x, bits_y, atb1, one);

    // linear order 0,1,2,3 -> BC1 codes 0,2,3,1; K:482-492
    PairWords w;
    {
        const u32 lo_bits = bits_x & 0x55555555u, hi_bits = bits_x & 0xAAAAAAAAu;
        w.a0 = ((u32)p1x << 16) + (u32)p0x;
        w.a1 = (hi_bits >> 1) + (hi_bits ^ (lo_bits << 1));
    }
    {
        const u32 lo_bits = bits_y & 0x55555555u, hi_bits = bits_y & 0xAAAAAAAAu;
        w.b0 = ((u32)p1y << 16) + (u32)p0y;
        w.b1 = (hi_bits >> 1) + (hi_bits ^ (lo_bits << 1));
    }
    return w;
}

// BC3 alpha half of both lanes; K:535-571.  al[k] = alpha of texel k (exact integers as floats).
ITW_HD PairWords bc3_alpha_pair(const f2 (&al)[16], const f2 one)
{
    f2 lo = splat2(255.0f), hi = splat2(0.0f);
#pragma unroll
    for (int k = 0; k < 16; k++) { lo = min2(lo, al[k]); hi = max2(hi, al[k]); }
    if (lo.x == hi.x) hi.x = lo.x + 0.1f;
    if (lo.y == hi.y) hi.y = lo.y + 0.1f;
    const f2 scale = mk2(7.0f / (hi.x - lo.x), 7.0f / (hi.y - lo.y));
    const f2 nlo = mk2(-lo.x, -lo.y);
    unsigned long long ix = 0ull, iy = 0ull;
#pragma unroll
    for (int k = 15; k >= 0; k--) {
        // (a - lo) * scale + 0.5, each step rounded as in K:553; 0.5 <= value <= 255*70 + 0.5, clamped to 0..7 after truncation
        const f2 proj = madd2(add2(al[k], nlo), scale, splat2(0.5f), one);
        const f2 t = trunc_clamp_magic(proj, 7.5f);
        int qx = 7 - (int)(float_bits(t.x) & 7u), qy = 7 - (int)(float_bits(t.y) & 7u);
        if (qx > 0) qx++;
        if (qx == 8) qx = 1;
        if (qy > 0) qy++;
        if (qy == 8) qy = 1;
        ix = (ix << 3) | (unsigned long long)qx;
        iy = (iy << 3) | (unsigned long long)qy;
    }
    PairWords w;
    // bytes: alpha0 = max, alpha1 = min (truncated, K:567), then 48 index bits
    w.a0 = (u32)(clampi(trunc_i(lo.x), 0, 255) * 256 + clampi(trunc_i(hi.x), 0, 255)) | ((u32)ix << 16);
    w.a1 = (u32)(ix >> 16);
    w.b0 = (u32)(clampi(trunc_i(lo.y), 0, 255) * 256 + clampi(trunc_i(hi.y), 0, 255)) | ((u32)iy << 16);
    w.b1 = (u32)(iy >> 16);
    return w;
}

// Two whole blocks: 2 x 16 packed RGBA8 texels in, 2 (BC1) or 4 (BC3) words out per block; K:573-596.  `px` = storage for the
// 48 converted texel-value pairs (stride ps, see ITW_PX).
template <bool kAlpha>
ITW_HD void bc1_bc3_encode_pair(const u32 (&ta)[16], const u32 (&tb)[16], u32 (&oa)[4], u32 (&ob)[4], f2* px, const int ps, const f2 one)
{
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 16; k++) ITW_PX(c, k) = bytes_to_f2(ta[k], tb[k], c);
    if (kAlpha) {                                   // alpha from the packed texels while they are still in registers
        f2 al[16];
#pragma unroll
        for (int k = 0; k < 16; k++) al[k] = bytes_to_f2(ta[k], tb[k], 3);
        const PairWords w = bc3_alpha_pair(al, one);
        oa[0] = w.a0; oa[1] = w.a1; ob[0] = w.b0; ob[1] = w.b1;
    }
    const PairWords w = bc1_colour_pair(px, ps, one);
    if (kAlpha) { oa[2] = w.a0; oa[3] = w.a1; ob[2] = w.b0; ob[3] = w.b1; }
    else { oa[0] = w.a0; oa[1] = w.a1; oa[2] = oa[3] = 0; ob[0] = w.b0; ob[1] = w.b1; ob[2] = ob[3] = 0; }
}

#if defined(__CUDACC__)
// CTA shape: 64 threads, 6 CTAs per SM = 12 warps; per CTA 24 KB of texel-value columns + one 8 KB stage buffer (+1 KB the
// system reserves per CTA: a seventh CTA does not fit into the 227 KB of an SM).
#ifndef ITW_BC1_THREADS
#define ITW_BC1_THREADS 64
#endif
#ifndef ITW_BC1_CTAS_PER_SM
#define ITW_BC1_CTAS_PER_SM 6
#endif
constexpr int kBc1PairThreads = ITW_BC1_THREADS;
constexpr int kBc1CtasPerSm = ITW_BC1_CTAS_PER_SM;
constexpr int kBc1TileBlocks = 2 * kBc1PairThreads;            // consecutive blocks per tile
constexpr int kBc1TileRowBytes = kBc1TileBlocks * 16;

// Persistent: CTA b encodes tiles b, b + gridDim.x, ...  Thread t owns blocks t and t + blockDim of the tile (consecutive
// threads read consecutive 16-byte pieces of the staged rows and write consecutive output blocks).  Per tile: wait for the TMA
// copy (cp.async.bulk, mbarrier) -> every thread takes its 2 x 16 packed texels into registers -> barrier -> thread 0 starts
// the TMA copy of the CTA's NEXT tile into the same stage buffer, which streams in behind the encode of this one.
// Needs 16-byte aligned surface rows (ptr and stride multiples of 16); other surfaces take bc1_bc3_kernel.
template <bool kAlpha>
__global__ void __launch_bounds__(kBc1PairThreads, kBc1CtasPerSm) bc1_bc3_pair_kernel(SurfaceView s, uint8_t* __restrict__ dst, long long nblocks, float one_arg)
{
    __shared__ __align__(128) unsigned char stage[4 * kBc1TileRowBytes];
    __shared__ __align__(16) f2 columns[48 * kBc1PairThreads];
    __shared__ __align__(8) unsigned long long full;
    const long long ntiles = (nblocks + kBc1TileBlocks - 1) / kBc1TileBlocks;
    auto prefetch = [&](long long tile) {
        if (tile >= ntiles) return;
        const long long first = tile * kBc1TileBlocks, left = nblocks - first;
        tma_prefetch_tile(stage, kBc1TileRowBytes, &full, s, first, (int)(left < kBc1TileBlocks ? left : kBc1TileBlocks), 16);
    };
    if (threadIdx.x == 0) mbar_init(&full, 1);
    __syncthreads();
    if (threadIdx.x == 0) prefetch(blockIdx.x);
    f2* px = columns + threadIdx.x;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        mbar_wait(&full, (unsigned)(it & 1));
        const long long first = tile * kBc1TileBlocks;
        const long long ida = first + threadIdx.x, idb = ida + kBc1PairThreads;
        u32 ta[16], tb[16], oa[4], ob[4];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint4 va = *reinterpret_cast<const uint4*>(stage + y * kBc1TileRowBytes + threadIdx.x * 16);
            const uint4 vb = *reinterpret_cast<const uint4*>(stage + y * kBc1TileRowBytes + (threadIdx.x + kBc1PairThreads) * 16);
            ta[4 * y + 0] = va.x; ta[4 * y + 1] = va.y; ta[4 * y + 2] = va.z; ta[4 * y + 3] = va.w;
            tb[4 * y + 0] = vb.x; tb[4 * y + 1] = vb.y; tb[4 * y + 2] = vb.z; tb[4 * y + 3] = vb.w;
        }
        __syncthreads();                                               // every thread has read the stage: it may be refilled
        if (threadIdx.x == 0) prefetch(tile + gridDim.x);
        // (blocks past the end of the surface read stale shared memory: any bit pattern is a valid input, nothing is stored)
        bc1_bc3_encode_pair<kAlpha>(ta, tb, oa, ob, px, kBc1PairThreads, splat2(one_arg));       // one_arg == 1.0f, see madd2 (itw_device.cuh)
        if (kAlpha) {
            if (ida < nblocks) reinterpret_cast<uint4*>(dst)[ida] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
            if (idb < nblocks) reinterpret_cast<uint4*>(dst)[idb] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        } else {
            if (ida < nblocks) reinterpret_cast<uint2*>(dst)[ida] = make_uint2(oa[0], oa[1]);
            if (idb < nblocks) reinterpret_cast<uint2*>(dst)[idb] = make_uint2(ob[0], ob[1]);
        }
    }
}
#endif

}  // namespace itw
