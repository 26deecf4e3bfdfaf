// bc67_core.cuh -- per-lane numerics shared by the BC7 and BC6H kernels
// (reference: kernel.ispc:760-971 PCA helpers, :976-1128 endpoint quantisation, :1133-1262 index
// search and least-squares refinement, :1694-1805 bitstream helpers; cited as K:line).
//
// Every function here is the work of ONE lane on ONE candidate of ONE 4x4 block: `px` points at
// the block's 64 planar floats (px[c*16 + k]) in shared memory, read by all lanes of the warp at
// the same address (broadcast, conflict-free).  Lanes never exchange partial float sums, so the
// reference's texel-order accumulation (rule F1) is preserved exactly.
#pragma once
#include "itw_device.cuh"

namespace itw {

// ---------------------------------------------------------------------------------------------
// symmetric mat-vec and power iteration; K:169-229.  Packing [xx xy xz xw yy yz yw zz zw ww].
// ---------------------------------------------------------------------------------------------
ITW_HD void sym_apply(float (&out)[4], const float (&m)[10], const float (&v)[4], int channels)
{
    if (channels == 3) {
        out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
        out[1] = m[1] * v[0] + m[4] * v[1] + m[5] * v[2];
        out[2] = m[2] * v[0] + m[5] * v[1] + m[7] * v[2];
        out[3] = 0.0f;
    } else {
        out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3] * v[3];
        out[1] = m[1] * v[0] + m[4] * v[1] + m[5] * v[2] + m[6] * v[3];
        out[2] = m[2] * v[0] + m[5] * v[1] + m[7] * v[2] + m[8] * v[3];
        out[3] = m[3] * v[0] + m[6] * v[1] + m[8] * v[2] + m[9] * v[3];
    }
}
// Start at all-ones, renormalise (1/sqrt, two exact ops) after every odd iteration; K:207-229
template <int kIterations>
ITW_HD void power_axis(float (&axis)[4], const float (&m)[10], int channels)
{
    float v[4] = {1.0f, 1.0f, 1.0f, 1.0f};
#pragma unroll
    for (int it = 0; it < kIterations; it++) {
        sym_apply(axis, m, v, channels);
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = axis[c];
        if (it & 1) {
            float n2 = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (c < channels) n2 += axis[c] * axis[c];
            float rn = 1.0f / sqrtf(n2);
#pragma unroll
            for (int c = 0; c < 4; c++) v[c] *= rn;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) axis[c] = v[c];
}

// ---------------------------------------------------------------------------------------------
// masked raw moments; K:763-803.  st[0..9] second order, st[10..13] sums, st[14] count.
// Every texel is visited and scaled by its 0/1 flag, exactly as the reference does.
// ---------------------------------------------------------------------------------------------
ITW_HD void masked_moments(float (&st)[15], const float* px, int mask, int channels)
{
#pragma unroll
    for (int i = 0; i < 15; i++) st[i] = 0.0f;
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        float flag = (float)((mask >> k) & 1);
        float r = px[k] * flag, g = px[16 + k] * flag, b = px[32 + k] * flag;
        st[14] += flag;
        st[10] += r; st[11] += g; st[12] += b;
        st[0] += r * r; st[1] += r * g; st[2] += r * b;
        st[4] += g * g; st[5] += g * b;
        st[7] += b * b;
        if (channels == 4) {
            float a = px[48 + k] * flag;
            st[13] += a;
            st[3] += r * a; st[6] += g * a; st[8] += b * a; st[9] += a * a;
        }
    }
}
// cov = sum(xy) - sum(x)*sum(y)/n; K:805-823.  Slots not owned by `channels` are zero.
ITW_HD void covariance_of(float (&cov)[10], const float (&st)[15], int channels)
{
    // x / n with the FMA-corrected quotient: equal to the IEEE quotient for all normal operands (proved by exhaustion,
    // tests/test_gpu_division.py; the numerators are 0 or products of texel sums, far from the subnormal range)
    const float n = st[14], rn = 1.0f / n;
    cov[0] = st[0] - div_by_rcp(st[10] * st[10], n, rn);
    cov[1] = st[1] - div_by_rcp(st[10] * st[11], n, rn);
    cov[2] = st[2] - div_by_rcp(st[10] * st[12], n, rn);
    cov[4] = st[4] - div_by_rcp(st[11] * st[11], n, rn);
    cov[5] = st[5] - div_by_rcp(st[11] * st[12], n, rn);
    cov[7] = st[7] - div_by_rcp(st[12] * st[12], n, rn);
    cov[3] = cov[6] = cov[8] = cov[9] = 0.0f;
    if (channels == 4) {
        cov[3] = st[3] - div_by_rcp(st[10] * st[13], n, rn);
        cov[6] = st[6] - div_by_rcp(st[11] * st[13], n, rn);
        cov[8] = st[8] - div_by_rcp(st[12] * st[13], n, rn);
        cov[9] = st[9] - div_by_rcp(st[13] * st[13], n, rn);
    }
}

// PCA line through the masked texels, endpoints at the extreme projections; K:834-905.
// clamp255 = K:896 block_segment (BC7); otherwise K:857 block_segment_core (BC6H).
// Writes ep[0..channels) and ep[4..4+channels).
ITW_HD void fit_segment_inl(float (&ep)[8], const float* px, int mask, int channels, bool clamp255)
{
    float st[15], cov[10], mean[4], axis[4];
    masked_moments(st, px, mask, channels);
    covariance_of(cov, st, channels);
    const float rcount = 1.0f / st[14];
#pragma unroll
    for (int c = 0; c < 4; c++) mean[c] = (c < channels) ? div_by_rcp(st[10 + c], st[14], rcount) : 0.0f;

    const float inv_var = 1.0f / (256.0f * 256.0f);
#pragma unroll
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = 0.001f * 0.001f;
    cov[0] += eps; cov[4] += eps; cov[7] += eps; cov[9] += eps;
    power_axis<8>(axis, cov, channels);

    float lo = inf_f(), hi = -inf_f();
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        float d = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < channels) d += axis[c] * (px[16 * c + k] - mean[c]);
        if ((mask >> k) & 1) {
            lo = min_sse(lo, d);
            hi = max_sse(hi, d);
        }
    }
    if (hi - lo < 1.0f) { lo -= 0.5f; hi += 0.5f; }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (c < channels) {
            float a = lo * axis[c] + mean[c], b = hi * axis[c] + mean[c];
            if (clamp255) { a = clamp_sse(a, 0.0f, 255.0f); b = clamp_sse(b, 0.0f, 255.0f); }
            ep[c] = a;
            ep[4 + c] = b;
        }
    }
}

// trace - lambda_max of a covariance (rescaled in place); K:907-939
ITW_HD float residual_bound(float (&cov)[10], int channels)
{
    const float inv_var = 1.0f / (256.0f * 256.0f);
#pragma unroll
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = 0.001f * 0.001f;
    cov[0] += eps; cov[4] += eps; cov[7] += eps;      // three diagonal slots only; K:918-920
    float axis[4], mv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    power_axis<4>(axis, cov, channels);
    sym_apply(mv, cov, axis, channels);
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (c < channels) s += sq(mv[c]);
    float lambda = sqrtf(s);
    float bound = cov[0] + cov[4] + cov[7];
    if (channels == 4) bound += cov[9];
    bound -= lambda;
    return max_sse(bound, 0.0f);
}
// Ranking key of a two-subset shape: shape + 64*(int)(256*sqrt(bound0+bound1)), where subset 1's
// moments are full - subset 0; K:952-971, :1403-1410
ITW_HD_NOINLINE int split_bound_key(const float* px, int shape, const float (&full)[15], int channels)
{
    float st[15], c1[10], c2[10];
    masked_moments(st, px, shape_mask(shape, 0), channels);
    covariance_of(c1, st, channels);
#pragma unroll
    for (int i = 0; i < 15; i++) st[i] = full[i] - st[i];
    covariance_of(c2, st, channels);
    float b = 0.0f;
    b += residual_bound(c1, channels);
    b += residual_bound(c2, channels);
    float bound = sqrtf(b) * 256.0f;
    return shape + (int)((unsigned)cvt_x86(bound) * 64u);   // wrapping multiply, as the ISPC int does
}

// Least-squares endpoints of one subset from its current indices; K:1198-1262
ITW_HD void solve_endpoints_inl(float (&ep)[8], const float* px, int bits, u32 idx0, u32 idx1, int mask, int channels)
{
    const float top = (float)((1 << bits) - 1);
    float atb1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float count = 0.0f, sq1 = 0.0f, sqq = 0.0f;
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        if (((mask >> k) & 1) == 0) continue;
        float q = (float)(((k < 8 ? idx0 : idx1) >> (4 * (k & 7))) & 15u);
        float x = (float)cvt_x86(top - q);
        sq1 += q;
        sqq += q * q;
        count += 1.0f;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < channels) {
                float v = px[16 * c + k];
                sum[c] += v;
                atb1[c] += x * v;
            }
    }
    float cxx = count * sq(top) - (2.0f * top) * sq1 + sqq;
    float cyy = sqq;
    float cxy = top * sq1 - sqq;
    float det = cxx * cyy - cxy * cxy;
    float scale = top / det;
    bool flat = fabsf(det) < 0.001f;
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (c < channels) {
            float atb2 = top * sum[c] - atb1[c];
            float a = (atb1[c] * cyy - atb2 * cxy) * scale;
            float b = (atb2 * cxx - atb1[c] * cxy) * scale;
            if (flat) { a = sum[c] / count; b = a; }
            ep[c] = a;
            ep[4 + c] = b;
        }
}

// ---------------------------------------------------------------------------------------------
// small BC7 format helpers; K:976-981
// ---------------------------------------------------------------------------------------------
ITW_HD int expand_bits(int v, int bits)
{
    int vv = v << (8 - bits);
    return vv + (int)((u32)vv >> bits);
}
ITW_HD int bc7_pairs(int mode) { return (mode == 0 || mode == 2) ? 3 : ((mode == 1 || mode == 3 || mode == 7) ? 2 : 1); }


// ---------------------------------------------------------------------------------------------
// bitstream helpers; K:1694-1805
// ---------------------------------------------------------------------------------------------
// Index payload: texel 0 and the anchors of subsets 1,2 are written one bit narrower (their MSB is
// zero after orientation).  Bit-identical to K's write-then-delete (K:1767-1805).
ITW_HD void put_indices(BitSink& s, u32 idx0, u32 idx1, int bits, int flips, int anchor1, int anchor2)
{
    const int top = (1 << bits) - 1;
    for (int k = 0; k < 16; k++) {
        int q = (int)(((k < 8 ? idx0 : idx1) >> (4 * (k & 7))) & 15u);
        if ((flips >> k) & 1) q = top - q;
        bool narrow = (k == 0) || (k == anchor1) || (k == anchor2);
        s.put(narrow ? bits - 1 : bits, (u32)q);
    }
}
// Single-subset variant (BC7 modes 4,5,6; BC6H one-region modes); K:1694-1706
ITW_HD void orient_single(int* q, int width, u32& idx0, u32& idx1, int bits)
{
    const int levels = 1 << bits;
    if ((int)(idx0 & 15u) >= levels / 2) {
        for (int c = 0; c < width; c++) { int t = q[c]; q[c] = q[width + c]; q[width + c] = t; }
        u32 all = 0x11111111u * (u32)(levels - 1);
        idx0 = all - idx0;
        idx1 = all - idx1;
    }
}

// ---------------------------------------------------------------------------------------------
// rank of a key among 64 (or 32) unique keys = its position after the reference's ascending
// selection sort (K:1365-1384); keys are unique because the shape id is their low 6 bits.
// ---------------------------------------------------------------------------------------------
ITW_HD int rank_of(const int* keys, int n, int i)
{
    int mine = keys[i], r = 0;
    for (int j = 0; j < n; j++) r += (keys[j] < mine) ? 1 : 0;
    return r;
}

}  // namespace itw
