// bc1_bc3.cuh -- BC1 / BC3 encoder kernels (reference: kernel.ispc:231-614, cited as K:line).
//
// Mapping: one thread owns one 4x4 block (the reference's own lane<->block mapping, K:600-604).
// A warp covers 32 horizontally adjacent blocks, so each of the four texel rows is one fully
// coalesced 512-byte request of 128-bit loads and the packed output is one coalesced 256-byte
// (BC1) or 512-byte (BC3) store.  BC1's covariance sums are inexact float sums whose order
// matters on high-variance blocks (K:377-417), so the per-block math stays sequential in one
// thread; there is no cross-lane traffic at all.
//
// Algorithmic traffic per block: 64 B read, 8 B (BC1) / 16 B (BC3) written.
#pragma once
#include "itw_device.cuh"

namespace itw {

// ---- RGB565 helpers; K:234-259 ----
ITW_HD int scale8(int a, int b) { int t = a * b + 128; return (t + (t >> 8)) >> 8; }
ITW_HD int pack565(float r, float g, float b)
{
    // callers clamp r,g,b to [0,255], so plain truncation equals the x86 conversion
    int v = (scale8(trunc_i(r), 31) << 11) + (scale8(trunc_i(g), 63) << 5) + scale8(trunc_i(b), 31);
    return v & 0xFFFF;
}
ITW_HD void unpack565(float c[3], int p)
{
    int b = p & 31, g = (p >> 5) & 63, r = (p >> 11) & 31;
    c[0] = (float)((r << 3) + (r >> 2));
    c[1] = (float)((g << 2) + (g >> 4));
    c[2] = (float)((b << 3) + (b >> 2));
}

// Linear 2-bit indices along p0 -> p1; K:308-344 (p0 == p1 -> NaN -> INT_MIN -> index 0)
ITW_HD u32 bc1_linear_indices(const float (&px)[3][16], int p0, int p1)
{
    float a[3], b[3], dir[3];
    unpack565(a, p0);
    unpack565(b, p1);
#pragma unroll
    for (int c = 0; c < 3; c++) dir[c] = b[c] - a[c];
    // Sums the reference starts from 0.0f start from their first term throughout this file: 0 + x differs from x only when x is
    // -0, and no consumer can see the sign of a zero here -- squares and sums of bytes are never -0; a projection d only meets
    // min/max, a difference, or d + bias with bias != -0; a covariance cross term is -0 only if every product is, which the
    // deviations from an exact mean (never all negative, never -0) rule out.
    float n2 = sq(dir[0]);
#pragma unroll
    for (int c = 1; c < 3; c++) n2 += sq(dir[c]);
    float inv = 1.0f / n2;
#pragma unroll
    for (int c = 0; c < 3; c++) dir[c] *= inv * 3.0f;
    float bias = 0.5f;
#pragma unroll
    for (int c = 0; c < 3; c++) bias -= a[c] * dir[c];
    u32 bits = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float d = px[0][k] * dir[0];
#pragma unroll
        for (int c = 1; c < 3; c++) d += px[c][k] * dir[c];
        // |d + bias| < 2^31; the only special value is NaN (p0 == p1), INT_MIN on x86 and 0 here: both clamp to 0
        int q = clampi(trunc_i(d + bias), 0, 3);
        bits |= (u32)q << (2 * k);      // == K's bits += q*4^k: the fields never overlap
    }
    return bits;
}

// The colour half shared by BC1 and BC3; K:494-533
// plane[c][i] = channel c of texels 4i..4i+3, one byte each: the two sums of the reference that are exact integers (the
// channel sums behind the mean, and the index-weighted sums of the refinement) are taken with IDP.4A on these words.
ITW_HD void bc1_colour_block(const float (&px)[3][16], const u32 (&plane)[3][4], u32& w0, u32& w1)
{
    // mean, then centred covariance accumulated in texel order; K:377-417
    float mean[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        u32 acc = 0u;                                            // sum of sixteen bytes: the reference's float sum is this integer
#pragma unroll
        for (int i = 0; i < 4; i++) acc = dp4a_u8(plane[c][i], 0x01010101u, acc);
        mean[c] = (float)(int)acc / 16.0f;
    }
    float crr = 0.0f, crg = 0.0f, crb = 0.0f, cgg = 0.0f, cgb = 0.0f, cbb = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float r = px[0][k] - mean[0], g = px[1][k] - mean[1], b = px[2][k] - mean[2];
        if (k == 0) { crr = r * r; crg = r * g; crb = r * b; cgg = g * g; cgb = g * b; cbb = b * b; }
        else {
            crr += r * r; crg += r * g; crb += r * b;
            cgg += g * g; cgb += g * b; cbb += b * b;
        }
    }
    const float eps = 0.001f;
    crr += eps; cgg += eps; cbb += eps;

    // 4 power iterations from (1,1,1), renormalised after iterations 1 and 3; K:184-205
    float v0 = 1.0f, v1 = 1.0f, v2 = 1.0f;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        float a0 = crr * v0 + crg * v1 + crb * v2;
        float a1 = crg * v0 + cgg * v1 + cgb * v2;
        float a2 = crb * v0 + cgb * v1 + cbb * v2;
        v0 = a0; v1 = a1; v2 = a2;
        if (it & 1) {
            float n2 = a0 * a0;
            n2 += a1 * a1; n2 += a2 * a2;
            float rn = 1.0f / sqrtf(n2);
            v0 *= rn; v1 *= rn; v2 *= rn;
        }
    }
    const float axis[3] = {v0, v1, v2};

    // extreme projections -> endpoints; K:274-306 (min starts at 65536, max at 0)
    float dmin = 65536.0f, dmax = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float d = (px[0][k] - mean[0]) * axis[0];
#pragma unroll
        for (int c = 1; c < 3; c++) d += (px[c][k] - mean[c]) * axis[c];
        dmin = fminf(dmin, d);            // d is finite: identical to the reference's (a<b)?a:b, one FMNMX
        dmax = fmaxf(dmax, d);
    }
    if (dmax - dmin < 1.0f) { dmin -= 0.5f; dmax += 0.5f; }
    float n2 = axis[0] * axis[0];
#pragma unroll
    for (int c = 1; c < 3; c++) n2 += axis[c] * axis[c];
    float inv = 1.0f / n2;
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        lo[c] = clamp_sse(mean[c] + dmin * inv * axis[c], 0.0f, 255.0f);
        hi[c] = clamp_sse(mean[c] + dmax * inv * axis[c], 0.0f, 255.0f);
    }
    int p0 = pack565(lo[0], lo[1], lo[2]), p1 = pack565(hi[0], hi[1], hi[2]);
    if (p0 < p1) { int t = p0; p0 = p1; p1 = t; }
    u32 bits = bc1_linear_indices(px, p0, p1);

    // one least-squares refinement pass; K:419-480, :524-530
    float ea[3], eb[3];
    if ((bits ^ (bits * 4u)) < 4u) {
#pragma unroll
        for (int c = 0; c < 3; c++) ea[c] = eb[c] = mean[c];
    } else {
        // sums of q and q^2 over the sixteen 2-bit indices are exact small integers: count bits instead of
        // adding floats (q = 2*hi + lo, q^2 = 4*hi + 4*hi*lo + lo).  x*px and its running sum (<= 12240) are
        // exact too, so the fused multiply-add gives the reference's value.
        const u32 lo_b = bits & 0x55555555u, hi_b = (bits >> 1) & 0x55555555u;
        const float sq1 = (float)(popcount32(lo_b) + 2 * popcount32(hi_b));
        const float sqq = (float)(popcount32(lo_b) + 4 * popcount32(hi_b) + 4 * popcount32(lo_b & hi_b));
        // sum of (3 - q) * texel: products and running sums (<= 12240) are exact integers in the reference's floats
        u32 isum[3] = {0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u32 t = (bits >> (8 * i)) & 255u;                    // four 2-bit indices ...
            t = (t | (t << 12)) & 0x000F000Fu;
            t = (t | (t << 6)) & 0x03030303u;                    // ... one per byte
            const u32 x = 0x03030303u - t;                       // 3 - q, bytewise, no borrows
#pragma unroll
            for (int c = 0; c < 3; c++) isum[c] = dp4a_u8(x, plane[c][i], isum[c]);
        }
        const float atb1[3] = {(float)(int)isum[0], (float)(int)isum[1], (float)(int)isum[2]};
        float cxx = 16.0f * 9.0f - 6.0f * sq1 + sqq;
        float cyy = sqq;
        float cxy = 3.0f * sq1 - sqq;
        float scale = 3.0f * (1.0f / (cxx * cyy - cxy * cxy));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float total = mean[c] * 16.0f;
            float atb2 = 3.0f * total - atb1[c];
            float a = (atb1[c] * cyy - atb2 * cxy) * scale;
            float b = (atb2 * cxx - atb1[c] * cxy) * scale;
            ea[c] = clamp_sse(a, 0.0f, 255.0f);
            eb[c] = clamp_sse(b, 0.0f, 255.0f);
        }
    }
    p0 = pack565(ea[0], ea[1], ea[2]);
    p1 = pack565(eb[0], eb[1], eb[2]);
    if (p0 < p1) { int t = p0; p0 = p1; p1 = t; }
    bits = bc1_linear_indices(px, p0, p1);

    // linear order 0,1,2,3 -> BC1 codes 0,2,3,1; K:482-492
    u32 lo_bits = bits & 0x55555555u, hi_bits = bits & 0xAAAAAAAAu;
    w0 = ((u32)p1 << 16) + (u32)p0;
    w1 = (hi_bits >> 1) + (hi_bits ^ (lo_bits << 1));
}

// BC3 alpha half; K:535-571
ITW_HD void bc3_alpha_block(const float (&a)[16], u32& w0, u32& w1)
{
    float lo = 255.0f, hi = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) { lo = fminf(lo, a[k]); hi = fmaxf(hi, a[k]); }
    if (lo == hi) hi = lo + 0.1f;
    unsigned long long idx = 0;
    float scale = 7.0f / (hi - lo);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float proj = (a[k] - lo) * scale + 0.5f;
        int q = clampi(trunc_i(proj), 0, 7);                 // 0.5 <= proj <= 255*70 + 0.5
        q = 7 - q;
        if (q > 0) q++;
        if (q == 8) q = 1;
        idx |= (unsigned long long)q << (3 * k);
    }
    // bytes: alpha0 = max, alpha1 = min, then 48 index bits
    u32 head = (u32)(clampi(trunc_i(lo), 0, 255) * 256 + clampi(trunc_i(hi), 0, 255));
    w0 = head | ((u32)idx << 16);
    w1 = (u32)(idx >> 16);
}

// Fetch the 16 texels of block (bx,by): four 128-bit loads when the surface is 16-byte aligned
// (any padded or tight RGBA8 surface whose stride is a multiple of 16), else byte-safe loads.
template <bool kVec16>
ITW_HD void fetch_rows_rgba8(u32 (&tex)[16], const SurfaceView& s, int bx, int by)
{
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint8_t* p = s.ptr + (size_t)(by * 4 + y) * (size_t)s.stride + (size_t)bx * 16;
        if (kVec16) {
#if defined(__CUDA_ARCH__)
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
            tex[4 * y + 0] = v.x; tex[4 * y + 1] = v.y; tex[4 * y + 2] = v.z; tex[4 * y + 3] = v.w;
#else
            for (int x = 0; x < 4; x++) tex[4 * y + x] = reinterpret_cast<const u32*>(p)[x];
#endif
        } else {
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const uint8_t* t = p + 4 * x;
                tex[4 * y + x] = (u32)t[0] | ((u32)t[1] << 8) | ((u32)t[2] << 16) | ((u32)t[3] << 24);
            }
        }
    }
}

// One whole block: 16 packed RGBA8 texels in, 2 (BC1) or 4 (BC3) words out; K:573-596
template <bool kAlpha>
ITW_HD void bc1_bc3_encode_block(const u32 (&tex)[16], u32 (&out)[4])
{
    float px[3][16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        px[0][k] = (float)(tex[k] & 255u);
        px[1][k] = (float)((tex[k] >> 8) & 255u);
        px[2][k] = (float)((tex[k] >> 16) & 255u);
    }
    u32 plane[3][4];                                             // 4x4 byte transposes: two PRMT stages per four texels
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u32 rg01 = byte_perm(tex[4 * i], tex[4 * i + 1], 0x5140u), rg23 = byte_perm(tex[4 * i + 2], tex[4 * i + 3], 0x5140u);
        const u32 ba01 = byte_perm(tex[4 * i], tex[4 * i + 1], 0x7362u), ba23 = byte_perm(tex[4 * i + 2], tex[4 * i + 3], 0x7362u);
        plane[0][i] = byte_perm(rg01, rg23, 0x5410u);
        plane[1][i] = byte_perm(rg01, rg23, 0x7632u);
        plane[2][i] = byte_perm(ba01, ba23, 0x5410u);
    }
    if (kAlpha) {
        float al[16];
#pragma unroll
        for (int k = 0; k < 16; k++) al[k] = (float)(tex[k] >> 24);
        bc3_alpha_block(al, out[0], out[1]);
        bc1_colour_block(px, plane, out[2], out[3]);
    } else {
        bc1_colour_block(px, plane, out[0], out[1]);
        out[2] = out[3] = 0;
    }
}

#if defined(__CUDACC__)
// CTAs of 128 threads per SM the register budget is cut for, measured on B200 (tools/bc1_variants.sh,
// profiles/r2_bc1_variants.txt): BC1 58.3 us with 4 (128 registers, 16 bytes of spills), 60.2 with 5 (96 registers), 62.3 with 6,
// 64.7 with 3; BC3 69.2 / 70.6 / 72.8 / 74.8 us.
#ifndef ITW_BC1_MIN_CTAS
#define ITW_BC1_MIN_CTAS(alpha) 4
#endif
template <bool kAlpha, bool kVec16>
__global__ void __launch_bounds__(128, ITW_BC1_MIN_CTAS(kAlpha)) bc1_bc3_kernel(SurfaceView s, uint8_t* __restrict__ dst)
{
    const int bw = s.width >> 2, bh = s.height >> 2;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)bw * bh) return;
    const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);

    u32 tex[16], out[4];
    fetch_rows_rgba8<kVec16>(tex, s, bx, by);
    bc1_bc3_encode_block<kAlpha>(tex, out);
    if (kAlpha) reinterpret_cast<uint4*>(dst)[id] = make_uint4(out[0], out[1], out[2], out[3]);
    else        reinterpret_cast<uint2*>(dst)[id] = make_uint2(out[0], out[1]);
}
#endif

}  // namespace itw
