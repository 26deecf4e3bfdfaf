/*
 * itw_oracle.cpp -- CPU ORACLE for the BCn hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A scalar, strict-IEEE restatement of the reference's ISPC encoder
 *   /root/reference/IntelCompressionPlugin/kernel.ispc   (cited below as K:line)
 *   /root/reference/3rdParty/Intel/Source/ispc_texcomp.cpp (profiles, cited as TCc:line)
 * One SIMD lane of the reference owns one 4x4 block and lanes never talk (K:600-604,
 * :2032-2036, :3134-3138), so a scalar per-block program is a faithful execution of it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the CPU baseline.  The product
 * (libitw_bcn.so) never links, loads or calls it.
 *
 * PINNING.  The reference ships no golden vectors, KATs or tests.  This
 * restatement is pinned two ways: (1) against oracle/_ref/libitw_ref.so -- the reference's OWN
 * kernel.ispc + ispc_texcomp.cpp compiled scalar by oracle/build_ref.py -- byte-for-byte over
 * every format/profile on random, gradient and degenerate inputs (tests/test_oracle_vs_ref.py);
 * (2) against the committed digests under tests/golden/ that were produced by that build.
 * What stays unpinned is the gap between any scalar strict-IEEE execution of kernel.ispc and the
 * shipped `ispc --opt=fast-math` SIMD binary (approximate rcp/rsqrt differ per CPU vendor); see
 * DESIGN.md "Canonical float model".
 *
 * Canonical float model: binary32 RNE, no FMA contraction (build with
 * -ffp-contract=off), source order of evaluation; rcp(x)=1/x, rsqrt(x)=1/sqrt(x);
 * float->int is x86 cvttss2si (NaN / out of range -> INT_MIN); min/max are SSE-ordered
 * ((a<b)?a:b / (a>b)?a:b); reads of never-written scratch see zero.
 *
 * The formulation deliberately differs from K where that gives an independent check: the BC7/BC6H
 * bit writer emits anchor indices at reduced width directly instead of K's write-then-delete
 * (K:1746-1805), and the BC6H header is produced from a field-layout table (the D3D BC6H format
 * definition) instead of K's 14 hand-written cases (K:2392-2980).
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "../include/itw_bcn.h"

namespace {

typedef uint32_t u32;

// ------------------------------------------------------------------------------------------
// canonical scalar helpers
// ------------------------------------------------------------------------------------------
const float kInf = std::numeric_limits<float>::infinity();

// float -> int with x86 cvttss2si semantics (rule F3)
inline int f2i(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
inline float fmin_sse(float a, float b) { return (a < b) ? a : b; }   // rule F4
inline float fmax_sse(float a, float b) { return (a > b) ? a : b; }
inline float fclamp(float v, float lo, float hi) { return fmin_sse(fmax_sse(v, lo), hi); }
inline int imin(int a, int b) { return (a < b) ? a : b; }
inline int imax(int a, int b) { return (a > b) ? a : b; }
inline int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); }
inline float sqf(float v) { return v * v; }
inline float rcp_exact(float x) { return 1.0f / x; }                   // rule F2
inline float rsqrt_exact(float x) { return 1.0f / sqrtf(x); }

// ------------------------------------------------------------------------------------------
// BC7 / BC6H partition data (format definition; same content as K:690-752 in another layout)
// ------------------------------------------------------------------------------------------
// two-subset shapes: bit k set <=> texel k belongs to subset 1
const uint16_t kShape2[64] = {
    0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80,
    0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8, 0xFF00, 0xFFF0, 0xF000,
    0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE,
    0x088C, 0x3110, 0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C,
    0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696, 0xA55A,
    0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660,
    0x0272, 0x04E4, 0x4E40, 0x2720, 0xC936, 0x936C, 0x39C6, 0x639C,
    0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22,
};
// three-subset shapes: 2 bits per texel (texel k in bits 2k..2k+1)
const u32 kShape3[64] = {
    0xAA685050, 0x6A5A5040, 0x5A5A4200, 0x5450A0A8, 0xA5A50000, 0xA0A05050, 0x5555A0A0, 0x5A5A5050,
    0xAA550000, 0xAA555500, 0xAAAA5500, 0x90909090, 0x94949494, 0xA4A4A4A4, 0xA9A59450, 0x2A0A4250,
    0xA5945040, 0x0A425054, 0xA5A5A500, 0x55A0A0A0, 0xA8A85454, 0x6A6A4040, 0xA4A45000, 0x1A1A0500,
    0x0050A4A4, 0xAAA59090, 0x14696914, 0x69691400, 0xA08585A0, 0xAA821414, 0x50A4A450, 0x6A5A0200,
    0xA9A58000, 0x5090A0A8, 0xA8A09050, 0x24242424, 0x00AA5500, 0x24924924, 0x24499224, 0x50A50A50,
    0x500AA550, 0xAAAA4444, 0x66660000, 0xA5A0A5A0, 0x50A050A0, 0x69286928, 0x44AAAA44, 0x66666600,
    0xAA444444, 0x54A854A8, 0x95809580, 0x96969600, 0xA85454A8, 0x80959580, 0xAA141414, 0x96960000,
    0xAAAA1414, 0xA05050A0, 0xA0A5A5A0, 0x96000000, 0x40804080, 0xA9A8A9A8, 0xAAAAAA44, 0x2A4A5254,
};
// anchor ("fix-up") texel of subset 1 for two-subset shapes
const uint8_t kAnchor2[64] = {
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2,
    15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6,
    6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15,
};
// anchor texels of subsets 1 and 2 for three-subset shapes
const uint8_t kAnchor3a[64] = {
    3, 3, 15, 15, 8, 3, 15, 15, 8, 8, 6, 6, 6, 5, 3, 3,
    3, 3, 8, 15, 3, 3, 6, 10, 5, 8, 8, 6, 8, 5, 15, 15,
    8, 15, 3, 5, 6, 10, 8, 15, 15, 3, 15, 5, 15, 15, 15, 15,
    3, 15, 5, 5, 5, 8, 5, 10, 5, 10, 8, 13, 15, 12, 3, 3,
};
const uint8_t kAnchor3b[64] = {
    15, 8, 8, 3, 15, 15, 3, 8, 15, 15, 15, 15, 15, 15, 15, 8,
    15, 8, 15, 3, 15, 8, 15, 8, 3, 15, 6, 10, 15, 15, 10, 8,
    15, 3, 15, 10, 10, 8, 9, 10, 6, 15, 8, 15, 3, 6, 6, 8,
    15, 3, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 3, 15, 15, 8,
};

// Shape ids follow K's convention: 0..63 two-subset, 64..127 three-subset (K:1313-1314).
u32 shape_pattern(int shape)                 // 2 bits per texel; K:688 get_pattern
{
    if (shape >= 64) return kShape3[shape - 64];
    u32 p = 0;
    for (int k = 0; k < 16; k++) p |= (u32)((kShape2[shape] >> k) & 1) << (2 * k);
    return p;
}
int shape_mask(int shape, int subset)        // 16-bit texel mask of a subset; K:712 get_pattern_mask
{
    u32 p = shape_pattern(shape);
    int m = 0;
    for (int k = 0; k < 16; k++)
        if ((int)((p >> (2 * k)) & 3) == subset) m |= 1 << k;
    return m;
}
void shape_anchors(int shape, int anchor[3])  // K:741 get_skips
{
    anchor[0] = 0;
    if (shape < 64) { anchor[1] = kAnchor2[shape]; anchor[2] = 0; }
    else            { anchor[1] = kAnchor3a[shape - 64]; anchor[2] = kAnchor3b[shape - 64]; }
}

// BC7 interpolation weights; K:675-686
const int kWeights2[4]  = {0, 21, 43, 64};
const int kWeights3[8]  = {0, 9, 18, 27, 37, 46, 55, 64};
const int kWeights4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
const int* weights_for(int bits) { return bits == 2 ? kWeights2 : (bits == 3 ? kWeights3 : kWeights4); }

// ------------------------------------------------------------------------------------------
// surface access; K:105-160
// ------------------------------------------------------------------------------------------
// Texels of block (bx,by) as planar floats px[c*16 + y*4 + x]; K:105-132
void fetch_rgba8(float* px, const rgba_surface* s, int bx, int by, int channels)
{
    for (int y = 0; y < 4; y++) {
        const uint8_t* row = s->ptr + (size_t)(by * 4 + y) * (size_t)s->stride;
        for (int x = 0; x < 4; x++) {
            const uint8_t* t = row + (size_t)(bx * 4 + x) * 4;
            for (int c = 0; c < channels; c++) px[16 * c + 4 * y + x] = (float)(int)t[c];
        }
    }
}
// RGBA16F: the three half bit patterns as integers 0..65535, alpha plane zero; K:134-151
void fetch_rgba16(float* px, const rgba_surface* s, int bx, int by)
{
    for (int y = 0; y < 4; y++) {
        const uint8_t* row = s->ptr + (size_t)(by * 4 + y) * (size_t)s->stride;
        for (int x = 0; x < 4; x++) {
            const uint8_t* t = row + (size_t)(bx * 4 + x) * 8;
            for (int c = 0; c < 3; c++) {
                uint16_t h;
                memcpy(&h, t + 2 * c, 2);
                px[16 * c + 4 * y + x] = (float)(int)h;
            }
            px[48 + 4 * y + x] = 0.0f;
        }
    }
}
// Blocks are stored in raster order with the pitch derived from src->width; K:153-160
void emit_block(uint8_t* dst, const rgba_surface* s, int bx, int by, const u32* words, int nwords)
{
    size_t off = ((size_t)by * (size_t)(s->width / 4) + (size_t)bx) * (size_t)nwords * 4;
    memcpy(dst + off, words, (size_t)nwords * 4);
}

// ------------------------------------------------------------------------------------------
// power iteration; K:162-229
// ------------------------------------------------------------------------------------------
// Symmetric matrices are packed [xx xy xz xw yy yz yw zz zw ww] (10 slots) as in K:169-182.
void sym_apply(float out[4], const float m[10], const float v[4], int channels)
{
    if (channels == 3) {                                    // K:169-174
        out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
        out[1] = m[1] * v[0] + m[4] * v[1] + m[5] * v[2];
        out[2] = m[2] * v[0] + m[5] * v[1] + m[7] * v[2];
    } else {                                                // K:176-182
        out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3] * v[3];
        out[1] = m[1] * v[0] + m[4] * v[1] + m[5] * v[2] + m[6] * v[3];
        out[2] = m[2] * v[0] + m[5] * v[1] + m[7] * v[2] + m[8] * v[3];
        out[3] = m[3] * v[0] + m[6] * v[1] + m[8] * v[2] + m[9] * v[3];
    }
}
// Dominant eigenvector estimate: start at all-ones, renormalise after every odd iteration;
// K:184-205 (3x3, BC1) and K:207-229 (3 or 4 channels).
void power_axis(float axis[4], const float m[10], int iterations, int channels)
{
    float v[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    for (int it = 0; it < iterations; it++) {
        sym_apply(axis, m, v, channels);
        for (int c = 0; c < channels; c++) v[c] = axis[c];
        if (it % 2 == 1) {
            float n2 = 0.0f;
            for (int c = 0; c < channels; c++) n2 += axis[c] * axis[c];
            float rn = rsqrt_exact(n2);
            for (int c = 0; c < channels; c++) v[c] *= rn;
        }
    }
    for (int c = 0; c < channels; c++) axis[c] = v[c];
}

// ==========================================================================================
// BC1 / BC3; K:231-614
// ==========================================================================================
int scale8(int a, int b) { int t = a * b + 128; return (t + (t >> 8)) >> 8; }      // K:234-238
int pack565(const float c[3])                                                       // K:240-248
{
    int r = scale8(f2i(c[0]), 31), g = scale8(f2i(c[1]), 63), b = scale8(f2i(c[2]), 31);
    return (int)(uint16_t)((r << 11) + (g << 5) + b);
}
void unpack565(float c[3], int p)                                                   // K:250-259
{
    int b = p & 31, g = (p >> 5) & 63, r = (p >> 11) & 31;
    c[0] = (float)((r << 3) + (r >> 2));
    c[1] = (float)((g << 2) + (g >> 4));
    c[2] = (float)((b << 3) + (b >> 2));
}

// Mean and centred covariance, accumulated texel by texel in order k=0..15; K:377-417.
// Covariance uses the 6-slot order [rr rg rb gg gb bb] of K:162-167.
void bc1_mean_covariance(float cov[6], float mean[3], const float* px)
{
    for (int c = 0; c < 3; c++) {
        float acc = 0.0f;
        for (int k = 0; k < 16; k++) acc += px[16 * c + k];
        mean[c] = acc / 16.0f;
    }
    for (int i = 0; i < 6; i++) cov[i] = 0.0f;
    for (int k = 0; k < 16; k++) {
        float r = px[k] - mean[0], g = px[16 + k] - mean[1], b = px[32 + k] - mean[2];
        cov[0] += r * r; cov[1] += r * g; cov[2] += r * b;
        cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
    }
}

// Project on the axis, take the extreme texels as endpoints; K:274-306 (min starts at 65536,
// max at 0 -- quirk Q6).
void bc1_span_endpoints(float lo[3], float hi[3], const float* px, const float axis[3], const float mean[3])
{
    float dmin = 65536.0f, dmax = 0.0f;
    for (int k = 0; k < 16; k++) {
        float d = 0.0f;
        for (int c = 0; c < 3; c++) d += (px[16 * c + k] - mean[c]) * axis[c];
        dmin = fmin_sse(dmin, d);
        dmax = fmax_sse(dmax, d);
    }
    if (dmax - dmin < 1.0f) { dmin -= 0.5f; dmax += 0.5f; }
    float n2 = 0.0f;
    for (int c = 0; c < 3; c++) n2 += axis[c] * axis[c];
    float inv = rcp_exact(n2);
    for (int c = 0; c < 3; c++) {
        lo[c] = fclamp(mean[c] + dmin * inv * axis[c], 0.0f, 255.0f);
        hi[c] = fclamp(mean[c] + dmax * inv * axis[c], 0.0f, 255.0f);
    }
}

// Linear 2-bit indices 0..3 along p0 -> p1; K:308-344.  p0 == p1 divides by zero, the NaN
// casts to INT_MIN and clamps to 0 (quirk Q7).
u32 bc1_linear_indices(const float* px, int p0, int p1)
{
    float a[3], b[3], dir[3];
    unpack565(a, p0);
    unpack565(b, p1);
    for (int c = 0; c < 3; c++) dir[c] = b[c] - a[c];
    float n2 = 0.0f;
    for (int c = 0; c < 3; c++) n2 += sqf(dir[c]);
    float inv = rcp_exact(n2);
    for (int c = 0; c < 3; c++) dir[c] *= inv * 3.0f;
    float bias = 0.5f;
    for (int c = 0; c < 3; c++) bias -= a[c] * dir[c];
    u32 bits = 0, scale = 1;
    for (int k = 0; k < 16; k++) {
        float d = 0.0f;
        for (int c = 0; c < 3; c++) d += px[16 * c + k] * dir[c];
        int q = iclamp(f2i(d + bias), 0, 3);
        bits += (u32)q * scale;
        scale *= 4;
    }
    return bits;
}

// One least-squares update of both endpoints from the current indices; K:419-480
void bc1_least_squares(int pe[2], const float* px, u32 bits, const float mean[3])
{
    float a[3], b[3];
    if ((bits ^ (bits * 4u)) < 4u) {                 // all sixteen indices equal; K:424-432
        for (int c = 0; c < 3; c++) a[c] = b[c] = mean[c];
    } else {
        float atb1[3] = {0.0f, 0.0f, 0.0f};
        float sq1 = 0.0f, sqq = 0.0f;
        u32 rest = bits;
        for (int k = 0; k < 16; k++) {
            float q = (float)(int)(rest & 3u);
            rest >>= 2;
            float x = 3.0f - q;
            sq1 += q;
            sqq += q * q;
            for (int c = 0; c < 3; c++) atb1[c] += x * px[16 * c + k];
        }
        float total[3], atb2[3];
        for (int c = 0; c < 3; c++) {
            total[c] = mean[c] * 16.0f;
            atb2[c] = 3.0f * total[c] - atb1[c];
        }
        float cxx = 16.0f * sqf(3.0f) - 6.0f * sq1 + sqq;    // K:463
        float cyy = sqq;
        float cxy = 3.0f * sq1 - sqq;
        float scale = 3.0f * rcp_exact(cxx * cyy - cxy * cxy);
        for (int c = 0; c < 3; c++) {
            a[c] = (atb1[c] * cyy - atb2[c] * cxy) * scale;
            b[c] = (atb2[c] * cxx - atb1[c] * cxy) * scale;
            a[c] = fclamp(a[c], 0.0f, 255.0f);
            b[c] = fclamp(b[c], 0.0f, 255.0f);
        }
    }
    pe[0] = pack565(a);
    pe[1] = pack565(b);
}

// The colour half shared by BC1 and BC3; K:494-533
void bc1_colour_block(const float* px, u32 out[2])
{
    float cov[6], mean[3];
    bc1_mean_covariance(cov, mean, px);
    const float eps = 0.001f;
    cov[0] += eps; cov[3] += eps; cov[5] += eps;

    // 4 power iterations on the 3x3 matrix; K:184-205 uses the 6-slot packing
    float m10[10] = {cov[0], cov[1], cov[2], 0.0f, cov[3], cov[4], 0.0f, cov[5], 0.0f, 0.0f};
    float axis[4];
    power_axis(axis, m10, 4, 3);

    float lo[3], hi[3];
    bc1_span_endpoints(lo, hi, px, axis, mean);
    int p[2] = {pack565(lo), pack565(hi)};
    if (p[0] < p[1]) { int t = p[0]; p[0] = p[1]; p[1] = t; }
    out[0] = (u32)((1 << 16) * p[1] + p[0]);
    out[1] = bc1_linear_indices(px, p[0], p[1]);

    // exactly one refinement pass; K:497, :524-530
    bc1_least_squares(p, px, out[1], mean);
    if (p[0] < p[1]) { int t = p[0]; p[0] = p[1]; p[1] = t; }
    out[0] = (u32)((1 << 16) * p[1] + p[0]);
    out[1] = bc1_linear_indices(px, p[0], p[1]);

    // linear order 0,1,2,3 -> BC1 codes 0,2,3,1; K:482-492
    u32 lo_bits = out[1] & 0x55555555u, hi_bits = out[1] & 0xAAAAAAAAu;
    out[1] = (hi_bits >> 1) + (hi_bits ^ (lo_bits << 1));
}

// BC3 alpha half; K:535-571 (endpoints truncated after the index search -- quirk Q9)
void bc3_alpha_block(const float* a, u32 out[2])
{
    float lo = 255.0f, hi = 0.0f;
    for (int k = 0; k < 16; k++) { lo = fmin_sse(lo, a[k]); hi = fmax_sse(hi, a[k]); }
    if (lo == hi) hi = lo + 0.1f;
    u32 idx[2] = {0, 0};
    float scale = 7.0f / (hi - lo);
    for (int k = 0; k < 16; k++) {
        float proj = (a[k] - lo) * scale + 0.5f;
        int q = iclamp(f2i(proj), 0, 7);
        q = 7 - q;
        if (q > 0) q++;
        if (q == 8) q = 1;
        idx[k / 8] |= (u32)q << ((k % 8) * 3);
    }
    out[0] = (u32)(iclamp(f2i(lo), 0, 255) * 256 + iclamp(f2i(hi), 0, 255));
    out[0] |= idx[0] << 16;
    out[1] = idx[0] >> 16;
    out[1] |= idx[1] << 8;
}

// ==========================================================================================
// BC7 / BC6H shared numerics; K:760-971, :1133-1262
// ==========================================================================================
// Raw moments over the masked texels.  slots 0..9 second order [xx xy xz xw yy yz yw zz zw ww],
// 10..13 sums, 14 count.  Every texel is visited and multiplied by its 0/1 flag; K:763-803.
void masked_moments(float st[15], const float* px, int mask, int channels)
{
    for (int i = 0; i < 15; i++) st[i] = 0.0f;
    for (int k = 0; k < 16; k++) {
        float flag = (float)((mask >> k) & 1);
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int c = 0; c < channels; c++) v[c] = px[16 * c + k];
        for (int c = 0; c < channels; c++) v[c] *= flag;
        st[14] += flag;
        st[10] += v[0]; st[11] += v[1]; st[12] += v[2];
        st[0] += v[0] * v[0]; st[1] += v[0] * v[1]; st[2] += v[0] * v[2];
        st[4] += v[1] * v[1]; st[5] += v[1] * v[2];
        st[7] += v[2] * v[2];
        if (channels == 4) {
            st[13] += v[3];
            st[3] += v[0] * v[3]; st[6] += v[1] * v[3]; st[8] += v[2] * v[3]; st[9] += v[3] * v[3];
        }
    }
}
// cov = E[xy]*n - sum(x)*sum(y)/n; K:805-823.  Slots not owned by `channels` stay zero (F6).
void covariance_of(float cov[10], const float st[15], int channels)
{
    for (int i = 0; i < 10; i++) cov[i] = 0.0f;
    cov[0] = st[0] - st[10] * st[10] / st[14];
    cov[1] = st[1] - st[10] * st[11] / st[14];
    cov[2] = st[2] - st[10] * st[12] / st[14];
    cov[4] = st[4] - st[11] * st[11] / st[14];
    cov[5] = st[5] - st[11] * st[12] / st[14];
    cov[7] = st[7] - st[12] * st[12] / st[14];
    if (channels == 4) {
        cov[3] = st[3] - st[10] * st[13] / st[14];
        cov[6] = st[6] - st[11] * st[13] / st[14];
        cov[8] = st[8] - st[12] * st[13] / st[14];
        cov[9] = st[9] - st[13] * st[13] / st[14];
    }
}

// PCA line through the masked texels, endpoints at the extreme projections; K:834-894.
// `clamp255` distinguishes K:896 block_segment (BC7) from K:857 block_segment_core (BC6H).
void fit_segment(float* ep, const float* px, int mask, int channels, bool clamp255)
{
    float st[15], cov[10], mean[4] = {0.0f, 0.0f, 0.0f, 0.0f}, axis[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    masked_moments(st, px, mask, channels);
    covariance_of(cov, st, channels);
    for (int c = 0; c < channels; c++) mean[c] = st[10 + c] / st[14];

    const float inv_var = 1.0f / (256.0f * 256.0f);           // K:842
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = sqf(0.001f);                            // K:848
    cov[0] += eps; cov[4] += eps; cov[7] += eps; cov[9] += eps;
    power_axis(axis, cov, 8, channels);

    float lo = kInf, hi = -kInf;                              // K:864-865 (1e99 -> inf)
    for (int k = 0; k < 16; k++) {
        if (((mask >> k) & 1) == 0) continue;
        float d = 0.0f;
        for (int c = 0; c < channels; c++) d += axis[c] * (px[16 * c + k] - mean[c]);
        lo = fmin_sse(lo, d);
        hi = fmax_sse(hi, d);
    }
    if (hi - lo < 1.0f) { lo -= 0.5f; hi += 0.5f; }
    for (int c = 0; c < channels; c++) {
        ep[c] = lo * axis[c] + mean[c];
        ep[4 + c] = hi * axis[c] + mean[c];
    }
    if (clamp255)
        for (int i = 0; i < 2; i++)
            for (int c = 0; c < channels; c++) ep[4 * i + c] = fclamp(ep[4 * i + c], 0.0f, 255.0f);
}

// trace - lambda_max estimate of a (destructively rescaled) covariance; K:907-939
float residual_bound(float cov[10], int channels)
{
    const float inv_var = 1.0f / (256.0f * 256.0f);
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = sqf(0.001f);
    cov[0] += eps; cov[4] += eps; cov[7] += eps;              // three diagonal slots only (K:918-920)
    float axis[4] = {0.0f, 0.0f, 0.0f, 0.0f}, mv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    power_axis(axis, cov, 4, channels);
    sym_apply(mv, cov, axis, channels);
    float s = 0.0f;
    for (int c = 0; c < channels; c++) s += sqf(mv[c]);
    float lambda = sqrtf(s);
    float bound = cov[0] + cov[4] + cov[7];
    if (channels == 4) bound += cov[9];
    bound -= lambda;
    return fmax_sse(bound, 0.0f);
}
// Bound for a two-way split: subset 0 from its own moments, subset 1 as full - subset 0; K:952-971
float split_bound(const float* px, int mask, const float full[15], int channels)
{
    float st[15], c1[10], c2[10];
    masked_moments(st, px, mask, channels);
    covariance_of(c1, st, channels);
    for (int i = 0; i < 15; i++) st[i] = full[i] - st[i];
    covariance_of(c2, st, channels);
    float b = 0.0f;
    b += residual_bound(c1, channels);
    b += residual_bound(c2, channels);
    return sqrtf(b) * 256.0f;
}

// Ascending selection sort of the first `take` keys; K:1365-1384
void select_smallest(int* keys, int n, int take)
{
    for (int k = 0; k < take; k++) {
        int best = k, bestv = keys[k];
        for (int i = k + 1; i < n; i++)
            if (bestv > keys[i]) { bestv = keys[i]; best = i; }
        keys[best] = keys[k];
        keys[k] = bestv;
    }
}

// Index search: project on the subset's segment, then compare the two neighbouring palette
// entries (decoded with the integer BC7 interpolation); K:1133-1193.  `pattern` = 2 bits/texel.
// The per-texel error is truncated through int (cvttss2si: quirk Q3 for BC6H) before summing.
float assign_indices(u32 idx[2], const float* px, int bits, const float* ep, u32 pattern, int channels)
{
    const int* w = weights_for(bits);
    const int levels = 1 << bits;
    float total = 0.0f;
    idx[0] = idx[1] = 0;
    for (int k = 0; k < 16; k++) {
        const float* e = ep + 8 * ((pattern >> (2 * k)) & 3);
        float proj = 0.0f, div = 0.0f;
        for (int c = 0; c < channels; c++) {
            proj += (px[16 * c + k] - e[c]) * (e[4 + c] - e[c]);
            div += sqf(e[4 + c] - e[c]);
        }
        proj /= div;
        int q1 = f2i(proj * (float)levels + 0.5f);
        q1 = iclamp(q1, 1, levels - 1);
        int w0 = w[q1 - 1], w1 = w[q1];
        float err0 = 0.0f, err1 = 0.0f;
        for (int c = 0; c < channels; c++) {
            float d0 = (float)f2i(((float)(64 - w0) * e[c] + (float)w0 * e[4 + c] + 32.0f) / 64.0f);
            float d1 = (float)f2i(((float)(64 - w1) * e[c] + (float)w1 * e[4 + c] + 32.0f) / 64.0f);
            err0 += sqf(d0 - px[16 * c + k]);
            err1 += sqf(d1 - px[16 * c + k]);
        }
        int best_err = f2i(err1), best_q = q1;
        if (err0 < err1) { best_err = f2i(err0); best_q = q1 - 1; }
        idx[k / 8] += (u32)best_q << (4 * (k % 8));
        total += (float)best_err;
    }
    return total;
}

// Least-squares endpoints of one subset from its current indices; K:1198-1262
void solve_endpoints(float* ep, const float* px, int bits, const u32 idx[2], int mask, int channels)
{
    const int levels = 1 << bits;
    float atb1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sum[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    float sq1 = 0.0f, sqq = 0.0f;
    for (int k = 0; k < 16; k++) {
        float q = (float)(int)((idx[k / 8] >> (4 * (k % 8))) & 15u);
        if (((mask >> k) & 1) == 0) continue;
        int x = f2i((float)(levels - 1) - q);
        sq1 += q;
        sqq += q * q;
        sum[4] += 1.0f;
        for (int c = 0; c < channels; c++) sum[c] += px[16 * c + k];
        for (int c = 0; c < channels; c++) atb1[c] += (float)x * px[16 * c + k];
    }
    float atb2[4];
    for (int c = 0; c < channels; c++) atb2[c] = (float)(levels - 1) * sum[c] - atb1[c];
    float cxx = sum[4] * sqf((float)(levels - 1)) - (float)(2 * (levels - 1)) * sq1 + sqq;
    float cyy = sqq;
    float cxy = (float)(levels - 1) * sq1 - sqq;
    float scale = (float)(levels - 1) / (cxx * cyy - cxy * cxy);
    for (int c = 0; c < channels; c++) {
        ep[c] = (atb1[c] * cyy - atb2[c] * cxy) * scale;
        ep[4 + c] = (atb2[c] * cxx - atb1[c] * cxy) * scale;
    }
    if (fabsf(cxx * cyy - cxy * cxy) < 0.001f) {             // degenerate: flatten to the mean
        for (int c = 0; c < channels; c++) {
            ep[c] = sum[c] / sum[4];
            ep[4 + c] = ep[c];
        }
    }
}

// ==========================================================================================
// BC7 endpoint quantisation; K:976-1128
// ==========================================================================================
int expand_bits(int v, int bits)                               // K:976-981
{
    int vv = v << (8 - bits);
    return vv + (int)((u32)vv >> bits);
}
inline int bc7_pairs(int mode) { const int t[8] = {3, 2, 3, 2, 1, 1, 1, 2}; return t[mode]; }

// modes 0,3,6,7: one p-bit per endpoint, chosen by the smaller squared error over the first
// `channels` components; K:983-1022.  All four components are produced either way.
void quant_unique_pbit(int* q, const float* ep, int mode, int channels)
{
    int bits = 7;
    if (mode == 0) bits = 4;
    if (mode == 7) bits = 5;
    const int levels = 1 << bits, levels2 = levels * 2 - 1;
    for (int i = 0; i < 2; i++) {
        int cand[8];
        for (int b = 0; b < 2; b++)
            for (int c = 0; c < 4; c++) {
                int v = f2i((ep[4 * i + c] / 255.0f * (float)levels2 - (float)b) / 2.0f + 0.5f) * 2 + b;
                cand[4 * b + c] = iclamp(v, b, levels2 - 1 + b);
            }
        float deq[8];
        for (int j = 0; j < 8; j++) deq[j] = (float)cand[j];
        if (mode == 0)
            for (int j = 0; j < 8; j++) deq[j] = (float)expand_bits(cand[j], 5);
        float e0 = 0.0f, e1 = 0.0f;
        for (int c = 0; c < channels; c++) {
            e0 += sqf(ep[4 * i + c] - deq[c]);
            e1 += sqf(ep[4 * i + c] - deq[4 + c]);
        }
        for (int c = 0; c < 4; c++) q[4 * i + c] = (e0 < e1) ? cand[c] : cand[4 + c];
    }
}
// mode 1: one p-bit shared by both endpoints of a subset, error over RGB; K:1024-1052
void quant_shared_pbit(int* q, const float* ep)
{
    int cand[16];
    for (int b = 0; b < 2; b++)
        for (int i = 0; i < 8; i++) {
            int v = f2i((ep[i] / 255.0f * 127.0f - (float)b) / 2.0f + 0.5f) * 2 + b;
            cand[8 * b + i] = iclamp(v, b, 126 + b);
        }
    float deq[16];
    for (int k = 0; k < 16; k++) deq[k] = (float)expand_bits(cand[k], 7);
    float e0 = 0.0f, e1 = 0.0f;
    for (int j = 0; j < 2; j++)
        for (int c = 0; c < 3; c++) {
            e0 += sqf(ep[4 * j + c] - deq[4 * j + c]);
            e1 += sqf(ep[4 * j + c] - deq[8 + 4 * j + c]);
        }
    for (int i = 0; i < 8; i++) q[i] = (e0 < e1) ? cand[i] : cand[8 + i];
}
// modes 2,4,5: no p-bit; K:1054-1065
void quant_plain(int* q, const float* ep, int mode)
{
    const int levels = 1 << ((mode == 5) ? 7 : 5);
    for (int i = 0; i < 8; i++) {
        int v = f2i(ep[i] / 255.0f * (float)(levels - 1) + 0.5f);
        q[i] = iclamp(v, 0, levels - 1);
    }
}
// Quantise every endpoint pair of a mode, then replace ep by the decoded values; K:1067-1128
void bc7_quantise(int* q, float* ep, int mode, int channels)
{
    const int pairs = bc7_pairs(mode);
    for (int j = 0; j < pairs; j++) {
        if (mode == 0 || mode == 3 || mode == 6 || mode == 7) quant_unique_pbit(q + 8 * j, ep + 8 * j, mode, channels);
        else if (mode == 1) quant_shared_pbit(q + 8 * j, ep + 8 * j);
        else quant_plain(q + 8 * j, ep + 8 * j, mode);
    }
    for (int i = 0; i < 8 * pairs; i++) {
        if (mode == 3 || mode == 6) ep[i] = (float)q[i];
        else if (mode == 1 || mode == 5) ep[i] = (float)expand_bits(q[i], 7);
        else if (mode == 7) ep[i] = (float)expand_bits(q[i], 6);
        else ep[i] = (float)expand_bits(q[i], 5);              // modes 0, 2, 4
    }
}

// ==========================================================================================
// bit writer shared by BC7 and BC6H
// ==========================================================================================
struct BitSink {
    u32 w[4];
    int pos;
    BitSink() : pos(0) { w[0] = w[1] = w[2] = w[3] = 0; }
    void put(int nbits, u32 v)                                 // LSB first; K:1735-1744
    {
        if (nbits == 0) return;
        if (nbits < 32) v &= (1u << nbits) - 1u;
        int word = pos >> 5, off = pos & 31;
        if (word < 4) w[word] |= v << off;
        if (off + nbits > 32 && word + 1 < 4) w[word + 1] |= v >> (32 - off);
        pos += nbits;
    }
};

// Index payload.  Texel 0 and the anchors of subsets 1,2 drop their (zero) MSB.  Equivalent to
// K:1767-1805, which writes full-width indices then deletes the anchor MSBs by shifting.
void put_indices(BitSink& s, const u32 idx[2], int bits, int flips, const int* anchors, int nanchors)
{
    const int levels = 1 << bits;
    for (int k = 0; k < 16; k++) {
        int q = (int)((idx[k / 8] >> (4 * (k % 8))) & 15u);
        if ((flips >> k) & 1) q = (levels - 1) - q;
        bool is_anchor = (k == 0);
        for (int a = 0; a < nanchors; a++) is_anchor = is_anchor || (anchors[a] == k);
        s.put(is_anchor ? bits - 1 : bits, (u32)q);
    }
}

// Make each subset's anchor index < levels/2 by swapping that subset's endpoints and mirroring
// its indices; returns the 16-bit texel mask of mirrored texels; K:1708-1733
int orient_subsets(int* q, const u32 idx[2], int bits, int pairs, int shape)
{
    int anchors[3];
    shape_anchors(shape, anchors);
    const int levels = 1 << bits;
    int flips = 0;
    for (int j = 0; j < pairs; j++) {
        int k0 = anchors[j];
        int v = (int)((idx[k0 >> 3] >> (4 * (k0 & 7))) & 15u);
        if (v >= levels / 2) {
            for (int c = 0; c < 4; c++) { int t = q[8 * j + c]; q[8 * j + c] = q[8 * j + 4 + c]; q[8 * j + 4 + c] = t; }
            flips |= shape_mask(shape, j);
        }
    }
    return flips;
}
// Single-subset variant used by modes 4,5,6 and BC6H one-region modes; K:1694-1706
void orient_single(int* q, int width, u32 idx[2], int bits)
{
    const int levels = 1 << bits;
    if ((int)(idx[0] & 15u) >= levels / 2) {
        for (int c = 0; c < width; c++) { int t = q[c]; q[c] = q[width + c]; q[width + c] = t; }
        for (int k = 0; k < 2; k++) idx[k] = (u32)(0x11111111u * (u32)(levels - 1)) - idx[k];
    }
}

// ==========================================================================================
// BC7 encoder; K:1264-2037
// ==========================================================================================
struct Bc7Block {
    float px[64];
    const bc7_enc_settings* cfg;
    float opaque_err;
    float best_err;
    u32 best[4];
};

void bc7_write_partitioned(u32 out[4], int* q, const u32 idx[2], int shape, int mode)   // K:1807-1877
{
    const int bits = (mode == 0 || mode == 1) ? 3 : 2;
    const int pairs = (mode == 0 || mode == 2) ? 3 : 2;
    const int channels = (mode == 7) ? 4 : 3;
    int flips = orient_subsets(q, idx, bits, pairs, shape);
    BitSink s;
    s.put(mode + 1, 1u << mode);
    s.put(mode == 0 ? 4 : 6, (u32)(shape & (mode == 0 ? 15 : 63)));
    for (int c = 0; c < channels; c++)
        for (int j = 0; j < pairs * 2; j++) {
            int v = q[4 * j + c];
            switch (mode) {
                case 0: s.put(4, (u32)(v >> 1)); break;
                case 1: s.put(6, (u32)(v >> 1)); break;
                case 2: s.put(5, (u32)v); break;
                case 3: s.put(7, (u32)(v >> 1)); break;
                default: s.put(5, (u32)(v >> 1)); break;      // mode 7
            }
        }
    if (mode == 1)
        for (int j = 0; j < 2; j++) s.put(1, (u32)(q[8 * j] & 1));
    if (mode == 0 || mode == 3 || mode == 7)
        for (int j = 0; j < pairs * 2; j++) s.put(1, (u32)(q[4 * j] & 1));
    int anchors[3];
    shape_anchors(shape, anchors);
    put_indices(s, idx, bits, flips, anchors + 1, pairs - 1);
    memcpy(out, s.w, 16);
}

// One partitioned mode over a candidate list, refinement on the winner, commit on strict
// improvement; K:1279-1363
void bc7_try_partitioned(Bc7Block& blk, int mode, const int* keys, int count)
{
    if (count == 0) return;
    const int bits = (mode == 0 || mode == 1) ? 3 : 2;
    const int pairs = (mode == 0 || mode == 2) ? 3 : 2;
    const int channels = (mode == 7) ? 4 : 3;

    int best_q[24] = {0};
    u32 best_idx[2] = {0, 0};
    int best_shape = -1;
    float best_err = kInf;
    for (int n = 0; n < count; n++) {
        int shape = keys[n] & 63;
        if (pairs == 3) shape += 64;
        float ep[24] = {0};
        int q[24] = {0};
        u32 idx[2];
        for (int j = 0; j < pairs; j++) fit_segment(ep + 8 * j, blk.px, shape_mask(shape, j), channels, true);
        bc7_quantise(q, ep, mode, channels);
        float err = assign_indices(idx, blk.px, bits, ep, shape_pattern(shape), channels);
        if (err < best_err) {
            memcpy(best_q, q, sizeof(int) * 8 * pairs);
            best_idx[0] = idx[0]; best_idx[1] = idx[1];
            best_shape = shape;
            best_err = err;
        }
    }
    for (int it = 0; it < blk.cfg->refineIterations[mode]; it++) {
        float ep[24] = {0};                                   // never-written slots read as zero (F6, quirk Q1)
        int q[24] = {0};
        u32 idx[2];
        for (int j = 0; j < pairs; j++) solve_endpoints(ep + 8 * j, blk.px, bits, best_idx, shape_mask(best_shape, j), channels);
        bc7_quantise(q, ep, mode, blk.cfg->channels);         // the PROFILE's channel count (K:1343)
        float err = assign_indices(idx, blk.px, bits, ep, shape_pattern(best_shape), channels);
        if (err < best_err) {
            memcpy(best_q, q, sizeof(int) * 8 * pairs);
            best_idx[0] = idx[0]; best_idx[1] = idx[1];
            best_err = err;
        }
    }
    if (mode != 7) best_err += blk.opaque_err;
    if (best_err < blk.best_err) {
        blk.best_err = best_err;
        bc7_write_partitioned(blk.best, best_q, best_idx, best_shape, mode);
    }
}

// Rank the 64 two-subset shapes by their split bound; key = shape + 64*(int)bound; K:1400-1410
void rank_two_subset_shapes(int keys[64], const float* px, int channels, int nshapes)
{
    float full[15];
    masked_moments(full, px, -1, channels);
    for (int s = 0; s < nshapes; s++) {
        float b = split_bound(px, shape_mask(s, 0), full, channels);
        keys[s] = s + f2i(b) * 64;
    }
}

// ---- modes 4/5: separate scalar channel; K:1437-1655 ----
void scalar_quantise(int q[2], float ep[2], int epbits)                              // K:1437-1447
{
    const int levels = 1 << epbits;
    for (int i = 0; i < 2; i++) {
        int v = f2i(ep[i] / 255.0f * (float)(levels - 1) + 0.5f);
        q[i] = iclamp(v, 0, levels - 1);
        ep[i] = (float)expand_bits(q[i], epbits);
    }
}
float scalar_assign(u32 idx[2], const float* a, int bits, const float ep[2])          // K:1498-1538
{
    const int* w = weights_for(bits);
    const int levels = 1 << bits;
    idx[0] = idx[1] = 0;
    float total = 0.0f;
    for (int k = 0; k < 16; k++) {
        float proj = (a[k] - ep[0]) / (ep[1] - ep[0] + 0.001f);
        int q1 = iclamp(f2i(proj * (float)levels + 0.5f), 1, levels - 1);
        int w0 = w[q1 - 1], w1 = w[q1];
        float d0 = (float)f2i(((float)(64 - w0) * ep[0] + (float)w0 * ep[1] + 32.0f) / 64.0f);
        float d1 = (float)f2i(((float)(64 - w1) * ep[0] + (float)w1 * ep[1] + 32.0f) / 64.0f);
        float err0 = 0.0f, err1 = 0.0f;
        err0 += sqf(d0 - a[k]);
        err1 += sqf(d1 - a[k]);
        int best_err = f2i(err1), best_q = q1;
        if (err0 < err1) { best_err = f2i(err0); best_q = q1 - 1; }
        idx[k / 8] += (u32)best_q << (4 * (k % 8));
        total += (float)best_err;
    }
    return total;
}
void scalar_solve(float ep[2], const float* a, int bits, const u32 idx[2])             // K:1449-1496
{
    const int levels = 1 << bits;
    float atb1 = 0.0f, sq1 = 0.0f, sqq = 0.0f, sum = 0.0f;
    for (int k = 0; k < 16; k++) {
        float q = (float)(int)((idx[k / 8] >> (4 * (k % 8))) & 15u);
        int x = f2i((float)(levels - 1) - q);
        sq1 += q;
        sqq += q * q;
        sum += a[k];
        atb1 += (float)x * a[k];
    }
    float atb2 = (float)(levels - 1) * sum - atb1;
    float cxx = 16.0f * sqf((float)(levels - 1)) - (float)(2 * (levels - 1)) * sq1 + sqq;
    float cyy = sqq;
    float cxy = (float)(levels - 1) * sq1 - sqq;
    float scale = (float)(levels - 1) / (cxx * cyy - cxy * cxy);
    ep[0] = (atb1 * cyy - atb2 * cxy) * scale;
    ep[1] = (atb2 * cxx - atb1 * cxy) * scale;
    ep[0] = fclamp(ep[0], 0.0f, 255.0f);
    ep[1] = fclamp(ep[1], 0.0f, 255.0f);
    if (fabsf(cxx * cyy - cxy * cxy) < 0.001f) {
        ep[0] = sum / 16.0f;
        ep[1] = ep[0];
    }
}
float scalar_channel(const Bc7Block& blk, u32 idx[2], int q[2], const float* a, int bits, int epbits)   // K:1540-1563
{
    float ep[2] = {255.0f, 0.0f};
    for (int k = 0; k < 16; k++) { ep[0] = fmin_sse(ep[0], a[k]); ep[1] = fmax_sse(ep[1], a[k]); }
    scalar_quantise(q, ep, epbits);
    float err = scalar_assign(idx, a, bits, ep);
    for (int it = 0; it < blk.cfg->refineIterations_channel; it++) {
        scalar_solve(ep, a, bits, idx);
        scalar_quantise(q, ep, epbits);
        err = scalar_assign(idx, a, bits, ep);
    }
    return err;
}

struct Mode45Pick { int q[8]; u32 idx[2]; int aq[2]; u32 aidx[2]; int rotation, swap; };

void bc7_mode45_candidate(const Bc7Block& blk, Mode45Pick& pick, float& pick_err, int mode, int rotation, int swap)   // K:1565-1621
{
    int bits = 2, abits = (mode == 4) ? 3 : 2;
    const int aepbits = (mode == 4) ? 6 : 8;
    if (swap == 1) { bits = 3; abits = 2; }
    float px[48];
    for (int k = 0; k < 16; k++) {
        for (int c = 0; c < 3; c++) px[16 * c + k] = blk.px[16 * c + k];
        if (rotation < 3) {
            if (blk.cfg->channels == 4) px[16 * rotation + k] = blk.px[48 + k];
            if (blk.cfg->channels == 3) px[16 * rotation + k] = 255.0f;
        }
    }
    float ep[8] = {0};
    int q[8] = {0};
    u32 idx[2];
    fit_segment(ep, px, -1, 3, true);
    bc7_quantise(q, ep, mode, 3);
    float err = assign_indices(idx, px, bits, ep, 0, 3);
    for (int it = 0; it < blk.cfg->refineIterations[mode]; it++) {
        solve_endpoints(ep, px, bits, idx, -1, 3);
        bc7_quantise(q, ep, mode, 3);
        err = assign_indices(idx, px, bits, ep, 0, 3);
    }
    int aq[2];
    u32 aidx[2];
    err += scalar_channel(blk, aidx, aq, blk.px + 16 * rotation, abits, aepbits);
    if (err < pick_err) {
        memcpy(pick.q, q, sizeof(q));
        pick.idx[0] = idx[0]; pick.idx[1] = idx[1];
        pick.aq[0] = aq[0]; pick.aq[1] = aq[1];
        pick.aidx[0] = aidx[0]; pick.aidx[1] = aidx[1];
        pick.rotation = rotation;
        pick.swap = swap;
        pick_err = err;
    }
}
void bc7_write_mode45(u32 out[4], const Mode45Pick& pick, int mode)                   // K:1879-1939
{
    int q[8], aq[2];
    u32 idx[2], aidx[2];
    memcpy(q, pick.q, sizeof(q));
    memcpy(aq, pick.aq, sizeof(aq));
    memcpy(idx, pick.idx, sizeof(idx));
    memcpy(aidx, pick.aidx, sizeof(aidx));
    const int bits = 2, abits = (mode == 4) ? 3 : 2;
    const int epbits = (mode == 4) ? 5 : 7, aepbits = (mode == 4) ? 6 : 8;
    if (!pick.swap) {
        orient_single(q, 4, idx, bits);
        orient_single(aq, 1, aidx, abits);
    } else {                                                  // index sets trade places; K:1903-1908
        u32 t0 = idx[0], t1 = idx[1];
        idx[0] = aidx[0]; idx[1] = aidx[1];
        aidx[0] = t0; aidx[1] = t1;
        orient_single(aq, 1, idx, bits);
        orient_single(q, 4, aidx, abits);
    }
    BitSink s;
    s.put(mode + 1, 1u << mode);
    s.put(2, (u32)((pick.rotation + 1) & 3));
    if (mode == 4) s.put(1, (u32)pick.swap);
    for (int c = 0; c < 3; c++) { s.put(epbits, (u32)q[c]); s.put(epbits, (u32)q[4 + c]); }
    s.put(aepbits, (u32)aq[0]);
    s.put(aepbits, (u32)aq[1]);
    put_indices(s, idx, bits, 0, 0, 0);
    put_indices(s, aidx, abits, 0, 0, 0);
    memcpy(out, s.w, 16);
}
void bc7_try_mode45(Bc7Block& blk)                                                     // K:1623-1655
{
    Mode45Pick pick;
    memset(&pick, 0, sizeof(pick));
    float pick_err = blk.best_err;
    const int first = blk.cfg->mode45_channel0, last = blk.cfg->channels;
    for (int r = first; r < last; r++) {
        bc7_mode45_candidate(blk, pick, pick_err, 4, r, 0);
        bc7_mode45_candidate(blk, pick, pick_err, 4, r, 1);
    }
    if (pick_err < blk.best_err) { blk.best_err = pick_err; bc7_write_mode45(blk.best, pick, 4); }
    for (int r = first; r < last; r++) bc7_mode45_candidate(blk, pick, pick_err, 5, r, 0);
    if (pick_err < blk.best_err) { blk.best_err = pick_err; bc7_write_mode45(blk.best, pick, 5); }
}

void bc7_try_mode6(Bc7Block& blk)                                                      // K:1657-1689, :1941-1964
{
    const int channels = blk.cfg->channels;
    float ep[8] = {0};
    int q[8] = {0};
    u32 idx[2];
    fit_segment(ep, blk.px, -1, channels, true);
    if (channels == 3) ep[3] = ep[7] = 255.0f;
    bc7_quantise(q, ep, 6, channels);
    float err = assign_indices(idx, blk.px, 4, ep, 0, channels);
    for (int it = 0; it < blk.cfg->refineIterations[6]; it++) {
        solve_endpoints(ep, blk.px, 4, idx, -1, channels);
        bc7_quantise(q, ep, 6, channels);
        err = assign_indices(idx, blk.px, 4, ep, 0, channels);
    }
    if (err < blk.best_err) {
        blk.best_err = err;
        orient_single(q, 4, idx, 4);
        BitSink s;
        s.put(7, 64u);
        for (int c = 0; c < 4; c++) { s.put(7, (u32)(q[c] >> 1)); s.put(7, (u32)(q[4 + c] >> 1)); }
        s.put(1, (u32)(q[0] & 1));
        s.put(1, (u32)(q[4] & 1));
        put_indices(s, idx, 4, 0, 0, 0);
        memcpy(blk.best, s.w, 16);
    }
}

void bc7_encode_block(Bc7Block& blk)                                                   // K:1970-1977, :2014-2028
{
    const bc7_enc_settings* cfg = blk.cfg;
    blk.best_err = kInf;
    memset(blk.best, 0, sizeof(blk.best));
    blk.opaque_err = 0.0f;                                                             // K:1267-1277
    if (cfg->channels != 3)
        for (int k = 0; k < 16; k++) blk.opaque_err += sqf(blk.px[48 + k] - 255.0f);

    if (cfg->mode_selection[0]) {                                                      // K:1386-1394
        int keys[64];
        for (int i = 0; i < 64; i++) keys[i] = i;
        bc7_try_partitioned(blk, 0, keys, 16);
        if (!cfg->skip_mode2) bc7_try_partitioned(blk, 2, keys, 64);
    }
    if (cfg->mode_selection[1]) {
        if (!(cfg->fastSkipTreshold_mode1 == 0 && cfg->fastSkipTreshold_mode3 == 0)) {  // K:1396-1415
            int keys[64];
            rank_two_subset_shapes(keys, blk.px, 3, 64);
            select_smallest(keys, 64, imax(cfg->fastSkipTreshold_mode1, cfg->fastSkipTreshold_mode3));
            bc7_try_partitioned(blk, 1, keys, cfg->fastSkipTreshold_mode1);
            bc7_try_partitioned(blk, 3, keys, cfg->fastSkipTreshold_mode3);
        }
        if (cfg->fastSkipTreshold_mode7 != 0) {                                        // K:1417-1435
            int keys[64];
            rank_two_subset_shapes(keys, blk.px, cfg->channels, 64);
            select_smallest(keys, 64, cfg->fastSkipTreshold_mode7);
            bc7_try_partitioned(blk, 7, keys, cfg->fastSkipTreshold_mode7);
        }
    }
    if (cfg->mode_selection[2]) bc7_try_mode45(blk);
    if (cfg->mode_selection[3]) bc7_try_mode6(blk);
}

// ==========================================================================================
// BC6H (UF16) encoder; K:2039-3139
// ==========================================================================================
// mode order follows K: 0..9 two-region (2,3,4 / 6,7,8 are the per-channel variants of K's
// "mode 2" / "mode 6"), 10..13 one-region.
const int kBc6Prefix[14] = {0, 1, 2, 6, 10, 14, 18, 22, 26, 30, 3, 7, 11, 15};         // K:2080-2088
const int kBc6Epb[14]    = {10, 7, 11, -1, -1, 9, 8, -1, -1, 6, 10, 11, 12, 16};      // K:2113-2125
// K:2090-2111: the float table is read back through an int, i.e. truncated (quirk Q4)
int bc6_span(int mode)
{
    const float f65535 = 65535.0f;
    float v;
    switch (mode) {
        case 0: v = 0.9f * f65535 / 64.0f; break;
        case 1: v = 0.9f * f65535 / 4.0f; break;
        case 2: v = 0.8f * f65535 / 256.0f; break;
        case 5: v = 0.9f * f65535 / 32.0f; break;
        case 6: v = 0.9f * f65535 / 16.0f; break;
        case 9: case 10: v = f65535; break;
        case 11: v = 0.95f * f65535 / 8.0f; break;
        case 12: v = 0.95f * f65535 / 32.0f; break;
        case 13: v = 6.0f; break;
        default: v = -1.0f; break;
    }
    return f2i(v);
}

/* Header layouts, LSB first after the 5 prefix bits' position 0 -- i.e. the full 82 (two-region)
 * or 65 (one-region) header bits of the D3D BC6H definition, restated from K's layout comments
 * (K:2412-2422, :2464-2474, :2552-2559, :2577-2584, :2602-2609, :2648-2657, :2740-2749,
 * :2764-2773, :2788-2797, :2817-2826, :2881-2886, :2918-2923, :2955-2960).
 * Token "xN.b" = bit b of component x (r,g,b) of endpoint N; "xN.a-b" = bits a..b in that order
 * (ascending or descending); "m.b" = bit b of the 5-bit prefix. */
const char* const kBc6Layout[14] = {
    /* 0*/ "m.0-1 g2.4 b2.4 b3.4 r0.0-9 g0.0-9 b0.0-9 r1.0-4 g3.4 g2.0-3 g1.0-4 b3.0 g3.0-3 b1.0-4 b3.1 b2.0-3 r2.0-4 b3.2 r3.0-4 b3.3",
    /* 1*/ "m.0-1 g2.5 g3.4 g3.5 r0.0-6 b3.0 b3.1 b2.4 g0.0-6 b2.5 b3.2 g2.4 b0.0-6 b3.3 b3.5 b3.4 r1.0-5 g2.0-3 g1.0-5 g3.0-3 b1.0-5 b2.0-3 r2.0-5 r3.0-5",
    /* 2*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-4 r0.10 g2.0-3 g1.0-3 g0.10 b3.0 g3.0-3 b1.0-3 b0.10 b3.1 b2.0-3 r2.0-4 b3.2 r3.0-4 b3.3",
    /* 3*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-3 r0.10 g3.4 g2.0-3 g1.0-4 g0.10 g3.0-3 b1.0-3 b0.10 b3.1 b2.0-3 r2.0-3 b3.0 b3.2 r3.0-3 g2.4 b3.3",
    /* 4*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-3 r0.10 b2.4 g2.0-3 g1.0-3 g0.10 b3.0 g3.0-3 b1.0-4 b0.10 b2.0-3 r2.0-3 b3.1 b3.2 r3.0-3 b3.4 b3.3",
    /* 5*/ "m.0-4 r0.0-8 b2.4 g0.0-8 g2.4 b0.0-8 b3.4 r1.0-4 g3.4 g2.0-3 g1.0-4 b3.0 g3.0-3 b1.0-4 b3.1 b2.0-3 r2.0-4 b3.2 r3.0-4 b3.3",
    /* 6*/ "m.0-4 r0.0-7 g3.4 b2.4 g0.0-7 b3.2 g2.4 b0.0-7 b3.3 b3.4 r1.0-5 g2.0-3 g1.0-4 b3.0 g3.0-3 b1.0-4 b3.1 b2.0-3 r2.0-5 r3.0-5",
    /* 7*/ "m.0-4 r0.0-7 b3.0 b2.4 g0.0-7 g2.5 g2.4 b0.0-7 g3.5 b3.4 r1.0-4 g3.4 g2.0-3 g1.0-5 g3.0-3 b1.0-4 b3.1 b2.0-3 r2.0-4 b3.2 r3.0-4 b3.3",
    /* 8*/ "m.0-4 r0.0-7 b3.1 b2.4 g0.0-7 b2.5 g2.4 b0.0-7 b3.5 b3.4 r1.0-4 g3.4 g2.0-3 g1.0-4 b3.0 g3.0-3 b1.0-5 b2.0-3 r2.0-4 b3.2 r3.0-4 b3.3",
    /* 9*/ "m.0-4 r0.0-5 g3.4 b3.0 b3.1 b2.4 g0.0-5 g2.5 b2.5 b3.2 g2.4 b0.0-5 g3.5 b3.3 b3.5 b3.4 r1.0-5 g2.0-3 g1.0-5 g3.0-3 b1.0-5 b2.0-3 r2.0-5 r3.0-5",
    /*10*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-9 g1.0-9 b1.0-9",
    /*11*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-8 r0.10 g1.0-8 g0.10 b1.0-8 b0.10",
    /*12*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-7 r0.11-10 g1.0-7 g0.11-10 b1.0-7 b0.11-10",
    /*13*/ "m.0-4 r0.0-9 g0.0-9 b0.0-9 r1.0-3 r0.15-10 g1.0-3 g0.15-10 b1.0-3 b0.15-10",
};

// Emit the header of `mode`.  Endpoints 1..3 are stored as differences from endpoint 0 (wrapped
// to the field width by the bit extraction itself) except in the absolute modes 9 and 10;
// K:2392-2980.
void bc6_put_header(BitSink& s, const int* q, int mode)
{
    const bool delta = !(mode == 9 || mode == 10);
    const char* p = kBc6Layout[mode];
    while (*p) {
        while (*p == ' ') p++;
        if (!*p) break;
        char comp = *p++;
        int value;
        if (comp == 'm') {
            value = kBc6Prefix[mode];
        } else {
            int e = *p++ - '0';
            int c = (comp == 'r') ? 0 : (comp == 'g' ? 1 : 2);
            value = q[4 * e + c];
            if (delta && e > 0) value -= q[c];
        }
        p++;                                                   // '.'
        int a = 0, b;
        while (*p >= '0' && *p <= '9') a = a * 10 + (*p++ - '0');
        b = a;
        if (*p == '-') { p++; b = 0; while (*p >= '0' && *p <= '9') b = b * 10 + (*p++ - '0'); }
        int step = (b >= a) ? 1 : -1;
        for (int bit = a;; bit += step) {
            s.put(1, (u32)((value >> bit) & 1));
            if (bit == b) break;
        }
    }
}

struct Bc6Block {
    float px[64];
    const bc6h_enc_settings* cfg;
    float best_err;
    u32 best[4];
    float lo[3], hi[3];
    float max_span;
    int max_span_idx;
    int mode, epb;
    int qbounds[8];
};

int bc6_dequant(int v, int bits)                                                       // K:2130-2137
{
    if (bits >= 15) return v;
    if (v == 0) return 0;
    if (v == (1 << bits) - 1) return 0xFFFF;
    return (int)(((u32)v * 2u + 1u) << (15 - bits));
}
void bc6_quant(int* q, const float* ep, int bits, int pairs)                           // K:2139-2148
{
    const int levels = 1 << bits;
    for (int i = 0; i < 8 * pairs; i++) {
        int v = f2i(ep[i] / (256.0f * 256.0f - 1.0f) * (float)(levels - 1) + 0.5f);
        q[i] = iclamp(v, 0, levels - 1);
    }
}
void bc6_quant_dequant(const Bc6Block& blk, int* q, float* ep, int pairs)              // K:2156-2169
{
    bc6_quant(q, ep, blk.epb, pairs);
    for (int i = 0; i < 2 * pairs; i++)
        for (int c = 0; c < 3; c++) q[4 * i + c] = iclamp(q[4 * i + c], blk.qbounds[c], blk.qbounds[4 + c]);
    for (int i = 0; i < 8 * pairs; i++) ep[i] = (float)bc6_dequant(q[i], blk.epb);
}

void bc6_write_two_region(u32 out[4], int* q, const u32 idx[2], int shape, int mode)  // K:2982-3010
{
    int flips = orient_subsets(q, idx, 3, 2, shape);
    BitSink s;
    bc6_put_header(s, q, mode);
    s.put(5, (u32)shape);
    int anchors[3];
    shape_anchors(shape, anchors);
    put_indices(s, idx, 3, flips, anchors + 1, 1);
    memcpy(out, s.w, 16);
}
void bc6_write_one_region(u32 out[4], int* q, u32 idx[2], int mode)                    // K:3012-3031
{
    orient_single(q, 4, idx, 4);
    BitSink s;
    bc6_put_header(s, q, mode);
    put_indices(s, idx, 4, 0, 0, 0);
    memcpy(out, s.w, 16);
}

void bc6_encode_two_region(Bc6Block& blk)                                              // K:2174-2273
{
    int keys[32];
    rank_two_subset_shapes(keys, blk.px, 3, 32);
    const int count = blk.cfg->fastSkipTreshold;
    select_smallest(keys, 32, count);
    if (count == 0) return;

    int best_q[24] = {0};
    u32 best_idx[2] = {0, 0};
    int best_shape = -1;
    float best_err = kInf;
    for (int n = 0; n < count; n++) {
        int shape = keys[n] & 31;
        float ep[16] = {0};
        int q[16] = {0};
        u32 idx[2];
        for (int j = 0; j < 2; j++) fit_segment(ep + 8 * j, blk.px, shape_mask(shape, j), 3, false);
        bc6_quant_dequant(blk, q, ep, 2);
        float err = assign_indices(idx, blk.px, 3, ep, shape_pattern(shape), 3);
        if (err < best_err) {
            memcpy(best_q, q, sizeof(q));
            best_idx[0] = idx[0]; best_idx[1] = idx[1];
            best_shape = shape;
            best_err = err;
        }
    }
    for (int it = 0; it < blk.cfg->refineIterations_2p; it++) {
        float ep[16] = {0};
        int q[16] = {0};
        u32 idx[2];
        for (int j = 0; j < 2; j++) solve_endpoints(ep + 8 * j, blk.px, 3, best_idx, shape_mask(best_shape, j), 3);
        bc6_quant_dequant(blk, q, ep, 2);
        float err = assign_indices(idx, blk.px, 3, ep, shape_pattern(best_shape), 3);
        if (err < best_err) {
            memcpy(best_q, q, sizeof(q));
            best_idx[0] = idx[0]; best_idx[1] = idx[1];
            best_err = err;
        }
    }
    if (best_err < blk.best_err) {
        blk.best_err = best_err;
        bc6_write_two_region(blk.best, best_q, best_idx, best_shape, blk.mode);
    }
}
void bc6_encode_one_region(Bc6Block& blk)                                              // K:2275-2300
{
    float ep[8] = {0};
    int q[8] = {0};
    u32 idx[2];
    fit_segment(ep, blk.px, -1, 3, false);
    bc6_quant_dequant(blk, q, ep, 1);
    float err = assign_indices(idx, blk.px, 4, ep, 0, 3);
    for (int it = 0; it < blk.cfg->refineIterations_1p; it++) {
        solve_endpoints(ep, blk.px, 4, idx, -1, 3);
        bc6_quant_dequant(blk, q, ep, 1);
        err = assign_indices(idx, blk.px, 4, ep, 0, 3);
    }
    if (err < blk.best_err) {
        blk.best_err = err;
        bc6_write_one_region(blk.best, q, idx, blk.mode);
    }
}

// Quantised per-channel windows the endpoints are clamped to, centred on the block's range;
// K:2302-2330.  Slots 3 and 7 are quantised from zero (F6, quirk Q2) and never consumed.
void bc6_set_qbounds(Bc6Block& blk, float span, int wide_channel)
{
    float bounds[8] = {0};
    for (int c = 0; c < 3; c++) {
        float sp = span;
        if (wide_channel >= 0) sp *= (c == wide_channel) ? 2.0f : 1.0f;
        float middle = (blk.lo[c] + blk.hi[c]) / 2.0f;
        bounds[c] = middle - sp / 2.0f;
        bounds[4 + c] = middle + sp / 2.0f;
    }
    bc6_quant(blk.qbounds, bounds, blk.epb, 1);
}
void bc6_consider_mode(Bc6Block& blk, int mode, bool encode, float margin)              // K:2332-2365
{
    const float span = (float)bc6_span(mode);
    if (blk.max_span * margin > span) return;
    blk.epb = kBc6Epb[mode];
    if (mode >= 10) {
        blk.mode = mode;
        bc6_set_qbounds(blk, span, -1);
        if (encode) bc6_encode_one_region(blk);
    } else if (mode <= 1 || mode == 5 || mode == 9) {
        blk.mode = mode;
        bc6_set_qbounds(blk, span, -1);
        if (encode) bc6_encode_two_region(blk);
    } else {
        blk.mode = mode + blk.max_span_idx;
        bc6_set_qbounds(blk, span, blk.max_span_idx);
        if (encode) bc6_encode_two_region(blk);
    }
}
void bc6_encode_block(Bc6Block& blk)                                                    // K:3036-3107
{
    const bc6h_enc_settings* cfg = blk.cfg;
    blk.best_err = kInf;
    memset(blk.best, 0, sizeof(blk.best));
    blk.mode = 0; blk.epb = 0;
    memset(blk.qbounds, 0, sizeof(blk.qbounds));
    for (int c = 0; c < 3; c++) { blk.lo[c] = 65535.0f; blk.hi[c] = 0.0f; }
    for (int c = 0; c < 3; c++)
        for (int k = 0; k < 16; k++) {
            float v = (blk.px[16 * c + k] / 31.0f) * 64.0f;                             // K:3048
            blk.px[16 * c + k] = v;
            blk.lo[c] = fmin_sse(blk.lo[c], v);
            blk.hi[c] = fmax_sse(blk.hi[c], v);
        }
    blk.max_span = 0.0f;
    blk.max_span_idx = 0;
    for (int c = 0; c < 3; c++) {
        float sp = blk.hi[c] - blk.lo[c];
        if (sp > blk.max_span) { blk.max_span_idx = c; blk.max_span = sp; }
    }

    if (cfg->slow_mode) {                                                               // K:3073-3085
        const int order[10] = {0, 1, 2, 5, 6, 9, 10, 11, 12, 13};
        for (int i = 0; i < 10; i++) bc6_consider_mode(blk, order[i], true, 0.0f);
    } else {                                                                            // K:3086-3106
        if (cfg->fastSkipTreshold > 0) {
            const float m12 = 1.0f / 1.2f;
            bc6_consider_mode(blk, 9, false, 0.0f);
            if (cfg->fast_mode) bc6_consider_mode(blk, 1, false, 1.0f);
            bc6_consider_mode(blk, 6, false, m12);
            bc6_consider_mode(blk, 5, false, m12);
            bc6_consider_mode(blk, 0, false, m12);
            bc6_consider_mode(blk, 2, false, 1.0f);
            bc6_encode_two_region(blk);
            if (!cfg->fast_mode) bc6_consider_mode(blk, 1, true, 0.0f);
        }
        bc6_consider_mode(blk, 10, false, 0.0f);
        bc6_consider_mode(blk, 11, false, 1.0f);
        bc6_consider_mode(blk, 12, false, 1.0f);
        bc6_consider_mode(blk, 13, false, 1.0f);
        bc6_encode_one_region(blk);
    }
}

// ==========================================================================================
// BC4 / BC5 (DirectXTex path; IntelPlugin.cpp:272 -> DirectXTex/BC4BC5.cpp, BC.h:727-856)
// PINNING: byte-identical to DirectXTex's own encoder bodies (oracle/build_ref_frontend.py -> tests/test_bc45_vs_directxtex.py);
// the byte -> float texel load (rule F7) is the one assumption.
// ==========================================================================================
// OptimizeAlpha<false> (DirectXTex/BC.h:727-856): Newton refinement of the two endpoints of a
// `steps`-entry ramp over values in [0,1].
void bc4_optimise(float* px, float* py, const float* pts, int steps)
{
    static const float c6[] = {5.0f / 5.0f, 4.0f / 5.0f, 3.0f / 5.0f, 2.0f / 5.0f, 1.0f / 5.0f, 0.0f / 5.0f};
    static const float d6[] = {0.0f / 5.0f, 1.0f / 5.0f, 2.0f / 5.0f, 3.0f / 5.0f, 4.0f / 5.0f, 5.0f / 5.0f};
    static const float c8[] = {7.0f / 7.0f, 6.0f / 7.0f, 5.0f / 7.0f, 4.0f / 7.0f, 3.0f / 7.0f, 2.0f / 7.0f, 1.0f / 7.0f, 0.0f / 7.0f};
    static const float d8[] = {0.0f / 7.0f, 1.0f / 7.0f, 2.0f / 7.0f, 3.0f / 7.0f, 4.0f / 7.0f, 5.0f / 7.0f, 6.0f / 7.0f, 7.0f / 7.0f};
    const float* pc = (steps == 6) ? c6 : c8;
    const float* pd = (steps == 6) ? d6 : d8;
    const float MAX_VALUE = 1.0f, MIN_VALUE = 0.0f;

    float fx = MAX_VALUE, fy = MIN_VALUE;
    if (steps == 8) {
        for (int i = 0; i < 16; i++) {
            if (pts[i] < fx) fx = pts[i];
            if (pts[i] > fy) fy = pts[i];
        }
    } else {
        for (int i = 0; i < 16; i++) {
            if (pts[i] < fx && pts[i] > MIN_VALUE) fx = pts[i];
            if (pts[i] > fy && pts[i] < MAX_VALUE) fy = pts[i];
        }
        if (fx == fy) fy = MAX_VALUE;
    }
    float fsteps = (float)(steps - 1);
    for (int iter = 0; iter < 8; iter++) {
        if ((fy - fx) < (1.0f / 256.0f)) break;
        float fscale = fsteps / (fy - fx);
        float pstep[8];
        for (int s = 0; s < steps; s++) pstep[s] = pc[s] * fx + pd[s] * fy;
        if (steps == 6) { pstep[6] = MIN_VALUE; pstep[7] = MAX_VALUE; }
        float dx = 0.0f, dy = 0.0f, d2x = 0.0f, d2y = 0.0f;
        for (int i = 0; i < 16; i++) {
            float fdot = (pts[i] - fx) * fscale;
            int istep;
            if (fdot <= 0.0f) istep = ((6 == steps) && (pts[i] <= fx * 0.5f)) ? 6 : 0;
            else if (fdot >= fsteps) istep = ((6 == steps) && (pts[i] >= (fy + 1.0f) * 0.5f)) ? 7 : (steps - 1);
            else istep = (int)(fdot + 0.5f);
            if (istep < steps) {
                float fdiff = pstep[istep] - pts[i];
                dx += pc[istep] * fdiff;
                d2x += pc[istep] * pc[istep];
                dy += pd[istep] * fdiff;
                d2y += pd[istep] * pd[istep];
            }
        }
        if (d2x > 0.0f) fx -= dx / d2x;
        if (d2y > 0.0f) fy -= dy / d2y;
        if (fx > fy) { float t = fx; fx = fy; fy = t; }
        if ((dx * dx < (1.0f / 64.0f)) && (dy * dy < (1.0f / 64.0f))) break;
    }
    *px = (fx < MIN_VALUE) ? MIN_VALUE : (fx > MAX_VALUE) ? MAX_VALUE : fx;
    *py = (fy < MIN_VALUE) ? MIN_VALUE : (fy > MAX_VALUE) ? MAX_VALUE : fy;
}
float bc4_palette_entry(int e0, int e1, int i)                    // BC4BC5.cpp:48-71
{
    if (i == 0) return (float)e0 / 255.0f;
    if (i == 1) return (float)e1 / 255.0f;
    float f0 = (float)e0 / 255.0f, f1 = (float)e1 / 255.0f;
    if (e0 > e1) {
        i -= 1;
        return (f0 * (float)(7 - i) + f1 * (float)i) / 7.0f;
    }
    if (i == 6) return 0.0f;
    if (i == 7) return 1.0f;
    i -= 1;
    return (f0 * (float)(5 - i) + f1 * (float)i) / 5.0f;
}
// One BC4U block from 16 values in [0,1]; BC4BC5.cpp:186-238, :314-337, :403-421
void bc4_encode_channel(uint8_t out[8], const float t[16])
{
    float bmax = t[0], bmin = t[0];
    for (int i = 0; i < 16; i++) {
        if (t[i] < bmin) bmin = t[i];
        else if (t[i] > bmax) bmax = t[i];
    }
    bool four_block = (0.0f == bmin || 1.0f == bmax);
    float fs, fe;
    int e0, e1;
    if (!four_block) {
        bc4_optimise(&fs, &fe, t, 8);
        int is = (uint8_t)(int)(fs * 255.0f), ie = (uint8_t)(int)(fe * 255.0f);
        e0 = ie; e1 = is;
    } else {
        bc4_optimise(&fs, &fe, t, 6);
        int is = (uint8_t)(int)(fs * 255.0f), ie = (uint8_t)(int)(fe * 255.0f);
        e1 = ie; e0 = is;
    }
    float pal[8];
    for (int i = 0; i < 8; i++) pal[i] = bc4_palette_entry(e0, e1, i);
    uint64_t data = (uint64_t)e0 | ((uint64_t)e1 << 8);
    for (int i = 0; i < 16; i++) {
        int best = 0;
        float best_d = 100000.0f;
        for (int j = 0; j < 8; j++) {
            float d = fabsf(pal[j] - t[i]);
            if (d < best_d) { best = j; best_d = d; }
        }
        data |= (uint64_t)best << (3 * i + 16);
    }
    memcpy(out, &data, 8);
}

}  // namespace

// ==========================================================================================
// exported entry points (oracle_ prefix so the library can be loaded beside the product)
// ==========================================================================================
extern "C" {

void oracle_CompressBlocksBC1(const rgba_surface* src, uint8_t* dst)                   // K:573-583, :598-605
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            float px[48];
            u32 w[2];
            fetch_rgba8(px, src, bx, by, 3);
            bc1_colour_block(px, w);
            emit_block(dst, src, bx, by, w, 2);
        }
}
void oracle_CompressBlocksBC3(const rgba_surface* src, uint8_t* dst)                   // K:585-596, :607-614
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            float px[64];
            u32 w[4];
            fetch_rgba8(px, src, bx, by, 4);
            bc3_alpha_block(px + 48, w);
            bc1_colour_block(px, w + 2);
            emit_block(dst, src, bx, by, w, 4);
        }
}
void oracle_CompressBlocksBC7(const rgba_surface* src, uint8_t* dst, bc7_enc_settings* settings)   // K:2014-2037
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            Bc7Block blk;
            blk.cfg = settings;
            fetch_rgba8(blk.px, src, bx, by, 4);
            bc7_encode_block(blk);
            emit_block(dst, src, bx, by, blk.best, 4);
        }
}
void oracle_CompressBlocksBC6H(const rgba_surface* src, uint8_t* dst, bc6h_enc_settings* settings) // K:3118-3139
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            Bc6Block blk;
            blk.cfg = settings;
            fetch_rgba16(blk.px, src, bx, by);
            bc6_encode_block(blk);
            emit_block(dst, src, bx, by, blk.best, 4);
        }
}
// DirectX::Compress -> _CompressBC -> D3DXEncodeBC4U (DirectXTexCompress.cpp:73-186, BC4BC5.cpp:403);
// texel float = byte * (1/255) (rule F7).
void oracle_CompressBlocksBC4(const rgba_surface* src, uint8_t* dst)
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            float px[64], t[16];
            uint8_t out[8];
            fetch_rgba8(px, src, bx, by, 1);
            for (int k = 0; k < 16; k++) t[k] = px[k] * (1.0f / 255.0f);
            bc4_encode_channel(out, t);
            memcpy(dst + ((size_t)by * (size_t)(src->width / 4) + (size_t)bx) * 8, out, 8);
        }
}
void oracle_CompressBlocksBC5(const rgba_surface* src, uint8_t* dst)                   // BC4BC5.cpp:481-512
{
    for (int by = 0; by < src->height / 4; by++)
        for (int bx = 0; bx < src->width / 4; bx++) {
            float px[64], t[16];
            uint8_t out[16];
            fetch_rgba8(px, src, bx, by, 2);
            for (int c = 0; c < 2; c++) {
                for (int k = 0; k < 16; k++) t[k] = px[16 * c + k] * (1.0f / 255.0f);
                bc4_encode_channel(out + 8 * c, t);
            }
            memcpy(dst + ((size_t)by * (size_t)(src->width / 4) + (size_t)bx) * 16, out, 16);
        }
}

// ---- profiles: TCc:20-410 restated as data -------------------------------------------------
// columns: channels | mode_selection[0..3] | skip2 | T1 T3 T7 | refine[0..7] | ch0 | refine_channel
// refine[7] is only written by the alpha profiles (TCc:191-365); -1 below = "left untouched".
struct Bc7ProfileRow { int channels; int sel[4]; int skip2; int t1, t3, t7; int refine[8]; int ch0; int rch; };
static const Bc7ProfileRow kBc7Profiles[10] = {
    /* ultrafast       TCc:20-50   */ {3, {0, 0, 0, 1}, 1, 3, 1, 0, {2, 2, 2, 1, 2, 2, 1, -1}, 0, 0},
    /* veryfast        TCc:52-82   */ {3, {0, 1, 0, 1}, 1, 3, 1, 0, {2, 2, 2, 1, 2, 2, 1, -1}, 0, 0},
    /* fast            TCc:84-120  */ {3, {0, 1, 0, 1}, 1, 12, 4, 0, {2, 2, 2, 1, 2, 2, 2, -1}, 0, 0},
    /* basic           TCc:122-154 */ {3, {1, 1, 1, 1}, 1, 12, 8, 0, {2, 2, 2, 2, 2, 2, 2, -1}, 0, 2},
    /* slow            TCc:156-189 */ {3, {1, 1, 1, 1}, 0, 64, 64, 0, {4, 4, 4, 4, 4, 4, 4, -1}, 0, 4},
    /* alpha_ultrafast TCc:191-224 */ {4, {0, 0, 1, 1}, 1, 0, 0, 4, {2, 1, 2, 1, 1, 1, 2, 2}, 3, 1},
    /* alpha_veryfast  TCc:226-259 */ {4, {0, 1, 1, 1}, 1, 0, 0, 4, {2, 1, 2, 1, 2, 2, 2, 2}, 3, 2},
    /* alpha_fast      TCc:261-294 */ {4, {0, 1, 1, 1}, 1, 4, 4, 8, {2, 1, 2, 1, 2, 2, 2, 2}, 3, 2},
    /* alpha_basic     TCc:296-329 */ {4, {1, 1, 1, 1}, 1, 12, 8, 8, {2, 2, 2, 2, 2, 2, 2, 2}, 0, 2},
    /* alpha_slow      TCc:331-365 */ {4, {1, 1, 1, 1}, 0, 64, 64, 64, {4, 4, 4, 4, 4, 4, 4, 4}, 0, 4},
};
static void fill_bc7(bc7_enc_settings* s, int row)
{
    const Bc7ProfileRow& r = kBc7Profiles[row];
    s->channels = r.channels;
    for (int i = 0; i < 4; i++) s->mode_selection[i] = r.sel[i] != 0;
    s->skip_mode2 = r.skip2 != 0;
    s->fastSkipTreshold_mode1 = r.t1;
    s->fastSkipTreshold_mode3 = r.t3;
    s->fastSkipTreshold_mode7 = r.t7;
    for (int i = 0; i < 8; i++)
        if (r.refine[i] >= 0) s->refineIterations[i] = r.refine[i];
    s->mode45_channel0 = r.ch0;
    s->refineIterations_channel = r.rch;
}
void oracle_GetProfile_ultrafast(bc7_enc_settings* s) { fill_bc7(s, 0); }
void oracle_GetProfile_veryfast(bc7_enc_settings* s) { fill_bc7(s, 1); }
void oracle_GetProfile_fast(bc7_enc_settings* s) { fill_bc7(s, 2); }
void oracle_GetProfile_basic(bc7_enc_settings* s) { fill_bc7(s, 3); }
void oracle_GetProfile_slow(bc7_enc_settings* s) { fill_bc7(s, 4); }
void oracle_GetProfile_alpha_ultrafast(bc7_enc_settings* s) { fill_bc7(s, 5); }
void oracle_GetProfile_alpha_veryfast(bc7_enc_settings* s) { fill_bc7(s, 6); }
void oracle_GetProfile_alpha_fast(bc7_enc_settings* s) { fill_bc7(s, 7); }
void oracle_GetProfile_alpha_basic(bc7_enc_settings* s) { fill_bc7(s, 8); }
void oracle_GetProfile_alpha_slow(bc7_enc_settings* s) { fill_bc7(s, 9); }

static void fill_bc6(bc6h_enc_settings* s, int slow, int fast, int skip, int r1, int r2)   // TCc:367-410
{
    s->slow_mode = slow != 0;
    s->fast_mode = fast != 0;
    s->fastSkipTreshold = skip;
    s->refineIterations_1p = r1;
    s->refineIterations_2p = r2;
}
void oracle_GetProfile_bc6h_veryfast(bc6h_enc_settings* s) { fill_bc6(s, 0, 1, 0, 0, 0); }
void oracle_GetProfile_bc6h_fast(bc6h_enc_settings* s) { fill_bc6(s, 0, 1, 2, 0, 1); }
void oracle_GetProfile_bc6h_basic(bc6h_enc_settings* s) { fill_bc6(s, 0, 0, 4, 2, 2); }
void oracle_GetProfile_bc6h_slow(bc6h_enc_settings* s) { fill_bc6(s, 1, 0, 10, 2, 2); }
void oracle_GetProfile_bc6h_veryslow(bc6h_enc_settings* s) { fill_bc6(s, 1, 0, 32, 2, 2); }

}  // extern "C"
