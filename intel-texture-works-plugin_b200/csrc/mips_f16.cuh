// mips_f16.cuh -- mip chain of RGBA16F textures (the BC6H save path), pinned to the reference's own filter code.
//
// For BC6H the plug-in forces DirectXTex's non-WIC generator (IntelPlugin.cpp:2117-2127: TEX_FILTER_DEFAULT |
// TEX_FILTER_SEPARATE_ALPHA | TEX_FILTER_FORCE_NON_WIC), i.e. code that IS in the reference tree:
//   GenerateMipMaps                DirectXTex/DirectXTexMipmaps.cpp:2611-2650   filter = BOX if width and height are powers
//                                                                               of two, else LINEAR -- chosen ONCE from level 0
//   _Generate2DMipsBoxFilter       :715-805    with AVERAGE4 (Filters.h:33-39):  ((p0 + p1) + p2) + p3) * 0.25
//   _Generate2DMipsLinearFilter    :809-905    with _CreateLinearFilter / BILINEAR_INTERPOLATE (Filters.h:60-112)
// Every level is read back from its stored halves (XMLoadHalf4) and written through XMStoreHalf4, so a level depends only
// on the previous level's half bits: one independent pass per level, one thread per output texel.  Float operations are
// the reference's, in its order, without contraction; the half conversions are the DirectXMath 3.06 restatement of
// frontend.cuh (same caveat: DirectXMath is not in the reference tree).  Padding to multiples of 4 (IntelPlugin.cpp:892-928)
// is the coordinate clamp used by mips.cuh.
//
// Reference quirk kept on purpose (bit-exact drop-in): in the box loop the fourth tap pointer `urow3` is computed ONCE as
// `urow1 + 1` (:738) and is only redirected when the width reaches 1 (:749-752).  When the HEIGHT reaches 1 first (wide
// power-of-two textures) `urow1` is redirected to `urow0` (:744-747) but `urow3` keeps pointing into the second scanline
// buffer, which is no longer loaded: the fourth tap then reads the LAST ROW of the last source level that still had two
// rows, at column 2x+1.  `stale_row` carries that row; a buffer that was never loaded (level 0 itself one texel high) reads
// as zero (rule F6).
#pragma once
#include "frontend.cuh"
#include "srgb_tables.cuh"

namespace itw {

struct MipTap { int u0, u1; float w0, w1; };
// _CreateLinearFilter for one destination coordinate, clamp addressing; Filters.h:67-100
ITW_HD MipTap mip_linear_tap(int source, int dest, int u)
{
    const float scale = (float)source / (float)dest;
    const float srcB = ((float)u + 0.5f) * scale + 0.5f;
    int isrcB = (int)srcB, isrcA = isrcB - 1;
    if (isrcA < 0) isrcA = 0;
    if (isrcB >= source) isrcB = source - 1;
    const float weight = 1.0f + (float)isrcB - srcB;
    return MipTap{isrcA, isrcB, weight, 1.0f - weight};
}
// the two-tap degenerate cases of the box loop: a 1-texel-wide (-high) source reads the same column (row) twice; :742-752
ITW_HD MipTap mip_box_tap(int source, int u)
{
    return MipTap{2 * u, (source > 1) ? 2 * u + 1 : 2 * u, 0.0f, 0.0f};
}
// ---- texel codecs ---------------------------------------------------------------------------------------------------
// kCodec 0: RGBA16F (8-byte texels), XMLoadHalf4 / XMStoreHalf4.
// kCodec 1: RGBA8 UNORM: XMLoadUByteN4 = byte * (1/255) (rule F7), XMStoreUByteN4 = saturate, * 255, round to nearest (the
//           store's rounding is DirectXMath's, outside the tree: assumed, like in the decoder tests).
// kCodec 2: RGBA8 UNORM_SRGB: RGB additionally through XMColorSRGBToRGB after the load and XMColorRGBToSRGB before the store
//           (DirectXTexConvert.cpp:2669-2685, :2757-2775) -- what the plug-in's chain is for *_SRGB encodings (IntelPlugin.cpp:
//           152-154 forces the scratch format, DirectXTexMipmaps.cpp:389-393 then avoids WIC).  Both functions are tables
//           derived from the oracle's own powf (srgb_tables.cuh, tools/gen_srgb_tables.py): 256 linear values of the bytes, 255
//           thresholds of the monotone float -> byte store.
ITW_TABLE_DECL(uint32_t, srgb_to_linear, 256)
ITW_TABLE_DECL(uint32_t, linear_threshold, 255)
ITW_TABLE_DECL(uint32_t, linear_base, ITW_SRGB_BASE_WORDS)

ITW_HD float mip_saturate(float v) { v = (v > 0.0f) ? v : 0.0f; return (v < 1.0f) ? v : 1.0f; }
ITW_HD u32 mip_unorm8_store(float v) { return (u32)trunc_i(mip_saturate(v) * 255.0f + 0.5f); }
ITW_HD u32 mip_srgb8_store(float v)
{
    // result = number of thresholds <= v (non-negative floats order like their bit patterns).  Two table reads: the byte at the
    // start of v's bucket (top bits of the float), then one threshold comparison -- inside a bucket the byte rises by at most one.
    const u32 bits = float_bits(mip_saturate(v));
    if (bits < ((u32)ITW_SRGB_BASE_FIRST_EXP << 23)) return 0u;                 // below 2^-13 every value stores 0
    const u32 bucket = (bits >> 15) - ((u32)ITW_SRGB_BASE_FIRST_EXP << 8);
    const u32 base = (ITW_TABLE(linear_base)[bucket >> 2] >> (8 * (bucket & 3u))) & 255u;
    return base + ((base < 255u && ITW_TABLE(linear_threshold)[base] <= bits) ? 1u : 0u);
}
template <int kCodec>
ITW_HD void mip_load(float (&v)[4], const uint8_t* row, int x)
{
    if (kCodec == 0) {
        const u32* p = reinterpret_cast<const u32*>(row + (size_t)x * 8);
        const u32 lo = p[0], hi = p[1];
        v[0] = front_float_from_half(lo & 0xFFFFu);
        v[1] = front_float_from_half(lo >> 16);
        v[2] = front_float_from_half(hi & 0xFFFFu);
        v[3] = front_float_from_half(hi >> 16);
    } else {
        const u32 t = *reinterpret_cast<const u32*>(row + (size_t)x * 4);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const u32 byte = (t >> (8 * c)) & 255u;
            v[c] = (kCodec == 2 && c < 3) ? bits_float(ITW_TABLE(srgb_to_linear)[byte]) : (float)byte * (1.0f / 255.0f);
        }
    }
}
template <int kCodec>
ITW_HD void mip_store(u32 (&out)[2], const float (&res)[4])
{
    if (kCodec == 0) {
        out[0] = front_half_from_float(res[0]) | (front_half_from_float(res[1]) << 16);
        out[1] = front_half_from_float(res[2]) | (front_half_from_float(res[3]) << 16);
    } else {
        u32 t = mip_unorm8_store(res[3]) << 24;
#pragma unroll
        for (int c = 0; c < 3; c++) t |= ((kCodec == 2) ? mip_srgb8_store(res[c]) : mip_unorm8_store(res[c])) << (8 * c);
        out[0] = t;
        out[1] = 0u;
    }
}
ITW_HD void mip_load_f16(float (&v)[4], const uint8_t* row, int x) { mip_load<0>(v, row, x); }
// One texel of the padded level (dw x dh valid, any x / y inside the padded storage) from the previous level's valid region
template <int kCodec>
ITW_HD void mip_float_texel(u32 (&out)[2], const uint8_t* src, int sw, int sh, long long sstride, int dw, int dh, int x, int y, bool box,
                            const uint8_t* stale_row)
{
    const int cx = mini(x, dw - 1), cy = mini(y, dh - 1);
    const MipTap tx = box ? mip_box_tap(sw, cx) : mip_linear_tap(sw, dw, cx);
    const MipTap ty = box ? mip_box_tap(sh, cy) : mip_linear_tap(sh, dh, cy);
    const uint8_t* r0 = src + (long long)ty.u0 * sstride;
    const uint8_t* r1 = src + (long long)ty.u1 * sstride;
    float a[4], b[4], c[4], d[4], res[4];
    mip_load<kCodec>(a, r0, tx.u0);       // (row u0, column u0)
    mip_load<kCodec>(b, r0, tx.u1);       // (row u0, column u1)
    mip_load<kCodec>(c, r1, tx.u0);       // (row u1, column u0)
    mip_load<kCodec>(d, r1, tx.u1);       // (row u1, column u1)
    if (box && sh <= 1 && sw > 1) {       // the stale fourth tap, see the header
        if (stale_row) mip_load<kCodec>(d, stale_row, 2 * cx + 1);
        else d[0] = d[1] = d[2] = d[3] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (box) {
            // AVERAGE4(urow0[x2], urow1[x2], urow0[x2+1], urow1[x2+1])
            res[i] = (((a[i] + c[i]) + b[i]) + d[i]) * 0.25f;
        } else {
            // y.w0 * (r0[x.u0]*x.w0 + r0[x.u1]*x.w1) + y.w1 * (r1[x.u0]*x.w0 + r1[x.u1]*x.w1)
            res[i] = (ty.w0 * (a[i] * tx.w0 + b[i] * tx.w1)) + (ty.w1 * (c[i] * tx.w0 + d[i] * tx.w1));
        }
    }
    mip_store<kCodec>(out, res);
}
ITW_HD void mip_f16_texel(u32 (&out)[2], const uint8_t* src, int sw, int sh, long long sstride, int dw, int dh, int x, int y, bool box,
                          const uint8_t* stale_row)
{
    mip_float_texel<0>(out, src, sw, sh, sstride, dw, dh, x, y, box, stale_row);
}

#if defined(__CUDACC__)
// grid: (ceil(pw/64), ph); thread = one padded output texel.  RGBA16F: 32 B read + 8 B written per texel, RGBA8: 16 B + 4 B.
template <int kCodec>
__global__ void __launch_bounds__(64) mip_float_kernel(const uint8_t* __restrict__ src, int sw, int sh, long long sstride, uint8_t* __restrict__ dst,
                                                       int dw, int dh, int pw, long long dstride, int box, const uint8_t* __restrict__ stale_row)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y;
    if (x >= pw) return;
    u32 out[2];
    mip_float_texel<kCodec>(out, src, sw, sh, sstride, dw, dh, x, y, box != 0, stale_row);
    if (kCodec == 0) *reinterpret_cast<uint2*>(dst + (long long)y * dstride + (long long)x * 8) = make_uint2(out[0], out[1]);
    else *reinterpret_cast<u32*>(dst + (long long)y * dstride + (long long)x * 4) = out[0];
}
// edge-replicating copy of level 0 into padded storage (8-byte texels)
__global__ void __launch_bounds__(64) pad_f16_kernel(const uint8_t* __restrict__ src, int sw, int sh, long long sstride, uint8_t* __restrict__ dst, int pw,
                                                     long long dstride)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y;
    if (x >= pw) return;
    const u32* p = reinterpret_cast<const u32*>(src + (long long)mini(y, sh - 1) * sstride + (long long)mini(x, sw - 1) * 8);
    *reinterpret_cast<uint2*>(dst + (long long)y * dstride + (long long)x * 8) = make_uint2(p[0], p[1]);
}
#endif

}  // namespace itw
