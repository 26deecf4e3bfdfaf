// mips.cuh -- on-GPU pre-pass of the save path: mip-chain generation with a 2x2 box filter
// and padding of every level to multiples of 4 by edge replication, so that the encoder never has to leave
// the device between IntelPlugin.cpp:2117 (GenerateMipMaps) and :2171 (SaveToDDSMemory).
//
//   * filter: out = (a + b + c + d + 2) >> 2 per 8-bit channel (round to nearest), the integer form of the box
//     filter of DirectXTex/DirectXTexMipmaps.cpp:715-805.  The plug-in itself asks for TEX_FILTER_DEFAULT, which
//     on Windows goes through WIC (not vendored, not restatable): parity of the mip CONTENT with the plug-in is
//     therefore unpinned; the contract here is the integer formula above (tests/test_mips.py, numpy restatement).
//   * level l has max(1, w>>l) x max(1, h>>l) texels (floor, like DirectXTex); a 1-texel-wide/high source
//     degenerates to the 2-tap average (a + b + 1) >> 1 through coordinate clamping.
//   * padding: texel (x, y) of the padded level reads (min(x, w-1), min(y, h-1)) -- exactly the edge replication
//     of DoPaddingToMultiplesOf4 (IntelPlugin.cpp:893-928) -- so box filter and padding are ONE bandwidth-bound
//     pass: 16 B read + 4 B written per output texel.
#pragma once
#include "itw_device.cuh"

namespace itw {

// 8-bit x4 box filter on packed RGBA
ITW_HD u32 box4_rgba8(u32 a, u32 b, u32 c, u32 d)
{
    const u32 m = 0x00FF00FFu;
    const u32 rb = ((a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u) >> 2;
    const u32 ga = (((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u) >> 2;
    return (rb & m) | ((ga & m) << 8);
}
// one output texel of a padded mip level; src = previous level (valid region sw x sh, tight or padded storage)
ITW_HD u32 mip_texel(const uint8_t* src, int sw, int sh, int sstride, int dw, int dh, int x, int y)
{
    const int cx = mini(x, dw - 1), cy = mini(y, dh - 1);
    const int x0 = mini(2 * cx, sw - 1), x1 = mini(2 * cx + 1, sw - 1);
    const int y0 = mini(2 * cy, sh - 1), y1 = mini(2 * cy + 1, sh - 1);
    const u32* r0 = reinterpret_cast<const u32*>(src + (size_t)y0 * (size_t)sstride);
    const u32* r1 = reinterpret_cast<const u32*>(src + (size_t)y1 * (size_t)sstride);
    return box4_rgba8(r0[x0], r0[x1], r1[x0], r1[x1]);
}

#if defined(__CUDACC__)
// grid: (ceil(pw/64), ph); block 64 threads: thread = one padded output texel
__global__ void __launch_bounds__(64) mip_box_rgba8_kernel(const uint8_t* __restrict__ src, int sw, int sh, int sstride,
                                                           uint8_t* __restrict__ dst, int dw, int dh, int pw, int dstride)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y;
    if (x >= pw) return;
    reinterpret_cast<u32*>(dst + (size_t)y * (size_t)dstride)[x] = mip_texel(src, sw, sh, sstride, dw, dh, x, y);
}
// Fast path for the large levels (>99 % of the traffic): source exactly 2x the destination, destination width a
// multiple of 4, both 16-byte aligned.  One thread makes 4 output texels from two 32-byte row segments:
// 128-bit loads and stores only, fully coalesced.  grid: (ceil(dw/4/128), dh); block 128.
__global__ void __launch_bounds__(128) mip_box_rgba8_x4_kernel(const uint8_t* __restrict__ src, int sstride,
                                                               uint8_t* __restrict__ dst, int dw, int dstride)
{
    const int q = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y;       // q = group of 4 output texels
    if (q * 4 >= dw) return;
    const uint4* r0 = reinterpret_cast<const uint4*>(src + (size_t)(2 * y) * (size_t)sstride) + 2 * q;
    const uint4* r1 = reinterpret_cast<const uint4*>(src + (size_t)(2 * y + 1) * (size_t)sstride) + 2 * q;
    const uint4 a0 = __ldg(r0), a1 = __ldg(r0 + 1), b0 = __ldg(r1), b1 = __ldg(r1 + 1);
    uint4 o;
    o.x = box4_rgba8(a0.x, a0.y, b0.x, b0.y);
    o.y = box4_rgba8(a0.z, a0.w, b0.z, b0.w);
    o.z = box4_rgba8(a1.x, a1.y, b1.x, b1.y);
    o.w = box4_rgba8(a1.z, a1.w, b1.z, b1.w);
    reinterpret_cast<uint4*>(dst + (size_t)y * (size_t)dstride)[q] = o;
}
// The small tail of the chain (every level of at most 64x64 padded texels) in ONE launch by one CTA: levels are
// produced one after the other with a block barrier in between (global writes of a CTA are visible to the same
// CTA after __syncthreads), which saves eight or nine ~3 us launches on a ~70 us pre-pass.
struct MipTail {
    int count;
    const uint8_t* src[12];
    uint8_t* dst[12];
    int sw[12], sh[12], sstride[12], dw[12], dh[12], pw[12], ph[12];
};
__global__ void __launch_bounds__(256) mip_tail_kernel(MipTail t)
{
    for (int l = 0; l < t.count; l++) {
        const int n = t.pw[l] * t.ph[l];
        for (int i = threadIdx.x; i < n; i += 256) {
            const int y = i / t.pw[l], x = i - y * t.pw[l];
            reinterpret_cast<u32*>(t.dst[l] + (size_t)y * (size_t)(t.pw[l] * 4))[x] =
                mip_texel(t.src[l], t.sw[l], t.sh[l], t.sstride[l], t.dw[l], t.dh[l], x, y);
        }
        __syncthreads();
    }
}
// edge-replicating copy of a (w x h) surface into its padded (pw x ph) storage (level 0 of an unpadded texture)
__global__ void __launch_bounds__(64) pad_rgba8_kernel(const uint8_t* __restrict__ src, int sw, int sh, int sstride,
                                                       uint8_t* __restrict__ dst, int pw, int dstride)
{
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y;
    if (x >= pw) return;
    const u32* r = reinterpret_cast<const u32*>(src + (size_t)mini(y, sh - 1) * (size_t)sstride);
    reinterpret_cast<u32*>(dst + (size_t)y * (size_t)dstride)[x] = r[mini(x, sw - 1)];
}
#endif

}  // namespace itw
