// bc4_bc5.cuh -- BC4 / BC5 (UNORM) encoder.  In the reference these two formats do not go through
// the ISPC kernels but through DirectX::Compress (IntelPlugin.cpp:272):
//   DirectXTex/DirectXTexCompress.cpp:73-186  (4x4 gather, texel -> float)
//   DirectXTex/BC4BC5.cpp:186-238 FindEndPointsBC4U, :314-337 FindClosestUNORM, :403/:481 encoders
//   DirectXTex/BC.h:727-856 OptimizeAlpha<false>
// One thread owns one block (BC5: both channels), loads/stores coalesced as in bc1_bc3.cuh.
// Texel float = byte * (1/255) (DESIGN.md rule F7).  Algorithmic traffic per block: 64 B read,
// 8 B (BC4) / 16 B (BC5) written.
#pragma once
#include "bc1_bc3.cuh"

namespace itw {

// 16 byte values of one channel -> one 8-byte BC4U block (two words).
// FindEndPointsBC4U (BC4BC5.cpp:186-238) + OptimizeAlpha<false> (BC.h:727-856) + FindClosestUNORM (:314-337).
// The reference has two code paths (6- and 8-step ramps); here the ramp length is DATA (`steps`), so the
// threads of a warp never diverge on it: blocks that touch 0 or 255 and blocks that do not run the same
// instruction stream.  The ramp coefficients (steps-1-s)/(steps-1) and s/(steps-1), compile-time constants
// in the reference (BC.h:730-733), are the same correctly rounded quotients computed at run time.
ITW_HD void bc4_encode_channel(const int (&v)[16], u32& w0, u32& w1)
{
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = (float)v[i] * (1.0f / 255.0f);
    float bmax = t[0], bmin = t[0];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (t[i] < bmin) bmin = t[i];
        else if (t[i] > bmax) bmax = t[i];
    }
    // blocks touching 0 or 1 use the 6-step ramp with explicit 0/1 codes; BC4BC5.cpp:206-237
    const bool six = (0.0f == bmin || 1.0f == bmax);
    const int steps = six ? 6 : 8;
    const float last = (float)(steps - 1), rlast = 1.0f / last;

    // starting interval; BC.h:748-776
    float lo = 1.0f, hi = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (t[i] < lo && (!six || t[i] > 0.0f)) lo = t[i];
        if (t[i] > hi && (!six || t[i] < 1.0f)) hi = t[i];
    }
    if (six && lo == hi) hi = 1.0f;

    // Newton iterations on the two endpoints; BC.h:778-851
#pragma unroll 1
    for (int iter = 0; iter < 8; iter++) {
        if ((hi - lo) < (1.0f / 256.0f)) break;
        const float scale = last / (hi - lo);
        float dlo = 0.0f, dhi = 0.0f, d2lo = 0.0f, d2hi = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float dot = (t[i] - lo) * scale;
            int step;
            if (dot <= 0.0f) step = (six && (t[i] <= lo * 0.5f)) ? 6 : 0;
            else if (dot >= last) step = (six && (t[i] >= (hi + 1.0f) * 0.5f)) ? 7 : (steps - 1);
            else step = (int)(dot + 0.5f);
            if (step < steps) {
                // integer / {5,7}: the exact quotient (tests/test_exact_division.py covers this domain)
                const float cl = div_by_rcp((float)(steps - 1 - step), last, rlast);
                const float ch = div_by_rcp((float)step, last, rlast);
                const float diff = (cl * lo + ch * hi) - t[i];
                dlo += cl * diff;
                d2lo += cl * cl;
                dhi += ch * diff;
                d2hi += ch * ch;
            }
        }
        if (d2lo > 0.0f) lo -= dlo / d2lo;
        if (d2hi > 0.0f) hi -= dhi / d2hi;
        if (lo > hi) { float f = lo; lo = hi; hi = f; }
        if ((dlo * dlo < (1.0f / 64.0f)) && (dhi * dhi < (1.0f / 64.0f))) break;
    }
    const float fs = (lo < 0.0f) ? 0.0f : ((lo > 1.0f) ? 1.0f : lo);
    const float fe = (hi < 0.0f) ? 0.0f : ((hi > 1.0f) ? 1.0f : hi);
    const int is = (int)(fs * 255.0f) & 255, ie = (int)(fe * 255.0f) & 255;      // (uint8_t) truncation
    const int e0 = six ? is : ie, e1 = six ? ie : is;                             // BC4BC5.cpp:222-236

    // palette; BC4BC5.cpp:48-71 (decode mode follows the stored endpoint order, not the encoder's intent)
    float pal[8];
    const float f0 = (float)e0 / 255.0f, f1 = (float)e1 / 255.0f;
    const bool eight = e0 > e1;
    const float n = eight ? 7.0f : 5.0f;
    pal[0] = f0;
    pal[1] = f1;
#pragma unroll
    for (int i = 2; i < 8; i++) {
        const float k = (float)(i - 1);
        float p = (f0 * (n - k) + f1 * k) / n;
        if (!eight && i == 6) p = 0.0f;
        if (!eight && i == 7) p = 1.0f;
        pal[i] = p;
    }
    unsigned long long data = (unsigned long long)e0 | ((unsigned long long)e1 << 8);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int best = 0;
        float best_d = 100000.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float d = fabsf(pal[j] - t[i]);
            if (d < best_d) { best = j; best_d = d; }
        }
        data |= (unsigned long long)best << (3 * i + 16);
    }
    w0 = (u32)data;
    w1 = (u32)(data >> 32);
}

template <bool kTwoChannels>
ITW_HD void bc4_bc5_encode_block(const u32 (&tex)[16], u32 (&out)[4])
{
    int v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (int)(tex[k] & 255u);
    bc4_encode_channel(v, out[0], out[1]);
    if (kTwoChannels) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = (int)((tex[k] >> 8) & 255u);
        bc4_encode_channel(v, out[2], out[3]);
    } else {
        out[2] = out[3] = 0;
    }
}

#if defined(__CUDACC__)
template <bool kTwoChannels, bool kVec16>
__global__ void __launch_bounds__(128) bc4_bc5_kernel(SurfaceView s, uint8_t* __restrict__ dst)
{
    const int bw = s.width >> 2, bh = s.height >> 2;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)bw * bh) return;
    const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);
    u32 tex[16], out[4];
    fetch_rows_rgba8<kVec16>(tex, s, bx, by);
    bc4_bc5_encode_block<kTwoChannels>(tex, out);
    if (kTwoChannels) reinterpret_cast<uint4*>(dst)[id] = make_uint4(out[0], out[1], out[2], out[3]);
    else              reinterpret_cast<uint2*>(dst)[id] = make_uint2(out[0], out[1]);
}
#endif

}  // namespace itw
