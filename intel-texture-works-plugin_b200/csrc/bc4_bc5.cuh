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

// Endpoint optimisation over a ramp of 6 or 8 steps (BC.h:727-856).  kSteps is a template
// parameter so that the ramp coefficient tables fold into immediates.
template <int kSteps>
ITW_HD void bc4_fit_ramp(float& out_lo, float& out_hi, const float (&t)[16])
{
    const float last = (float)(kSteps - 1);
    float coef_lo[8], coef_hi[8];
#pragma unroll
    for (int s = 0; s < kSteps; s++) {
        coef_lo[s] = (float)(kSteps - 1 - s) / last;
        coef_hi[s] = (float)s / last;
    }
    float lo = 1.0f, hi = 0.0f;
    if (kSteps == 8) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (t[i] < lo) lo = t[i];
            if (t[i] > hi) hi = t[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (t[i] < lo && t[i] > 0.0f) lo = t[i];
            if (t[i] > hi && t[i] < 1.0f) hi = t[i];
        }
        if (lo == hi) hi = 1.0f;
    }
    for (int iter = 0; iter < 8; iter++) {
        if ((hi - lo) < (1.0f / 256.0f)) break;
        const float scale = last / (hi - lo);
        float ramp[8];
#pragma unroll
        for (int s = 0; s < kSteps; s++) ramp[s] = coef_lo[s] * lo + coef_hi[s] * hi;
        float dlo = 0.0f, dhi = 0.0f, d2lo = 0.0f, d2hi = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float dot = (t[i] - lo) * scale;
            int step;
            if (dot <= 0.0f) step = ((kSteps == 6) && (t[i] <= lo * 0.5f)) ? 6 : 0;
            else if (dot >= last) step = ((kSteps == 6) && (t[i] >= (hi + 1.0f) * 0.5f)) ? 7 : (kSteps - 1);
            else step = (int)(dot + 0.5f);
            if (step < kSteps) {
                // select the step's ramp value / coefficients without dynamic register indexing
                float r = ramp[0], cl = coef_lo[0], ch = coef_hi[0];
#pragma unroll
                for (int s = 1; s < kSteps; s++)
                    if (step == s) { r = ramp[s]; cl = coef_lo[s]; ch = coef_hi[s]; }
                const float diff = r - t[i];
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
    out_lo = (lo < 0.0f) ? 0.0f : ((lo > 1.0f) ? 1.0f : lo);
    out_hi = (hi < 0.0f) ? 0.0f : ((hi > 1.0f) ? 1.0f : hi);
}

// 16 byte values of one channel -> one 8-byte BC4U block (two words)
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
    float fs, fe;
    int e0, e1;
    if (!six) {
        bc4_fit_ramp<8>(fs, fe, t);
        e1 = (int)(fs * 255.0f) & 255;          // (uint8_t) truncation; values are in [0,255]
        e0 = (int)(fe * 255.0f) & 255;
    } else {
        bc4_fit_ramp<6>(fs, fe, t);
        e0 = (int)(fs * 255.0f) & 255;
        e1 = (int)(fe * 255.0f) & 255;
    }
    // palette; BC4BC5.cpp:48-71
    float pal[8];
    const float f0 = (float)e0 / 255.0f, f1 = (float)e1 / 255.0f;
    pal[0] = f0;
    pal[1] = f1;
    if (e0 > e1) {
#pragma unroll
        for (int i = 2; i < 8; i++) pal[i] = (f0 * (float)(8 - i) + f1 * (float)(i - 1)) / 7.0f;
    } else {
#pragma unroll
        for (int i = 2; i < 6; i++) pal[i] = (f0 * (float)(6 - i) + f1 * (float)(i - 1)) / 5.0f;
        pal[6] = 0.0f;
        pal[7] = 1.0f;
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
