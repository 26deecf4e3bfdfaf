// check_f32x2.cu -- GPU check of the packed two-lane float operations (FADD2 / FMUL2 / FFMA2, sm_100a) against their scalar
// IEEE forms, bit for bit, over random operands of every magnitude class; and of the two-blocks-per-thread BC1 / BC3 encoder
// (csrc/bc1_pair.cuh) against the one-block form (csrc/bc1_bc3.cuh) on a smooth gradient, printing the inputs of differing blocks.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -o check_f32x2 check_f32x2.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "experiments/bc1_pair.cuh"

using namespace itw;

__device__ unsigned long long g_bad[8];
__device__ unsigned g_ex[8][6];

__device__ __forceinline__ unsigned rnd(unsigned long long& s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(s >> 32);
}
// operands of class k: 0 any finite bit pattern, 1 small integers, 2 [0,256) multiples of 1/16, 3 denormals / tiny, 4 around 2^23
__device__ float operand(unsigned long long& s, int k)
{
    unsigned r = rnd(s);
    if (k == 0) { unsigned e = (r >> 23) & 0xFF; if (e == 0xFF) r &= ~(1u << 30); return __uint_as_float(r); }
    if (k == 1) return (float)((int)(r % 1021) - 510);
    if (k == 2) return (float)(r % 4096) * 0.0625f * ((r & 0x80000000u) ? -1.0f : 1.0f);
    if (k == 3) return __uint_as_float(r & 0x80FFFFFFu);
    return 8388608.0f + (float)(r % 1024) * ((r & 0x40000000u) ? 1.0f : 0.0009765625f);
}
__device__ void report(int slot, unsigned a, unsigned b, unsigned c, unsigned got, unsigned want)
{
    const unsigned long long n = atomicAdd(&g_bad[slot], 1ull);
    if (n == 0) { g_ex[slot][0] = a; g_ex[slot][1] = b; g_ex[slot][2] = c; g_ex[slot][3] = got; g_ex[slot][4] = want; }
}
__global__ void prim_kernel(int iters)
{
    unsigned long long s = (blockIdx.x * 1024ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    for (int it = 0; it < iters; it++) {
        const int ka = it % 5, kb = (it / 5) % 5, kc = (it / 25) % 5;
        const float a0 = operand(s, ka), a1 = operand(s, ka), b0 = operand(s, kb), b1 = operand(s, kb), c0 = operand(s, kc), c1 = operand(s, kc);
        const float2 A = make_float2(a0, a1), B = make_float2(b0, b1), C = make_float2(c0, c1);
        float2 r;
        r = __fadd2_rn(A, B);
        if (__float_as_uint(r.x) != __float_as_uint(__fadd_rn(a0, b0)) && !(r.x != r.x)) report(0, __float_as_uint(a0), __float_as_uint(b0), 0, __float_as_uint(r.x), __float_as_uint(__fadd_rn(a0, b0)));
        if (__float_as_uint(r.y) != __float_as_uint(__fadd_rn(a1, b1)) && !(r.y != r.y)) report(0, __float_as_uint(a1), __float_as_uint(b1), 1, __float_as_uint(r.y), __float_as_uint(__fadd_rn(a1, b1)));
        r = __fmul2_rn(A, B);
        if (__float_as_uint(r.x) != __float_as_uint(__fmul_rn(a0, b0)) && !(r.x != r.x)) report(1, __float_as_uint(a0), __float_as_uint(b0), 0, __float_as_uint(r.x), __float_as_uint(__fmul_rn(a0, b0)));
        if (__float_as_uint(r.y) != __float_as_uint(__fmul_rn(a1, b1)) && !(r.y != r.y)) report(1, __float_as_uint(a1), __float_as_uint(b1), 1, __float_as_uint(r.y), __float_as_uint(__fmul_rn(a1, b1)));
        r = __ffma2_rn(A, B, C);
        if (__float_as_uint(r.x) != __float_as_uint(__fmaf_rn(a0, b0, c0)) && !(r.x != r.x)) report(2, __float_as_uint(a0), __float_as_uint(b0), __float_as_uint(c0), __float_as_uint(r.x), __float_as_uint(__fmaf_rn(a0, b0, c0)));
        if (__float_as_uint(r.y) != __float_as_uint(__fmaf_rn(a1, b1, c1)) && !(r.y != r.y)) report(2, __float_as_uint(a1), __float_as_uint(b1), __float_as_uint(c1), __float_as_uint(r.y), __float_as_uint(__fmaf_rn(a1, b1, c1)));
        r = __fadd2_rz(A, B);
        if (__float_as_uint(r.x) != __float_as_uint(__fadd_rz(a0, b0)) && !(r.x != r.x)) report(3, __float_as_uint(a0), __float_as_uint(b0), 0, __float_as_uint(r.x), __float_as_uint(__fadd_rz(a0, b0)));
        if (__float_as_uint(r.y) != __float_as_uint(__fadd_rz(a1, b1)) && !(r.y != r.y)) report(3, __float_as_uint(a1), __float_as_uint(b1), 1, __float_as_uint(r.y), __float_as_uint(__fadd_rz(a1, b1)));
        // NaN-ness must agree too
        r = __fmul2_rn(A, B);
        if ((r.x != r.x) != (__fmul_rn(a0, b0) != __fmul_rn(a0, b0))) report(4, __float_as_uint(a0), __float_as_uint(b0), 0, __float_as_uint(r.x), __float_as_uint(__fmul_rn(a0, b0)));
        const unsigned w0 = rnd(s), w1 = rnd(s);
        const int c = it & 3;
        const f2 cv = bytes_to_f2(w0, w1, c);
        if (cv.x != (float)((w0 >> (8 * c)) & 255u) || cv.y != (float)((w1 >> (8 * c)) & 255u)) report(5, w0, w1, c, __float_as_uint(cv.x), 0);
        // the magic truncation against the plain conversion, on values around the clamp range
        const float v0 = (float)((int)(rnd(s) % 4096) - 1024) * 0.00390625f, v1 = (float)((int)(rnd(s) % 4096) - 1024) * 0.00390625f;
        const f2 t = trunc_clamp_magic(mk2(v0, v1), 3.5f);
        const int q0 = min(max(__float2int_rz(v0), 0), 3), q1 = min(max(__float2int_rz(v1), 0), 3);
        if ((int)(__float_as_uint(t.x) & 3u) != q0 || (int)(__float_as_uint(t.y) & 3u) != q1) report(6, __float_as_uint(v0), __float_as_uint(v1), 0, __float_as_uint(t.x), (unsigned)q0);
    }
}

__global__ void bc1_kernel(const u32* tex, int nblocks, u32* out_scalar, u32* out_pair, float one)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= nblocks) return;
    const int ia = 2 * i, ib = (2 * i + 1 < nblocks) ? 2 * i + 1 : 2 * i;
    u32 ta[16], tb[16], oa[4], ob[4], sa[4], sb[4];
    for (int k = 0; k < 16; k++) { ta[k] = tex[ia * 16 + k]; tb[k] = tex[ib * 16 + k]; }
    bc1_bc3_encode_block<true>(ta, sa);
    bc1_bc3_encode_block<true>(tb, sb);
    f2 px[48];
    bc1_bc3_encode_pair<true>(ta, tb, oa, ob, px, 1, splat2(one));
    for (int k = 0; k < 4; k++) { out_scalar[ia * 4 + k] = sa[k]; out_scalar[ib * 4 + k] = sb[k]; out_pair[ia * 4 + k] = oa[k]; out_pair[ib * 4 + k] = ob[k]; }
}

int main()
{
    prim_kernel<<<296, 256>>>(4000);
    cudaDeviceSynchronize();
    unsigned long long bad[8];
    unsigned ex[8][6];
    cudaMemcpyFromSymbol(bad, g_bad, sizeof(bad));
    cudaMemcpyFromSymbol(ex, g_ex, sizeof(ex));
    const char* names[] = {"fadd2_rn", "fmul2_rn", "ffma2_rn", "fadd2_rz", "fmul2 NaN-ness", "bytes_to_f2", "trunc_clamp_magic", ""};
    for (int i = 0; i < 7; i++)
        printf("%-20s mismatches %llu   first: a=%08x b=%08x c=%08x got=%08x want=%08x\n", names[i], bad[i], ex[i][0], ex[i][1], ex[i][2], ex[i][3], ex[i][4]);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));

    // gradient images of several sizes (synth.gradient_rgba8) + a low-variance noise image
    for (int test = 0; test < 3; test++) {
        const int n = test == 0 ? 64 : (test == 1 ? 512 : 256);
        std::vector<u32> tex((size_t)n * n);
        std::vector<u32> img((size_t)n * n);
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                u32 r, g, b, a = 255;
                if (test < 2) { r = (255u * x) / (n - 1); g = (255u * y) / (n - 1); b = (255u * (x + y)) / (2 * n - 2); }
                else { r = 100 + ((x * 7 + y * 3) % 9); g = 90 + ((x * 5 + y * 11) % 7); b = 80 + ((x + y) % 5); a = 200 + (x % 13); }
                img[(size_t)y * n + x] = r | (g << 8) | (b << 16) | (a << 24);
            }
        const int bw = n / 4, nb = bw * bw;
        for (int by = 0; by < bw; by++)
            for (int bx = 0; bx < bw; bx++)
                for (int k = 0; k < 16; k++) tex[((size_t)by * bw + bx) * 16 + k] = img[(size_t)(by * 4 + k / 4) * n + bx * 4 + (k % 4)];
        u32 *d_tex, *d_a, *d_b;
        cudaMalloc(&d_tex, tex.size() * 4); cudaMalloc(&d_a, (size_t)nb * 16); cudaMalloc(&d_b, (size_t)nb * 16);
        cudaMemcpy(d_tex, tex.data(), tex.size() * 4, cudaMemcpyHostToDevice);
        bc1_kernel<<<(nb / 2 + 63) / 64, 64>>>(d_tex, nb, d_a, d_b, 1.0f);
        std::vector<u32> a((size_t)nb * 4), b((size_t)nb * 4);
        cudaMemcpy(a.data(), d_a, (size_t)nb * 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(b.data(), d_b, (size_t)nb * 16, cudaMemcpyDeviceToHost);
        int diff = 0;
        for (int i = 0; i < nb; i++)
            if (memcmp(&a[(size_t)i * 4], &b[(size_t)i * 4], 16)) {
                if (diff < 4) {
                    printf("test %d block %d: scalar %08x %08x %08x %08x  pair %08x %08x %08x %08x\n  texels:", test, i, a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3],
                           b[i * 4], b[i * 4 + 1], b[i * 4 + 2], b[i * 4 + 3]);
                    for (int k = 0; k < 16; k++) printf(" %08x", tex[(size_t)i * 16 + k]);
                    printf("\n");
                }
                diff++;
            }
        printf("BC3 pair vs one-block form, test %d (%dx%d): %d of %d blocks differ (%s)\n", test, n, n, diff, nb, cudaGetErrorString(cudaGetLastError()));
        cudaFree(d_tex); cudaFree(d_a); cudaFree(d_b);
    }
    return 0;
}
