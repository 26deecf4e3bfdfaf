// div_sweep.cu -- TEST-ONLY proof by exhaustion for itw_device.cuh's div_by_rcp on ARBITRARY normal floats.
//
// q' = fma(fma(-q, d, x), r, q) with q = x * r and r = RN(1 / d) is used instead of an IEEE division where the operands
// are not confined to a small domain (bc6h.cuh: projection / squared segment length).  Correct rounding of a quotient
// is invariant under scaling x or d by powers of two (no overflow / underflow in the ranges used), so it suffices to
// compare against the IEEE quotient for every pair of significands: x = 1.mx, d = 1.md, mx, md in [0, 2^23) -- 2^46
// pairs (x < d and x >= d both occur, so both result binades are covered).  Built by tests/gpu_checks/build.py; run by
// tests/test_gpu_division.py on the GPU (about a minute).  The product never links this file.
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) sweep_kernel(unsigned md_begin, unsigned md_count, unsigned long long* bad, unsigned* examples)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= md_count) return;
    const unsigned md = md_begin + i;
    const float d = __uint_as_float(0x3F800000u | md);
    const float r = __frcp_rn(d);                                  // correctly rounded reciprocal, as 1.0f / d in the kernels
    unsigned long long local = 0;
    for (unsigned mx = 0; mx < (1u << 23); mx++) {
        const float x = __uint_as_float(0x3F800000u | mx);
        const float q = __fmul_rn(x, r);
        const float fast = __fmaf_rn(__fmaf_rn(-q, d, x), r, q);
        const float ieee = __fdiv_rn(x, d);
        if (__float_as_uint(fast) != __float_as_uint(ieee)) {
            local++;
            const unsigned long long slot = atomicAdd(bad, 1ull);
            if (slot < 16) { examples[2 * slot] = md; examples[2 * slot + 1] = mx; }
        }
    }
    (void)local;
}

// Sweeps divisor significands [md_begin, md_begin + md_count); returns 0 on success.  bad_total / examples are host pointers.
extern "C" int div_sweep(unsigned md_begin, unsigned md_count, unsigned long long* bad_total, unsigned* examples32)
{
    unsigned long long* d_bad = nullptr;
    unsigned* d_ex = nullptr;
    if (cudaMalloc(&d_bad, 8) != cudaSuccess || cudaMalloc(&d_ex, 32 * 4) != cudaSuccess) return -1;
    cudaMemset(d_bad, 0, 8);
    cudaMemset(d_ex, 0, 32 * 4);
    const unsigned chunk = 1u << 18;
    for (unsigned off = 0; off < md_count; off += chunk) {
        const unsigned n = (md_count - off < chunk) ? (md_count - off) : chunk;
        sweep_kernel<<<(n + 255) / 256, 256>>>(md_begin + off, n, d_bad, d_ex);
        if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    }
    cudaMemcpy(bad_total, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(examples32, d_ex, 32 * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    cudaFree(d_ex);
    return 0;
}
