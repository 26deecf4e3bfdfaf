// ubench_pipes.cu -- issue-rate microbenchmark of the instructions the encoders lean on (sm_100a): scalar FFMA / FMUL / FADD
// against the packed two-lane forms FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2 ..., new on Blackwell), the conversions on the XU
// pipe (I2F.U8, F2I), IDP.4A and a shared-memory load.  Each kernel runs 8 independent dependency chains per thread,
// 1024 threads per SM, and reports warp-instructions per clock per SM from clock64() deltas.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_pipes ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, float seed)
{
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed;
    __syncthreads();
    float2 a[kChains];
    unsigned u[kChains];
    for (int i = 0; i < kChains; i++) { a[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i); u[i] = threadIdx.x * 2654435761u + i; }
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) {
            if (OP == 0) a[i].x = __fmaf_rn(a[i].x, m.x, c.x);
            if (OP == 1) a[i] = __ffma2_rn(a[i], m, c);
            if (OP == 2) a[i].x = __fmul_rn(a[i].x, m.x);
            if (OP == 3) a[i] = __fmul2_rn(a[i], m);
            if (OP == 4) a[i].x = __fadd_rn(a[i].x, c.x);
            if (OP == 5) a[i] = __fadd2_rn(a[i], c);
            if (OP == 6) { a[i].x = (float)(u[i] & 255u); u[i] = u[i] * 3u + __float_as_uint(a[i].x); }          // I2F.U8 + IMAD
            if (OP == 7) { u[i] = (unsigned)__float2int_rz(a[i].x) + u[i]; a[i].x = __uint_as_float((u[i] & 0x7fffffu) | 0x3f800000u); }   // F2I + IADD + LOP3
            if (OP == 8) u[i] = __dp4a(u[i], 0x01020304u, u[i]);
            if (OP == 9) { a[i].x = sm[(u[i] + it) & 1023]; u[i] += __float_as_uint(a[i].x); }                 // LDS + IADD (+ address)
            if (OP == 10) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); u[i] = __dp4a(u[i], 0x01020304u, u[i]); }      // FFMA + IDP co-issue
            if (OP == 11) { a[i] = __ffma2_rn(a[i], m, c); u[i] = __dp4a(u[i], 0x01020304u, u[i]); }            // FFMA2 + IDP co-issue
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
    for (int i = 0; i < kChains; i++) s += a[i].x + a[i].y + __uint_as_float(u[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, float* out, long long* cyc, int sms)
{
    bench<OP><<<sms, 1024>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms, 1024>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double winst = 32.0 * kIters * kChains * per_iter;     // warp instructions of interest per SM (32 warps)
    printf("%-28s %8.3f warp-inst/clk/SM  (%.0f cycles, %.3f ms, err=%d)\n", name, winst / avg, avg, ms, (int)cudaGetLastError());
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 256);
    printf("SMs %d; numbers count ONLY the named instruction(s) (loop overhead excluded from the count, included in the time)\n", sms);
    run<0>("FFMA", 1, out, cyc, sms);
    run<1>("FFMA2 (f32x2)", 1, out, cyc, sms);
    run<2>("FMUL", 1, out, cyc, sms);
    run<3>("FMUL2 (f32x2)", 1, out, cyc, sms);
    run<4>("FADD", 1, out, cyc, sms);
    run<5>("FADD2 (f32x2)", 1, out, cyc, sms);
    run<6>("I2F.U8 + IMAD", 2, out, cyc, sms);
    run<7>("F2I + IADD + LOP3", 3, out, cyc, sms);
    run<8>("IDP.4A", 1, out, cyc, sms);
    run<9>("LDS + IADD3 (+addr)", 2, out, cyc, sms);
    run<10>("FFMA + IDP.4A", 2, out, cyc, sms);
    run<11>("FFMA2 + IDP.4A", 2, out, cyc, sms);
    return 0;
}
