// itw_bcn.cu -- host side of libitw_bcn.so: the reference's C-ABI (include/itw_bcn.h section 1,
// replacing 3rdParty/Intel/Source/ispc_texcomp.cpp:20-435) plus the additive entry points.
//
// There is NO CPU fallback in this library: every encode runs the sm_100a kernels or fails with a
// message retrievable through itw_get_last_error().
//
// Threading: the reference is called concurrently from up to 64 pool threads on disjoint row
// bands (win32Threads.cpp:211-274).  All mutable state here is thread_local (one CUDA stream,
// one set of staging buffers and timing events per calling thread), so calls are re-entrant.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/itw_bcn.h"
#include "bc4_bc5.cuh"
#include "mips.cuh"
#include "decode.cuh"
#include "frontend.cuh"
#include "mips_f16.cuh"
#include "itw_params.h"

using namespace itw;

namespace {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_default_device{-1};      // itw_set_devices() with ONE device: the process-wide default

struct ThreadCtx {
    int device = -1;             // device the resources below belong to
    int wanted_device = -1;      // itw_set_device() request, -1 = keep the current device
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    uint8_t* d_in = nullptr;  size_t d_in_cap = 0;
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    uint8_t* d_mid = nullptr; size_t d_mid_cap = 0;      // itw_encode_pixels: the converted surface between the two kernels
    // band pipeline of the compute-bound encoders (encode_banded): copy-in / copy-out streams and per-band events
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t band_in[4] = {nullptr, nullptr, nullptr, nullptr}, band_done[4] = {nullptr, nullptr, nullptr, nullptr};
    int sm_count = 0;
    std::string err;
    // itw_encode_batch: three copy/compute lanes so that H2D of tile i+1, the kernel of tile i and
    // D2H of tile i-1 overlap (tile streaming)
    struct Lane {
        cudaStream_t stream = nullptr;
        uint8_t* d_in = nullptr;  size_t d_in_cap = 0;
        uint8_t* d_out = nullptr; size_t d_out_cap = 0;
        bool busy = false;
    } lanes[3];
    // deferred mode (itw_begin_deferred / itw_flush): host calls are enqueued on the lanes round-robin and drained by itw_flush
    bool deferred = false;
    unsigned deferred_next = 0;
    int deferred_rc = 0;
    float pool_ms = -1.0f;       // multi-device call: max over the devices of their kernel time
    void release()
    {
        // best effort: at process exit the runtime may already be unloading, errors are ignored
        if (stream) cudaStreamDestroy(stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        cudaFree(d_in); cudaFree(d_out); cudaFree(d_mid);
        d_mid = nullptr; d_mid_cap = 0;
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
        copy_in = copy_out = nullptr;
        for (int i = 0; i < 4; i++) {
            if (band_in[i]) cudaEventDestroy(band_in[i]);
            if (band_done[i]) cudaEventDestroy(band_done[i]);
            band_in[i] = band_done[i] = nullptr;
        }
        for (auto& l : lanes) { if (l.stream) cudaStreamDestroy(l.stream); cudaFree(l.d_in); cudaFree(l.d_out); l = Lane(); }
        stream = nullptr; ev0 = ev1 = nullptr; d_in = d_out = nullptr; d_in_cap = d_out_cap = 0;
        cudaGetLastError();
    }
    ~ThreadCtx() { release(); }            // pool threads come and go (win32Threads.cpp:98-190)
};
thread_local ThreadCtx tls;

int fail(const char* what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof(buf), "%s", what);
    tls.err = buf;
    return -1;
}
#define ITW_CUDA(call)                                             \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail(#call, e_);             \
    } while (0)

int ensure_ctx()
{
    ThreadCtx& c = tls;
    int dev = 0;
    if (c.wanted_device >= 0) {
        ITW_CUDA(cudaSetDevice(c.wanted_device));
        dev = c.wanted_device;
    } else if (g_default_device.load(std::memory_order_relaxed) >= 0) {
        dev = g_default_device.load(std::memory_order_relaxed);
        ITW_CUDA(cudaSetDevice(dev));
    } else {
        ITW_CUDA(cudaGetDevice(&dev));
    }
    if (c.device == dev && c.stream) return 0;
    if (c.stream) {                       // device changed: drop the old resources
        c.release();
        c.device = -1;
        c.timed = false;
    }
    cudaDeviceProp prop;
    ITW_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail("libitw_bcn needs an sm_100a (Blackwell) device");
    c.sm_count = prop.multiProcessorCount;
    ITW_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    ITW_CUDA(cudaEventCreate(&c.ev0));
    ITW_CUDA(cudaEventCreate(&c.ev1));
    c.device = dev;
    return 0;
}
int grow(uint8_t*& p, size_t& cap, size_t need)
{
    if (need <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    ITW_CUDA(cudaMalloc(&p, need));
    cap = need;
    return 0;
}
bool is_device_pointer(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct FormatInfo { int bpb; int texel_bytes; };
bool format_info(int format, FormatInfo& f)
{
    switch (format) {
        case ITW_FORMAT_BC1: case ITW_FORMAT_BC4: f = {8, 4}; return true;
        case ITW_FORMAT_BC3: case ITW_FORMAT_BC5: case ITW_FORMAT_BC7: f = {16, 4}; return true;
        case ITW_FORMAT_BC6H: case ITW_FORMAT_BC6H_SF16: f = {16, 8}; return true;
        default: return false;
    }
}
// 0 = encode it, 1 = nothing to do, -1 = error.  Zero-height bands are legitimate: CompressImageMT hands them to the
// threads beyond height/4 (win32Threads.cpp:217-230) and the reference's loops simply do not execute (kernel.ispc:600).
int check_surface(const rgba_surface* s, const FormatInfo& f)
{
    if (!s) return fail("null surface");
    if (s->height == 0 && s->width >= 0 && (s->width & 3) == 0) return 1;
    if (!s->ptr) return fail("null surface");
    if (s->width <= 0 || s->height < 0 || (s->width & 3) || (s->height & 3))
        return fail("surface width/height must be positive multiples of 4 (ispc_texcomp.h:93-95)");
    if ((long long)s->stride < (long long)s->width * f.texel_bytes) return fail("surface stride smaller than a texel row");
    return 0;
}

// Enqueue the kernel for one device-resident surface on `stream`.
int launch(int format, const SurfaceView& v, uint8_t* d_dst, const void* settings, cudaStream_t stream)
{
    if (format == ITW_FORMAT_BC6H_SF16) format = ITW_FORMAT_BC6H;     // the plug-in encodes SF16 with the unsigned encoder (IntelPlugin.cpp:840-843)
    const long long nblocks = (long long)(v.width >> 2) * (v.height >> 2);
    const bool vec16 = ((reinterpret_cast<uintptr_t>(v.ptr) | (uintptr_t)v.stride) & 15u) == 0;
    const unsigned grid1 = (unsigned)((nblocks + 127) / 128);
    switch (format) {
        case ITW_FORMAT_BC1:
            if (vec16) bc1_bc3_kernel<false, true><<<grid1, 128, 0, stream>>>(v, d_dst);
            else       bc1_bc3_kernel<false, false><<<grid1, 128, 0, stream>>>(v, d_dst);
            break;
        case ITW_FORMAT_BC3:
            if (vec16) bc1_bc3_kernel<true, true><<<grid1, 128, 0, stream>>>(v, d_dst);
            else       bc1_bc3_kernel<true, false><<<grid1, 128, 0, stream>>>(v, d_dst);
            break;
        case ITW_FORMAT_BC4:
            if (vec16) bc4_bc5_kernel<false, true><<<grid1, 128, 0, stream>>>(v, d_dst);
            else       bc4_bc5_kernel<false, false><<<grid1, 128, 0, stream>>>(v, d_dst);
            break;
        case ITW_FORMAT_BC5:
            if (vec16) bc4_bc5_kernel<true, true><<<grid1, 128, 0, stream>>>(v, d_dst);
            else       bc4_bc5_kernel<true, false><<<grid1, 128, 0, stream>>>(v, d_dst);
            break;
        case ITW_FORMAT_BC7: {
            if (!settings) return fail("CompressBlocksBC7: null settings");
            const Bc7Params P = bc7_params_from(*static_cast<const bc7_enc_settings*>(settings));
            if (const char* why = bc7_params_check(P)) return fail(why);
            // one 16-warp CTA per SM, persistent over the rounds; TMA-staged variant when the surface allows it.  A warp takes 16
            // blocks per round when that still gives every SM a full CTA, else 8 (small surfaces / row bands: more warps busy)
            const bool tma = vec16;                                       // 16-byte aligned rows
            const long long cap = (long long)tls.sm_count;
            const int per_warp = (nblocks >= cap * kBc7WarpsPerCta * kBc7Super) ? kBc7Super : kBc7Batch;
            const long long tile = (long long)kBc7WarpsPerCta * per_warp;
            const long long want = (nblocks + tile - 1) / tile;
            const unsigned grid = (unsigned)(want < cap ? want : cap);
            // per-device attribute; cheap enough to set on every (millisecond-scale) launch
            if (tma) {
                ITW_CUDA(cudaFuncSetAttribute(bc7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBc7SmemBytes));
                bc7_kernel<true><<<grid, kBc7WarpsPerCta * 32, kBc7SmemBytes, stream>>>(v, d_dst, P, nblocks, per_warp);
            } else {
                ITW_CUDA(cudaFuncSetAttribute(bc7_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBc7SmemBytes));
                bc7_kernel<false><<<grid, kBc7WarpsPerCta * 32, kBc7SmemBytes, stream>>>(v, d_dst, P, nblocks, per_warp);
            }
            break;
        }
        case ITW_FORMAT_BC6H: {
            if (!settings) return fail("CompressBlocksBC6H: null settings");
            const Bc6Params P = bc6_params_from(*static_cast<const bc6h_enc_settings*>(settings));
            if (const char* why = bc6_params_check(P)) return fail(why);
            const long long want = (nblocks + kBc6TileBlocks - 1) / kBc6TileBlocks;
            const long long cap = (long long)tls.sm_count;
            const unsigned grid = (unsigned)(want < cap ? want : cap);
            if (vec16) {
                ITW_CUDA(cudaFuncSetAttribute(bc6h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBc6SmemBytes));
                bc6h_kernel<true><<<grid, kBc6WarpsPerCta * 32, kBc6SmemBytes, stream>>>(v, d_dst, P, nblocks);
            } else {
                ITW_CUDA(cudaFuncSetAttribute(bc6h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBc6SmemBytes));
                bc6h_kernel<false><<<grid, kBc6WarpsPerCta * 32, kBc6SmemBytes, stream>>>(v, d_dst, P, nblocks);
            }
            break;
        }
        default: return fail("unknown format");
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ITW_CUDA(cudaGetLastError());
    return 0;
}

int encode_many(int format, const rgba_surface* srcs, uint8_t* const* dsts, int count, const void* settings);

// Host surface -> host blocks for the COMPUTE-bound encoders (BC7, BC6H): the surface is cut into four row bands whose
// H2D copies, kernels and D2H copies run on three streams, so that all but the first (small) band's input copy and all
// but the last band's output copy hide behind the kernels.  (For BC1-BC5 the copies ARE the cost -- see encode_single.)
// Bands are whole block rows; the kernels see each band as an independent surface, exactly as the reference's own
// callers do (win32Threads.cpp:217-230), so the output is unchanged.
int encode_banded(int format, const rgba_surface* src, uint8_t* dst, const void* settings, const FormatInfo& f)
{
    ThreadCtx& c = tls;
    if (!c.copy_in) {
        ITW_CUDA(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking));
        ITW_CUDA(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) {
            ITW_CUDA(cudaEventCreateWithFlags(&c.band_in[i], cudaEventDisableTiming));
            ITW_CUDA(cudaEventCreateWithFlags(&c.band_done[i], cudaEventDisableTiming));
        }
    }
    const size_t row_bytes = (size_t)src->width * f.texel_bytes;
    const size_t block_row_bytes = (size_t)(src->width >> 2) * f.bpb;
    const int block_rows = src->height >> 2;
    if (grow(c.d_in, c.d_in_cap, row_bytes * src->height)) return -1;
    if (grow(c.d_out, c.d_out_cap, block_row_bytes * block_rows)) return -1;
    // 1/16 of the rows first (its copy is the only exposed one), then three equal parts
    int first[5];
    first[0] = 0;
    first[1] = block_rows / 16 > 0 ? block_rows / 16 : 1;
    const int rest = block_rows - first[1];
    first[2] = first[1] + rest / 3;
    first[3] = first[1] + (2 * rest) / 3;
    first[4] = block_rows;
    // A launch works in rounds of `wave` blocks (every SM one tile); a band that ends inside a round leaves SMs idle until
    // the band's kernel is over -- four ragged bands cost BC7 4 % of a 4096 x 4096 surface.  When the surface is several waves
    // long the band borders are therefore put on wave borders (rounded down to whole block rows): the bands together then
    // need as many rounds as one launch over the whole surface.
    {
        const long long bw = src->width >> 2;
        const long long wave = (long long)c.sm_count * (format == ITW_FORMAT_BC7 ? kBc7TileBlocks : kBc6TileBlocks);
        const long long total = bw * block_rows, waves = total / wave;
        if (waves >= 4 && wave >= bw) {
            const long long w0 = waves / 16 > 0 ? waves / 16 : 1;                   // waves of the first band
            const long long part = (waves - w0 + 2) / 3;                           // ... of the second and the third
            first[1] = (int)(w0 * wave / bw);
            first[2] = (int)((w0 + part) * wave / bw);
            first[3] = (int)((w0 + 2 * part) * wave / bw);
            if (first[3] > block_rows) first[3] = block_rows;
            if (first[2] > first[3]) first[2] = first[3];
        }
    }
    // Every failure after the first enqueue leaves through `drain`: async copies into the caller's dst and kernels
    // on the three streams must not stay in flight when the call returns -1 (the caller may free src / dst).
    int rc = 0;
    auto step = [&](cudaError_t e, const char* what) { if (rc == 0 && e != cudaSuccess) rc = fail(what, e); return rc == 0; };
    for (int b = 0; b < 4 && rc == 0; b++) {
        const int r0 = first[b], r1 = first[b + 1];
        if (r1 <= r0) continue;
        const size_t rows = (size_t)(r1 - r0) * 4;
        uint8_t* d_band = c.d_in + (size_t)r0 * 4 * row_bytes;
        const uint8_t* h_band = src->ptr + (size_t)r0 * 4 * (size_t)src->stride;
        if ((size_t)src->stride == row_bytes) {
            if (!step(cudaMemcpyAsync(d_band, h_band, row_bytes * rows, cudaMemcpyHostToDevice, c.copy_in), "band H2D")) break;
        } else {
            if (!step(cudaMemcpy2DAsync(d_band, row_bytes, h_band, (size_t)src->stride, row_bytes, rows, cudaMemcpyHostToDevice, c.copy_in), "band H2D")) break;
        }
        if (!step(cudaEventRecord(c.band_in[b], c.copy_in), "cudaEventRecord")) break;
        if (!step(cudaStreamWaitEvent(c.stream, c.band_in[b], 0), "cudaStreamWaitEvent")) break;
        if (b == 0 && !step(cudaEventRecord(c.ev0, c.stream), "cudaEventRecord")) break;
        const SurfaceView v{d_band, src->width, (int)rows, (int)row_bytes};
        uint8_t* d_blocks = c.d_out + (size_t)r0 * block_row_bytes;
        if (launch(format, v, d_blocks, settings, c.stream)) { rc = -1; break; }           // bad settings or a launch failure
        if (!step(cudaEventRecord(c.band_done[b], c.stream), "cudaEventRecord")) break;
        if (!step(cudaStreamWaitEvent(c.copy_out, c.band_done[b], 0), "cudaStreamWaitEvent")) break;
        if (!step(cudaMemcpyAsync(dst + (size_t)r0 * block_row_bytes, d_blocks, (size_t)(r1 - r0) * block_row_bytes, cudaMemcpyDeviceToHost,
                                  c.copy_out), "band D2H")) break;
    }
    if (rc == 0 && step(cudaEventRecord(c.ev1, c.stream), "cudaEventRecord")) c.timed = true;
    // drain (also on error): nothing may stay in flight towards dst, and d_in / d_out must be idle for the next call
    const cudaError_t e0 = cudaStreamSynchronize(c.copy_in), e1 = cudaStreamSynchronize(c.stream), e2 = cudaStreamSynchronize(c.copy_out);
    if (rc == 0) { step(e0, "cudaStreamSynchronize"); step(e1, "cudaStreamSynchronize"); step(e2, "cudaStreamSynchronize"); }
    else cudaGetLastError();
    return rc;
}

// Deferred mode (itw_begin_deferred .. itw_flush): a host -> host encode is enqueued on one of three lanes (stream +
// staging buffers) and the call returns; the H2D copy of call i+1 overlaps the kernel of call i and the D2H copy of call
// i-1.  The caller keeps src and dst valid until itw_flush().  This is what makes the reference's unchanged slice loop
// (IntelPlugin.cpp:851-879: one CompressImageMT per 256 K-texel slice) cheap: no synchronisation per slice.
int encode_deferred(int format, const rgba_surface* src, uint8_t* dst, const void* settings, const FormatInfo& f)
{
    ThreadCtx& c = tls;
    ThreadCtx::Lane& L = c.lanes[c.deferred_next++ % 3];
    if (!L.stream) ITW_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    if (L.busy) {                                      // its buffers are still in flight from three calls ago
        ITW_CUDA(cudaStreamSynchronize(L.stream));
        L.busy = false;
    }
    const size_t row_bytes = (size_t)src->width * f.texel_bytes;
    const size_t out_bytes = (size_t)(src->width >> 2) * (src->height >> 2) * f.bpb;
    if (grow(L.d_in, L.d_in_cap, row_bytes * src->height)) return -1;
    if (grow(L.d_out, L.d_out_cap, out_bytes)) return -1;
    if ((size_t)src->stride == row_bytes)
        ITW_CUDA(cudaMemcpyAsync(L.d_in, src->ptr, row_bytes * src->height, cudaMemcpyHostToDevice, L.stream));
    else
        ITW_CUDA(cudaMemcpy2DAsync(L.d_in, row_bytes, src->ptr, (size_t)src->stride, row_bytes, (size_t)src->height, cudaMemcpyHostToDevice, L.stream));
    L.busy = true;
    const SurfaceView v{L.d_in, src->width, src->height, (int)row_bytes};
    if (launch(format, v, L.d_out, settings, L.stream)) return -1;
    ITW_CUDA(cudaMemcpyAsync(dst, L.d_out, out_bytes, cudaMemcpyDeviceToHost, L.stream));
    c.timed = false;
    return 0;
}
int drain_lanes(ThreadCtx& c)
{
    int rc = 0;
    for (auto& L : c.lanes)
        if (L.busy) {
            if (cudaStreamSynchronize(L.stream) != cudaSuccess && rc == 0) rc = fail("cudaStreamSynchronize");
            L.busy = false;
        }
    return rc;
}

// One device: src and dst may each be host or device memory.
int encode_single(int format, const rgba_surface* src, uint8_t* dst, const void* settings, const FormatInfo& f)
{
    if (ensure_ctx()) return -1;
    ThreadCtx& c = tls;
    // (Splitting one large host surface into pipelined row bands was measured and rejected for BC1-BC5: the H2D copy is
    //  >90 % of the BC1 end-to-end time, so overlap buys <0.1 ms while eight smaller copies cost 0.3 ms.)
    const size_t row_bytes = (size_t)src->width * f.texel_bytes;
    const size_t out_bytes = (size_t)(src->width >> 2) * (src->height >> 2) * f.bpb;
    const bool src_dev = is_device_pointer(src->ptr), dst_dev = is_device_pointer(dst);
    if (!src_dev && !dst_dev && c.deferred) {
        const int rc = encode_deferred(format, src, dst, settings, f);
        if (rc && c.deferred_rc == 0) c.deferred_rc = rc;
        return rc;
    }
    if (!src_dev && !dst_dev && (format == ITW_FORMAT_BC7 || format == ITW_FORMAT_BC6H) && src->height >= 256)
        return encode_banded(format, src, dst, settings, f);

    // Device-resident operands were produced by the caller's own streams.  The legacy default stream orders
    // after every blocking stream (torch's default stream included), which gives the synchronous call the
    // semantics a caller expects; pure host calls use this thread's private non-blocking stream.
    cudaStream_t s = (src_dev || dst_dev) ? cudaStreamLegacy : c.stream;
    SurfaceView v{src->ptr, src->width, src->height, src->stride};
    if (!src_dev) {                      // H2D of the tightly packed rows (pinned sources copy at PCIe rate)
        if (grow(c.d_in, c.d_in_cap, row_bytes * src->height)) return -1;
        if ((size_t)src->stride == row_bytes)   // tightly packed rows: one linear copy (full PCIe rate from pinned memory)
            ITW_CUDA(cudaMemcpyAsync(c.d_in, src->ptr, row_bytes * src->height, cudaMemcpyHostToDevice, s));
        else
            ITW_CUDA(cudaMemcpy2DAsync(c.d_in, row_bytes, src->ptr, (size_t)src->stride, row_bytes, (size_t)src->height,
                                       cudaMemcpyHostToDevice, s));
        v.ptr = c.d_in;
        v.stride = (int)row_bytes;
    }
    uint8_t* d_dst = dst;
    const bool dst_ok = dst_dev && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    if (!dst_ok) {
        if (grow(c.d_out, c.d_out_cap, out_bytes)) return -1;
        d_dst = c.d_out;
    }
    // From here on work is in flight on `s`: every failure still leaves through the synchronisation below (the caller may free
    // src / dst as soon as the call returns), like encode_banded.
    int rc = 0;
    auto step = [&](cudaError_t e, const char* what) { if (rc == 0 && e != cudaSuccess) rc = fail(what, e); return rc == 0; };
    step(cudaEventRecord(c.ev0, s), "cudaEventRecord");
    if (rc == 0 && launch(format, v, d_dst, settings, s)) rc = -1;                       // bad settings or a launch failure
    if (rc == 0 && step(cudaEventRecord(c.ev1, s), "cudaEventRecord")) c.timed = true;
    if (rc == 0 && !dst_ok)
        step(cudaMemcpyAsync(dst, d_dst, out_bytes, dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
    const cudaError_t e = cudaStreamSynchronize(s);
    if (rc == 0) step(e, "cudaStreamSynchronize");
    else cudaGetLastError();
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-device engine (the B200 analogue of the reference's thread pool, win32Threads.cpp:98-274).
// itw_set_devices() creates one worker thread per selected GPU; every worker owns that device's ThreadCtx (streams,
// staging buffers).  A host -> host call is then cut into one row band per device (whole block rows, like the
// reference's per-thread bands, win32Threads.cpp:217-230 -- bands are independent surfaces, so the bytes are those
// of a single-device encode), every worker runs its band through the single-device path above (H2D, kernels and D2H
// on its own PCIe link) and writes straight into the caller's dst.  One process, one call, no collective.
// ---------------------------------------------------------------------------------------------------------------
class DevicePool {
public:
    explicit DevicePool(const std::vector<int>& devices)
    {
        for (int d : devices) workers_.emplace_back(new Worker(d));
    }
    ~DevicePool()
    {
        for (auto& w : workers_) {
            { std::lock_guard<std::mutex> g(w->m); w->quit = true; }
            w->cv.notify_all();
            w->thread.join();
        }
    }
    int size() const { return (int)workers_.size(); }
    int device(int i) const { return workers_[i]->device; }
    // Run jobs[i] on worker i (empty functions are skipped); returns 0 or -1 with the first error text in `err`
    // and the largest kernel time in `ms`.
    int run(const std::vector<std::function<int()>>& jobs, std::string& err, float& ms)
    {
        std::lock_guard<std::mutex> serial(run_mutex_);           // concurrent callers take turns
        for (size_t i = 0; i < workers_.size() && i < jobs.size(); i++) {
            Worker& w = *workers_[i];
            if (!jobs[i]) continue;
            { std::lock_guard<std::mutex> g(w.m); w.job = &jobs[i]; w.done = false; }
            w.cv.notify_all();
        }
        int rc = 0;
        ms = -1.0f;
        for (size_t i = 0; i < workers_.size() && i < jobs.size(); i++) {
            Worker& w = *workers_[i];
            if (!jobs[i]) continue;
            std::unique_lock<std::mutex> g(w.m);
            w.cv.wait(g, [&] { return w.done; });
            if (w.rc != 0 && rc == 0) { rc = -1; err = w.err; }
            if (w.ms > ms) ms = w.ms;
        }
        return rc;
    }

private:
    struct Worker {
        int device;
        std::mutex m;
        std::condition_variable cv;
        const std::function<int()>* job = nullptr;
        bool done = true, quit = false;
        int rc = 0;
        float ms = -1.0f;
        std::string err;
        std::thread thread;
        explicit Worker(int dev) : device(dev), thread([this] { loop(); }) {}
        void loop()
        {
            tls.wanted_device = device;                           // this thread's ThreadCtx lives on `device`
            for (;;) {
                const std::function<int()>* j;
                {
                    std::unique_lock<std::mutex> g(m);
                    cv.wait(g, [&] { return quit || job; });
                    if (quit) break;
                    j = job;
                }
                tls.err.clear();
                const int r = (*j)();
                float t = -1.0f;
                if (tls.timed && cudaEventElapsedTime(&t, tls.ev0, tls.ev1) != cudaSuccess) { cudaGetLastError(); t = -1.0f; }
                {
                    std::lock_guard<std::mutex> g(m);
                    rc = r; err = tls.err; ms = t; job = nullptr; done = true;
                }
                cv.notify_all();
            }
            tls.release();                                        // free this device's buffers on the owning thread
        }
    };
    std::vector<std::unique_ptr<Worker>> workers_;
    std::mutex run_mutex_;
};
std::mutex g_pool_mutex;
std::shared_ptr<DevicePool> g_pool;                               // null = single-device behaviour
std::shared_ptr<DevicePool> current_pool()
{
    std::lock_guard<std::mutex> g(g_pool_mutex);
    return g_pool;
}
// smallest band worth a device of its own: below this the per-device launch + copy latency outweighs the split
constexpr long long kPoolMinBlocksPerDevice = 4096;

int encode_fanout(DevicePool& pool, int format, const rgba_surface* src, uint8_t* dst, const void* settings, const FormatInfo& f)
{
    const int block_rows = src->height >> 2;
    const long long blocks_per_row = src->width >> 2;
    int n = pool.size();
    while (n > 1 && ((long long)block_rows * blocks_per_row) / n < kPoolMinBlocksPerDevice) n--;
    if (n > block_rows) n = block_rows;
    const size_t block_row_bytes = (size_t)blocks_per_row * f.bpb;
    std::vector<rgba_surface> bands((size_t)n);
    std::vector<std::function<int()>> jobs((size_t)pool.size());
    for (int i = 0; i < n; i++) {
        const int r0 = (int)((long long)block_rows * i / n), r1 = (int)((long long)block_rows * (i + 1) / n);
        bands[i] = rgba_surface{src->ptr + (size_t)r0 * 4 * (size_t)src->stride, src->width, (r1 - r0) * 4, src->stride};
        uint8_t* out = dst + (size_t)r0 * block_row_bytes;
        const rgba_surface* band = &bands[i];
        if (r1 > r0) jobs[i] = [=, &f]() { return encode_single(format, band, out, settings, f); };
    }
    std::string err;
    float ms = -1.0f;
    const int rc = pool.run(jobs, err, ms);
    tls.timed = false;
    tls.pool_ms = ms;
    if (rc) return fail(err.c_str());
    return 0;
}

// The CompressBlocks* path.
int encode_any(int format, const rgba_surface* src, uint8_t* dst, const void* settings)
{
    tls.err.clear();
    tls.pool_ms = -1.0f;
    if (format == ITW_FORMAT_BC6H_SF16) format = ITW_FORMAT_BC6H;      // one encoder for both (IntelPlugin.cpp:840-843)
    FormatInfo f;
    if (!format_info(format, f)) return fail("unknown format");
    const int chk = check_surface(src, f);
    if (chk) return chk < 0 ? -1 : 0;
    if (!dst) return fail("null dst");
    if (std::shared_ptr<DevicePool> pool = current_pool()) {
        if (pool->size() > 1 && !tls.deferred && (long long)(src->width >> 2) * (src->height >> 2) >= 2 * kPoolMinBlocksPerDevice &&
            !is_device_pointer(src->ptr) && !is_device_pointer(dst))
            return encode_fanout(*pool, format, src, dst, settings, f);
    }
    return encode_single(format, src, dst, settings, f);
}


// ---- decoders (decode.cuh): blocks -> RGBA8 (BC1/3/4/5/7) or RGBA16F (BC6H) surface; either side host or device ----
template <int kFormat>
void launch_decode_as(bool vec16, unsigned grid, cudaStream_t s, const uint8_t* blocks, uint8_t* dst, int w, int h, long long stride)
{
    if (vec16) decode_kernel<kFormat, true><<<grid, 128, 0, s>>>(blocks, dst, w, h, stride);
    else       decode_kernel<kFormat, false><<<grid, 128, 0, s>>>(blocks, dst, w, h, stride);
}
int launch_decode(int format, const uint8_t* d_blocks, uint8_t* d_dst, int w, int h, long long stride, cudaStream_t s)
{
    const long long nblocks = (long long)(w >> 2) * (h >> 2);
    const unsigned grid = (unsigned)((nblocks + 127) / 128);
    const bool vec16 = ((reinterpret_cast<uintptr_t>(d_dst) | (uintptr_t)stride) & 15u) == 0;
    switch (format) {
        case ITW_FORMAT_BC1: launch_decode_as<ITW_FORMAT_BC1>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC3: launch_decode_as<ITW_FORMAT_BC3>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC4: launch_decode_as<ITW_FORMAT_BC4>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC5: launch_decode_as<ITW_FORMAT_BC5>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC6H: launch_decode_as<ITW_FORMAT_BC6H>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC6H_SF16: launch_decode_as<ITW_FORMAT_BC6H_SF16>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        case ITW_FORMAT_BC7: launch_decode_as<ITW_FORMAT_BC7>(vec16, grid, s, d_blocks, d_dst, w, h, stride); break;
        default: return fail("unknown format");
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ITW_CUDA(cudaGetLastError());
    return 0;
}
int decode_any(int format, const uint8_t* blocks, const rgba_surface* dst)
{
    tls.err.clear();
    FormatInfo f;
    if (!format_info(format, f)) return fail("unknown format");
    if (const int chk = check_surface(dst, f)) return chk < 0 ? -1 : 0;
    if (!blocks) return fail("null blocks");
    if (ensure_ctx()) return -1;
    ThreadCtx& c = tls;
    const size_t row_bytes = (size_t)dst->width * f.texel_bytes;
    const size_t in_bytes = (size_t)(dst->width >> 2) * (dst->height >> 2) * f.bpb;
    const bool src_dev = is_device_pointer(blocks), dst_dev = is_device_pointer(dst->ptr);
    cudaStream_t s = (src_dev || dst_dev) ? cudaStreamLegacy : c.stream;       // same ordering rule as encode_any
    const uint8_t* d_blocks = blocks;
    if (!src_dev || (reinterpret_cast<uintptr_t>(blocks) & 15u)) {
        if (grow(c.d_in, c.d_in_cap, in_bytes)) return -1;
        ITW_CUDA(cudaMemcpyAsync(c.d_in, blocks, in_bytes, src_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        d_blocks = c.d_in;
    }
    uint8_t* d_dst = dst->ptr;
    long long stride = dst->stride;
    const bool dst_ok = dst_dev && (((uintptr_t)dst->ptr | (uintptr_t)dst->stride) & 3u) == 0;
    if (!dst_ok) {
        if (grow(c.d_out, c.d_out_cap, row_bytes * dst->height)) return -1;
        d_dst = c.d_out;
        stride = (long long)row_bytes;
    }
    int rc = 0;                                        // work is in flight from here on: every exit synchronises (see encode_single)
    auto step = [&](cudaError_t e, const char* what) { if (rc == 0 && e != cudaSuccess) rc = fail(what, e); return rc == 0; };
    step(cudaEventRecord(c.ev0, s), "cudaEventRecord");
    if (rc == 0 && launch_decode(format, d_blocks, d_dst, dst->width, dst->height, stride, s)) rc = -1;
    if (rc == 0 && step(cudaEventRecord(c.ev1, s), "cudaEventRecord")) c.timed = true;
    if (rc == 0 && !dst_ok)
        step(cudaMemcpy2DAsync(dst->ptr, (size_t)dst->stride, d_dst, row_bytes, row_bytes, (size_t)dst->height,
                               dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s), "cudaMemcpy2DAsync");
    const cudaError_t e = cudaStreamSynchronize(s);
    if (rc == 0) step(e, "cudaStreamSynchronize");
    else cudaGetLastError();
    return rc;
}


// ---- pixel-format front end (frontend.cuh) ----
int front_params(FrontParams& P, int format, const itw_pixel_source* src, uint32_t flags)
{
    FormatInfo f;
    if (!format_info(format, f)) return fail("unknown format");
    if (!src || !src->data) return fail("null pixel source");
    if (src->width <= 0 || src->height <= 0) return fail("pixel source: width/height must be positive");
    if (src->planes < 1 || src->planes > 4) return fail("pixel source: planes must be 1..4");
    if (src->depth != 8 && src->depth != 16 && src->depth != 32) return fail("pixel source: depth must be 8, 16 or 32");
    if (flags & ~31u) return fail("unknown front-end flag");
    const int esize = src->depth / 8;
    const long long tight = (long long)src->width * src->planes * esize;
    const long long row_bytes = src->row_bytes ? src->row_bytes : tight;
    if (row_bytes < tight) return fail("pixel source: row_bytes smaller than a texel row");
    if ((reinterpret_cast<uintptr_t>(src->data) | (uintptr_t)row_bytes) & (uintptr_t)(esize - 1)) return fail("pixel source: misaligned elements");
    const int family = (format == ITW_FORMAT_BC6H || format == ITW_FORMAT_BC6H_SF16) ? 2 : ((format == ITW_FORMAT_BC4 || format == ITW_FORMAT_BC5) ? 1 : 0);
    if (flags & ITW_FRONT_HAS_ALPHA) {
        // the converters read plane 3 (IntelPlugin.cpp:310, :758 ...); the 32-bit HDR one reads plane 2 (:361)
        const int need = (family == 2 && src->depth == 32) ? 3 : 4;
        if (src->planes < need) return fail("ITW_FRONT_HAS_ALPHA needs an alpha plane in the source");
    }
    P = FrontParams{static_cast<const uint8_t*>(src->data), src->width, src->height, src->planes, src->depth, row_bytes, family, flags};
    return 0;
}
template <int kDepth>
void launch_front_x4(const FrontParams& P, dim3 grid, cudaStream_t s, uint8_t* d_dst, int dw, int dh, long long dstride)
{
    switch (P.planes) {
        case 1: front_kernel_x4<kDepth, 1><<<grid, 256, 0, s>>>(P, d_dst, dw, dh, dstride); break;
        case 2: front_kernel_x4<kDepth, 2><<<grid, 256, 0, s>>>(P, d_dst, dw, dh, dstride); break;
        case 3: front_kernel_x4<kDepth, 3><<<grid, 256, 0, s>>>(P, d_dst, dw, dh, dstride); break;
        default: front_kernel_x4<kDepth, 4><<<grid, 256, 0, s>>>(P, d_dst, dw, dh, dstride); break;
    }
}
int launch_front(const FrontParams& P, uint8_t* d_dst, int dw, int dh, long long dstride, cudaStream_t s)
{
    if (dh > 65535) return fail("front end: height above 65535");
    const bool quads = (dw & 3) == 0 && ((reinterpret_cast<uintptr_t>(d_dst) | (uintptr_t)dstride) & 15u) == 0 &&
                       ((reinterpret_cast<uintptr_t>(P.data) | (uintptr_t)P.row_bytes) & 3u) == 0;
    if (quads) {
        dim3 grid((unsigned)(((dw + 7) / 8 + 255) / 256), (unsigned)dh);      // two quads per thread
        if (P.depth == 8) launch_front_x4<8>(P, grid, s, d_dst, dw, dh, dstride);
        else if (P.depth == 16) launch_front_x4<16>(P, grid, s, d_dst, dw, dh, dstride);
        else launch_front_x4<32>(P, grid, s, d_dst, dw, dh, dstride);
    } else {
        dim3 grid((unsigned)((dw + 255) / 256), (unsigned)dh);
        front_kernel<<<grid, 256, 0, s>>>(P, d_dst, dw, dh, dstride);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ITW_CUDA(cudaGetLastError());
    return 0;
}
// Stage a host pixel source in d_in (tight rows); P.data / P.row_bytes are redirected.
int stage_pixels(ThreadCtx& c, FrontParams& P, cudaStream_t s)
{
    const size_t tight = (size_t)P.width * P.planes * (P.depth / 8);
    if (grow(c.d_in, c.d_in_cap, tight * P.height)) return -1;
    ITW_CUDA(cudaMemcpy2DAsync(c.d_in, tight, P.data, (size_t)P.row_bytes, tight, (size_t)P.height, cudaMemcpyHostToDevice, s));
    P.data = c.d_in;
    P.row_bytes = (long long)tight;
    return 0;
}
int convert_any(int format, const itw_pixel_source* src, uint32_t flags, const rgba_surface* dst)
{
    tls.err.clear();
    FrontParams P;
    if (front_params(P, format, src, flags)) return -1;
    if (!dst || !dst->ptr) return fail("null surface");
    const int texel = (P.family == 2) ? 8 : 4;
    const int pw = (P.width + 3) & ~3, ph = (P.height + 3) & ~3;
    if (!((dst->width == P.width && dst->height == P.height) || (dst->width == pw && dst->height == ph)))
        return fail("itw_convert_pixels: dst must have the source size or that size rounded up to multiples of 4");
    if ((long long)dst->stride < (long long)dst->width * texel) return fail("surface stride smaller than a texel row");
    if (ensure_ctx()) return -1;
    ThreadCtx& c = tls;
    const bool src_dev = is_device_pointer(P.data), dst_dev = is_device_pointer(dst->ptr);
    cudaStream_t s = (src_dev || dst_dev) ? cudaStreamLegacy : c.stream;
    if (!src_dev && stage_pixels(c, P, s)) return -1;
    uint8_t* d_dst = dst->ptr;
    long long stride = dst->stride;
    const bool dst_ok = dst_dev && (((uintptr_t)dst->ptr | (uintptr_t)dst->stride) & (uintptr_t)(texel - 1)) == 0;
    const size_t row_bytes = (size_t)dst->width * texel;
    if (!dst_ok) {
        if (grow(c.d_out, c.d_out_cap, row_bytes * dst->height)) return -1;
        d_dst = c.d_out;
        stride = (long long)row_bytes;
    }
    ITW_CUDA(cudaEventRecord(c.ev0, s));
    if (launch_front(P, d_dst, dst->width, dst->height, stride, s)) return -1;
    ITW_CUDA(cudaEventRecord(c.ev1, s));
    c.timed = true;
    if (!dst_ok)
        ITW_CUDA(cudaMemcpy2DAsync(dst->ptr, (size_t)dst->stride, d_dst, row_bytes, row_bytes, (size_t)dst->height,
                                   dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    ITW_CUDA(cudaStreamSynchronize(s));
    return 0;
}
// convert + pad + encode; the intermediate surface never leaves the GPU
int encode_pixels_any(int format, const itw_pixel_source* src, uint32_t flags, const void* settings, uint8_t* dst)
{
    tls.err.clear();
    FrontParams P;
    if (front_params(P, format, src, flags)) return -1;
    if (!dst) return fail("null dst");
    FormatInfo f;
    format_info(format, f);
    if (ensure_ctx()) return -1;
    ThreadCtx& c = tls;
    const int pw = (P.width + 3) & ~3, ph = (P.height + 3) & ~3;
    const size_t mid_row = (size_t)pw * f.texel_bytes, out_bytes = (size_t)(pw >> 2) * (ph >> 2) * f.bpb;
    const bool src_dev = is_device_pointer(P.data), dst_dev = is_device_pointer(dst);
    cudaStream_t s = (src_dev || dst_dev) ? cudaStreamLegacy : c.stream;
    if (!src_dev && stage_pixels(c, P, s)) return -1;
    if (grow(c.d_mid, c.d_mid_cap, mid_row * ph)) return -1;
    uint8_t* d_dst = dst;
    const bool dst_ok = dst_dev && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    if (!dst_ok) {
        if (grow(c.d_out, c.d_out_cap, out_bytes)) return -1;
        d_dst = c.d_out;
    }
    ITW_CUDA(cudaEventRecord(c.ev0, s));
    if (launch_front(P, c.d_mid, pw, ph, (long long)mid_row, s)) return -1;
    const SurfaceView v{c.d_mid, pw, ph, (int)mid_row};
    if (launch(format, v, d_dst, settings, s)) return -1;
    ITW_CUDA(cudaEventRecord(c.ev1, s));
    c.timed = true;
    if (!dst_ok)
        ITW_CUDA(cudaMemcpyAsync(dst, d_dst, out_bytes, dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    ITW_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace

extern "C" {

// ---- profiles (ispc_texcomp.cpp:20-410) ----
void GetProfile_ultrafast(bc7_enc_settings* s) { bc7_fill_profile(s, 0); }
void GetProfile_veryfast(bc7_enc_settings* s) { bc7_fill_profile(s, 1); }
void GetProfile_fast(bc7_enc_settings* s) { bc7_fill_profile(s, 2); }
void GetProfile_basic(bc7_enc_settings* s) { bc7_fill_profile(s, 3); }
void GetProfile_slow(bc7_enc_settings* s) { bc7_fill_profile(s, 4); }
void GetProfile_alpha_ultrafast(bc7_enc_settings* s) { bc7_fill_profile(s, 5); }
void GetProfile_alpha_veryfast(bc7_enc_settings* s) { bc7_fill_profile(s, 6); }
void GetProfile_alpha_fast(bc7_enc_settings* s) { bc7_fill_profile(s, 7); }
void GetProfile_alpha_basic(bc7_enc_settings* s) { bc7_fill_profile(s, 8); }
void GetProfile_alpha_slow(bc7_enc_settings* s) { bc7_fill_profile(s, 9); }
void GetProfile_bc6h_veryfast(bc6h_enc_settings* s) { bc6_fill_profile(s, 0); }
void GetProfile_bc6h_fast(bc6h_enc_settings* s) { bc6_fill_profile(s, 1); }
void GetProfile_bc6h_basic(bc6h_enc_settings* s) { bc6_fill_profile(s, 2); }
void GetProfile_bc6h_slow(bc6h_enc_settings* s) { bc6_fill_profile(s, 3); }
void GetProfile_bc6h_veryslow(bc6h_enc_settings* s) { bc6_fill_profile(s, 4); }

// ---- the reference's encode entry points (ispc_texcomp.cpp:417-435) ----
void CompressBlocksBC1(const rgba_surface* src, uint8_t* dst) { encode_any(ITW_FORMAT_BC1, src, dst, nullptr); }
void CompressBlocksBC3(const rgba_surface* src, uint8_t* dst) { encode_any(ITW_FORMAT_BC3, src, dst, nullptr); }
void CompressBlocksBC6H(const rgba_surface* src, uint8_t* dst, bc6h_enc_settings* settings) { encode_any(ITW_FORMAT_BC6H, src, dst, settings); }
void CompressBlocksBC7(const rgba_surface* src, uint8_t* dst, bc7_enc_settings* settings) { encode_any(ITW_FORMAT_BC7, src, dst, settings); }

// ---- additive ----
void CompressBlocksBC4(const rgba_surface* src, uint8_t* dst) { encode_any(ITW_FORMAT_BC4, src, dst, nullptr); }
void CompressBlocksBC5(const rgba_surface* src, uint8_t* dst) { encode_any(ITW_FORMAT_BC5, src, dst, nullptr); }

int itw_bytes_per_block(int format)
{
    FormatInfo f;
    return format_info(format, f) ? f.bpb : 0;
}

int itw_encode_device(int format, const rgba_surface* src, uint8_t* dst, const void* settings, void* cuda_stream)
{
    tls.err.clear();
    FormatInfo f;
    if (!format_info(format, f)) return fail("unknown format");
    if (const int chk = check_surface(src, f)) return chk < 0 ? -1 : 0;
    if (!dst) return fail("null dst");
    if (reinterpret_cast<uintptr_t>(dst) & 15u) return fail("itw_encode_device: dst must be 16-byte aligned");
    if (ensure_ctx()) return -1;
    tls.timed = false;
    SurfaceView v{src->ptr, src->width, src->height, src->stride};
    return launch(format, v, dst, settings, static_cast<cudaStream_t>(cuda_stream));
}

int itw_decode(int format, const uint8_t* blocks, const rgba_surface* dst) { return decode_any(format, blocks, dst); }
int itw_convert_pixels(int format, const itw_pixel_source* src, uint32_t flags, const rgba_surface* dst) { return convert_any(format, src, flags, dst); }
int itw_encode_pixels(int format, const itw_pixel_source* src, uint32_t flags, const void* settings, uint8_t* dst_blocks)
{
    return encode_pixels_any(format, src, flags, settings, dst_blocks);
}

int itw_encode_batch(int format, const rgba_surface* srcs, uint8_t* const* dsts, int count, const void* settings)
{
    tls.err.clear();
    return encode_many(format, srcs, dsts, count, settings);
}

}  // extern "C"

namespace {
// Tiles i = first, first+step, ... of the batch on THIS thread's device, pipelined over the three lanes.
int encode_many_single(int format, const rgba_surface* srcs, uint8_t* const* dsts, int count, const void* settings, int first, int step)
{
    FormatInfo f;
    if (!format_info(format, f)) return fail("unknown format");
    if (ensure_ctx()) return -1;
    ThreadCtx& c = tls;
    int rc = drain_lanes(c);
    // Device-resident operands were produced by the caller's blocking streams: make every lane wait for what is already
    // enqueued on the legacy default stream (the ordering rule of CompressBlocks*, see itw_bcn.h), once per call.
    bool any_dev = false;
    for (int i = first; i < count && !any_dev; i += step) any_dev = is_device_pointer(srcs[i].ptr) || is_device_pointer(dsts[i]);
    if (any_dev && rc == 0) {
        if (cudaEventRecord(c.ev0, cudaStreamLegacy) != cudaSuccess) rc = fail("cudaEventRecord");
        for (auto& L : c.lanes) {
            if (rc) break;
            if (!L.stream && cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail("cudaStreamCreate"); break; }
            if (cudaStreamWaitEvent(L.stream, c.ev0, 0) != cudaSuccess) rc = fail("cudaStreamWaitEvent");
        }
    }
    unsigned turn = 0;
    for (int i = first; i < count && rc == 0; i += step) {
        const rgba_surface* src = &srcs[i];
        uint8_t* dst = dsts[i];
        const int chk = check_surface(src, f);
        if (chk < 0) { rc = -1; break; }
        if (chk > 0) continue;
        if (!dst) { rc = fail("null dst"); break; }
        ThreadCtx::Lane& L = c.lanes[turn++ % 3];
        if (!L.stream && cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail("cudaStreamCreate"); break; }
        if (L.busy) {                                  // its buffers are still in flight from three tiles ago
            if (cudaStreamSynchronize(L.stream) != cudaSuccess) { rc = fail("cudaStreamSynchronize"); break; }
            L.busy = false;
        }
        const size_t row_bytes = (size_t)src->width * f.texel_bytes;
        const size_t out_bytes = (size_t)(src->width >> 2) * (src->height >> 2) * f.bpb;
        const bool src_dev = is_device_pointer(src->ptr), dst_dev = is_device_pointer(dst);
        SurfaceView v{src->ptr, src->width, src->height, src->stride};
        cudaError_t e = cudaSuccess;
        if (!src_dev) {
            if (grow(L.d_in, L.d_in_cap, row_bytes * src->height)) { rc = -1; break; }
            if ((size_t)src->stride == row_bytes)
                e = cudaMemcpyAsync(L.d_in, src->ptr, row_bytes * src->height, cudaMemcpyHostToDevice, L.stream);
            else
                e = cudaMemcpy2DAsync(L.d_in, row_bytes, src->ptr, (size_t)src->stride, row_bytes, (size_t)src->height,
                                      cudaMemcpyHostToDevice, L.stream);
            if (e != cudaSuccess) { rc = fail("itw_encode_batch H2D", e); break; }
            v.ptr = L.d_in;
            v.stride = (int)row_bytes;
        }
        uint8_t* d_dst = dst;
        const bool dst_ok = dst_dev && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
        if (!dst_ok) {
            if (grow(L.d_out, L.d_out_cap, out_bytes)) { rc = -1; break; }
            d_dst = L.d_out;
        }
        L.busy = true;
        if (launch(format, v, d_dst, settings, L.stream)) { rc = -1; break; }
        if (!dst_ok) {
            e = cudaMemcpyAsync(dst, d_dst, out_bytes, dst_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, L.stream);
            if (e != cudaSuccess) { rc = fail("itw_encode_batch D2H", e); break; }
        }
    }
    if (drain_lanes(c) && rc == 0) rc = -1;               // always drain, also on error
    tls.timed = false;
    return rc;
}
int encode_many(int format, const rgba_surface* srcs, uint8_t* const* dsts, int count, const void* settings)
{
    tls.pool_ms = -1.0f;
    if (count < 0 || (count > 0 && (!srcs || !dsts))) return fail("itw_encode_batch: bad arguments");
    // tile stream over several devices: host tiles are dealt round-robin, every device runs its
    // own three-lane pipeline; no collective
    if (std::shared_ptr<DevicePool> pool = current_pool()) {
        bool host_only = count >= 2 && pool->size() > 1;
        for (int i = 0; i < count && host_only; i++)
            host_only = srcs[i].ptr && dsts[i] && !is_device_pointer(srcs[i].ptr) && !is_device_pointer(dsts[i]);
        if (host_only) {
            const int n = pool->size() < count ? pool->size() : count;
            std::vector<std::function<int()>> jobs((size_t)pool->size());
            for (int k = 0; k < n; k++) jobs[k] = [=]() { return encode_many_single(format, srcs, dsts, count, settings, k, n); };
            std::string err;
            float ms = -1.0f;
            const int rc = pool->run(jobs, err, ms);
            tls.timed = false;
            if (rc) return fail(err.c_str());
            return 0;
        }
    }
    return encode_many_single(format, srcs, dsts, count, settings, 0, 1);
}
}  // namespace

extern "C" {

int itw_set_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return fail("itw_set_device: no such device");
    tls.wanted_device = device;
    return 0;
}

int itw_set_devices(const int* devices, int count)
{
    tls.err.clear();
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail("itw_set_devices: no CUDA device"); }
    if (count < 0 || (count > 0 && !devices)) return fail("itw_set_devices: bad arguments");
    std::vector<int> list(devices, devices + count);
    for (size_t i = 0; i < list.size(); i++) {
        if (list[i] < 0 || list[i] >= n) return fail("itw_set_devices: no such device");
        for (size_t j = 0; j < i; j++) if (list[j] == list[i]) return fail("itw_set_devices: duplicate device");
    }
    std::shared_ptr<DevicePool> fresh, old;
    if (list.size() >= 2) fresh = std::make_shared<DevicePool>(list);
    {
        std::lock_guard<std::mutex> g(g_pool_mutex);
        old.swap(g_pool);
        g_pool = fresh;
        g_default_device.store(list.size() == 1 ? list[0] : -1, std::memory_order_relaxed);
    }
    old.reset();                                          // joins the previous workers (after in-flight calls released it)
    return 0;
}
int itw_get_devices(int* devices, int capacity)
{
    std::shared_ptr<DevicePool> pool = current_pool();
    if (!pool) {
        const int d = g_default_device.load(std::memory_order_relaxed);
        if (d >= 0 && devices && capacity > 0) devices[0] = d;
        return d >= 0 ? 1 : 0;
    }
    for (int i = 0; i < pool->size() && devices && i < capacity; i++) devices[i] = pool->device(i);
    return pool->size();
}

int itw_begin_deferred(void)
{
    tls.err.clear();
    if (ensure_ctx()) return -1;
    tls.deferred = true;
    tls.deferred_rc = 0;
    return 0;
}
int itw_flush(void)
{
    ThreadCtx& c = tls;
    int rc = c.deferred_rc;
    const std::string first = c.err;                      // the first failure of a deferred call is the one reported
    if (drain_lanes(c) && rc == 0) rc = -1;
    else if (rc) c.err = first;
    c.deferred = false;
    c.deferred_rc = 0;
    return rc;
}

// ---- the reference's coarse seam (3rdParty/Intel/Source/win32Threads.h:24, :52-80; win32Threads.cpp:192-330) ----
// CompressImageBC* build the profile on the stack and call CompressBlocks*, exactly like win32Threads.cpp:289-330.
// CompressImageMT / ST take the same arguments as the reference's; here BOTH hand the whole surface to the encoder in
// one call: the reference's per-thread row bands (win32Threads.cpp:217-230) exist to occupy CPU cores, while one GPU
// launch (or one band per selected device, itw_set_devices) already covers the surface -- the bytes are the same
// because bands are independent surfaces.
void CompressImageBC1(const rgba_surface* input, uint8_t* output) { CompressBlocksBC1(input, output); }
void CompressImageBC3(const rgba_surface* input, uint8_t* output) { CompressBlocksBC3(input, output); }
#define ITW_IMAGE_BC7(profile)                                                             \
    void CompressImageBC7_##profile(const rgba_surface* input, uint8_t* output)            \
    {                                                                                      \
        bc7_enc_settings settings;                                                         \
        GetProfile_##profile(&settings);                                                   \
        CompressBlocksBC7(input, output, &settings);                                       \
    }
#define ITW_IMAGE_BC6H(profile)                                                            \
    void CompressImageBC6H_##profile(const rgba_surface* input, uint8_t* output)           \
    {                                                                                      \
        bc6h_enc_settings settings;                                                        \
        GetProfile_bc6h_##profile(&settings);                                              \
        CompressBlocksBC6H(input, output, &settings);                                      \
    }
ITW_IMAGE_BC7(ultrafast) ITW_IMAGE_BC7(veryfast) ITW_IMAGE_BC7(fast) ITW_IMAGE_BC7(basic) ITW_IMAGE_BC7(slow)
ITW_IMAGE_BC7(alpha_ultrafast) ITW_IMAGE_BC7(alpha_veryfast) ITW_IMAGE_BC7(alpha_fast) ITW_IMAGE_BC7(alpha_basic) ITW_IMAGE_BC7(alpha_slow)
ITW_IMAGE_BC6H(veryfast) ITW_IMAGE_BC6H(fast) ITW_IMAGE_BC6H(basic) ITW_IMAGE_BC6H(slow) ITW_IMAGE_BC6H(veryslow)
#undef ITW_IMAGE_BC7
#undef ITW_IMAGE_BC6H

int GetBytesPerBlock(int format)                          // win32Threads.cpp:192-209: everything but BC3/BC6H/BC7 is 8
{
    switch (format) {
        case 77: case 78: case 98: case 99: case 95: case 96: return 16;
        default: return 8;
    }
}
bool CompressImageMT(const rgba_surface* input, uint8_t* output, CompressionFunc* cmpFunc, int compformat)
{
    (void)compformat;
    if (!cmpFunc) { fail("CompressImageMT: null compression function"); return false; }
    (*cmpFunc)(input, output);
    return true;                                          // like the reference (win32Threads.cpp:248, :281); errors: itw_get_last_error()
}
bool CompressImageST(const rgba_surface* input, uint8_t* output, CompressionFunc* cmpFunc, int compformat)
{
    return CompressImageMT(input, output, cmpFunc, compformat);
}
// The reference's thread-pool life cycle (win32Threads.h:52-55) mapped onto the device pool: InitWin32Threads selects
// every visible GPU, DestroyThreads returns to the single-device behaviour, GetProcessorCount reports the GPUs.
int GetProcessorCount(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    return n > 0 ? n : 1;
}
void InitWin32Threads(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); fail("InitWin32Threads: no CUDA device"); return; }
    std::vector<int> all((size_t)n);
    for (int i = 0; i < n; i++) all[i] = i;
    if (n >= 2) itw_set_devices(all.data(), n);
}
void DestroyThreads(void) { itw_set_devices(nullptr, 0); }

const char* itw_get_last_error(void) { return tls.err.c_str(); }

void itw_release(void)
{
    if (tls.stream) cudaStreamSynchronize(tls.stream);
    tls.release();
    tls.device = -1;
    tls.timed = false;
}

uint64_t itw_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

float itw_last_kernel_ms(void)
{
    float ms = -1.0f;
    if (tls.pool_ms >= 0.0f) return tls.pool_ms;
    if (tls.timed && cudaEventElapsedTime(&ms, tls.ev0, tls.ev1) != cudaSuccess) { cudaGetLastError(); ms = -1.0f; }
    return ms;
}

}  // extern "C"

#include "itw_dds.inc"
#include "itw_mips.inc"
#include "itw_shard.inc"
