/*
 * itw_bcn.h -- C-ABI of the B200-native BCn block encoder (libitw_bcn.so).
 *
 * Drop-in boundary: every struct and entry point in the first section has the exact layout,
 * name, argument meaning and calling convention of the reference's
 *     3rdParty/Intel/Source/ispc_texcomp.h:19-50  (structs)
 *     3rdParty/Intel/Source/ispc_texcomp.h:67-87  (GetProfile_*)
 *     3rdParty/Intel/Source/ispc_texcomp.h:104-107 (CompressBlocksBC1/BC3/BC6H/BC7)
 * so a host that links ispc_texcomp today (IntelPlugin.cpp's save path through
 * win32Threads.cpp:289-330) can link this library instead.  The second section is additive.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 */
#ifndef ITW_BCN_H
#define ITW_BCN_H

#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Section 1 -- the reference's own ABI
 * ------------------------------------------------------------------------------------------- */

/* ispc_texcomp.h:19-25.  `stride` is in BYTES.  LDR surfaces are RGBA8 (R in byte 0), HDR
 * surfaces are RGBA16F (half R,G,B at +0,+2,+4; A ignored).  width/height multiples of 4. */
typedef struct rgba_surface {
    uint8_t* ptr;
    int32_t  width;
    int32_t  height;
    int32_t  stride;
} rgba_surface;

/* ispc_texcomp.h:27-41 (64 bytes) */
typedef struct bc7_enc_settings {
    bool mode_selection[4];      /* {modes 0+2, modes 1+3+7, modes 4+5, mode 6} */
    int  refineIterations[8];    /* per BC7 mode */
    bool skip_mode2;
    int  fastSkipTreshold_mode1; /* (sic) number of ranked 2-subset partitions tried */
    int  fastSkipTreshold_mode3;
    int  fastSkipTreshold_mode7;
    int  mode45_channel0;        /* first rotation tried by modes 4/5 */
    int  refineIterations_channel;
    int  channels;               /* 3 = ignore alpha, 4 = RGBA */
} bc7_enc_settings;

/* ispc_texcomp.h:43-50 (16 bytes) */
typedef struct bc6h_enc_settings {
    bool slow_mode;
    bool fast_mode;
    int  refineIterations_1p;
    int  refineIterations_2p;
    int  fastSkipTreshold;
} bc6h_enc_settings;

/* ispc_texcomp.h:67-71 -- RGB profiles (alpha ignored); ispc_texcomp.cpp:20-189 */
void GetProfile_ultrafast(bc7_enc_settings* settings);
void GetProfile_veryfast(bc7_enc_settings* settings);
void GetProfile_fast(bc7_enc_settings* settings);
void GetProfile_basic(bc7_enc_settings* settings);
void GetProfile_slow(bc7_enc_settings* settings);
/* ispc_texcomp.h:74-78 -- RGBA profiles; ispc_texcomp.cpp:191-365 */
void GetProfile_alpha_ultrafast(bc7_enc_settings* settings);
void GetProfile_alpha_veryfast(bc7_enc_settings* settings);
void GetProfile_alpha_fast(bc7_enc_settings* settings);
void GetProfile_alpha_basic(bc7_enc_settings* settings);
void GetProfile_alpha_slow(bc7_enc_settings* settings);
/* ispc_texcomp.h:81-85 -- BC6H profiles; ispc_texcomp.cpp:367-410 */
void GetProfile_bc6h_veryfast(bc6h_enc_settings* settings);
void GetProfile_bc6h_fast(bc6h_enc_settings* settings);
void GetProfile_bc6h_basic(bc6h_enc_settings* settings);
void GetProfile_bc6h_slow(bc6h_enc_settings* settings);
void GetProfile_bc6h_veryslow(bc6h_enc_settings* settings);

/* ispc_texcomp.h:104-107; ispc_texcomp.cpp:417-435 -> kernel.ispc:598/607/3132/2030.
 * dst receives (width/4)*(height/4) blocks, tightly packed in raster block order
 * (8 bytes/block for BC1, 16 for BC3/BC6H/BC7).  `src->ptr` and `dst` may each be host or
 * device memory (detected per call).  The call is synchronous.  With a device operand the work is
 * issued on the legacy default stream, i.e. it is ordered after work already enqueued on blocking
 * streams; producers on NON-blocking streams must be synchronised by the caller (or use
 * itw_encode_device, which is stream-ordered).
 * void return as in the reference: failures are reported through itw_get_last_error(). */
void CompressBlocksBC1(const rgba_surface* src, uint8_t* dst);
void CompressBlocksBC3(const rgba_surface* src, uint8_t* dst);
void CompressBlocksBC6H(const rgba_surface* src, uint8_t* dst, bc6h_enc_settings* settings);
void CompressBlocksBC7(const rgba_surface* src, uint8_t* dst, bc7_enc_settings* settings);

/* ---------------------------------------------------------------------------------------------
 * Section 2 -- additive entry points (not in the reference header)
 * ------------------------------------------------------------------------------------------- */

/* BC4 (R -> 8 B/block) and BC5 (R,G -> 16 B/block) of an RGBA8 surface.  In the reference these
 * formats go through DirectX::Compress (IntelPlugin.cpp:272) to D3DXEncodeBC4U / D3DXEncodeBC5U
 * (DirectXTex/BC4BC5.cpp:403, :481); same surface/dst conventions as above. */
void CompressBlocksBC4(const rgba_surface* src, uint8_t* dst);
void CompressBlocksBC5(const rgba_surface* src, uint8_t* dst);

/* Format selector for the generic entry points below; values are the DXGI_FORMAT enumerants the
 * plug-in switches on (IntelPlugin.cpp:820-847, win32Threads.cpp:192-209). */
enum {
    ITW_FORMAT_BC1 = 71,  /* DXGI_FORMAT_BC1_UNORM  */
    ITW_FORMAT_BC3 = 77,  /* DXGI_FORMAT_BC3_UNORM  */
    ITW_FORMAT_BC4 = 80,  /* DXGI_FORMAT_BC4_UNORM  */
    ITW_FORMAT_BC5 = 83,  /* DXGI_FORMAT_BC5_UNORM  */
    ITW_FORMAT_BC6H = 95, /* DXGI_FORMAT_BC6H_UF16  */
    ITW_FORMAT_BC6H_SF16 = 96, /* DXGI_FORMAT_BC6H_SF16: ENCODED by the same unsigned encoder, as the plug-in does
                                  (IntelPlugin.cpp:840-843); DECODED with the signed rules (DirectXTexCompress.cpp:414) */
    ITW_FORMAT_BC7 = 98   /* DXGI_FORMAT_BC7_UNORM  */
};

/* Bytes per 4x4 block for a format (win32Threads.cpp:192-209); 0 for an unknown format. */
int itw_bytes_per_block(int format);

/* Device-resident encode on an explicit CUDA stream (cudaStream_t passed as void*; NULL = the
 * legacy default stream).  src->ptr and dst MUST be device pointers; nothing is copied and the
 * call returns as soon as the kernel is enqueued.  `settings` is a bc7_enc_settings* for BC7, a
 * bc6h_enc_settings* for BC6H and ignored otherwise; it is read before the call returns.
 * Returns 0 on success, a negative value on error (see itw_get_last_error). */
int itw_encode_device(int format, const rgba_surface* src, uint8_t* dst, const void* settings,
                      void* cuda_stream);

/* Batched encode of `count` independent surfaces (config C5's tile stream).  Host surfaces are pipelined
 * H2D / encode / D2H over internal streams (and dealt round-robin over the devices of itw_set_devices);
 * dst[i] receives surface i's blocks.  Synchronous.  Ordering: device-resident operands are ordered after the work
 * already enqueued on the legacy default stream / blocking streams at the time of the call (as for
 * CompressBlocks*); producers on NON-blocking streams must be synchronised by the caller.  The same rule holds for the
 * device-resident surfaces of itw_dds_encode_file / itw_dds_encode_texture / itw_dds_encode_pixels. */
int itw_encode_batch(int format, const rgba_surface* srcs, uint8_t* const* dsts, int count,
                     const void* settings);

/* Select the CUDA device used by this thread's subsequent calls (default: current device). */
int itw_set_device(int device);

/* Multi-GPU in ONE process (the analogue of the reference's thread pool, win32Threads.cpp:98-274, which
 * fans row bands of one surface over its workers, :211-249).  After itw_set_devices(list, n >= 2) every HOST -> HOST
 * call of CompressBlocks* / CompressImage* is cut into one band of whole block rows per device; each device copies
 * in, encodes and copies out its band on its own PCIe link, straight into the caller's dst.  itw_encode_batch deals
 * its tiles round-robin over the devices (config C5's tile stream).  Bands and tiles are independent, so the output is
 * byte-identical to the single-device one; there is no collective.  Device-resident operands stay on their device.
 * n == 1 makes that device the process-wide default; n == 0 restores the single-device behaviour (the calling
 * thread's current device / itw_set_device).  Process-wide; not to be called concurrently with encodes.
 * Returns 0 on success. */
int itw_set_devices(const int* devices, int count);
/* The devices selected by itw_set_devices (written to devices[0..capacity)); returns their number, 0 if none. */
int itw_get_devices(int* devices, int capacity);

/* Deferred mode for hosts that keep the reference's slice loop (IntelPlugin.cpp:851-879: one CompressImageMT per
 * 256 K-texel slice of the image).  Between itw_begin_deferred() and itw_flush(), host -> host encodes issued by the
 * calling thread are only ENQUEUED (copy-in, kernel, copy-out on internal streams, pipelined over three lanes) and
 * return at once; src and dst must stay valid and untouched until itw_flush(), which waits for all of them and returns
 * 0 or the first failure (text in itw_get_last_error).  Two added lines in the host: begin before the loop, flush after. */
int itw_begin_deferred(void);
int itw_flush(void);

/* ---- the reference's coarse seam: 3rdParty/Intel/Source/win32Threads.h:24, :52-80 (win32Threads.cpp:192-330) ----
 * Same names and argument meaning; BYTE is spelled uint8_t and DXGI_FORMAT is passed as int (the enum's underlying
 * type).  CompressImageBC* = GetProfile + CompressBlocks* (win32Threads.cpp:289-330).  CompressImageMT / ST call
 * `cmpFunc` ONCE for the whole surface: the reference's per-thread bands (win32Threads.cpp:217-230) become one GPU
 * launch, or one band per device after itw_set_devices / InitWin32Threads -- same bytes, bands being independent. */
typedef void(CompressionFunc)(const rgba_surface* input, uint8_t* output);
int  GetProcessorCount(void);      /* visible CUDA devices, 1 or more (win32Threads.h:52) */
void InitWin32Threads(void);       /* select every visible GPU (itw_set_devices) -- the pool the reference builds here */
void DestroyThreads(void);         /* back to single-device behaviour */
int  GetBytesPerBlock(int dxgi_format);
bool CompressImageMT(const rgba_surface* input, uint8_t* output, CompressionFunc* cmpFunc, int compformat);
bool CompressImageST(const rgba_surface* input, uint8_t* output, CompressionFunc* cmpFunc, int compformat);
void CompressImageBC1(const rgba_surface* input, uint8_t* output);
void CompressImageBC3(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_ultrafast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_veryfast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_fast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_basic(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_slow(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_alpha_ultrafast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_alpha_veryfast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_alpha_fast(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_alpha_basic(const rgba_surface* input, uint8_t* output);
void CompressImageBC7_alpha_slow(const rgba_surface* input, uint8_t* output);
void CompressImageBC6H_veryfast(const rgba_surface* input, uint8_t* output);
void CompressImageBC6H_fast(const rgba_surface* input, uint8_t* output);
void CompressImageBC6H_basic(const rgba_surface* input, uint8_t* output);
void CompressImageBC6H_slow(const rgba_surface* input, uint8_t* output);
void CompressImageBC6H_veryslow(const rgba_surface* input, uint8_t* output);

/* Last error message of the calling thread ("" if none).  The CompressBlocks* entry points keep
 * the reference's void signature, so this is the only error channel for them. */
const char* itw_get_last_error(void);

/* Free the calling thread's device buffers, streams and events now (they are otherwise kept for the next
 * call and freed when the thread exits); the next call re-creates what it needs. */
void itw_release(void);

/* Number of kernel launches issued by this library since load (all threads); bench evidence. */
uint64_t itw_kernel_launch_count(void);

/* Device time in milliseconds of the most recent call issued by this thread (encode, decode, convert),
 * measured with CUDA events on the launching stream around the kernel(s) only (no copies).  For host
 * surfaces BC7 / BC6H run as four pipelined row bands: the span then covers the four kernels including
 * the short waits for the later bands' input copies.  Valid after the call has completed. */
float itw_last_kernel_ms(void);

/* ---------------------------------------------------------------------------------------------
 * Section 3 -- DDS container: what IntelPlugin.cpp:2171 obtains from
 * DirectX::SaveToDDSMemory (DirectXTex/DirectXTexDDS.cpp:1611-1815) with header rules of
 * _EncodeDDSHeader (:441-675) and the structures of DirectXTex/DDS.h:40-235.  2-D textures and
 * cube maps of block-compressed formats only.
 * ------------------------------------------------------------------------------------------- */
typedef struct itw_dds_desc {
    uint32_t width, height;   /* top level, texels (need not be multiples of 4) */
    uint32_t mip_levels;      /* >= 1 */
    uint32_t array_size;      /* images per mip level: 1, or 6*n for cube maps */
    uint32_t dxgi_format;     /* DXGI_FORMAT value, e.g. 71 BC1_UNORM, 72 BC1_UNORM_SRGB, 77/78 BC3, 80 BC4_UNORM,
                                 83 BC5_UNORM, 95 BC6H_UF16, 96 BC6H_SF16, 98/99 BC7 */
    uint32_t is_cubemap;      /* 0 / 1 */
} itw_dds_desc;

/* 128 (legacy FourCC header: BC1/BC3/BC4U/BC5U, single image or one cube map) or 148 (with the
 * 'DX10' extension: sRGB variants, BC6H, BC7, arrays); 0 if the description is not supported. */
size_t itw_dds_header_bytes(const itw_dds_desc* desc);
/* Bytes of one image (array item `item`, mip `mip`) and its offset from the start of the file;
 * images are stored item-major, mip-minor, each tightly packed (DirectXTexDDS.cpp:1676-1720). */
size_t itw_dds_image_bytes(const itw_dds_desc* desc, uint32_t mip);
size_t itw_dds_image_offset(const itw_dds_desc* desc, uint32_t item, uint32_t mip);
size_t itw_dds_file_bytes(const itw_dds_desc* desc);
/* Write magic + DDS_HEADER (+ DDS_HEADER_DXT10).  Returns the header size, 0 on error. */
size_t itw_dds_write_header(const itw_dds_desc* desc, uint8_t* dst, size_t capacity);
/* Parse a header written by this library or by DirectXTex.  Returns the payload offset, 0 on error. */
size_t itw_dds_read_header(const uint8_t* src, size_t size, itw_dds_desc* desc);
/* The whole save path for one texture: encode array_size*mip_levels surfaces (item-major, mip-minor,
 * each already padded to multiples of 4 as IntelPlugin.cpp:893-928 does) straight into `file`
 * (capacity >= itw_dds_file_bytes) behind the header.  `settings` as for itw_encode_device.
 * Returns the file size, 0 on error. */
size_t itw_dds_encode_file(const itw_dds_desc* desc, const rgba_surface* images, const void* settings,
                           uint8_t* file, size_t capacity);

/* ---------------------------------------------------------------------------------------------
 * Section 4 -- on-GPU pre-pass: mip chain + pad-to-4.
 * Level l has max(1,w>>l) x max(1,h>>l) texels and is STORED padded to multiples of 4 by edge replication
 * (IntelPlugin.cpp:893-928), rows tightly packed -- ready to be handed to CompressBlocks*.
 * Filters = DirectXTex's own (non-WIC) generators, which are in the reference tree (DirectXTexMipmaps.cpp:715-905,
 * Filters.h:33-112): the BOX filter ((p0+p1)+p2+p3)*0.25 when width and height of level 0 are powers of two, the two-tap
 * LINEAR filter otherwise (GenerateMipMaps' default choice, :2611-2616), every level re-read from the stored previous one.
 *   itw_generate_mips_device       RGBA8 as R8G8B8A8_UNORM.  For exact 2:1 levels the float box on byte/255 values stored
 *                                  with round-to-nearest equals the integer (a+b+c+d+2)>>2, which is what runs there.
 *                                  (The plug-in's default for UNORM encodings is the WIC scaler, which is NOT in the tree:
 *                                  parity with WIC is unpinned; this is DirectXTex's TEX_FILTER_FORCE_NON_WIC result.)
 *   itw_generate_mips_device_srgb  RGBA8 as R8G8B8A8_UNORM_SRGB: filtered in linear light (XMColorSRGBToRGB after the load,
 *                                  XMColorRGBToSRGB before the store, DirectXTexConvert.cpp:2669-2685, :2757-2775) -- what the
 *                                  plug-in gets for BC1/BC3/BC7 *_SRGB encodings (IntelPlugin.cpp:152-154; WIC is bypassed for
 *                                  sRGB formats, DirectXTexMipmaps.cpp:389-393).  itw_dds_encode_texture / _pixels use it for
 *                                  dxgi_format 72 / 78 / 99.
 *   itw_generate_mips_device_f16   RGBA16F, the BC6H save path (TEX_FILTER_FORCE_NON_WIC, IntelPlugin.cpp:2117-2127).
 * Bit-exact to those generator bodies (cut and compiled by oracle/build_ref_frontend.py), including their stale fourth tap
 * on wide power-of-two textures (csrc/mips_f16.cuh).  Outside the tree and therefore assumed: XMLoadUByteN4 = byte*(1/255),
 * XMStoreUByteN4 = round to nearest, the half conversions of DirectXMath 3.06, the C library's powf.
 * ------------------------------------------------------------------------------------------- */
/* Bytes of device scratch for the padded levels first_level..levels-1 of a w x h RGBA8 texture. */
size_t itw_mip_scratch_bytes(int width, int height, int levels, int first_level);
/* Build padded levels 1..levels-1 (and a padded copy of level 0 if its size is not a multiple of 4) in
 * `scratch` (device, >= itw_mip_scratch_bytes(w, h, levels, pad0 ? 0 : 1)) from the device surface
 * `level0` (width/height = the true texture size, any stride).  out[l] receives the padded surface of
 * level l (out[0] = *level0 when no padding is needed).  Enqueued on `cuda_stream`; returns 0 on success. */
int itw_generate_mips_device(const rgba_surface* level0, int levels, rgba_surface* out, uint8_t* scratch,
                             void* cuda_stream);
/* Same arguments, sRGB-correct RGBA8 chain / RGBA16F chain (8-byte texels: scratch needs TWICE itw_mip_scratch_bytes(...)). */
int itw_generate_mips_device_srgb(const rgba_surface* level0, int levels, rgba_surface* out, uint8_t* scratch,
                                  void* cuda_stream);
int itw_generate_mips_device_f16(const rgba_surface* level0, int levels, rgba_surface* out, uint8_t* scratch,
                                 void* cuda_stream);
/* The complete save path of one texture (IntelPlugin.cpp:2117-2171): `tops` = array_size HOST or
 * device surfaces holding level 0 only (RGBA8; RGBA16F for BC6H_UF16 / _SF16); mips are generated on the
 * GPU, every level is encoded, and the DDS file is written to `file` (host).  Returns the file size, 0 on
 * error. */
size_t itw_dds_encode_texture(const itw_dds_desc* desc, const rgba_surface* tops, const void* settings,
                              uint8_t* file, size_t capacity);

/* ---------------------------------------------------------------------------------------------
 * Section 5 -- decoders: what the plug-in's preview obtains from
 * DirectX::Decompress (IntelPlugin.cpp:1051-1066; DirectXTex/DirectXTexCompress.cpp:358-468 ->
 * D3DXDecodeBC1/3/4U/5U/6HU/7, BC.cpp:897, BC4BC5.cpp:369-400, BC6HBC7.cpp:2879-2900).
 * `blocks` holds (width/4)*(height/4) blocks in raster order; dst->ptr is WRITTEN: RGBA8 texels
 * (4 B) for BC1/BC3/BC4/BC5/BC7 (BC4: r,0,0,255; BC5: r,g,0,255), RGBA16F texels (8 B, half bit
 * patterns, alpha = 1.0) for BC6H_UF16 / BC6H_SF16 (signed decode: D3DXDecodeBC6HS).  width/height multiples of 4; either side may be host or
 * device memory; synchronous.  BC7 and BC6H are exact by the format definition; BC1-BC5 equal
 * DirectXTex's float decoders rounded to nearest (csrc/decode.cuh).  Returns 0 on success.
 * ------------------------------------------------------------------------------------------- */
int itw_decode(int format, const uint8_t* blocks, const rgba_surface* dst);

/* ---------------------------------------------------------------------------------------------
 * Section 6 -- pixel-format front end and the image-level entry: what the plug-in
 * does between Photoshop's buffer and CompressBlocks*:
 *   CopyDataForEncoding (IntelPlugin.cpp:85-181) -> ConvertToBC{,4or5,6}From{8,16,32}Bit (:291-433,
 *   :741-810) with the scalar rules of IntelPlugin.h:41-96; FlipXYChannelNormalMap (:1504-1546);
 *   NormalizeNormalMapChain (:1551-1612); DoPaddingToMultiplesOf4 (:892-928).
 * All of it is per-texel work, fused here into ONE bandwidth-bound pass (the padding is a
 * coordinate clamp), optionally followed by the encoder without leaving the GPU.
 * ------------------------------------------------------------------------------------------- */
typedef struct itw_pixel_source {  /* Photoshop's FormatRecord as the converters read it */
    const void* data;              /* interleaved: element (x, y, c) at row y, index x*planes + c   */
    int32_t width, height;         /* texels; need NOT be multiples of 4                            */
    int32_t planes;                /* channels per texel in `data`, 1..4 (hiPlane - loPlane + 1)    */
    int32_t depth;                 /* 8; 16 = Photoshop's 0..32768 integers; 32 = float, gamma 1.0   */
    int64_t row_bytes;             /* bytes between rows; 0 = tightly packed (the plug-in's case)    */
} itw_pixel_source;
enum {
    ITW_FRONT_HAS_ALPHA = 1,   /* plane 3 is alpha (else alpha = 255 / 1.0); needs planes == 4          */
    ITW_FRONT_GAMMA = 2,       /* 32-bit LDR only: v^(1/2.2) before the byte conversion (IntelPlugin.h:71) */
    ITW_FRONT_FLIP_X = 4,      /* normal maps: r = 255 - r   /  half(1 - r)                              */
    ITW_FRONT_FLIP_Y = 8,      /* normal maps: g = 255 - g   /  half(1 - g)                              */
    ITW_FRONT_NORMALIZE = 16   /* normal maps: rgb re-normalised around 128 (LDR) / to unit length (HDR) */
};
/* Convert `src` for encoding to `format` (ITW_FORMAT_*): BC6H -> RGBA16F texels, everything else ->
 * RGBA8; missing planes are 0 (BC4/BC5: a copy of plane 0).  dst->width/height must be the source
 * size or that size rounded up to multiples of 4 (edge texels are replicated).  Either side may be
 * host or device memory; synchronous.  Returns 0 on success. */
int itw_convert_pixels(int format, const itw_pixel_source* src, uint32_t flags, const rgba_surface* dst);
/* The image-level entry (the seam of CompressImageST, win32Threads.h:57-58, moved above the
 * conversion): convert + pad + encode in one call; `dst_blocks` receives ((w+3)/4)*((h+3)/4)
 * blocks.  `settings` as for itw_encode_device.  Returns 0 on success. */
int itw_encode_pixels(int format, const itw_pixel_source* src, uint32_t flags, const void* settings,
                      uint8_t* dst_blocks);

/* The plug-in's whole save path (IntelPlugin.cpp:2062-2171) in one call: `sources` = array_size pixel
 * sources (level 0 of a 2-D texture, or the faces of a cube map) of desc->width x desc->height texels.
 * CopyDataForEncoding + FlipXYChannelNormalMap -> mip chain (as itw_dds_encode_texture) ->
 * NormalizeNormalMapChain on EVERY level when ITW_FRONT_NORMALIZE is set (the plug-in normalises after
 * the chain exists, :2149-2152) -> pad -> encode -> DDS blob in `file` (host).  Returns the file size, 0 on
 * error. */
size_t itw_dds_encode_pixels(const itw_dds_desc* desc, const itw_pixel_source* sources, uint32_t flags,
                             const void* settings, uint8_t* file, size_t capacity);

/* ---------------------------------------------------------------------------------------------
 * Section 7 -- row-sharded encode over several GPUs, one process per GPU:
 * "images shard row-wise across the GPUs with a single NCCL all-gather of the packed output".
 * Band rule = the reference's thread split (win32Threads.cpp:217-230).  Every rank holds its rows of level 0
 * of an RGBA8 texture on its device; mips are made band-locally while a level's band is a whole number of block
 * rows, each band is encoded where it lies, ONE ncclAllGather moves the packed bands (plus the few KiB of raw
 * texels the tiny remaining levels are filtered from), and every rank ends up with the complete packed chain --
 * byte-identical to the single-GPU encode of the same texture.  NCCL is bound at run time (libnccl.so.2, or the
 * path in ITW_NCCL_LIB); with no communicator the same call runs on one GPU.
 * ------------------------------------------------------------------------------------------- */
typedef struct itw_shard_plan {
    int32_t  nranks, rank, levels;
    int32_t  band_levels;         /* leading levels whose row bands are whole block rows on every rank            */
    int32_t  band_y0, band_y1;    /* rows of level 0 owned by `rank`                                              */
    uint64_t slot_bytes;          /* per-rank payload of the all-gather (multiple of 16)                          */
    uint64_t chain_bytes;         /* the complete packed chain, levels 0..levels-1 back to back (= DDS payload)   */
    uint64_t level_offset[16];    /* of level l in the chain                                                      */
    uint64_t level_bytes[16];
    uint64_t send_offset[16];     /* of this rank's band of level l inside its slot (l < band_levels)             */
    uint64_t band_bytes[16];      /* one rank's band of level l; rank r's band sits at level_offset + r*band_bytes */
    uint64_t texel_offset;        /* raw texels of level band_levels-1 inside the slot ...                        */
    uint64_t texel_bytes;         /* ... 0 when every level is band-local                                         */
} itw_shard_plan;
/* Pure arithmetic (no CUDA call): the layout used by itw_encode_mip_chain_sharded.  width/height multiples of 4,
 * levels <= 16, height / nranks a multiple of 4 rows.  Returns 0 on success. */
int itw_shard_plan_make(int format, int width, int height, int levels, int nranks, int rank, itw_shard_plan* plan);
/* NCCL bootstrap: rank 0 obtains an id (ncclGetUniqueId), the host passes it to the other ranks by its own means
 * (MPI, a torch.distributed broadcast, a file), every rank calls itw_shard_init on the thread and device it encodes
 * with (ncclCommInitRank).  itw_shard_finalize destroys the calling thread's communicator. */
int itw_shard_unique_id(uint8_t id[128]);
int itw_shard_init(int rank, int nranks, const uint8_t id[128]);
void itw_shard_finalize(void);
/* band0 = this rank's rows [band_y0, band_y1) of level 0 (device memory, any 4-byte aligned stride; width x
 * (band_y1-band_y0)); chain = device buffer of plan.chain_bytes (16-byte aligned) that receives the COMPLETE packed
 * chain on every rank.  Everything is enqueued on `cuda_stream` (cudaStream_t as void*), nothing is synchronised;
 * the scratch belongs to the calling thread, so consecutive calls of one thread must use one stream (or be separated
 * by a synchronisation).  `settings` as for itw_encode_device.  Returns 0 on success. */
int itw_encode_mip_chain_sharded(int format, const rgba_surface* band0, int width, int height, int levels,
                                 const void* settings, uint8_t* chain, void* cuda_stream);

#ifdef __cplusplus
}
#endif

#endif /* ITW_BCN_H */
