#!/usr/bin/env python3
"""Build oracle/_ref/libitw_ref_frontend.so FROM THE REFERENCE'S OWN SOURCES (test infrastructure only).

The pixel-format front end of the plug-in is plain C++ inside two Photoshop-SDK-bound files:

  IntelCompressionPlugin/IntelPlugin.h:41-96     FloatToByte, ConvertTo8Bit x3, ConvertTo16Bit x3
  IntelCompressionPlugin/IntelPlugin.cpp         ConvertToBC{,4or5,6}From{8,16,32}Bit (:291-433, :741-810),
                                                 FlipXYChannelNormalMap (:1504-1546), NormalizeNormalMapChain (:1551-1612),
                                                 DoPaddingToMultiplesOf4 (:892-928)

and, for the RGBA16F mip chain of the BC6H save path (TEX_FILTER_FORCE_NON_WIC, IntelPlugin.cpp:2117-2127),

  3rdParty/DirectXTex/DirectXTex/DirectXTexMipmaps.cpp   _Generate2DMipsBoxFilter (:715-805), _Generate2DMipsLinearFilter (:809-905)
  3rdParty/DirectXTex/DirectXTex/Filters.h               AVERAGE4, LinearFilter, _CreateLinearFilter, BILINEAR_INTERPOLATE (:29-112)

(shimmed there: XMVECTOR and its + and * operators as four scalar IEEE operations, and _LoadScanlineLinear /
_StoreScanlineLinear for R16G16B16A16_FLOAT as the half conversions below -- in the reference they are XMLoadHalf4 / XMStoreHalf4),

and the BC4 / BC5 encoders the plug-in reaches through DirectX::Compress (IntelPlugin.cpp:272):

  3rdParty/DirectXTex/DirectXTex/BC4BC5.cpp   the whole DirectX namespace body: BC4_UNORM, FindEndPointsBC4U, FindClosestUNORM,
                                              D3DXEncodeBC4U / BC5U (and the decoders / SNORM variants, unused here)
  3rdParty/DirectXTex/DirectXTex/BC.h         template OptimizeAlpha (:727-856)

(shimmed there: XMVectorGetX / XMVectorSet / XMStoreFloat4A as plain member access; the texel floats handed to the encoders are
byte * (1/255), rule F7 -- in the reference they come from XMLoadUByteN4 inside _CompressBC, DirectXMath again),

and the BC1 / BC3 DECODERS of the preview path (DirectX::Decompress, IntelPlugin.cpp:1051-1066):

  3rdParty/DirectXTex/DirectXTex/BC.cpp       DecodeBC1 (:322-370), D3DXDecodeBC1 (:722-726), D3DXDecodeBC3 (:897-936)
  3rdParty/DirectXTex/DirectXTex/BC.h         struct D3DX_BC1 / D3DX_BC3 (:282-301)

(shimmed there: XMLoadU565, XMVectorSwizzle, XMVectorSelect, XMVectorLerp -- "(V1 - V0) * t + V0", a multiply then an add --,
XMVectorZero, XMVectorSetW).  The decoders return FLOAT texels; the final store to RGBA8 is DirectXMath's XMStoreUByteN4, whose
rounding is outside the tree: the tests compare against round-to-nearest of these floats.

Neither file compiles here (Photoshop SDK, Windows, DirectXTex), but these functions only touch a handful of
fields.  This recipe cuts exactly those function bodies out of the files WHERE THEY LIE under /root/reference
(never copied into the repo), puts them in a temp directory behind a small shim that declares the fields they
use (`ps.formatRecord->{depth,data,imageSize}`, `ps.data->{FlipX,FlipY,encoding_g}`, DirectXTex's `Image`), and
compiles them with g++ -O2 -ffp-contract=off into oracle/_ref/ (git-ignored; travels to the GPU box).

What the shim has to SUPPLY, because the reference takes it from the Windows SDK and not from its own tree:
  * DirectX::PackedVector::XMConvertFloatToHalf / XMConvertHalfToFloat.  Restated from the published DirectXMath
    3.06 source (Windows 8.1 SDK -- the SDK of the reference's v120 toolset; no F16C path in that version):
    round-to-nearest-even on the rebiased bit pattern, |x| > 0x47FFEFFF -> 0x7FFF, denormals by a truncating
    shift, and a HalfToFloat that treats exponent 31 as an ordinary binade.  Later DirectXMath versions differ
    above 65504, for NaN and (F16C builds) for half-denormal rounding: half conversion is UNPINNED there; finite
    values whose result is a normal half are round-to-nearest-even in every version.
  * pow(double,double): the reference links the MSVC CRT, this build links glibc.  Both are faithfully rounded;
    a byte could differ only where pow(v, 1/2.2)*255 lies within an ulp of an integer.  UNPINNED to that extent.
"""
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ITW_REFERENCE_ROOT", "/root/reference")
PLUGIN_H = os.path.join(REF_ROOT, "IntelCompressionPlugin", "IntelPlugin.h")
PLUGIN_CPP = os.path.join(REF_ROOT, "IntelCompressionPlugin", "IntelPlugin.cpp")
MIPMAPS_CPP = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "DirectXTexMipmaps.cpp")
FILTERS_H = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "Filters.h")
CONVERT_CPP = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "DirectXTexConvert.cpp")
BC45_CPP = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "BC4BC5.cpp")
BC_H = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "BC.h")
BC_CPP = os.path.join(REF_ROOT, "3rdParty", "DirectXTex", "DirectXTex", "BC.cpp")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libitw_ref_frontend.so")

MEMBERS = ["ConvertToBC6From8Bit", "ConvertToBC6From16Bit", "ConvertToBC6From32Bit",
           "ConvertToBC4or5From8Bit", "ConvertToBC4or5From16Bit", "ConvertToBC4or5From32Bit",
           "ConvertToBCFrom8Bit", "ConvertToBCFrom16Bit", "ConvertToBCFrom32Bit",
           "FlipXYChannelNormalMap", "NormalizeNormalMapChain", "DoPaddingToMultiplesOf4"]
INLINES = ["F16toF32", "F32toF16", "FloatToByte", "F16toByte"]          # + every ConvertTo8Bit / ConvertTo16Bit overload

SHIM_TOP = r"""
// generated by oracle/build_ref_frontend.py -- shim around function bodies cut from the reference
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
typedef unsigned char unsigned8;
typedef unsigned short unsigned16;
typedef uint8_t uint8;
typedef uint16_t uint16;
enum DXGI_FORMAT { DXGI_FORMAT_R16G16B16A16_FLOAT = 10, DXGI_FORMAT_R8G8B8A8_UNORM = 28, DXGI_FORMAT_R8G8B8A8_UNORM_SRGB = 29, DXGI_FORMAT_BC6H_UF16 = 95 };
struct rgba_surface { uint8_t* ptr; int32_t width, height, stride; };
namespace DirectX {
struct Image { size_t width, height; DXGI_FORMAT format; size_t rowPitch, slicePitch; uint8_t* pixels; };
struct ScratchImage {
    Image* images; size_t count;
    const Image* GetImages() const { return images; }
    size_t GetImageCount() const { return count; }
};
namespace PackedVector {
// DirectXMath 3.06 (Windows 8.1 SDK), DirectXPackedVector.inl, scalar path -- restated, see the recipe's header
inline float XMConvertHalfToFloat(unsigned short Value)
{
    uint32_t Mantissa = (uint32_t)(Value & 0x03FF);
    uint32_t Exponent;
    if ((Value & 0x7C00) != 0) Exponent = (uint32_t)((Value >> 10) & 0x1F);
    else if (Mantissa != 0) {
        Exponent = 1;
        do { Exponent--; Mantissa <<= 1; } while ((Mantissa & 0x0400) == 0);
        Mantissa &= 0x03FF;
    } else Exponent = (uint32_t)-112;
    uint32_t Result = ((uint32_t)(Value & 0x8000) << 16) | ((Exponent + 112) << 23) | (Mantissa << 13);
    float f; memcpy(&f, &Result, 4); return f;
}
inline unsigned short XMConvertFloatToHalf(float Value)
{
    uint32_t Result, IValue; memcpy(&IValue, &Value, 4);
    uint32_t Sign = (IValue & 0x80000000U) >> 16U;
    IValue = IValue & 0x7FFFFFFFU;
    if (IValue > 0x47FFEFFFU) Result = 0x7FFFU;
    else {
        if (IValue < 0x38800000U) {
            uint32_t Shift = 113U - (IValue >> 23U);
            IValue = (Shift < 32U) ? ((0x800000U | (IValue & 0x7FFFFFU)) >> Shift) : 0U;   // shift >= 32 is undefined in the original
        } else IValue += 0xC8000000U;
        Result = ((IValue + 0x0FFFU + ((IValue >> 13U) & 1U)) >> 13U) & 0x7FFFU;
    }
    return (unsigned short)(Result | Sign);
}
}  // namespace PackedVector
}  // namespace DirectX
using namespace DirectX;
struct FakePoint { int h, v; };
struct FakeFormatRecord { int depth; void* data; FakePoint imageSize; };
struct FakeData { bool FlipX, FlipY; int encoding_g; };
struct FakeGlobals { FakeFormatRecord* formatRecord; FakeData* data; };
"""

SHIM_CLASS = r"""
class IntelPlugin {
public:
    FakeGlobals ps;
    bool ConvertToBC6From8Bit(unsigned16*, int, bool);
    bool ConvertToBC6From16Bit(unsigned16*, int, bool);
    bool ConvertToBC6From32Bit(unsigned16*, int, bool);
    bool ConvertToBC4or5From8Bit(unsigned8*, int, bool);
    bool ConvertToBC4or5From16Bit(unsigned8*, int, bool);
    bool ConvertToBC4or5From32Bit(unsigned8*, int, bool, bool);
    bool ConvertToBCFrom8Bit(unsigned8*, int, bool);
    bool ConvertToBCFrom16Bit(unsigned8*, int, bool);
    bool ConvertToBCFrom32Bit(unsigned8*, int, bool, bool);
    void FlipXYChannelNormalMap(ScratchImage*);
    void NormalizeNormalMapChain(ScratchImage*);
    rgba_surface DoPaddingToMultiplesOf4(const rgba_surface&);
};
"""

SHIM_BOTTOM = r"""
// ---- C entry point: the order of IntelPlugin.cpp's save path ---------------------------------------------
//   CopyDataForEncoding (:85-181, the depth/format dispatch is restated here) -> FlipXYChannelNormalMap (:2100-2105)
//   -> NormalizeNormalMapChain (:2151-2152) -> DoPaddingToMultiplesOf4 (:244-248)
// flags: 1 has alpha, 2 gamma correct, 4 flip X, 8 flip Y, 16 normalise.  `dst` = tightly packed RGBA8 / RGBA16F
// of ((w+3)&~3) x ((h+3)&~3) texels.
extern "C" int ref_convert_pixels(int format, const void* data, int w, int h, int planes, int depth, unsigned flags, uint8_t* dst)
{
    IntelPlugin p;
    FakeFormatRecord rec{depth, const_cast<void*>(data), {w, h}};
    FakeData dat{(flags & 4) != 0, (flags & 8) != 0, format};
    p.ps.formatRecord = &rec;
    p.ps.data = &dat;
    const bool alpha = (flags & 1) != 0, gamma = (flags & 2) != 0;
    const bool hdr = (format == 95);
    const size_t texel = hdr ? 8 : 4;
    uint8_t* top = new uint8_t[(size_t)w * h * texel];
    bool ok = false;
    if (hdr) {
        unsigned16* t = reinterpret_cast<unsigned16*>(top);
        ok = depth == 8 ? p.ConvertToBC6From8Bit(t, planes, alpha) : depth == 16 ? p.ConvertToBC6From16Bit(t, planes, alpha)
                        : p.ConvertToBC6From32Bit(t, planes, alpha);
    } else if (format == 80 || format == 83) {
        ok = depth == 8 ? p.ConvertToBC4or5From8Bit(top, planes, alpha) : depth == 16 ? p.ConvertToBC4or5From16Bit(top, planes, alpha)
                        : p.ConvertToBC4or5From32Bit(top, planes, alpha, gamma);
    } else {
        ok = depth == 8 ? p.ConvertToBCFrom8Bit(top, planes, alpha) : depth == 16 ? p.ConvertToBCFrom16Bit(top, planes, alpha)
                        : p.ConvertToBCFrom32Bit(top, planes, alpha, gamma);
    }
    if (!ok) { delete[] top; return -1; }
    Image img{(size_t)w, (size_t)h, hdr ? DXGI_FORMAT_R16G16B16A16_FLOAT : DXGI_FORMAT_R8G8B8A8_UNORM, (size_t)w * texel, (size_t)w * h * texel, top};
    ScratchImage scratch{&img, 1};
    if (flags & 12) p.FlipXYChannelNormalMap(&scratch);
    if (flags & 16) p.NormalizeNormalMapChain(&scratch);
    rgba_surface in{top, w, h, (int32_t)(w * texel)};
    if ((w | h) & 3) {
        rgba_surface out = p.DoPaddingToMultiplesOf4(in);
        memcpy(dst, out.ptr, (size_t)out.stride * out.height);
        delete[] out.ptr;
    } else memcpy(dst, top, (size_t)w * h * texel);
    delete[] top;
    return 0;
}
extern "C" unsigned short ref_float_to_half(float v) { return F32toF16(v); }
extern "C" float ref_half_to_float(unsigned short h) { return F16toF32(h); }
"""


MIP_SHIM_TOP = r"""
// ---- DirectXTex mip filters: shim ------------------------------------------------------------------------
#include <memory>
#include <new>
#include <utility>
#include <vector>
#include <cstdlib>
#define _In_
#define _In_z_
#define _Out_writes_(x)
#define _In_reads_(x)
typedef long HRESULT;
typedef unsigned long DWORD;
typedef const void* LPCVOID;
typedef void* LPVOID;
#define S_OK 0L
#define E_FAIL (-1L)
#define E_INVALIDARG (-2L)
#define E_OUTOFMEMORY (-3L)
#define E_POINTER (-4L)
#define FAILED(hr) ((hr) < 0)
#undef assert
#define assert(x) ((void)0)
enum { TEX_FILTER_WRAP_U = 0x1, TEX_FILTER_WRAP_V = 0x2, TEX_FILTER_WRAP_W = 0x4 };
struct XMVECTOR { float f[4]; };
struct XMVECTORF32 { float f[4]; operator XMVECTOR() const { return XMVECTOR{{f[0], f[1], f[2], f[3]}}; } };
#define XMGLOBALCONST static const
inline XMVECTOR XMVectorAdd(XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[0] + b.f[0], a.f[1] + b.f[1], a.f[2] + b.f[2], a.f[3] + b.f[3]}}; }
inline XMVECTOR XMVectorMultiply(XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[0] * b.f[0], a.f[1] * b.f[1], a.f[2] * b.f[2], a.f[3] * b.f[3]}}; }
inline XMVECTOR operator+(XMVECTOR a, XMVECTOR b) { return XMVectorAdd(a, b); }
inline XMVECTOR operator*(XMVECTOR a, float s) { return XMVECTOR{{a.f[0] * s, a.f[1] * s, a.f[2] * s, a.f[3] * s}}; }
inline XMVECTOR operator*(float s, XMVECTOR a) { return XMVECTOR{{a.f[0] * s, a.f[1] * s, a.f[2] * s, a.f[3] * s}}; }
// zero-filled: the box generator reads a scanline buffer it never loaded when level 0 is one texel high (see
// csrc/mips_f16.cuh); the canonical model defines never-written scratch as zero (rule F6)
inline void* _aligned_malloc(size_t n, size_t a) { void* p = nullptr; if (posix_memalign(&p, a, n)) return nullptr; memset(p, 0, n); return p; }
struct aligned_deleter { void operator()(void* p) { free(p); } };
typedef std::unique_ptr<XMVECTOR, aligned_deleter> ScopedAlignedArrayXMVECTOR;
struct MipMeta { size_t width, height; };
struct MipChain {                      // stands in for DirectX::ScratchImage inside the two generators
    std::vector<Image> images; MipMeta meta;
    const Image* GetImages() const { return images.data(); }
    const MipMeta& GetMetadata() const { return meta; }
    const Image* GetImage(size_t level, size_t, size_t) const { return &images[level]; }
};
#define ScratchImage MipChain
// ---- what XMColorSRGBToRGB / XMColorRGBToSRGB (cut from DirectXTexConvert.cpp below) need from DirectXMath ----
typedef const XMVECTOR FXMVECTOR;
struct XMVECTORU32 { uint32_t u[4]; };
static const XMVECTORU32 g_XMSelect1110 = {{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u}};
static inline float xm_mask(bool on) { uint32_t u = on ? 0xFFFFFFFFu : 0u; float f; memcpy(&f, &u, 4); return f; }
static inline bool xm_on(float f) { uint32_t u; memcpy(&u, &f, 4); return u != 0; }
inline XMVECTOR XMVectorSaturate(XMVECTOR v) { XMVECTOR r; for (int i = 0; i < 4; i++) { float x = v.f[i]; x = (x > 0.0f) ? x : 0.0f; r.f[i] = (x < 1.0f) ? x : 1.0f; } return r; }
inline XMVECTOR XMVectorPow(XMVECTOR a, XMVECTOR b) { XMVECTOR r; for (int i = 0; i < 4; i++) r.f[i] = powf(a.f[i], b.f[i]); return r; }   // DirectXMath: powf per component
inline XMVECTOR XMVectorGreater(XMVECTOR a, XMVECTOR b) { XMVECTOR r; for (int i = 0; i < 4; i++) r.f[i] = xm_mask(a.f[i] > b.f[i]); return r; }
inline XMVECTOR XMVectorLess(XMVECTOR a, XMVECTOR b) { XMVECTOR r; for (int i = 0; i < 4; i++) r.f[i] = xm_mask(a.f[i] < b.f[i]); return r; }
inline XMVECTOR XMVectorSelect(XMVECTOR a, XMVECTOR b, XMVECTOR c) { XMVECTOR r; for (int i = 0; i < 4; i++) r.f[i] = xm_on(c.f[i]) ? b.f[i] : a.f[i]; return r; }   // all-or-nothing masks only
inline XMVECTOR XMVectorSelect(XMVECTOR a, XMVECTOR b, const XMVECTORU32& c) { XMVECTOR r; for (int i = 0; i < 4; i++) r.f[i] = c.u[i] ? b.f[i] : a.f[i]; return r; }
inline XMVECTOR operator*(XMVECTOR a, XMVECTOR b) { return XMVectorMultiply(a, b); }
inline XMVECTOR operator-(XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[0] - b.f[0], a.f[1] - b.f[1], a.f[2] - b.f[2], a.f[3] - b.f[3]}}; }
//@SRGB_FUNCTIONS@
enum { TEX_FILTER_SRGB_IN = 0x1000000, TEX_FILTER_SRGB_OUT = 0x2000000, TEX_FILTER_SRGB = 0x3000000 };
// Scanline load / store (DirectXTexConvert.cpp:2688-2830).  The format dispatch is restated; the per-format conversions are
//   R16G16B16A16_FLOAT  XMLoadHalf4 / XMStoreHalf4            (restated above)
//   R8G8B8A8_UNORM      XMLoadUByteN4: byte * (1/255)         (rule F7: DirectXMath, not in the tree)
//                       XMStoreUByteN4: saturate, * 255, round to nearest (x + 0.5 truncated) -- the store rounding of this
//                       DirectXMath version is outside the tree: ASSUMED, as for the decoders (tests/test_decode.py)
//   ..._UNORM_SRGB      the same bytes with XMColorSRGBToRGB after the load and XMColorRGBToSRGB before the store -- the
//                       reference's own definitions (DirectXTexConvert.cpp:2669-2685, :2757-2775), cut above
static inline float unorm8_load(uint8_t b) { return (float)b * (1.0f / 255.0f); }
static inline uint8_t unorm8_store(float v) { v = (v > 0.0f) ? v : 0.0f; v = (v < 1.0f) ? v : 1.0f; return (uint8_t)(int)(v * 255.0f + 0.5f); }
static bool _LoadScanlineLinear(XMVECTOR* d, size_t count, LPCVOID src, size_t, DXGI_FORMAT format, DWORD flags)
{
    if (format == DXGI_FORMAT_R16G16B16A16_FLOAT) {
        const unsigned short* s = static_cast<const unsigned short*>(src);
        for (size_t i = 0; i < count; i++)
            for (int c = 0; c < 4; c++) d[i].f[c] = PackedVector::XMConvertHalfToFloat(s[4 * i + c]);
        return true;
    }
    if (format == DXGI_FORMAT_R8G8B8A8_UNORM_SRGB) flags |= TEX_FILTER_SRGB;
    const uint8_t* s = static_cast<const uint8_t*>(src);
    for (size_t i = 0; i < count; i++) {
        for (int c = 0; c < 4; c++) d[i].f[c] = unorm8_load(s[4 * i + c]);
        if (flags & TEX_FILTER_SRGB_IN) d[i] = XMColorSRGBToRGB(d[i]);
    }
    return true;
}
static bool _StoreScanlineLinear(LPVOID dst, size_t, DXGI_FORMAT format, XMVECTOR* s, size_t count, DWORD flags)
{
    if (format == DXGI_FORMAT_R16G16B16A16_FLOAT) {
        unsigned short* d = static_cast<unsigned short*>(dst);
        for (size_t i = 0; i < count; i++)
            for (int c = 0; c < 4; c++) d[4 * i + c] = PackedVector::XMConvertFloatToHalf(s[i].f[c]);
        return true;
    }
    if (format == DXGI_FORMAT_R8G8B8A8_UNORM_SRGB) flags |= TEX_FILTER_SRGB;
    uint8_t* d = static_cast<uint8_t*>(dst);
    for (size_t i = 0; i < count; i++) {
        if (flags & TEX_FILTER_SRGB_OUT) s[i] = XMColorRGBToSRGB(s[i]);
        for (int c = 0; c < 4; c++) d[4 * i + c] = unorm8_store(s[i].f[c]);
    }
    return true;
}
inline static bool ispow2(size_t x) { return ((x != 0) && !(x & (x - 1))); }
"""

MIP_SHIM_BOTTOM = r"""
#undef ScratchImage
// level 0 (w x h RGBA16F, tight) -> `levels` tightly packed levels written one after the other into `out`;
// the BOX / LINEAR choice restates GenerateMipMaps, DirectXTexMipmaps.cpp:2611-2650
extern "C" int ref_mip_chain_f16(const unsigned short* level0, int w, int h, int levels, unsigned short* out)
{
    MipChain chain;
    chain.meta = MipMeta{(size_t)w, (size_t)h};
    size_t off = 0;
    for (int l = 0; l < levels; l++) {
        const size_t lw = (w >> l) ? (w >> l) : 1, lh = (h >> l) ? (h >> l) : 1;
        chain.images.push_back(Image{lw, lh, DXGI_FORMAT_R16G16B16A16_FLOAT, lw * 8, lw * lh * 8, reinterpret_cast<uint8_t*>(out + off)});
        off += lw * lh * 4;
    }
    memcpy(out, level0, (size_t)w * h * 8);
    if (levels < 2) return 0;
    const bool box = ispow2((size_t)w) && ispow2((size_t)h);
    return (int)(box ? _Generate2DMipsBoxFilter((size_t)levels, 0, chain, 0) : _Generate2DMipsLinearFilter((size_t)levels, 0, chain, 0));
}
// RGBA8 chain through the SAME two generators: srgb = 0 -> R8G8B8A8_UNORM, 1 -> R8G8B8A8_UNORM_SRGB (what the plug-in's scratch
// image is for *_SRGB encodings, IntelPlugin.cpp:152-154; _UseWICFiltering then returns false, DirectXTexMipmaps.cpp:389-393)
extern "C" int ref_mip_chain_rgba8(const uint8_t* level0, int w, int h, int levels, int srgb, uint8_t* out)
{
    MipChain chain;
    chain.meta = MipMeta{(size_t)w, (size_t)h};
    const DXGI_FORMAT fmt = srgb ? DXGI_FORMAT_R8G8B8A8_UNORM_SRGB : DXGI_FORMAT_R8G8B8A8_UNORM;
    size_t off = 0;
    for (int l = 0; l < levels; l++) {
        const size_t lw = (w >> l) ? (w >> l) : 1, lh = (h >> l) ? (h >> l) : 1;
        chain.images.push_back(Image{lw, lh, fmt, lw * 4, lw * lh * 4, out + off});
        off += lw * lh * 4;
    }
    memcpy(out, level0, (size_t)w * h * 4);
    if (levels < 2) return 0;
    const bool box = ispow2((size_t)w) && ispow2((size_t)h);
    return (int)(box ? _Generate2DMipsBoxFilter((size_t)levels, 0, chain, 0) : _Generate2DMipsLinearFilter((size_t)levels, 0, chain, 0));
}
// the scalar pieces, for the table generator (tools/gen_srgb_tables.py) and the tests
extern "C" float ref_srgb_to_linear(float v) { return XMColorSRGBToRGB(XMVECTOR{{v, v, v, v}}).f[0]; }
extern "C" int ref_linear_to_srgb8(float v) { return unorm8_store(XMColorRGBToSRGB(XMVECTOR{{v, v, v, v}}).f[0]); }
extern "C" float ref_unorm8_load(int b) { return unorm8_load((uint8_t)b); }
extern "C" int ref_unorm8_store(float v) { return unorm8_store(v); }
"""


BC45_SHIM_TOP = r"""
// ---- DirectXTex BC4 / BC5 encoders: shim -----------------------------------------------------------------
#define _Out_
#define _Inout_
#define _Use_decl_annotations_
#define UNREFERENCED_PARAMETER(x) ((void)(x))
#define NUM_PIXELS_PER_BLOCK 16
#define _isnan(x) ((x) != (x))            // MSVC CRT name, used by the SNORM helpers
struct XMFLOAT4A { float x, y, z, w; };
inline float XMVectorGetX(XMVECTOR v) { return v.f[0]; }
inline XMVECTOR XMVectorSet(float x, float y, float z, float w) { return XMVECTOR{{x, y, z, w}}; }
inline void XMStoreFloat4A(XMFLOAT4A* d, XMVECTOR v) { d->x = v.f[0]; d->y = v.f[1]; d->z = v.f[2]; d->w = v.f[3]; }
namespace DirectX {
"""

BC45_SHIM_BOTTOM = r"""
// `texels` = 16 floats per channel, already byte * (1/255)
extern "C" void ref_encode_bc4u(const float* r, uint8_t* out8)
{
    XMVECTOR c[16];
    for (int i = 0; i < 16; i++) c[i] = XMVECTOR{{r[i], 0.0f, 0.0f, 1.0f}};
    DirectX::D3DXEncodeBC4U(out8, c, 0);
}
extern "C" void ref_encode_bc5u(const float* r, const float* g, uint8_t* out16)
{
    XMVECTOR c[16];
    for (int i = 0; i < 16; i++) c[i] = XMVECTOR{{r[i], g[i], 0.0f, 1.0f}};
    DirectX::D3DXEncodeBC5U(out16, c, 0);
}
// decoded channel values as floats (D3DXDecodeBC4U, BC4BC5.cpp:373-386): 16 floats
extern "C" void ref_decode_bc4u(const uint8_t* in8, float* out16)
{
    XMVECTOR c[16];
    DirectX::D3DXDecodeBC4U(c, in8);
    for (int i = 0; i < 16; i++) out16[i] = c[i].f[0];
}
"""


def cut_bc45(cpp, bch):
    m = re.search(r"^template\s*<bool bRange>\s*void OptimizeAlpha\s*\(", bch, re.M)
    if not m:
        raise RuntimeError("OptimizeAlpha not found in BC.h")
    optimize = cut_function(bch, m.start())
    a = cpp.index("namespace DirectX")
    a = cpp.index("{", a) + 1
    b = cpp.rindex("} // namespace")
    return optimize + "\n" + cpp[a:b] + "\n}  // namespace DirectX\n"


BC13_SHIM_TOP = r"""
// ---- DirectXTex BC1 / BC3 decoders: shim -----------------------------------------------------------------
#define _In_
struct XMU565 { uint16_t v; };
static const XMVECTORF32 g_XMIdentityR3 = {0.0f, 0.0f, 0.0f, 1.0f};      // XMVECTORU32 / g_XMSelect1110 / XMVectorSelect: the mip shim above
inline XMVECTOR XMLoadU565(const XMU565* p) { return XMVECTOR{{(float)(p->v & 31), (float)((p->v >> 5) & 63), (float)((p->v >> 11) & 31), 0.0f}}; }
template <int A, int B, int C, int D> inline XMVECTOR XMVectorSwizzle(XMVECTOR v) { return XMVECTOR{{v.f[A], v.f[B], v.f[C], v.f[D]}}; }
inline XMVECTOR XMVectorLerp(XMVECTOR v0, XMVECTOR v1, float t)          // DirectXMath: Length = V1 - V0; Length * t + V0 (mul, add)
{
    XMVECTOR r;
    for (int i = 0; i < 4; i++) { float l = v1.f[i] - v0.f[i]; float m = l * t; r.f[i] = m + v0.f[i]; }
    return r;
}
inline XMVECTOR XMVectorZero() { return XMVECTOR{{0.0f, 0.0f, 0.0f, 0.0f}}; }
inline XMVECTOR XMVectorSetW(XMVECTOR v, float w) { v.f[3] = w; return v; }
#pragma pack(push, 1)
"""

BC13_SHIM_BOTTOM = r"""
// 16 texels x (r, g, b, a) floats
extern "C" void ref_decode_bc1(const uint8_t* in8, float* out64)
{
    XMVECTOR c[16];
    D3DXDecodeBC1(c, in8);
    for (int i = 0; i < 16; i++) for (int k = 0; k < 4; k++) out64[4 * i + k] = c[i].f[k];
}
extern "C" void ref_decode_bc3(const uint8_t* in16, float* out64)
{
    XMVECTOR c[16];
    D3DXDecodeBC3(c, in16);
    for (int i = 0; i < 16; i++) for (int k = 0; k < 4; k++) out64[4 * i + k] = c[i].f[k];
}
"""


def cut_bc13(cpp, bch):
    a = bch.index("struct D3DX_BC1")
    b = bch.index("#pragma pack(pop)", a)
    structs = bch[a:b] + "#pragma pack(pop)\n"
    out = [structs]
    for pat in (r"^inline static void DecodeBC1\s*\(", r"^void D3DXDecodeBC1\s*\(", r"^void D3DXDecodeBC3\s*\("):
        m = re.search(pat, cpp, re.M)
        if not m:
            raise RuntimeError(pat + " not found in BC.cpp")
        out.append(cut_function(cpp, m.start()))
    return "\n\n".join(out)


def cut_srgb(conv):
    """XMColorRGBToSRGB / XMColorSRGBToRGB as DirectXTexConvert.cpp defines them (its DIRECTX_MATH_VERSION < 306 branch; DirectXMath
    3.06 ships the same two functions)."""
    out = []
    for name in ("XMColorRGBToSRGB", "XMColorSRGBToRGB"):
        m = re.search(r"^static inline XMVECTOR " + name + r"\s*\(", conv, re.M)
        if not m:
            raise RuntimeError(name + " not found in DirectXTexConvert.cpp")
        out.append(cut_function(conv, m.start()))
    return "\n".join(out)


def cut_static_function(src, name):
    m = re.search(r"^static\s+HRESULT\s+" + name + r"\s*\(", src, re.M)
    if not m:
        raise RuntimeError(f"{name} not found in the reference")
    return cut_function(src, m.start())


def cut_filters(hdr):
    a = hdr.index("// Box filtering helpers")
    b = hdr.index("// Cubic filtering helpers")
    a = hdr.rfind("//----", 0, a)
    b = hdr.rfind("//----", 0, b)
    return hdr[a:b]


def cut_function(src, start):
    """Text of the function whose signature starts at `start` (through its matching closing brace)."""
    i = src.index("{", start)
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[start:j + 1]
        j += 1


def cut_members(cpp):
    out = []
    for name in MEMBERS:
        m = re.search(r"^[A-Za-z_][\w \t\*&:]*\bIntelPlugin::" + name + r"\s*\(", cpp, re.M)
        if not m:
            raise RuntimeError(f"IntelPlugin::{name} not found in the reference")
        out.append(cut_function(cpp, m.start()))
    return "\n\n".join(out)


def cut_inlines(hdr):
    out = []
    for m in re.finditer(r"^inline\s+[\w \t]+?\b(\w+)\s*\([^)]*\)\s*$", hdr, re.M):
        if m.group(1) in INLINES or m.group(1) in ("ConvertTo8Bit", "ConvertTo16Bit"):
            out.append(cut_function(hdr, m.start()))
    if len(out) < 10:
        raise RuntimeError("conversion helpers not found in IntelPlugin.h")
    return "\n\n".join(out)


def build(verbose=True):
    have_ref = os.path.exists(PLUGIN_CPP)
    if os.path.exists(OUT_SO) and not have_ref:
        return OUT_SO                      # GPU box: prebuilt .so travelled with the snapshot
    if not have_ref:
        raise FileNotFoundError(f"reference not present at {REF_ROOT} and no prebuilt {OUT_SO}")
    if os.path.exists(OUT_SO) and os.path.getmtime(OUT_SO) > max(os.path.getmtime(__file__), os.path.getmtime(PLUGIN_CPP)):
        return OUT_SO
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="itw_ref_front_")
    try:
        hdr = open(PLUGIN_H, encoding="utf-8", errors="replace").read()
        cpp = open(PLUGIN_CPP, encoding="utf-8", errors="replace").read()
        mips = open(MIPMAPS_CPP, encoding="utf-8", errors="replace").read()
        filt = open(FILTERS_H, encoding="utf-8", errors="replace").read()
        conv = open(CONVERT_CPP, encoding="utf-8", errors="replace").read()
        unit = (SHIM_TOP + cut_inlines(hdr) + SHIM_CLASS + cut_members(cpp) + SHIM_BOTTOM + MIP_SHIM_TOP.replace("//@SRGB_FUNCTIONS@", cut_srgb(conv))
                + cut_filters(filt)
                + cut_static_function(mips, "_Generate2DMipsBoxFilter") + "\n" + cut_static_function(mips, "_Generate2DMipsLinearFilter")
                + MIP_SHIM_BOTTOM
                + BC45_SHIM_TOP + cut_bc45(open(BC45_CPP, encoding="utf-8", errors="replace").read(),
                                           open(BC_H, encoding="utf-8", errors="replace").read()) + BC45_SHIM_BOTTOM
                + BC13_SHIM_TOP + cut_bc13(open(BC_CPP, encoding="utf-8", errors="replace").read(),
                                           open(BC_H, encoding="utf-8", errors="replace").read()) + BC13_SHIM_BOTTOM)
        path = os.path.join(tmp, "frontend_ref.cpp")
        open(path, "w").write(unit)
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-fno-fast-math", "-mfpmath=sse", "-msse2",
               path, "-o", OUT_SO]
        if verbose:
            print("[build_ref_frontend]", " ".join(cmd))
        subprocess.check_call(cmd)
        if os.environ.get("ITW_KEEP_REF_TMP"):
            print("[build_ref_frontend] kept", tmp)
    finally:
        if not os.environ.get("ITW_KEEP_REF_TMP"):
            shutil.rmtree(tmp, ignore_errors=True)
    return OUT_SO


if __name__ == "__main__":
    print(build())
