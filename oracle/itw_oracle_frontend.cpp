/*
 * itw_oracle_frontend.cpp -- CPU ORACLE for the pixel-format front end.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restates, as the whole-image passes the plug-in runs on the host,
 *   IntelPlugin.h:41-96                  FloatToByte, ConvertTo8Bit x3, ConvertTo16Bit x3     (cited IPh:line)
 *   IntelPlugin.cpp:85-181               CopyDataForEncoding (dispatch on format / depth)      (cited IP:line)
 *   IntelPlugin.cpp:291-433, :741-810    ConvertToBC{,4or5,6}From{8,16,32}Bit
 *   IntelPlugin.cpp:1504-1546            FlipXYChannelNormalMap
 *   IntelPlugin.cpp:1551-1612            NormalizeNormalMapChain
 *   IntelPlugin.cpp:892-928              DoPaddingToMultiplesOf4
 * Built into libitw_oracle.so by oracle/Makefile.  Only tests/ and __graft_entry__.smoke() may load it.
 *
 * PINNING.  tests/test_frontend.py compares this restatement byte for byte with oracle/_ref/libitw_ref_frontend.so,
 * i.e. with the reference's OWN function bodies cut from IntelPlugin.h / IntelPlugin.cpp by oracle/build_ref_frontend.py,
 * over every depth x plane count x format family x flag combination.  Two ingredients are NOT in the reference tree
 * and are therefore unpinned beyond that build: DirectXMath's half conversions (restated here from the published
 * DirectXMath 3.06 scalar source, see build_ref_frontend.py) and the C library's pow().
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/itw_bcn.h"

namespace {

uint16_t half_from_float(float value)                  // XMConvertFloatToHalf, DirectXMath 3.06 scalar path
{
    uint32_t bits;
    memcpy(&bits, &value, 4);
    const uint32_t sign = (bits & 0x80000000u) >> 16;
    bits &= 0x7FFFFFFFu;
    uint32_t result;
    if (bits > 0x47FFEFFFu) result = 0x7FFFu;
    else {
        if (bits < 0x38800000u) {
            const uint32_t shift = 113u - (bits >> 23);
            bits = (shift < 32u) ? ((0x800000u | (bits & 0x7FFFFFu)) >> shift) : 0u;
        } else bits += 0xC8000000u;
        result = ((bits + 0x0FFFu + ((bits >> 13) & 1u)) >> 13) & 0x7FFFu;
    }
    return (uint16_t)(result | sign);
}
float float_from_half(uint16_t h)                      // XMConvertHalfToFloat, DirectXMath 3.06 scalar path
{
    uint32_t mant = h & 0x3FFu, exp;
    if (h & 0x7C00u) exp = (h >> 10) & 31u;
    else if (mant) {
        exp = 1;
        do { exp--; mant <<= 1; } while (!(mant & 0x400u));
        mant &= 0x3FFu;
    } else exp = (uint32_t)-112;
    const uint32_t bits = ((uint32_t)(h & 0x8000u) << 16) | ((exp + 112u) << 23) | (mant << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
uint8_t float_to_byte(double v)                        // IPh:41-48; out-of-range cast = cvttsd2si low byte (NaN -> 0)
{
    if (v > 1) return 255;
    if (v < 0) return 0;
    if (v != v) return 0;
    return (uint8_t)(int)(v * 255);
}
uint8_t to8(uint8_t v) { return v; }                                           // IPh:56-59
uint8_t to8(uint16_t v) { return float_to_byte(v / 32768.0); }                 // IPh:60-66
uint8_t to8(float v, bool gamma)                                               // IPh:67-76
{
    double d = v;
    if (gamma) d = pow(d, 1 / 2.2);
    return float_to_byte(d);
}
uint16_t to16(uint8_t v) { return half_from_float(v / 255.f); }                // IPh:79-82
uint16_t to16(uint16_t v) { return half_from_float((float)(v / 32768.0)); }   // IPh:83-89
uint16_t to16(float v) { return half_from_float(v); }                          // IPh:90-98

template <class T>
const T* element(const itw_pixel_source* s, long long row_bytes, int x, int y)
{
    return reinterpret_cast<const T*>(static_cast<const uint8_t*>(s->data) + y * row_bytes) + (long long)x * s->planes;
}

}  // namespace

extern "C" int oracle_itw_convert_pixels(int format, const itw_pixel_source* s, uint32_t flags, const rgba_surface* dst)
{
    const int w = s->width, h = s->height, planes = s->planes;
    const bool alpha = flags & ITW_FRONT_HAS_ALPHA, gamma = flags & ITW_FRONT_GAMMA;
    const bool hdr = format == ITW_FORMAT_BC6H, copy0 = (format == ITW_FORMAT_BC4 || format == ITW_FORMAT_BC5);
    const long long row_bytes = s->row_bytes ? s->row_bytes : (long long)w * planes * (s->depth / 8);
    const int texel = hdr ? 8 : 4;
    std::vector<uint8_t> top((size_t)w * h * texel);

    // pass 1: CopyDataForEncoding, IP:85-181
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (hdr) {                                                         // IP:291-366
                uint16_t* t = reinterpret_cast<uint16_t*>(top.data()) + ((size_t)y * w + x) * 4;
                for (int c = 0; c < 4; c++) {
                    // alpha: plane 3, except the 32-bit converter, which reads plane 2 (IP:361)
                    const int sc = (c == 3 && s->depth == 32) ? 2 : c;
                    const bool present = (c == 3) ? alpha : (c < planes);
                    uint16_t v = (c == 3) ? half_from_float(1.f) : 0;
                    if (present) {
                        if (s->depth == 8) v = to16(element<uint8_t>(s, row_bytes, x, y)[sc]);
                        else if (s->depth == 16) v = to16(element<uint16_t>(s, row_bytes, x, y)[sc]);
                        else v = to16(element<float>(s, row_bytes, x, y)[sc]);
                    }
                    t[c] = v;
                }
            } else {                                                           // IP:741-810, :368-433
                uint8_t* t = top.data() + ((size_t)y * w + x) * 4;
                for (int c = 0; c < 4; c++) {
                    const bool present = (c == 3) ? alpha : (c < planes);
                    uint8_t v = (c == 3) ? 255 : (copy0 ? t[0] : 0);
                    if (present) {
                        if (s->depth == 8) v = to8(element<uint8_t>(s, row_bytes, x, y)[c]);
                        else if (s->depth == 16) v = to8(element<uint16_t>(s, row_bytes, x, y)[c]);
                        else v = to8(element<float>(s, row_bytes, x, y)[c], gamma);
                    }
                    t[c] = v;
                }
            }
        }
    // pass 2: FlipXYChannelNormalMap, IP:1504-1546
    if (flags & (ITW_FRONT_FLIP_X | ITW_FRONT_FLIP_Y))
        for (size_t i = 0; i < (size_t)w * h; i++) {
            if (hdr) {
                uint16_t* t = reinterpret_cast<uint16_t*>(top.data()) + i * 4;
                const float r = float_from_half(t[0]), g = float_from_half(t[1]);
                if (flags & ITW_FRONT_FLIP_X) t[0] = half_from_float(1.f - r);
                if (flags & ITW_FRONT_FLIP_Y) t[1] = half_from_float(1.f - g);
            } else {
                uint8_t* t = top.data() + i * 4;
                if (flags & ITW_FRONT_FLIP_X) t[0] = 255 - t[0];
                if (flags & ITW_FRONT_FLIP_Y) t[1] = 255 - t[1];
            }
        }
    // pass 3: NormalizeNormalMapChain, IP:1551-1612
    if (flags & ITW_FRONT_NORMALIZE)
        for (size_t i = 0; i < (size_t)w * h; i++) {
            if (hdr) {
                uint16_t* t = reinterpret_cast<uint16_t*>(top.data()) + i * 4;
                const float r = float_from_half(t[0]), g = float_from_half(t[1]), b = float_from_half(t[2]);
                float m = sqrtf(r * r + g * g + b * b);
                if (m > 0) {
                    m = 1.0f / m;
                    t[0] = half_from_float(r * m); t[1] = half_from_float(g * m); t[2] = half_from_float(b * m);
                } else { t[0] = half_from_float(0); t[1] = half_from_float(0); t[2] = half_from_float(1); }
            } else {
                uint8_t* t = top.data() + i * 4;
                const float r = (float)(t[0] - 128), g = (float)(t[1] - 128), b = (float)(t[2] - 128);
                float m = sqrtf(r * r + g * g + b * b);
                if (m > 0) {
                    m = 127 / m;
                    t[0] = (uint8_t)(int)(r * m + 128); t[1] = (uint8_t)(int)(g * m + 128); t[2] = (uint8_t)(int)(b * m + 128);
                } else { t[0] = 128; t[1] = 128; t[2] = 255; }
            }
        }
    // pass 4: DoPaddingToMultiplesOf4, IP:892-928 (or a plain copy when dst has the source size)
    for (int y = 0; y < dst->height; y++) {
        const int sy = y < h ? y : h - 1;
        for (int x = 0; x < dst->width; x++) {
            const int sx = x < w ? x : w - 1;
            memcpy(dst->ptr + (size_t)y * dst->stride + (size_t)x * texel, top.data() + ((size_t)sy * w + sx) * texel, texel);
        }
    }
    return 0;
}

// =====================================================================================================================
// RGBA16F mip chain of the BC6H save path: DirectXTex's non-WIC generators, restated in their scanline form.
//   GenerateMipMaps                 DirectXTex/DirectXTexMipmaps.cpp:2611-2650 (BOX when both sizes are powers of two, else LINEAR)
//   _Generate2DMipsBoxFilter        :715-805, AVERAGE4 Filters.h:33-39
//   _Generate2DMipsLinearFilter     :809-905, _CreateLinearFilter / BILINEAR_INTERPOLATE Filters.h:60-112
// Pinned by tests/test_mips_f16.py against the reference's own function bodies (oracle/build_ref_frontend.py).
// `out` receives `levels` tightly packed levels one after the other.
// =====================================================================================================================
namespace {
struct Tap { size_t u0, u1; float w0, w1; };
void linear_taps(size_t source, size_t dest, std::vector<Tap>& lf)                 // Filters.h:67-100, clamp addressing
{
    lf.resize(dest);
    const float scale = float(source) / float(dest);
    for (size_t u = 0; u < dest; ++u) {
        const float srcB = (float(u) + 0.5f) * scale + 0.5f;
        ptrdiff_t isrcB = ptrdiff_t(srcB), isrcA = isrcB - 1;
        if (isrcA < 0) isrcA = 0;
        if (size_t(isrcB) >= source) isrcB = ptrdiff_t(source) - 1;
        const float weight = 1.0f + float(isrcB) - srcB;
        lf[u] = Tap{size_t(isrcA), size_t(isrcB), weight, 1.0f - weight};
    }
}
// ---- texel codecs of the three chains ----
// RGBA16F: XMLoadHalf4 / XMStoreHalf4 (float_from_half / half_from_float above).
// RGBA8 UNORM: XMLoadUByteN4 = byte * (1/255) (rule F7); XMStoreUByteN4 = saturate, * 255, round to nearest -- ASSUMED, the store's
//   rounding is DirectXMath's (not in the tree); the decoder tests make the same assumption.
// RGBA8 UNORM_SRGB: the same bytes with XMColorSRGBToRGB after the load and XMColorRGBToSRGB before the store, restated from
//   DirectXTexConvert.cpp:2669-2685, :2757-2775 (float constants 1.f/12.92f ..., powf per component, alpha untouched).
inline float saturate_f(float v) { v = (v > 0.0f) ? v : 0.0f; return (v < 1.0f) ? v : 1.0f; }
inline float srgb_to_linear(float s)
{
    const float v = saturate_f(s);
    const float v0 = v * (1.f / 12.92f);
    const float v1 = powf((v + 0.055f) * (1.f / 1.055f), 2.4f);
    return (v > 0.04045f) ? v1 : v0;
}
inline float linear_to_srgb(float l)
{
    const float v = saturate_f(l);
    const float v0 = v * 12.92f;
    const float v1 = 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
    return (v < 0.0031308f) ? v0 : v1;
}
inline uint8_t unorm8_store(float v) { return (uint8_t)(int)(saturate_f(v) * 255.0f + 0.5f); }

struct CodecF16 {
    typedef uint16_t T;
    static void load(std::vector<float>& row, const T* src, size_t width) { for (size_t i = 0; i < width * 4; i++) row[i] = float_from_half(src[i]); }
    static T store(float v, int) { return half_from_float(v); }
};
struct CodecUnorm8 {
    typedef uint8_t T;
    static void load(std::vector<float>& row, const T* src, size_t width) { for (size_t i = 0; i < width * 4; i++) row[i] = (float)src[i] * (1.0f / 255.0f); }
    static T store(float v, int) { return unorm8_store(v); }
};
struct CodecSrgb8 {
    typedef uint8_t T;
    static void load(std::vector<float>& row, const T* src, size_t width)
    {
        for (size_t i = 0; i < width * 4; i++) {
            const float f = (float)src[i] * (1.0f / 255.0f);
            row[i] = ((i & 3) == 3) ? f : srgb_to_linear(f);
        }
    }
    static T store(float v, int c) { return unorm8_store(c == 3 ? v : linear_to_srgb(v)); }
};

// The two generators in their scanline form; `out` receives `levels` tightly packed levels one after the other.
template <class Codec>
int mip_chain(const typename Codec::T* level0, int w, int h, int levels, typename Codec::T* out)
{
    typedef typename Codec::T T;
    memcpy(out, level0, (size_t)w * h * 4 * sizeof(T));
    const bool box = w > 0 && h > 0 && !(w & (w - 1)) && !(h & (h - 1));
    size_t width = (size_t)w, height = (size_t)h;
    const T* src = out;
    T* dst = out + width * height * 4;
    // the box loop's scanline buffers live across levels (:730-739); `second` is only reloaded while the source has two rows,
    // and the fourth tap keeps reading it when the height reaches one before the width does (reference behaviour)
    std::vector<float> first((size_t)w * 4, 0.0f), second((size_t)w * 4, 0.0f), r0((size_t)w * 4), r1((size_t)w * 4);
    std::vector<Tap> lx, ly;
    for (int level = 1; level < levels; level++) {
        const size_t nwidth = width > 1 ? width >> 1 : 1, nheight = height > 1 ? height >> 1 : 1;
        if (box) {
            for (size_t y = 0; y < nheight; y++) {
                Codec::load(first, src + (height > 1 ? 2 * y : y) * width * 4, width);
                if (height > 1) Codec::load(second, src + (2 * y + 1) * width * 4, width);
                const std::vector<float>& u1 = height > 1 ? second : first;        // urow1
                for (size_t x = 0; x < nwidth; x++) {
                    const size_t x2 = x << 1;
                    for (int c = 0; c < 4; c++) {
                        const float p0 = first[x2 * 4 + c], p1 = u1[x2 * 4 + c];
                        const float p2 = width > 1 ? first[(x2 + 1) * 4 + c] : first[x2 * 4 + c];            // urow2
                        const float p3 = width > 1 ? second[(x2 + 1) * 4 + c] : u1[x2 * 4 + c];              // urow3 (stale when height <= 1)
                        float v = p0 + p1;
                        v = v + p2;
                        v = v + p3;
                        dst[(y * nwidth + x) * 4 + c] = Codec::store(v * 0.25f, c);
                    }
                }
            }
        } else {
            linear_taps(width, nwidth, lx);
            linear_taps(height, nheight, ly);
            for (size_t y = 0; y < nheight; y++) {
                Codec::load(r0, src + ly[y].u0 * width * 4, width);
                Codec::load(r1, src + ly[y].u1 * width * 4, width);
                for (size_t x = 0; x < nwidth; x++)
                    for (int c = 0; c < 4; c++) {
                        const float a = r0[lx[x].u0 * 4 + c] * lx[x].w0 + r0[lx[x].u1 * 4 + c] * lx[x].w1;
                        const float b = r1[lx[x].u0 * 4 + c] * lx[x].w0 + r1[lx[x].u1 * 4 + c] * lx[x].w1;
                        dst[(y * nwidth + x) * 4 + c] = Codec::store((ly[y].w0 * a) + (ly[y].w1 * b), c);
                    }
            }
        }
        src = dst;
        dst += nwidth * nheight * 4;
        width = nwidth;
        height = nheight;
    }
    return 0;
}
}  // namespace

extern "C" int oracle_mip_chain_f16(const uint16_t* level0, int w, int h, int levels, uint16_t* out) { return mip_chain<CodecF16>(level0, w, h, levels, out); }
// RGBA8 chain of the LDR save path, non-WIC generators: srgb = 1 is what the plug-in gets for *_SRGB encodings (IntelPlugin.cpp:152-154
// overrides the scratch format, _UseWICFiltering then returns false, DirectXTexMipmaps.cpp:389-393); srgb = 0 is the same code on
// R8G8B8A8_UNORM (TEX_FILTER_FORCE_NON_WIC) -- the plug-in's default for UNORM encodings is WIC, which is outside the tree.
extern "C" int oracle_mip_chain_rgba8(const uint8_t* level0, int w, int h, int levels, int srgb, uint8_t* out)
{
    return srgb ? mip_chain<CodecSrgb8>(level0, w, h, levels, out) : mip_chain<CodecUnorm8>(level0, w, h, levels, out);
}
// scalar pieces for tools/gen_srgb_tables.py (the product's tables are derived from THESE functions, i.e. from the C library's powf)
extern "C" float oracle_srgb_to_linear(float v) { return srgb_to_linear(v); }
extern "C" int oracle_linear_to_srgb8(float v) { return unorm8_store(linear_to_srgb(v)); }

// ConvertTo8Bit(double, gammaCorrect = true) of one float, exported for tools/gen_gamma_table.py (which derives the product's
// threshold table from THIS function, i.e. from the C library's pow) and for tests.
extern "C" int oracle_gamma_byte(float v) { return to8(v, true); }
