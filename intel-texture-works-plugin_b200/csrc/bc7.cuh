// bc7.cuh -- BC7 encoder (reference: kernel.ispc:616-2037, cited as K:line).
//
// Mapping.  One WARP owns kBc7Super = 16 consecutive 4x4 blocks per round (kBc7Batch = 8 on surfaces too small to give
// every SM such a CTA).  The shape / ranking phases hold one HALF of eight blocks at a time, as two GROUPS of kBc7Slots = 4
// that reuse the per-group scratch; the chain phase -- 14 to 18 roles per block, a quarter of the work -- runs ONCE for all
// sixteen blocks, two lanes per block, so that a pass is two equally long roles wide (DESIGN.md section 4 has the numbers).
// The 32 lanes are spread over the blocks' independent work items instead of over texels:
//   * shape phases: lane <-> (block slot, partition shape).  A lane fits the PCA segments of its
//     shape ONCE and evaluates both BC7 modes that use that shape (0 and 2 share the three-subset
//     shapes, 1 and 3 the two-subset shapes; the reference refits per mode, K:1279-1297, with
//     identical results because the fit does not depend on the mode);
//   * ranking phase: lane <-> (block slot, two-subset shape) for the 64 PCA split bounds;
//   * chain phase: lane <-> (block slot, role).  The per-mode refinement chains, the mode 4/5
//     rotation / index-swap candidates and mode 6 are independent of one another in the reference
//     (they only meet in a strict `<` on the final error, K:1358, :1638, :1650, :1684), so they
//     all run side by side through ONE generic chain routine, and the block's winner is the first
//     minimum in the reference's order: mode 0, 2, 1, 3, 7, 4(rotation, swap), 5(rotation), 6.
//
// Exactness.  BC7 texels are 8-bit integers, so most of the reference's float arithmetic is exact
// integer arithmetic in disguise.  Wherever every intermediate is an integer below 2^24 the kernel
// uses packed-byte integer instructions instead -- same values, ~6x fewer instructions:
//   * raw moments of a subset (K:763-803): IDP.4A dot products over channel-planar packed bytes;
//   * index search (K:1133-1193): the projection numerator / denominator are integer dot
//     products (then ONE float division, evaluated as the FMA-corrected quotient that equals the
//     IEEE one -- tests/test_exact_division.py, tests/test_gpu_division.py; the two-bit search of the shape phases by two
//     integer thresholds instead), the two candidate palette entries are integer interpolations (the reference truncates
//     them through int, K:1172-1173) and their squared errors are VABSDIFF4 + IDP.4A;
//   * least-squares sums (K:1198-1230): IDP.4A over index bytes.
// Everything that rounds (covariance, power iteration, endpoint solve, quantisation) follows the
// reference's float expression order exactly (DESIGN.md "Canonical float model").
//
// Structure for speed: the non-inlined routines exchange their results BY VALUE (small structs stay
// in registers across calls; arrays passed by pointer would live in local memory), the inner routines are branch-free
// (rotation views via PRMT selectors, the fourth component via zero operands), and chain tasks are scheduled role-major.
//
// The phases are per-lane functions separated by barriers; tests/emu drives the same
// functions lane by lane on the CPU (test-only).
#pragma once
#include "bc67_core.cuh"
#include "itw_masks3.cuh"

// unroll factors of the two hottest loops (tuned on B200, tools/tune_unroll.sh)
#ifndef ITW_BC7_ASSIGN_UNROLL
#define ITW_BC7_ASSIGN_UNROLL 16
#endif
#ifndef ITW_BC7_POWER_UNROLL
#define ITW_BC7_POWER_UNROLL 8
#endif

namespace itw {

// distinct subset masks of the three-subset shapes and the mask ids of each shape (tools/gen_mask_tables.py)
ITW_TABLE_DECL(uint16_t, mask3_unique, ITW_MASK3_COUNT)
ITW_TABLE_DECL(uint8_t, shape3_mask_id, 64 * 3)

// bc7_enc_settings (ispc_texcomp.h:27-41) flattened to ints for the device
struct Bc7Params {
    int sel[4];
    int refine[8];
    int skip2, t1, t3, t7, ch0, rch, channels;
};

constexpr int kBc7Slots = 4;          // blocks per group (ranking / shape phases share per-group scratch)
constexpr int kBc7Batch = 8;          // blocks per HALF: what the shape phases hold in Bc7Warp::blk at a time (two groups)
constexpr int kBc7Super = 16;         // blocks per warp and round when the surface is large: two halves, ONE chain phase
constexpr int kBc7MaxRoles = 18;      // 5 partitioned modes + up to 8 mode-4 + 4 mode-5 + mode 6
constexpr int kErrNone = 0x7fffffff;  // "no result": loses every strict < comparison

// One 4x4 block in shared memory
struct Bc7Block {
    u32 tex[16];       // packed RGBA8 of texel k (R in byte 0)
    u32 plane[4][4];   // plane[c][i] = channel c of texels 4i..4i+3, one byte each
    u32 rgb[16];       // tex[k] & 0x00FFFFFF: what the three-channel index searches read (three quarters of all searches)
    u32 pad;           // 49 words: word k of the eight blocks of a half lands in eight different banks (in the chain phase
                       // adjacent lanes hold different blocks; with 32 words every block read was a bank conflict)
};
// Per-warp scratch in shared memory
struct Bc7Warp {
    Bc7Block blk[kBc7Batch];                   // the half the shape phases are working on
    union {
        // per-GROUP scratch (slot = block index inside the group), reused by every group in turn
        int cand_err[kBc7Slots][2][64];        // errors of the mode-slot pair being evaluated: [first/second][list position]
        int keys[kBc7Slots][64];               // split-bound keys of the set being ranked: dead once `order` is written, and
                                               // the candidate errors of the previous pair have been consumed by then
        // end of the chain phase -> store phase (the shape phases are over by then)
        struct {
            int err[32], role[32];             // best (error, role) of the roles a lane ran for ITS block ...
            u32 code[32][4];                   // ... and the 128 bits of that candidate
        } fin;
    };
    uint8_t order[kBc7Slots][2][64];           // shapes in ascending key order: [0] RGB (modes 1,3), [1] profile channels (mode 7)
    int win_shape[kBc7Super][5];               // winning shape id per block of the round and mode slot, -1 = none
    u32 palette[40][32];                       // lane-private scratch of the index search, [entry][lane]: 24 palette
                                               // entries, then 5 per-subset constants x 3 subsets
    union {
        // decoded endpoints (A, B as RGBA bytes) of every DISTINCT three-subset mask, for the two blocks being processed:
        // mode 2 for all 140 masks, mode 0 for the 36 masks of shapes 0..15 (bc7_phase_masks3 -> bc7_phase_shapes3)
        struct {
            u32 mode2[2][ITW_MASK3_COUNT][2];
            u32 mode0[2][ITW_MASK3_FIRST][2];
        } ends3;
        // chain phase of a two-half round: the FIRST half again (re-read from the staged tile once its slot in blk[] has
        // been taken by the second half; the mask tables are dead by then)
        Bc7Block parked[kBc7Batch];
    };
    int nvalid;                                // valid blocks in blk[]
    int half;                                  // which half of the round blk[] holds (0 / 1)
    int ntotal;                                // valid blocks of the round (chain and store phases)
    int nparked;                               // blocks in parked[]: the round's blocks [0, nparked); blk[] holds the rest
};
// mode slots m = 0..4 <-> BC7 modes {0, 2, 1, 3, 7}: the reference's evaluation order
ITW_HD int bc7_slot_mode(int m) { return (m == 0) ? 0 : ((m == 1) ? 2 : ((m == 2) ? 1 : ((m == 3) ? 3 : 7))); }
ITW_HD int bc7_mode_bits(int mode) { return (mode == 0 || mode == 1) ? 3 : ((mode == 6) ? 4 : 2); }

// number of candidates mode slot m evaluates under the profile; K:1386-1435
ITW_HD int bc7_slot_count(const Bc7Params& P, int m)
{
    if (m == 0) return P.sel[0] ? 16 : 0;
    if (m == 1) return (P.sel[0] && !P.skip2) ? 64 : 0;
    if (m == 2) return P.sel[1] ? P.t1 : 0;
    if (m == 3) return P.sel[1] ? P.t3 : 0;
    return P.sel[1] ? P.t7 : 0;
}
// shape id of list position n of mode slot m
ITW_HD int bc7_slot_shape(const Bc7Warp& W, int slot, int m, int n)
{
    if (m <= 1) return 64 + n;                                  // modes 0/2 walk the 3-subset table in order
    return W.order[slot][(m == 4) ? 1 : 0][n] & 63;             // modes 1/3/7 walk the ranked list
}

// ---------------------------------------------------------------------------------------------
// Views.  Modes 4/5 encode the block with one colour channel swapped with alpha ("rotation");
// for RGB profiles the swapped-in plane is the constant 255 (K:1579-1584).  rot = 3: no swap.
// ---------------------------------------------------------------------------------------------
struct View {
    const Bc7Block* b;
    int rot;      // 0..2 rotated channel, 3 = identity
    int alpha;    // 1: the swapped-in plane is alpha; 0: it is 255
};
ITW_HD u32 view_tex(const View& v, int k)
{
    const u32 t = v.b->tex[k];
    // byte `rot` <- byte 3 of the 2nd operand; rot = 3 with alpha reproduces t, so only "rot = 3 without alpha" (identity
    // views of RGB profiles) needs the plain selector
    const u32 swap_sel = (0x3210u & ~(0xFu << (4 * v.rot))) | (7u << (4 * v.rot));
    const u32 sel = (v.rot < 3) ? swap_sel : 0x3210u;
    return byte_perm(t, v.alpha ? t : 0xFFFFFFFFu, sel);
}
ITW_HD u32 view_plane(const View& v, int c, int i)          // branch-free: one load from a selected plane, one select
{
    const bool swapped = (c == v.rot);
    const u32 p = v.b->plane[swapped ? 3 : c][i];
    return (swapped && !v.alpha) ? 0xFFFFFFFFu : p;
}
ITW_HD u32 nibble_to_bytemask(u32 nib) { return (((nib & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu; }

// ---------------------------------------------------------------------------------------------
// raw moments of the masked texels as exact integers; slots as K:763-803
// ---------------------------------------------------------------------------------------------
ITW_HD void moments_u8(int (&st)[15], const u32 (&P)[4][4], u32 mask, int channels)
{
    u32 s[15];
#pragma unroll
    for (int i = 0; i < 15; i++) s[i] = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u32 bm = nibble_to_bytemask(mask >> (4 * i));
        const u32 r = P[0][i] & bm, g = P[1][i] & bm, b = P[2][i] & bm;
        s[10] = dp4a_u8(r, 0x01010101u, s[10]);
        s[11] = dp4a_u8(g, 0x01010101u, s[11]);
        s[12] = dp4a_u8(b, 0x01010101u, s[12]);
        s[0] = dp4a_u8(r, P[0][i], s[0]);
        s[1] = dp4a_u8(r, P[1][i], s[1]);
        s[2] = dp4a_u8(r, P[2][i], s[2]);
        s[4] = dp4a_u8(g, P[1][i], s[4]);
        s[5] = dp4a_u8(g, P[2][i], s[5]);
        s[7] = dp4a_u8(b, P[2][i], s[7]);
        if (channels == 4) {
            const u32 a = P[3][i] & bm;
            s[13] = dp4a_u8(a, 0x01010101u, s[13]);
            s[3] = dp4a_u8(r, P[3][i], s[3]);
            s[6] = dp4a_u8(g, P[3][i], s[6]);
            s[8] = dp4a_u8(b, P[3][i], s[8]);
            s[9] = dp4a_u8(a, P[3][i], s[9]);
        }
    }
    s[14] = (u32)popcount16(mask & 0xFFFFu);
#pragma unroll
    for (int i = 0; i < 15; i++) st[i] = (int)s[i];
}
ITW_HD void load_planes(u32 (&P)[4][4], const View& v)
{
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) P[c][i] = view_plane(v, c, i);
}

// Power iteration on a packed symmetric matrix [xx xy xz xw yy yz yw zz zw ww]; K:207-229.  Unroll factor measured
// (tools/tune_unroll.sh): rolled by 4 in round 1, when the kernel was instruction-cache bound; fully unrolled is 0.4 % faster now.
template <int CH, int kIterations>
ITW_HD void bc7_power_axis(float (&axis)[4], const float (&m)[10])
{
    float v0 = 1.0f, v1 = 1.0f, v2 = 1.0f, v3 = 1.0f;
    ITW_UNROLL(ITW_BC7_POWER_UNROLL)
    for (int it = 0; it < kIterations; it++) {
        float a0, a1, a2, a3 = 0.0f;
        if (CH == 3) {
            a0 = m[0] * v0 + m[1] * v1 + m[2] * v2;
            a1 = m[1] * v0 + m[4] * v1 + m[5] * v2;
            a2 = m[2] * v0 + m[5] * v1 + m[7] * v2;
        } else {
            a0 = m[0] * v0 + m[1] * v1 + m[2] * v2 + m[3] * v3;
            a1 = m[1] * v0 + m[4] * v1 + m[5] * v2 + m[6] * v3;
            a2 = m[2] * v0 + m[5] * v1 + m[7] * v2 + m[8] * v3;
            a3 = m[3] * v0 + m[6] * v1 + m[8] * v2 + m[9] * v3;
        }
        v0 = a0; v1 = a1; v2 = a2; v3 = a3;
        if (it & 1) {                                  // renormalise every other iteration: 1/sqrt, two exact ops
            float n2 = a0 * a0;
            n2 += a1 * a1;
            n2 += a2 * a2;
            if (CH == 4) n2 += a3 * a3;
            const float rn = 1.0f / sqrtf(n2);
            v0 *= rn; v1 *= rn; v2 *= rn; v3 *= rn;
        }
    }
    axis[0] = v0; axis[1] = v1; axis[2] = v2; axis[3] = v3;
}
// cov = sum(xy) - sum(x)*sum(y)/n on exact integer moments; K:805-823.  sum(x)*sum(y) < 2^24 is an
// exact integer, so the division by the count can use the slow-path-free exact quotient.
template <int CH>
ITW_HD void bc7_covariance(float (&cov)[10], float (&mean)[4], const int (&st)[15])
{
    const float n = (float)st[14], rn = 1.0f / n;
    float s[4];
#pragma unroll
    for (int c = 0; c < 4; c++) s[c] = (float)st[10 + c];
#pragma unroll
    for (int i = 0; i < 10; i++) cov[i] = 0.0f;
    cov[0] = (float)st[0] - div_by_rcp(s[0] * s[0], n, rn);
    cov[1] = (float)st[1] - div_by_rcp(s[0] * s[1], n, rn);
    cov[2] = (float)st[2] - div_by_rcp(s[0] * s[2], n, rn);
    cov[4] = (float)st[4] - div_by_rcp(s[1] * s[1], n, rn);
    cov[5] = (float)st[5] - div_by_rcp(s[1] * s[2], n, rn);
    cov[7] = (float)st[7] - div_by_rcp(s[2] * s[2], n, rn);
    if (CH == 4) {
        cov[3] = (float)st[3] - div_by_rcp(s[0] * s[3], n, rn);
        cov[6] = (float)st[6] - div_by_rcp(s[1] * s[3], n, rn);
        cov[8] = (float)st[8] - div_by_rcp(s[2] * s[3], n, rn);
        cov[9] = (float)st[9] - div_by_rcp(s[3] * s[3], n, rn);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) mean[c] = (c < CH) ? div_by_rcp(s[c], n, rn) : 0.0f;
}
// By-value working set of the shape and chain phases: nvcc keeps these small structs in registers across the
// non-inlined calls, whereas arrays handed over by pointer live in local memory (see bc6h.cuh).
struct Bc7Seg { float v[8]; };                       // endpoints A (r,g,b,a) and B (r,g,b,a) of one subset
struct Bc7Packed { u32 dec_a, dec_b, q_a, q_b; };    // decoded A, B and quantised A, B as RGBA bytes
struct Bc7Search { int err; u32 idx0, idx1; };

template <int CH>
ITW_HD void bc7_fit_impl(float (&ep)[8], const View& v, u32 mask)
{
    u32 P[4][4];
    load_planes(P, v);
    int ist[15];
    moments_u8(ist, P, mask, CH);
    float cov[10], mean[4], axis[4];
    bc7_covariance<CH>(cov, mean, ist);
    const float inv_var = 1.0f / (256.0f * 256.0f);
#pragma unroll
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = 0.001f * 0.001f;
    cov[0] += eps; cov[4] += eps; cov[7] += eps; cov[9] += eps;
    bc7_power_axis<CH, 8>(axis, cov);

    float lo = inf_f(), hi = -inf_f();
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // the reference starts the sum from 0.0f; 0 + x differs from x only in the sign of a zero, which nothing below can
        // observe (see the next comment), so the sum starts from the first product
        float d = axis[0] * ((float)((P[0][k >> 2] >> (8 * (k & 3))) & 255u) - mean[0]);
#pragma unroll
        for (int c = 1; c < CH; c++) d += axis[c] * ((float)((P[c][k >> 2] >> (8 * (k & 3))) & 255u) - mean[c]);
        // d is finite here (8-bit texels, finite axis), so the reference's (a<b)?a:b equals fminf/fmaxf up to
        // the sign of a zero, which the affine map below cannot observe: one FMNMX instead of FSETP+FSEL
        if ((mask >> k) & 1u) {
            lo = fminf(lo, d);
            hi = fmaxf(hi, d);
        }
    }
    if (hi - lo < 1.0f) { lo -= 0.5f; hi += 0.5f; }
#pragma unroll
    for (int c = 0; c < CH; c++) {
        ep[c] = clamp_sse(lo * axis[c] + mean[c], 0.0f, 255.0f);
        ep[4 + c] = clamp_sse(hi * axis[c] + mean[c], 0.0f, 255.0f);
    }
}
// PCA line through the masked texels, clamped to [0,255]; K:834-905.  v[0..3] = A, v[4..7] = B;
// components >= channels are zero (the reference's never-written slots, rule F6).
ITW_HD_NOINLINE Bc7Seg bc7_fit(const Bc7Block* blk, int rot, int alpha, u32 mask, int channels)
{
    const View v{blk, rot, alpha};
    Bc7Seg seg;
#pragma unroll
    for (int i = 0; i < 8; i++) seg.v[i] = 0.0f;
    if (channels == 4) bc7_fit_impl<4>(seg.v, v, mask);
    else bc7_fit_impl<3>(seg.v, v, mask);
    return seg;
}

// trace - lambda_max of a covariance; K:907-939 (eps on three diagonal slots only, K:918-920)
template <int CH>
ITW_HD float bc7_residual_bound(float (&cov)[10])
{
    const float inv_var = 1.0f / (256.0f * 256.0f);
#pragma unroll
    for (int i = 0; i < 10; i++) cov[i] *= inv_var;
    const float eps = 0.001f * 0.001f;
    cov[0] += eps; cov[4] += eps; cov[7] += eps;
    float axis[4];
    bc7_power_axis<CH, 4>(axis, cov);
    float mv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (CH == 3) {
        mv[0] = cov[0] * axis[0] + cov[1] * axis[1] + cov[2] * axis[2];
        mv[1] = cov[1] * axis[0] + cov[4] * axis[1] + cov[5] * axis[2];
        mv[2] = cov[2] * axis[0] + cov[5] * axis[1] + cov[7] * axis[2];
    } else {
        mv[0] = cov[0] * axis[0] + cov[1] * axis[1] + cov[2] * axis[2] + cov[3] * axis[3];
        mv[1] = cov[1] * axis[0] + cov[4] * axis[1] + cov[5] * axis[2] + cov[6] * axis[3];
        mv[2] = cov[2] * axis[0] + cov[5] * axis[1] + cov[7] * axis[2] + cov[8] * axis[3];
        mv[3] = cov[3] * axis[0] + cov[6] * axis[1] + cov[8] * axis[2] + cov[9] * axis[3];
    }
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < CH; c++) sum += sq(mv[c]);
    float bound = cov[0] + cov[4] + cov[7];
    if (CH == 4) bound += cov[9];
    bound -= sqrtf(sum);
    return max_sse(bound, 0.0f);
}
// Ranking key of a two-subset shape (K:952-971, :1403-1410); subset 1 = full - subset 0
template <int CH>
ITW_HD int bc7_split_key_impl(const Bc7Block* blk, int shape)
{
    const View v{blk, 3, 1};
    u32 P[4][4];
    load_planes(P, v);
    int full[15], part[15], rest[15];
    moments_u8(full, P, 0xFFFFu, CH);
    moments_u8(part, P, (u32)shape_mask(shape, 0), CH);
#pragma unroll
    for (int i = 0; i < 15; i++) rest[i] = full[i] - part[i];            // exact: the reference subtracts exact floats
    float c1[10], c2[10], mean[4];
    bc7_covariance<CH>(c1, mean, part);
    bc7_covariance<CH>(c2, mean, rest);
    float b = 0.0f;
    b += bc7_residual_bound<CH>(c1);
    b += bc7_residual_bound<CH>(c2);
    return shape + (int)((unsigned)cvt_x86(sqrtf(b) * 256.0f) * 64u);
}
ITW_HD_NOINLINE int bc7_split_key(const Bc7Block* blk, int shape, int channels)
{
    return (channels == 4) ? bc7_split_key_impl<4>(blk, shape) : bc7_split_key_impl<3>(blk, shape);
}

// Quantise one endpoint pair; K:983-1128.  out[0],out[1] = decoded A,B as RGBA bytes, out[2],out[3] =
// quantised A,B as RGBA bytes.  `channels` = components that vote on the p-bit (K:1011-1020); all four
// components are always produced (a never-written component is quantised from 0, rule F6).
// One rolled loop over the 8 components serves all three families:
//   modes 0,3,6,7  one p-bit per endpoint   K:983-1022   (the vote compares against the RAW quantised value
//                                                         except in mode 0 -- reference behaviour, K:1003-1009)
//   mode 1         one p-bit per pair        K:1024-1052  (a single running error sum over both endpoints)
//   modes 2,4,5    no p-bit                  K:1054-1065
ITW_HD_NOINLINE Bc7Packed bc7_quantise(Bc7Seg seg, int mode, int channels)
{
    const float (&ep)[8] = seg.v;
    // per-mode constants from nibble tables (mode 0 in the low nibble): nested conditionals on `mode` compile to a jump table
    // that cost more than a tenth of this routine
    const int family = (int)((0x00220210u >> (4 * mode)) & 3u);           // 1: mode 1; 2: modes 2, 4, 5; else 0
    // stored bits per component including the p-bit: 2^qbits - 1 is K's `levels2` (p-bit modes) or `levels-1`
    const int qbits = (int)((0x68758575u >> (4 * mode)) & 15u);           // modes 0,2,4: 5; 1,5: 7; 7: 6; 3,6: 8
    const int top = (1 << qbits) - 1;
    const float ftop = (float)top;
    const int vote_bits = (int)((0x88888875u >> (4 * mode)) & 15u);       // mode 0: 5; mode 1: 7; else 8 (expand_bits(v, 8) == v)
    const int votes = (mode == 1) ? 3 : channels;
    // Conversions: |ep| < 2^31 always holds here -- the fit clamps to [0,255] and the least-squares
    // solve divides by an exact non-zero INTEGER determinant (bc7_solve), which bounds |ep| by ~1.2e7 --
    // so the x86 overflow rule of cvt_x86() can never trigger and a plain truncation is identical.
    // t*0.5 and (t-1)*0.5 are exact (power-of-two scaling), so fusing the +0.5 rounds once, exactly like
    // the reference's separate multiply and add.
    // Modes 0-3 never emit a 4th component and their index search ignores it; it only matters when it votes
    // (alpha profiles refining modes 0/3, quirk Q1).  Otherwise it is skipped (its byte stays 0).
    const bool four = !(mode <= 3 && votes <= 3);
    // Branch-free per component: the no-p-bit family is the p-bit formula with scale 1 instead of 1/2 (t*1 + 0.5 rounds
    // once, like t + 0.5) and both candidates equal.  Components 0..2 always take part and always vote (votes >= 3);
    // only the fourth is conditional.
    const bool plain = (family == 2);
    const float half = plain ? 1.0f : 0.5f;
    const int step = plain ? 1 : 2, top0 = plain ? top : top - 1;
    u32 cand0[2] = {0u, 0u}, cand1[2] = {0u, 0u};
    bool pick1[2] = {false, false};
    float e0 = 0.0f, e1 = 0.0f;
    if (plain) {
        // no p-bit (modes 2, 4, 5; K:1054-1065): one candidate per component, no vote -- a third of all quantisations
#pragma unroll
        for (int i = 0; i < 2; i++) {
            u32 c0 = 0u;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (c == 3 && !four) continue;
                const float t = div255(ep[4 * i + c]) * ftop;
                c0 |= (u32)clampi(trunc_i(fma_rn(t, 1.0f, 0.5f)), 0, top) << (8 * c);
            }
            cand0[i] = c0;
        }
    } else
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (family == 0) { e0 = 0.0f; e1 = 0.0f; }
        u32 c0 = 0u, c1 = 0u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c == 3 && !four) continue;
            const float x = ep[4 * i + c];
            const float t = div255(x) * ftop;
            const int v0 = clampi(trunc_i(fma_rn(t, half, 0.5f)) * step, 0, top0);
            const int v1p = clampi(trunc_i(fma_rn(t - 1.0f, 0.5f, 0.5f)) * 2 + 1, 1, top);      // ((t - 1)/2 + 0.5) truncated, *2 + 1
            const int v1 = plain ? v0 : v1p;
            c0 |= (u32)v0 << (8 * c);
            c1 |= (u32)v1 << (8 * c);
            const float d0 = sq(x - (float)expand_bits(v0, vote_bits)), d1 = sq(x - (float)expand_bits(v1, vote_bits));
            if (c < 3 || votes == 4) { e0 += d0; e1 += d1; }
        }
        const bool p = !(e0 < e1);
        if (i == 0) { cand0[0] = c0; cand1[0] = c1; pick1[0] = p; }
        else        { cand0[1] = c0; cand1[1] = c1; pick1[1] = p; }
    }
    if (family == 1) pick1[0] = pick1[1];                                // decided after both endpoints
    if (family == 2) pick1[0] = pick1[1] = false;
    // decode all four bytes at once (K:1093-1122): v << (8-d) | that >> d, bytewise
    const int dbits = qbits;                                             // every mode decodes the bits it stores
    const u32 lowmask = 0x01010101u * (0xFFu >> dbits);
    u32 dec[2], qq[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const u32 q = pick1[i] ? cand1[i] : cand0[i];
        const u32 vv = q << (8 - dbits);
        dec[i] = vv + ((vv >> dbits) & lowmask);
        qq[i] = q;
    }
    return Bc7Packed{dec[0], dec[1], qq[0], qq[1]};
}

// Index search; K:1133-1193.  ends[2j], ends[2j+1] = decoded endpoints A,B of subset j (RGBA bytes);
// chmask zeroes the channels that do not take part (0x00FFFFFF for three-channel modes).
// Returns the summed error (exact integer) and the sixteen 4-bit indices in idx[0..1].
// bits_flags = index bits, plus kAssignThresholds when the caller allows the integer-threshold form of the two-bit search.
// Only the shape phases do: there every lane of the warp searches the same mode at the same time, whereas a chain-phase
// pass mixes two- and three-bit roles and two texel loops would run one after the other.
constexpr int kAssignThresholds = 8;
ITW_HD_NOINLINE Bc7Search bc7_assign(u32 (*pal)[32], int lane, const Bc7Block* blk, int rot, int alpha, int bits_flags, int pairs,
                                     u32 pattern, u32 e0, u32 e1, u32 e2, u32 e3, u32 e4, u32 e5, u32 chmask)
{
    const View v{blk, rot, alpha};
    const int bits = bits_flags & 7;
    const bool thresholds = (bits_flags == (2 | kAssignThresholds)) && rot == 3;
    const int levels = 1 << bits;
    // per-subset constants go to lane-private shared memory too and are fetched by subset id in the texel
    // loop: the load/store pipe is nearly idle in this kernel while the ALU pipe (selects) is the busiest
    u32 cur_a = e0, cur_b = e1, next_a = e2, next_b = e3;        // endpoints walk through two register pairs: no select by j
    for (int j = 0; j < pairs; j++) {
        const u32 a = cur_a & chmask, b = cur_b & chmask;
        cur_a = next_a; cur_b = next_b; next_a = e4; next_b = e5;
        const u32 aa = dp4a_u8(a, a, 0u), ab = dp4a_u8(a, b, 0u), bb = dp4a_u8(b, b, 0u);
        const int idiv = (int)(bb - 2u * ab + aa);               // sum of squared differences, exact
        pal[24 + 5 * j + 0][lane] = a;
        pal[24 + 5 * j + 1][lane] = b;
        pal[24 + 5 * j + 2][lane] = ab - aa;
        if (thresholds) {
            // two-bit search by thresholds: q1 >= m  <=>  8 num >= (2m - 1) div  <=>  num >= ceil((2m - 1) div / 8), m = 2, 3
            // (see the texel loop); coincident endpoints give 0/0 = NaN -> q1 = 1 in the reference: thresholds out of reach
            pal[24 + 5 * j + 3][lane] = (u32)(idiv ? (3 * idiv + 7) >> 3 : 0x7fffffff);
            pal[24 + 5 * j + 4][lane] = (u32)(idiv ? (5 * idiv + 7) >> 3 : 0x7fffffff);
        } else {
            const float fdiv = (float)idiv;
            pal[24 + 5 * j + 3][lane] = float_bits(fdiv);
            pal[24 + 5 * j + 4][lane] = float_bits(1.0f / fdiv);  // inf when the endpoints coincide (-> NaN below, as 0/0)
        }
        // Palette, K:1172: ((64 - w) a + w b + 32) >> 6 per channel, two channels per 32-bit word.  Written as
        // (64 a + 32) + w (b - a): the packed difference may borrow across the 16-bit lanes, but every lane of the SUM is the
        // non-negative 14-bit value above, so the word arithmetic is exact.  Entries 0 and levels-1 are the endpoints.
        const u32 m = 0x00FF00FFu;
        const u32 a_rb = a & m, a_ga = (a >> 8) & m;
        const u32 base_rb = (a_rb << 6) + 0x00200020u, base_ga = (a_ga << 6) + 0x00200020u;
        const u32 d_rb = (b & m) - a_rb, d_ga = ((b >> 8) & m) - a_ga;
        pal[j * levels][lane] = a;
        pal[j * levels + levels - 1][lane] = b;
        for (int q = 1; q < levels - 1; q++) {
            const u32 w = (u32)bc7_weight(bits, q);
            pal[j * levels + q][lane] = (((base_rb + w * d_rb) >> 6) & m) | (((base_ga + w * d_ga) << 2) & 0xFF00FF00u);
        }
    }
    const float flevels = (float)levels;
    int total = 0;
    u32 out0 = 0u, out1 = 0u;
    if (thresholds) {
        // The reference's q1 = clamp((int)(RN(RN(num / div) * 4) + 0.5), 1, 3) is a non-decreasing step function of the
        // integer num; with tau = (2m - 1) / 8 (a float): RN(p * 4 + 0.5) >= m <=> p >= tau (the sum is exact below the next
        // binade, and the binade boundary itself is an integer m), and RN(num / div) >= tau <=> num / div >= tau because a
        // quotient below tau stays at least 1 / (8 div) > 2^-21 away from it while half an ulp of tau is below 2^-24
        // (div <= 4 * 255^2 < 2^18).  tests/test_exact_division.py checks both sides of every step for every possible div.
        const u32* tx = (chmask == 0x00FFFFFFu) ? blk->rgb : blk->tex;
        ITW_UNROLL(ITW_BC7_ASSIGN_UNROLL)
        for (int k = 0; k < 16; k++) {
            const u32 t = tx[k];
            const int j = (int)((pattern >> (2 * k)) & 3u);
            const u32 ea = pal[24 + 5 * j + 0][lane], eb = pal[24 + 5 * j + 1][lane];
            const int cj = (int)pal[24 + 5 * j + 2][lane];
            const int t2 = (int)pal[24 + 5 * j + 3][lane], t3 = (int)pal[24 + 5 * j + 4][lane];
            const int num = (int)dp4a_u8(t, eb, 0u) - (int)dp4a_u8(t, ea, 0u) - cj;
            const int q1 = 1 + ((num >= t2) ? 1 : 0) + ((num >= t3) ? 1 : 0);
            const u32 p0 = pal[j * 4 + q1 - 1][lane], p1 = pal[j * 4 + q1][lane];
            const u32 d0 = absdiff_u8x4(p0, t), d1 = absdiff_u8x4(p1, t);
            const int e0 = (int)dp4a_u8(d0, d0, 0u), e1 = (int)dp4a_u8(d1, d1, 0u);
            const bool first = e0 < e1;
            total += first ? e0 : e1;
            const u32 bq = (u32)(first ? q1 - 1 : q1) << (4 * (k & 7));
            if (k < 8) out0 += bq; else out1 += bq;
        }
    } else
    if (rot == 3) {
        // identity view (every shape-phase search, the partitioned and mode-6 chains): the texel words are read as stored
        const u32* tx = (chmask == 0x00FFFFFFu) ? blk->rgb : blk->tex;
        ITW_UNROLL(ITW_BC7_ASSIGN_UNROLL)
        for (int k = 0; k < 16; k++) {
            const u32 t = tx[k];
            const int j = (int)((pattern >> (2 * k)) & 3u);
            const u32 ea = pal[24 + 5 * j + 0][lane], eb = pal[24 + 5 * j + 1][lane];
            const int cj = (int)pal[24 + 5 * j + 2][lane];
            const float dj = bits_float(pal[24 + 5 * j + 3][lane]), rj = bits_float(pal[24 + 5 * j + 4][lane]);
            // sum_c (t_c - a_c)(b_c - a_c): integer, |value| < 2^18, so the float it converts to is the
            // reference's float sum; the division of K:1158 is the exact FMA-corrected quotient (proved
            // equal to IEEE num/div on this integer domain, tests/test_exact_division.py)
            const int num = (int)dp4a_u8(t, eb, 0u) - (int)dp4a_u8(t, ea, 0u) - cj;
            const float proj = div_by_rcp((float)num, dj, rj);
            // |proj*levels| < 2^23, so truncation never overflows; NaN (coincident endpoints, 0/0) becomes
            // INT_MIN on x86 and 0 here -- both clamp to 1 (K:1160-1161).  proj*levels is exact, so the fused
            // form rounds once like the reference's multiply-then-add.
            const int q1 = clampi(trunc_i(fma_rn(proj, flevels, 0.5f)), 1, levels - 1);
            const u32 p0 = pal[j * levels + q1 - 1][lane], p1 = pal[j * levels + q1][lane];
            const u32 d0 = absdiff_u8x4(p0, t), d1 = absdiff_u8x4(p1, t);
            const int e0 = (int)dp4a_u8(d0, d0, 0u), e1 = (int)dp4a_u8(d1, d1, 0u);
            const bool first = e0 < e1;
            total += first ? e0 : e1;
            const u32 bq = (u32)(first ? q1 - 1 : q1) << (4 * (k & 7));
            if (k < 8) out0 += bq; else out1 += bq;
        }
    } else {                                                       // rotated views of modes 4 / 5: same loop, texels through view_tex
        ITW_UNROLL(ITW_BC7_ASSIGN_UNROLL)
        for (int k = 0; k < 16; k++) {
            const u32 t = view_tex(v, k) & chmask;
            const int j = (int)((pattern >> (2 * k)) & 3u);
            const u32 ea = pal[24 + 5 * j + 0][lane], eb = pal[24 + 5 * j + 1][lane];
            const int cj = (int)pal[24 + 5 * j + 2][lane];
            const float dj = bits_float(pal[24 + 5 * j + 3][lane]), rj = bits_float(pal[24 + 5 * j + 4][lane]);
            const int num = (int)dp4a_u8(t, eb, 0u) - (int)dp4a_u8(t, ea, 0u) - cj;
            const float proj = div_by_rcp((float)num, dj, rj);
            const int q1 = clampi(trunc_i(fma_rn(proj, flevels, 0.5f)), 1, levels - 1);
            const u32 p0 = pal[j * levels + q1 - 1][lane], p1 = pal[j * levels + q1][lane];
            const u32 d0 = absdiff_u8x4(p0, t), d1 = absdiff_u8x4(p1, t);
            const int e0 = (int)dp4a_u8(d0, d0, 0u), e1 = (int)dp4a_u8(d1, d1, 0u);
            const bool first = e0 < e1;
            total += first ? e0 : e1;
            const u32 bq = (u32)(first ? q1 - 1 : q1) << (4 * (k & 7));
            if (k < 8) out0 += bq; else out1 += bq;
        }
    }
    return Bc7Search{total, out0, out1};
}

// Least-squares endpoints of one subset from its indices; K:1198-1262.  The sums are exact integers.
// Components >= channels are zero.
ITW_HD_NOINLINE Bc7Seg bc7_solve(const Bc7Block* blk, int rot, int alpha, int bits, u32 idx0, u32 idx1, u32 mask, int channels)
{
    Bc7Seg seg;
#pragma unroll
    for (int i = 0; i < 8; i++) seg.v[i] = 0.0f;
    float (&ep)[8] = seg.v;
    const View v{blk, rot, alpha};
    const u32 top = (1u << bits) - 1u;
    u32 sq1 = 0u, sqq = 0u, sum[4] = {0u, 0u, 0u, 0u}, atb1[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 n = (((i < 2) ? idx0 : idx1) >> (16 * (i & 1))) & 0xFFFFu;     // four 4-bit indices
        n = (n | (n << 8)) & 0x00FF00FFu;
        n = (n | (n << 4)) & 0x0F0F0F0Fu;                                  // ... one per byte
        const u32 bm = nibble_to_bytemask(mask >> (4 * i));
        const u32 qm = n & bm, ones = bm & 0x01010101u;
        const u32 xm = ((top * 0x01010101u) & bm) - qm;                    // (levels-1) - q, bytewise, no borrows
        sq1 = dp4a_u8(qm, 0x01010101u, sq1);
        sqq = dp4a_u8(qm, qm, sqq);
#pragma unroll
        for (int c = 0; c < 4; c++) {                           // channels >= 3; an unused fourth channel sums zeros
            const u32 p = (c < 3 || channels == 4) ? view_plane(v, c, i) : 0u;
            sum[c] = dp4a_u8(ones, p, sum[c]);
            atb1[c] = dp4a_u8(xm, p, atb1[c]);
        }
    }
    const float ftop = (float)top, count = (float)popcount16(mask & 0xFFFFu);
    const float fsq1 = (float)sq1, fsqq = (float)sqq;
    float cxx = count * sq(ftop) - (2.0f * ftop) * fsq1 + fsqq;
    float cyy = fsqq;
    float cxy = ftop * fsq1 - fsqq;
    float det = cxx * cyy - cxy * cxy;
    float scale = ftop / det;
    bool flat = fabsf(det) < 0.001f;
    const float rcount = 1.0f / count;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float fs = (float)sum[c], fa1 = (float)atb1[c];
        float atb2 = ftop * fs - fa1;
        float a = (fa1 * cyy - atb2 * cxy) * scale;
        float b = (atb2 * cxx - fa1 * cxy) * scale;
        // integer sum / integer count (1..16): the exact FMA-corrected quotient (tests/test_exact_division.py); a plain
        // IEEE division would take its slow path for every zero sum
        if (flat) { a = div_by_rcp(fs, count, rcount); b = a; }
        const bool used = (c < 3 || channels == 4);
        ep[c] = used ? a : 0.0f;
        ep[4 + c] = used ? b : 0.0f;
    }
    return seg;
}

// ---------------------------------------------------------------------------------------------
// 128-bit layouts; K:1694-1964.  Q[j][0/1] = quantised endpoints A/B of subset j as RGBA bytes.
// ---------------------------------------------------------------------------------------------
ITW_HD u32 qcomp(u32 packed, int c) { return (packed >> (8 * c)) & 255u; }

ITW_HD void bc7_write_partitioned(u32* out, u32 (&Q)[3][2], u32 idx0, u32 idx1, int shape, int mode)
{
    const int bits = bc7_mode_bits(mode), pairs = bc7_pairs(mode), channels = (mode == 7) ? 4 : 3;
    const int half = (1 << bits) / 2;
    int flips = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {                               // anchor index MSB must be 0; K:1708-1733
        if (j >= pairs) continue;
        const int k0 = shape_anchor(shape, j);
        const int vv = (int)(((k0 < 8 ? idx0 : idx1) >> (4 * (k0 & 7))) & 15u);
        if (vv >= half) {
            const u32 t = Q[j][0]; Q[j][0] = Q[j][1]; Q[j][1] = t;
            flips |= shape_mask(shape, j);
        }
    }
    BitSink s;
    s.reset();
    s.put(mode + 1, 1u << mode);
    s.put(mode == 0 ? 4 : 6, (u32)(shape & (mode == 0 ? 15 : 63)));
    const int width = (mode == 0) ? 4 : ((mode == 1) ? 6 : ((mode == 3) ? 7 : 5));
    const int drop = (mode == 2) ? 0 : 1;                      // p-bit modes store the value without its LSB
    for (int c = 0; c < channels; c++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j >= pairs) continue;
            s.put(width, qcomp(Q[j][0], c) >> drop);
            s.put(width, qcomp(Q[j][1], c) >> drop);
        }
    }
    if (mode == 1) { s.put(1, Q[0][0] & 1u); s.put(1, Q[1][0] & 1u); }
    if (mode == 0 || mode == 3 || mode == 7) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j >= pairs) continue;
            s.put(1, Q[j][0] & 1u);
            s.put(1, Q[j][1] & 1u);
        }
    }
    put_indices(s, idx0, idx1, bits, flips, shape_anchor(shape, 1), (pairs == 3) ? shape_anchor(shape, 2) : -1);
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}
// single-subset orientation on packed endpoints; K:1694-1706
ITW_HD void orient_packed(u32& qa, u32& qb, u32& idx0, u32& idx1, int bits)
{
    const int levels = 1 << bits;
    if ((int)(idx0 & 15u) >= levels / 2) {
        const u32 t = qa; qa = qb; qb = t;
        const u32 all = 0x11111111u * (u32)(levels - 1);
        idx0 = all - idx0;
        idx1 = all - idx1;
    }
}
ITW_HD void bc7_write_mode45(u32* out, u32 qa, u32 qb, u32 i0, u32 i1, int aq0, int aq1, u32 a0, u32 a1, int mode,
                             int rotation, int swap)
{
    const int epbits = (mode == 4) ? 5 : 7, aepbits = (mode == 4) ? 6 : 8;
    const int cbits = 2, sbits = (mode == 4) ? 3 : 2;          // widths of the first / second index set
    int aq[2] = {aq0, aq1};
    if (!swap) {
        orient_packed(qa, qb, i0, i1, cbits);
        orient_single(aq, 1, a0, a1, sbits);
    } else {                                                    // the two index sets trade places; K:1903-1908
        u32 t0 = i0, t1 = i1;
        i0 = a0; i1 = a1;
        a0 = t0; a1 = t1;
        orient_single(aq, 1, i0, i1, cbits);
        orient_packed(qa, qb, a0, a1, sbits);
    }
    BitSink s;
    s.reset();
    s.put(mode + 1, 1u << mode);
    s.put(2, (u32)((rotation + 1) & 3));
    if (mode == 4) s.put(1, (u32)swap);
    for (int c = 0; c < 3; c++) { s.put(epbits, qcomp(qa, c)); s.put(epbits, qcomp(qb, c)); }
    s.put(aepbits, (u32)aq[0]);
    s.put(aepbits, (u32)aq[1]);
    put_indices(s, i0, i1, cbits, 0, -1, -1);
    put_indices(s, a0, a1, sbits, 0, -1, -1);
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}
ITW_HD void bc7_write_mode6(u32* out, u32 qa, u32 qb, u32 i0, u32 i1)
{
    orient_packed(qa, qb, i0, i1, 4);
    BitSink s;
    s.reset();
    s.put(7, 64u);
    for (int c = 0; c < 4; c++) { s.put(7, qcomp(qa, c) >> 1); s.put(7, qcomp(qb, c) >> 1); }
    s.put(1, qa & 1u);
    s.put(1, qb & 1u);
    put_indices(s, i0, i1, 4, 0, -1, -1);
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}

// ---------------------------------------------------------------------------------------------
// scalar channel of modes 4/5; K:1437-1563.  `a` = byte `shift/8` of the ORIGINAL texels.
// ---------------------------------------------------------------------------------------------
// After quantisation the two endpoints are integers, the texels are integers and the weights are
// integers, so the index search and the least-squares sums are exact integer arithmetic; the projection
// (x - e0)/(e1 - e0 + 0.001f) keeps the reference's float expression (K:1510) with the exact quotient
// (domain proved in tests/test_exact_division.py).
// The sixteen texels of the channel are the four packed words of its plane (pl[i] byte j = texel 4i+j); the palette
// (at most 8 byte-sized entries) is packed into one 64-bit register and read with a shift.
ITW_HD void scalar_quantise(int (&q)[2], int (&e)[2], const float (&ep)[2], int epbits)
{
    const int top = (1 << epbits) - 1;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        q[i] = clampi(trunc_i(div255(ep[i]) * (float)top + 0.5f), 0, top);   // ep in [0,255]: no overflow
        e[i] = expand_bits(q[i], epbits);
    }
}
ITW_HD int scalar_assign(u32& idx0, u32& idx1, const u32 (&pl)[4], int bits, const int (&e)[2])
{
    const int levels = 1 << bits;
    const float flevels = (float)levels;
    const float den = (float)(e[1] - e[0]) + 0.001f, rden = 1.0f / den;
    unsigned long long pal = 0ull;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int w = bc7_weight(bits, q);
        const u32 v = (q < levels) ? (u32)(((64 - w) * e[0] + w * e[1] + 32) >> 6) : 0u;
        pal |= (unsigned long long)v << (8 * q);
    }
    u32 out[2] = {0u, 0u};
    int total = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll 1
        for (int i = 0; i < 2; i++) {                          // one plane word = four texels per iteration
            const u32 word = pl[2 * h + i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int a = (int)((word >> (8 * j)) & 255u);
                const float proj = div_by_rcp((float)(a - e[0]), den, rden);
                const int q1 = clampi(trunc_i(fma_rn(proj, flevels, 0.5f)), 1, levels - 1);
                const u32 two = (u32)(pal >> (8 * (q1 - 1)));  // entries q1-1 and q1
                const int d0 = (int)(two & 255u) - a, d1 = (int)((two >> 8) & 255u) - a;
                const int err0 = d0 * d0, err1 = d1 * d1;
                const bool first = err0 < err1;
                total += first ? err0 : err1;
                out[h] += (u32)(first ? q1 - 1 : q1) << (16 * i + 4 * j);
            }
        }
    }
    idx0 = out[0];
    idx1 = out[1];
    return total;
}
ITW_HD void scalar_solve(float (&ep)[2], const u32 (&pl)[4], int bits, u32 idx0, u32 idx1)
{
    const u32 utop = (1u << bits) - 1u;
    u32 usq1 = 0u, usqq = 0u, usum = 0u, uatb1 = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 n = (((i < 2) ? idx0 : idx1) >> (16 * (i & 1))) & 0xFFFFu;     // four 4-bit indices ...
        n = (n | (n << 8)) & 0x00FF00FFu;
        n = (n | (n << 4)) & 0x0F0F0F0Fu;                                  // ... one per byte
        const u32 xm = utop * 0x01010101u - n;                             // (levels-1) - q, bytewise, no borrows
        usq1 = dp4a_u8(n, 0x01010101u, usq1);
        usqq = dp4a_u8(n, n, usqq);
        usum = dp4a_u8(pl[i], 0x01010101u, usum);
        uatb1 = dp4a_u8(xm, pl[i], uatb1);
    }
    const float top = (float)utop, atb1 = (float)uatb1, sq1 = (float)usq1, sqq = (float)usqq, sum = (float)usum;
    float atb2 = top * sum - atb1;
    float cxx = 16.0f * sq(top) - (2.0f * top) * sq1 + sqq;
    float cyy = sqq;
    float cxy = top * sq1 - sqq;
    float det = cxx * cyy - cxy * cxy;
    float scale = top / det;
    ep[0] = clamp_sse((atb1 * cyy - atb2 * cxy) * scale, 0.0f, 255.0f);
    ep[1] = clamp_sse((atb2 * cxx - atb1 * cxy) * scale, 0.0f, 255.0f);
    if (fabsf(det) < 0.001f) {
        ep[0] = sum * 0.0625f;                                             // sum / 16, exact
        ep[1] = ep[0];
    }
}
struct Bc7Scalar { int err, q0, q1; u32 idx0, idx1; };
ITW_HD_NOINLINE Bc7Scalar bc7_scalar_channel(const Bc7Block* blk, int rotation, int abits, int aepbits, int rch)
{
    u32 pl[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pl[i] = blk->plane[rotation][i];
    int lo = 255, hi = 0;                                        // K:1542-1548
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int a = (int)((pl[i] >> (8 * j)) & 255u);
            lo = mini(lo, a);
            hi = maxi(hi, a);
        }
    float ep[2] = {(float)lo, (float)hi};
    int q[2], e[2];
    u32 a0, a1;
    scalar_quantise(q, e, ep, aepbits);
    int err = scalar_assign(a0, a1, pl, abits, e);
#pragma unroll 1
    for (int it = 0; it < rch; it++) {
        scalar_solve(ep, pl, abits, a0, a1);
        scalar_quantise(q, e, ep, aepbits);
        err = scalar_assign(a0, a1, pl, abits, e);
    }
    return Bc7Scalar{err, q[0], q[1], a0, a1};
}

// ---------------------------------------------------------------------------------------------
// The generic chain: initial fit + refinement + encode of one (block, role); K:1299-1363 (modes
// 0-3,7), :1565-1655 (modes 4,5), :1657-1689 (mode 6).
// ---------------------------------------------------------------------------------------------
struct Role {
    int kind;      // 0 partitioned, 1 mode 4/5, 2 mode 6
    int mode, shape, rotation, swap;
};
ITW_HD int bc7_rotations(const Bc7Params& P) { return P.sel[2] ? maxi(P.channels - P.ch0, 0) : 0; }
ITW_HD int bc7_role_count(const Bc7Params& P) { return 5 + 3 * bc7_rotations(P) + (P.sel[3] ? 1 : 0); }
// The chain phase walks only the roles that can produce a candidate: a partitioned mode whose list is empty under the profile
// (mode 7 of the RGB profiles, ...) would otherwise idle one lane of every pass.  i-th active role, ascending.
ITW_HD int bc7_active_roles(const Bc7Params& P)
{
    int n = bc7_role_count(P) - 5;
    for (int m = 0; m < 5; m++) n += (bc7_slot_count(P, m) > 0) ? 1 : 0;
    return n;
}
ITW_HD int bc7_active_role(const Bc7Params& P, int i)
{
    for (int m = 0; m < 5; m++)
        if (bc7_slot_count(P, m) > 0) { if (i == 0) return m; i--; }
    return 5 + i;
}

struct Bc7Result { int err; u32 code[4]; };
ITW_HD_NOINLINE Bc7Result bc7_chain(Bc7Warp& W, const Bc7Params& P, int lane, int slot, int r)
{
    Bc7Result res;
    res.err = kErrNone;
    res.code[0] = res.code[1] = res.code[2] = res.code[3] = 0u;
    const Bc7Block* blk = (slot < W.nparked) ? &W.parked[slot] : &W.blk[slot - W.nparked];
    const int nrot = bc7_rotations(P);
    Role role;
    role.rotation = 3; role.swap = 0; role.shape = 0;
    if (r < 5) {
        role.kind = 0;
        role.mode = bc7_slot_mode(r);
        role.shape = W.win_shape[slot][r];
        if (role.shape < 0) return res;
    } else if (r < 5 + 2 * nrot) {
        role.kind = 1; role.mode = 4; role.rotation = P.ch0 + ((r - 5) >> 1); role.swap = (r - 5) & 1;
    } else if (r < 5 + 3 * nrot) {
        role.kind = 1; role.mode = 5; role.rotation = P.ch0 + (r - 5 - 2 * nrot);
    } else {
        role.kind = 2; role.mode = 6;
    }
    const int mode = role.mode;
    const int pairs = bc7_pairs(mode);
    int bits = bc7_mode_bits(mode);
    if (role.kind == 1 && role.swap) bits = 3;
    // channels fitted / searched, and the count that votes on p-bits during refinement (K:1343)
    const int channels = (role.kind == 2) ? P.channels : ((mode == 7) ? 4 : 3);
    const int vote_refine = (role.kind == 0) ? P.channels : channels;
    const u32 chmask = (channels == 4) ? 0xFFFFFFFFu : 0x00FFFFFFu;
    // view: rotation 3 (alpha itself rotated "with itself") is the identity
    const int rot = (role.kind == 1 && role.rotation < 3) ? role.rotation : 3;
    const int alpha = (P.channels == 4) ? 1 : 0;
    const u32 pattern = (role.kind == 0) ? shape_pattern(role.shape) : 0u;

    u32 ends[6], Q[3][2];
    // initial candidate
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Q[j][0] = Q[j][1] = 0u;
        ends[2 * j] = ends[2 * j + 1] = 0u;
    }
    float tail_a = 0.0f, tail_b = 0.0f;                          // ep[3], ep[7] carried between iterations (quirk Q5)
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (j >= pairs) continue;
        const u32 mask = (role.kind == 0) ? (u32)shape_mask(role.shape, j) : 0xFFFFu;
        Bc7Seg seg = bc7_fit(blk, rot, alpha, mask, channels);   // slots >= channels are zero (F6)
        if (role.kind == 2 && channels == 3) seg.v[3] = seg.v[7] = 255.0f;     // K:1664-1667
        const Bc7Packed pk = bc7_quantise(seg, mode, channels);
        tail_a = (float)(pk.dec_a >> 24); tail_b = (float)(pk.dec_b >> 24);
        ends[2 * j] = pk.dec_a; ends[2 * j + 1] = pk.dec_b;
        Q[j][0] = pk.q_a; Q[j][1] = pk.q_b;
    }
    Bc7Search best = bc7_assign(W.palette, lane, blk, rot, alpha, bits, pairs, pattern, ends[0], ends[1], ends[2], ends[3], ends[4],
                                ends[5], chmask);

    const int refine = P.refine[mode];
    for (int it = 0; it < refine; it++) {
        u32 nQ[3][2];
#pragma unroll
        for (int j = 0; j < 3; j++) nQ[j][0] = nQ[j][1] = 0u;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j >= pairs) continue;
            const u32 mask = (role.kind == 0) ? (u32)shape_mask(role.shape, j) : 0xFFFFu;
            Bc7Seg seg = bc7_solve(blk, rot, alpha, bits, best.idx0, best.idx1, mask, channels);
            // K's arrays live across iterations in modes 4-6: a fourth slot the solve does not write keeps the
            // previous iteration's decoded value
            if (role.kind != 0 && channels < 4) { seg.v[3] = tail_a; seg.v[7] = tail_b; }
            const Bc7Packed pk = bc7_quantise(seg, mode, vote_refine);
            tail_a = (float)(pk.dec_a >> 24); tail_b = (float)(pk.dec_b >> 24);
            ends[2 * j] = pk.dec_a; ends[2 * j + 1] = pk.dec_b;
            nQ[j][0] = pk.q_a; nQ[j][1] = pk.q_b;
        }
        const Bc7Search found = bc7_assign(W.palette, lane, blk, rot, alpha, bits, pairs, pattern, ends[0], ends[1], ends[2], ends[3],
                                           ends[4], ends[5], chmask);
        // partitioned modes keep the best iterate (K:1348); modes 4,5,6 keep the last (K:1598-1603, :1677-1682)
        if (role.kind != 0 || found.err < best.err) {
#pragma unroll
            for (int j = 0; j < 3; j++) { Q[j][0] = nQ[j][0]; Q[j][1] = nQ[j][1]; }
            best = found;
        }
    }
    int best_err = best.err;
    const u32 best_idx[2] = {best.idx0, best.idx1};

    u32* out = res.code;
    if (role.kind == 0) {
        if (mode != 7 && P.channels != 3) {                       // opaque error of the dropped alpha; K:1267-1277, :1356
            int opaque = 0;
            for (int k = 0; k < 16; k++) { const int d = (int)(blk->tex[k] >> 24) - 255; opaque += d * d; }
            best_err += opaque;
        }
        bc7_write_partitioned(out, Q, best_idx[0], best_idx[1], role.shape, mode);
    } else if (role.kind == 1) {
        const int abits = (mode == 4 && !role.swap) ? 3 : 2, aepbits = (mode == 4) ? 6 : 8;
        const Bc7Scalar sc = bc7_scalar_channel(blk, role.rotation, abits, aepbits, P.rch);
        best_err += sc.err;
        bc7_write_mode45(out, Q[0][0], Q[0][1], best_idx[0], best_idx[1], sc.q0, sc.q1, sc.idx0, sc.idx1, mode, role.rotation,
                         role.swap);
    } else {
        bc7_write_mode6(out, Q[0][0], Q[0][1], best_idx[0], best_idx[1]);
    }
    res.err = best_err;
    return res;
}

// =============================================================================================
// Warp program: per-lane phase functions.  A phase reads what earlier phases wrote to W and
// writes disjoint locations; the caller separates phases with a warp barrier.
// =============================================================================================
ITW_HD int bc7_group_count(const Bc7Warp& W, int g) { return mini(maxi(W.nvalid - kBc7Slots * g, 0), kBc7Slots); }   // valid blocks of group g

// Loads half `half` (blocks 8*half .. 8*half+7 of the round's `ntotal`) into blk[].
ITW_HD void bc7_phase_load(int lane, Bc7Warp& W, const SurfaceView& s, long long first_block, int ntotal, int half)
{
    const int bw = s.width >> 2;
    const int nvalid = mini(maxi(ntotal - kBc7Batch * half, 0), kBc7Batch);
    for (int t = lane; t < nvalid * 16; t += 32) {
        const int slot = t >> 4, k = t & 15;
        const long long id = first_block + kBc7Batch * half + slot;
        const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);
        const uint8_t* p = s.ptr + (size_t)(by * 4 + (k >> 2)) * (size_t)s.stride + (size_t)(bx * 4 + (k & 3)) * 4;
        const u32 rgb = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16);
        W.blk[slot].rgb[k] = rgb;
        W.blk[slot].tex[k] = rgb | ((u32)p[3] << 24);
    }
    if (lane == 0) { W.nvalid = nvalid; W.half = half; W.ntotal = ntotal; W.nparked = 0; }
    for (int t = lane; t < kBc7Batch * 5; t += 32) W.win_shape[kBc7Batch * half + t / 5][t % 5] = -1;
}
// After the shape phases of the second half: the first half is read again (texels and channel planes) for the chain phase.
ITW_HD void bc7_phase_park(int lane, Bc7Warp& W, const SurfaceView& s, long long first_block, int ntotal)
{
    if (ntotal <= kBc7Batch) return;                       // single-half round: blk[] still holds it
    const int bw = s.width >> 2;
    for (int t = lane; t < kBc7Batch * 16; t += 32) {
        const int slot = t >> 4, k = t & 15;
        const long long id = first_block + slot;
        const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);
        const uint8_t* p = s.ptr + (size_t)(by * 4 + (k >> 2)) * (size_t)s.stride + (size_t)(bx * 4 + (k & 3)) * 4;
        const u32 rgb = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16);
        W.parked[slot].rgb[k] = rgb;
        W.parked[slot].tex[k] = rgb | ((u32)p[3] << 24);
        // plane word (c, i) = channel c of texels 4i..4i+3: here c = k & 3, i = k >> 2 (any bijection of 16 onto (c, i) does)
        const int c = k & 3, i = k >> 2;
        u32 v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int kk = 4 * i + j;
            v |= (u32)s.ptr[(size_t)(by * 4 + (kk >> 2)) * (size_t)s.stride + (size_t)(bx * 4 + (kk & 3)) * 4 + c] << (8 * j);
        }
        W.parked[slot].plane[c][i] = v;
    }
    if (lane == 0) W.nparked = kBc7Batch;
}
// channel planes from the packed texels (a 4x4 byte transpose per group of four texels)
ITW_HD void bc7_phase_planes(int lane, Bc7Warp& W)
{
    for (int t = lane; t < W.nvalid * 16; t += 32) {
        const int slot = t >> 4, c = (t >> 2) & 3, i = t & 3;
        const u32* x = &W.blk[slot].tex[4 * i];
        u32 v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v |= ((x[j] >> (8 * c)) & 255u) << (8 * j);
        W.blk[slot].plane[c][i] = v;
    }
}
// One shape, both modes of a mode-slot pair (ma, mb): fits once, quantises and searches per mode.  `slot` = position inside
// the group (scratch index), `blk` = the block.
ITW_HD_NOINLINE void bc7_eval_shape(Bc7Warp& W, int lane, int slot, const Bc7Block* blk, int shape, int n, int ma, bool do_a, int mb, bool do_b)
{
    const int mode_a = bc7_slot_mode(ma), mode_b = bc7_slot_mode(mb);
    const int pairs = bc7_pairs(mode_a);
    const int channels = (mode_a == 7) ? 4 : 3;
    const u32 chmask = (channels == 4) ? 0xFFFFFFFFu : 0x00FFFFFFu;
    u32 ends_a[6], ends_b[6];
#pragma unroll
    for (int i = 0; i < 6; i++) ends_a[i] = ends_b[i] = 0u;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (j >= pairs) continue;
        const Bc7Seg seg = bc7_fit(blk, 3, 1, (u32)shape_mask(shape, j), channels);
        if (do_a) {
            const Bc7Packed pk = bc7_quantise(seg, mode_a, channels);
            ends_a[2 * j] = pk.dec_a; ends_a[2 * j + 1] = pk.dec_b;
        }
        if (do_b) {
            const Bc7Packed pk = bc7_quantise(seg, mode_b, channels);
            ends_b[2 * j] = pk.dec_a; ends_b[2 * j + 1] = pk.dec_b;
        }
    }
    const u32 pattern = shape_pattern(shape);
    if (do_a)
        W.cand_err[slot][0][n] = bc7_assign(W.palette, lane, blk, 3, 1, bc7_mode_bits(mode_a) | kAssignThresholds, pairs, pattern, ends_a[0], ends_a[1], ends_a[2],
                                            ends_a[3], ends_a[4], ends_a[5], chmask).err;
    if (do_b)
        W.cand_err[slot][1][n] = bc7_assign(W.palette, lane, blk, 3, 1, bc7_mode_bits(mode_b) | kAssignThresholds, pairs, pattern, ends_b[0], ends_b[1], ends_b[2],
                                            ends_b[3], ends_b[4], ends_b[5], chmask).err;
}
// ---- three-subset shapes (modes 0 and 2) -------------------------------------------------------------------
// The 64 shapes x 3 subsets use only 140 distinct texel masks, and the PCA fit and the endpoint quantisation of a subset
// depend on its mask alone (the reference refits per shape and per mode with identical results, K:1279-1297).  So, two
// blocks at a time (`duo` = 0..3: blocks 2*duo, 2*duo+1 of the batch): phase A fits and quantises every distinct mask once
// (lane <-> (block, mask)); phase B runs the index search of every shape from the stored endpoints (lane <-> (block, shape)).
// Same values as evaluating shape by shape, 25 % fewer fits and quantisations.
ITW_HD void bc7_phase_masks3(int lane, Bc7Warp& W, const Bc7Params& P, int duo)
{
    const int ca = bc7_slot_count(P, 0), cb = bc7_slot_count(P, 1);
    const int nmask = (cb > 0) ? ITW_MASK3_COUNT : ITW_MASK3_FIRST;
    const int nslots = mini(maxi(W.nvalid - 2 * duo, 0), 2);
    for (int t = lane; t < nslots * nmask; t += 32) {
        const int s = t / nmask, u = t - s * nmask;
        const Bc7Block* blk = &W.blk[2 * duo + s];
        const Bc7Seg seg = bc7_fit(blk, 3, 1, (u32)ITW_TABLE(mask3_unique)[u], 3);
        if (cb > 0) {
            const Bc7Packed pk = bc7_quantise(seg, 2, 3);
            W.ends3.mode2[s][u][0] = pk.dec_a;
            W.ends3.mode2[s][u][1] = pk.dec_b;
        }
        if (ca > 0 && u < ITW_MASK3_FIRST) {
            const Bc7Packed pk = bc7_quantise(seg, 0, 3);
            W.ends3.mode0[s][u][0] = pk.dec_a;
            W.ends3.mode0[s][u][1] = pk.dec_b;
        }
    }
}
ITW_HD void bc7_phase_shapes3(int lane, Bc7Warp& W, const Bc7Params& P, int duo)
{
    const int ca = bc7_slot_count(P, 0), cb = bc7_slot_count(P, 1);
    const int count = maxi(ca, cb);
    const int nslots = mini(maxi(W.nvalid - 2 * duo, 0), 2);
    for (int t = lane; t < nslots * count; t += 32) {
        const int s = t / count, n = t - s * count;
        const int slot = (2 * duo + s) & (kBc7Slots - 1);             // scratch index inside the group
        const Bc7Block* blk = &W.blk[2 * duo + s];
        const u32 pattern = shape_pattern(64 + n);
        const int u0 = ITW_TABLE(shape3_mask_id)[3 * n], u1 = ITW_TABLE(shape3_mask_id)[3 * n + 1], u2 = ITW_TABLE(shape3_mask_id)[3 * n + 2];
        if (n < ca)
            W.cand_err[slot][0][n] = bc7_assign(W.palette, lane, blk, 3, 1, 3, 3, pattern, W.ends3.mode0[s][u0][0], W.ends3.mode0[s][u0][1],
                                                W.ends3.mode0[s][u1][0], W.ends3.mode0[s][u1][1], W.ends3.mode0[s][u2][0],
                                                W.ends3.mode0[s][u2][1], 0x00FFFFFFu).err;
        if (n < cb)
            W.cand_err[slot][1][n] = bc7_assign(W.palette, lane, blk, 3, 1, 2 | kAssignThresholds, 3, pattern, W.ends3.mode2[s][u0][0], W.ends3.mode2[s][u0][1],
                                                W.ends3.mode2[s][u1][0], W.ends3.mode2[s][u1][1], W.ends3.mode2[s][u2][0],
                                                W.ends3.mode2[s][u2][1], 0x00FFFFFFu).err;
    }
}

// shapes of a pair of mode slots that walk the same list, for group g: (2,3) ranked two-subset, (4,4) mode 7
ITW_HD void bc7_phase_shapes(int lane, Bc7Warp& W, const Bc7Params& P, int g, int ma, int mb)
{
    const int ca = bc7_slot_count(P, ma), cb = (mb != ma) ? bc7_slot_count(P, mb) : 0;
    const int both = mini(ca, cb), count = maxi(ca, cb);
    // list positions [0, both) run both modes, [both, count) only the longer list's mode; tasks are ordered
    // so that the "both" positions of all blocks come first and a warp never mixes the two kinds
    const int nv = bc7_group_count(W, g);
    const int nboth = nv * both, nall = nv * count;
    for (int t = lane; t < nall; t += 32) {
        int slot, n;
        if (t < nboth) { slot = t / both; n = t - slot * both; }
        else { const int u = t - nboth, rest = count - both; slot = u / rest; n = both + (u - slot * rest); }
        bc7_eval_shape(W, lane, slot, &W.blk[kBc7Slots * g + slot], bc7_slot_shape(W, slot, ma, n), n, ma, n < ca, mb, n < cb);
    }
}
// split-bound keys of the 64 two-subset shapes; set 0 = RGB (modes 1,3), set 1 = profile channels (mode 7)
ITW_HD bool bc7_needs_keys(const Bc7Params& P, int set)
{
    if (!P.sel[1]) return false;
    return set == 0 ? !(P.t1 == 0 && P.t3 == 0) : (P.t7 != 0);
}
ITW_HD void bc7_phase_keys(int lane, Bc7Warp& W, const Bc7Params& P, int g, int set)
{
    const int channels = (set == 0) ? 3 : P.channels;
    for (int t = lane; t < bc7_group_count(W, g) * 64; t += 32) {
        const int slot = t >> 6, shape = t & 63;
        W.keys[slot][shape] = bc7_split_key(&W.blk[kBc7Slots * g + slot], shape, channels);
    }
}
ITW_HD void bc7_phase_rank(int lane, Bc7Warp& W, int g, int set)
{
    for (int t = lane; t < bc7_group_count(W, g) * 64; t += 32) {
        const int slot = t >> 6, i = t & 63;
        W.order[slot][set][rank_of(W.keys[slot], 64, i)] = (uint8_t)(W.keys[slot][i] & 63);
    }
}
// first minimum of the candidate lists just evaluated for mode slots (ma, mb) of group g, as a SHAPE id; K:1320 (strict <)
ITW_HD void bc7_phase_winners(int lane, Bc7Warp& W, const Bc7Params& P, int g, int ma, int mb)
{
    const int nm = (mb != ma) ? 2 : 1;
    for (int t = lane; t < bc7_group_count(W, g) * nm; t += 32) {
        const int slot = t / nm, which = t - slot * nm;
        const int m = which ? mb : ma;
        const int count = bc7_slot_count(P, m);
        int best = -1, best_err = kErrNone;
        for (int n = 0; n < count; n++) {
            const int e = W.cand_err[slot][which][n];
            if (e < best_err) { best_err = e; best = n; }
        }
        W.win_shape[kBc7Batch * W.half + kBc7Slots * g + slot][m] = (best < 0) ? -1 : bc7_slot_shape(W, slot, m, best);
    }
}
// Chain phase over the whole round.  With nb = 16 (two halves) or 8 blocks, lane L works for block L % nb in every pass and
// runs the active roles L / nb, L / nb + 32 / nb, ...: a pass holds 32 / nb consecutive roles for all blocks, so its lanes run
// the same KIND of role (partitioned / mode 4-5 / mode 6) almost everywhere, and a lane keeps the best candidate of ITS block
// in registers.  Sixteen blocks make a pass two roles wide: (mode 0, mode 2), (mode 1, mode 3), (rotation, rotation), ... are
// equally long, where four-wide passes pair three-subset with two-subset chains and leave the last pass half empty.
// The first strict minimum in role order is kept (K:1358, :1638, :1650, :1684): a lane meets its roles in ascending order,
// the store phase compares (error, role).
ITW_HD void bc7_phase_chains(int lane, Bc7Warp& W, const Bc7Params& P)
{
    const int nactive = bc7_active_roles(P);
    const int nb = (W.ntotal > kBc7Batch) ? kBc7Super : kBc7Batch;
    const int slot = lane & (nb - 1);
    int best_err = kErrNone, best_role = 0;
    u32 code[4] = {0u, 0u, 0u, 0u};
    if (slot < W.ntotal)
        for (int i = lane / nb; i < nactive; i += 32 / nb) {
            const int r = bc7_active_role(P, i);
            const Bc7Result res = bc7_chain(W, P, lane, slot, r);
            if (res.err < best_err) {
                best_err = res.err; best_role = r;
#pragma unroll
                for (int j = 0; j < 4; j++) code[j] = res.code[j];
            }
        }
    W.fin.err[lane] = best_err;
    W.fin.role[lane] = best_role;
#pragma unroll
    for (int i = 0; i < 4; i++) W.fin.code[lane][i] = code[i];
}
// the block's winner among the lanes that worked for it, then the 16-byte store; K:2027
ITW_HD void bc7_phase_store(int lane, Bc7Warp& W, const Bc7Params& P, uint8_t* dst, long long first_block)
{
    (void)P;
    const int nb = (W.ntotal > kBc7Batch) ? kBc7Super : kBc7Batch;
    for (int t = lane; t < W.ntotal; t += 32) {
        int best = t;
        for (int j = 1; j < 32 / nb; j++) {
            const int l = t + nb * j;
            if (W.fin.err[l] < W.fin.err[best] || (W.fin.err[l] == W.fin.err[best] && W.fin.role[l] < W.fin.role[best])) best = l;
        }
        u32* out = reinterpret_cast<u32*>(dst + (size_t)(first_block + t) * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = W.fin.code[best][i];
    }
}

// The program of one round (up to kBc7Super blocks: `nvalid` of them from `first_block` of `surf`, results to block
// `out_block` of `dst`), as a list of (phase, barrier) pairs.
#define ITW_BC7_GROUP_PROGRAM(PHASE, g)                                                \
    if (P.sel[0]) {                                                                    \
        PHASE(bc7_phase_masks3(lane, W, P, 2 * (g)));                                  \
        PHASE(bc7_phase_shapes3(lane, W, P, 2 * (g)));                                 \
        PHASE(bc7_phase_masks3(lane, W, P, 2 * (g) + 1));                              \
        PHASE(bc7_phase_shapes3(lane, W, P, 2 * (g) + 1));                             \
        PHASE(bc7_phase_winners(lane, W, P, g, 0, 1));                                 \
    }                                                                                  \
    if (bc7_needs_keys(P, 0)) {                                                        \
        PHASE(bc7_phase_keys(lane, W, P, g, 0));                                       \
        PHASE(bc7_phase_rank(lane, W, g, 0));                                          \
        PHASE(bc7_phase_shapes(lane, W, P, g, 2, 3));                                  \
        PHASE(bc7_phase_winners(lane, W, P, g, 2, 3));                                 \
    }                                                                                  \
    if (bc7_needs_keys(P, 1)) {                                                        \
        PHASE(bc7_phase_keys(lane, W, P, g, 1));                                       \
        PHASE(bc7_phase_rank(lane, W, g, 1));                                          \
        PHASE(bc7_phase_shapes(lane, W, P, g, 4, 4));                                  \
        PHASE(bc7_phase_winners(lane, W, P, g, 4, 4));                                 \
    }
// The halves and the groups are LOOPS, not copies: four inlined copies of every shape phase made the kernel body 6.6 k
// instructions (105 KB) -- the phases of one copy are what the instruction cache should hold.
#define ITW_BC7_SHAPE_PROGRAM(PHASE)                                                   \
    ITW_UNROLL(1)                                                                      \
    for (int half = 0; half < ((per_warp > kBc7Batch) ? 2 : 1); half++) {              \
        PHASE(bc7_phase_load(lane, W, surf, first_block, nvalid, half));               \
        PHASE(bc7_phase_planes(lane, W));                                              \
        ITW_UNROLL(1)                                                                  \
        for (int g = 0; g < kBc7Batch / kBc7Slots; g++) {                              \
            ITW_BC7_GROUP_PROGRAM(PHASE, g)                                            \
        }                                                                              \
    }                                                                                  \
    if (per_warp > kBc7Batch) {                                                        \
        PHASE(bc7_phase_park(lane, W, surf, first_block, nvalid));                     \
    }
#define ITW_BC7_CHAIN_PROGRAM(PHASE)                                                   \
    PHASE(bc7_phase_chains(lane, W, P));                                               \
    PHASE(bc7_phase_store(lane, W, P, dst, out_block));
#define ITW_BC7_PROGRAM(PHASE)                                                         \
    ITW_BC7_SHAPE_PROGRAM(PHASE)                                                       \
    ITW_BC7_CHAIN_PROGRAM(PHASE)

#if defined(__CUDACC__)
// All warps of a CTA walk the phases in lock step (block barrier between phases) and one CTA fills an
// SM: at any moment every warp of the SM is inside the same few hundred instructions, which is what
// the instruction cache needs -- the first version of this kernel lost 90 % of its issue slots to
// instruction fetch (profiles/r1_bc7_v1_ncu.txt).  Work per phase is the same for every warp, so the
// barriers cost little.
constexpr int kBc7WarpsPerCta = 16;
constexpr size_t kBc7SmemBytes = sizeof(Bc7Warp) * kBc7WarpsPerCta;

// kTma: the consecutive blocks a CTA works on in one round (16 warps x per_warp blocks) are fetched by the TMA engine
// (cp.async.bulk, SASS UBLKCP) into a 4-row shared-memory tile, signalled through an mbarrier; the load phases read the tile
// from shared memory.  One tile buffer is enough: it is free again once the last load phase of a round is over, and the next
// tile then streams in behind the whole chain phase.  Needs 16-byte aligned surface rows; other surfaces use the plain
// global-load variant.
// per_warp = kBc7Super for surfaces large enough to give every warp of the grid 16 blocks, kBc7Batch for small ones (more warps
// busy; the chain phase then runs four roles wide).
constexpr int kBc7TileBlocks = kBc7WarpsPerCta * kBc7Super;              // 256
constexpr int kBc7TileRowBytes = kBc7TileBlocks * 16;                     // 4096

template <bool kTma>
__global__ void __launch_bounds__(kBc7WarpsPerCta * 32, 1)
bc7_kernel(SurfaceView gsurf, uint8_t* __restrict__ dst, Bc7Params P, long long nblocks, int per_warp)
{
    extern __shared__ __align__(16) unsigned char bc7_smem[];
    __shared__ __align__(128) unsigned char stage[kTma ? 4 * kBc7TileRowBytes : 16];
    __shared__ __align__(8) unsigned long long full;
    Bc7Warp& W = reinterpret_cast<Bc7Warp*>(bc7_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nbatches = (nblocks + per_warp - 1) / per_warp;
    const long long nwarps = (long long)gridDim.x * kBc7WarpsPerCta;
    const long long rounds = (nbatches + nwarps - 1) / nwarps;            // same trip count for every warp of the CTA
    const int tile_blocks = kBc7WarpsPerCta * per_warp;

    auto tile_first = [&](long long round) { return ((long long)blockIdx.x * kBc7WarpsPerCta + round * nwarps) * per_warp; };
    auto prefetch = [&](long long round) {                                // one thread feeds the TMA engine
        const long long fb = tile_first(round);
        if (round >= rounds || fb >= nblocks) return;
        const long long left = nblocks - fb;
        tma_prefetch_tile(stage, kBc7TileRowBytes, &full, gsurf, fb, (int)(left < tile_blocks ? left : tile_blocks), 16);
    };
    if (kTma) {
        if (threadIdx.x == 0) mbar_init(&full, 1);
        __syncthreads();
        if (threadIdx.x == 0) prefetch(0);
    }
    for (long long round = 0; round < rounds; round++) {
        const long long batch = (long long)blockIdx.x * kBc7WarpsPerCta + warp + round * nwarps;
        const long long out_block = batch * per_warp;
        const long long left = nblocks - out_block;
        const int nvalid = (int)(left <= 0 ? 0 : (left < per_warp ? left : per_warp));
        SurfaceView surf = gsurf;
        long long first_block = out_block;
        if (kTma) {
            if (tile_first(round) < nblocks) mbar_wait(&full, (unsigned)(round & 1));
            surf = SurfaceView{stage, kBc7TileBlocks * 4, 4, kBc7TileRowBytes};
            first_block = (long long)warp * per_warp;                    // block index inside the staged tile
        }
#define ITW_PHASE_DEVICE(call) call; __syncthreads()
        ITW_BC7_SHAPE_PROGRAM(ITW_PHASE_DEVICE)
        if (kTma && threadIdx.x == 0) prefetch(round + 1);               // every load phase of this round is behind a barrier
        ITW_BC7_CHAIN_PROGRAM(ITW_PHASE_DEVICE)
#undef ITW_PHASE_DEVICE
    }
}
#endif

}  // namespace itw
