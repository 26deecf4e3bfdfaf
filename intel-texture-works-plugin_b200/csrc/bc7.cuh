// bc7.cuh -- BC7 encoder (reference: kernel.ispc:616-2037, cited as K:line).
//
// Mapping.  One WARP owns a batch of kSlots 4x4 blocks; its 32 lanes are spread over the block's
// independent work items instead of over texels:
//   * candidate phases: lane <-> (block slot, partition candidate) -- each lane fits the PCA
//     segments of its candidate, quantises the endpoints and runs the index search alone, reading
//     the block's texels from shared memory (broadcast), so the reference's texel-order float sums
//     are reproduced bit for bit;
//   * ranking phase: lane <-> (block slot, two-subset shape) for the 64 PCA split bounds;
//   * chain phase: lane <-> (block slot, mode) -- the per-mode refinement chains, the mode 4/5
//     rotation/index-swap candidates and mode 6 are independent of one another in the reference
//     (each only competes through a strict `<` on the final error, K:1358, :1638, :1650, :1684),
//     so they run side by side and the winner is the first minimum in the reference's order
//     mode 0, 2, 1, 3, 7, 4(rot,swap), 5(rot), 6.
// Batching kSlots blocks per warp keeps the lanes of the short phases (16 mode-0 candidates, <=18
// chains per block) busy.  All candidate errors are exact integers < 2^23 (K:1178-1189 sums
// truncated per-texel errors), so "first minimum in list order" is a plain scan.
//
// The phases are written as per-lane functions separated by warp barriers; the same functions are
// driven lane by lane on the CPU by tests/emu (test-only).
#pragma once
#include "bc67_core.cuh"

namespace itw {

// bc7_enc_settings (ispc_texcomp.h:27-41) flattened to ints for the device
struct Bc7Params {
    int sel[4];
    int refine[8];
    int skip2, t1, t3, t7, ch0, rch, channels;
};

constexpr int kBc7Slots = 4;          // blocks per warp batch
constexpr int kBc7MaxRoles = 18;      // 5 partitioned modes + up to 8 mode-4 + 4 mode-5 + mode 6

// Per-warp scratch in shared memory
struct Bc7Warp {
    float px[kBc7Slots][64];                   // planar texels: px[c*16 + k]
    float cand_err[kBc7Slots][5][64];          // per (mode slot m, list position)
    int keys[kBc7Slots][2][64];                // split-bound keys: [0] RGB (modes 1,3), [1] profile channels (mode 7)
    int order[kBc7Slots][2][64];               // keys in ascending order
    int win_pos[kBc7Slots][5];                 // winning list position per mode slot, -1 = none
    float res_err[kBc7Slots][kBc7MaxRoles];
    u32 res_code[kBc7Slots][kBc7MaxRoles][4];
    int nvalid;
};
// mode slots m = 0..4 <-> BC7 modes {0, 2, 1, 3, 7}: the reference's evaluation order
ITW_HD int bc7_slot_mode(int m) { return (m == 0) ? 0 : ((m == 1) ? 2 : ((m == 2) ? 1 : ((m == 3) ? 3 : 7))); }

ITW_HD int bc7_mode_bits(int mode) { return (mode == 0 || mode == 1) ? 3 : 2; }
ITW_HD int bc7_mode_channels(int mode) { return (mode == 7) ? 4 : 3; }

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

// Fit + quantise + index search of one partition candidate; K:1279-1297
ITW_HD float bc7_eval_partitioned(const float* px, int mode, int shape, int* q, u32& idx0, u32& idx1)
{
    const int pairs = bc7_pairs(mode), channels = bc7_mode_channels(mode);
    float ep[24];
#pragma unroll
    for (int i = 0; i < 24; i++) ep[i] = 0.0f;                 // never-written slots read as zero (F6)
    for (int j = 0; j < pairs; j++) fit_segment(ep + 8 * j, px, shape_mask(shape, j), channels, true);
    for (int j = 0; j < pairs; j++) bc7_quantise_pair(q + 8 * j, ep + 8 * j, mode, channels);
    return assign_indices(idx0, idx1, px, bc7_mode_bits(mode), ep, shape_pattern(shape), channels);
}

// ---- 128-bit layouts; K:1807-1964 ----
ITW_HD void bc7_write_partitioned(u32 (&out)[4], int* q, u32 idx0, u32 idx1, int shape, int mode)
{
    const int bits = bc7_mode_bits(mode), pairs = bc7_pairs(mode), channels = bc7_mode_channels(mode);
    int flips = orient_subsets(q, idx0, idx1, bits, pairs, shape);
    BitSink s;
    s.reset();
    s.put(mode + 1, 1u << mode);
    s.put(mode == 0 ? 4 : 6, (u32)(shape & (mode == 0 ? 15 : 63)));
    const int width = (mode == 0) ? 4 : ((mode == 1) ? 6 : ((mode == 3) ? 7 : 5));
    const int drop = (mode == 2) ? 0 : 1;                      // p-bit modes store the value without its LSB
    for (int c = 0; c < channels; c++)
        for (int j = 0; j < pairs * 2; j++) s.put(width, (u32)(q[4 * j + c] >> drop));
    if (mode == 1)
        for (int j = 0; j < 2; j++) s.put(1, (u32)(q[8 * j] & 1));
    if (mode == 0 || mode == 3 || mode == 7)
        for (int j = 0; j < pairs * 2; j++) s.put(1, (u32)(q[4 * j] & 1));
    put_indices(s, idx0, idx1, bits, flips, shape_anchor(shape, 1), (pairs == 3) ? shape_anchor(shape, 2) : -1);
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}

// ---- chain: refinement of one partitioned mode's winner; K:1329-1362 ----
ITW_HD void bc7_chain_partitioned(Bc7Warp& W, const Bc7Params& P, int slot, int m)
{
    float& out_err = W.res_err[slot][m];
    out_err = inf_f();
    const int pos = W.win_pos[slot][m];
    if (pos < 0) return;
    const float* px = W.px[slot];
    const int mode = bc7_slot_mode(m);
    const int bits = bc7_mode_bits(mode), pairs = bc7_pairs(mode), channels = bc7_mode_channels(mode);
    const int shape = bc7_slot_shape(W, slot, m, pos);

    int best_q[24];
    u32 best_i0, best_i1;
    // the winner's integers are recomputed here instead of being carried out of the candidate phase
    float best_err = bc7_eval_partitioned(px, mode, shape, best_q, best_i0, best_i1);

    for (int it = 0; it < P.refine[mode]; it++) {
        float ep[24];
        int q[24];
#pragma unroll
        for (int i = 0; i < 24; i++) ep[i] = 0.0f;
        for (int j = 0; j < pairs; j++) solve_endpoints(ep + 8 * j, px, bits, best_i0, best_i1, shape_mask(shape, j), channels);
        for (int j = 0; j < pairs; j++) bc7_quantise_pair(q + 8 * j, ep + 8 * j, mode, P.channels);   // profile's channels; K:1343
        u32 i0, i1;
        float err = assign_indices(i0, i1, px, bits, ep, shape_pattern(shape), channels);
        if (err < best_err) {
            for (int i = 0; i < 8 * pairs; i++) best_q[i] = q[i];
            best_i0 = i0; best_i1 = i1;
            best_err = err;
        }
    }
    if (mode != 7) {                                           // opaque error of the dropped alpha; K:1267-1277, :1356
        float opaque = 0.0f;
        if (P.channels != 3)
            for (int k = 0; k < 16; k++) opaque += sq(px[48 + k] - 255.0f);
        best_err += opaque;
    }
    out_err = best_err;
    bc7_write_partitioned(W.res_code[slot][m], best_q, best_i0, best_i1, shape, mode);
}

// ---- scalar channel of modes 4/5; K:1437-1563 ----
ITW_HD void scalar_quantise(int (&q)[2], float (&ep)[2], int epbits)
{
    const int top = (1 << epbits) - 1;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        q[i] = clampi(cvt_x86(ep[i] / 255.0f * (float)top + 0.5f), 0, top);
        ep[i] = (float)expand_bits(q[i], epbits);
    }
}
ITW_HD float scalar_assign(u32& idx0, u32& idx1, const float* a, int bits, const float (&ep)[2])
{
    const int levels = 1 << bits;
    u32 out[2] = {0u, 0u};
    float total = 0.0f;
    for (int k = 0; k < 16; k++) {
        float proj = (a[k] - ep[0]) / (ep[1] - ep[0] + 0.001f);
        int q1 = clampi(cvt_x86(proj * (float)levels + 0.5f), 1, levels - 1);
        float fw0 = (float)bc7_weight(bits, q1 - 1), fw1 = (float)bc7_weight(bits, q1);
        float d0 = (float)cvt_x86(((64.0f - fw0) * ep[0] + fw0 * ep[1] + 32.0f) / 64.0f);
        float d1 = (float)cvt_x86(((64.0f - fw1) * ep[0] + fw1 * ep[1] + 32.0f) / 64.0f);
        float err0 = sq(d0 - a[k]), err1 = sq(d1 - a[k]);
        int best_err = cvt_x86(err1), best_q = q1;
        if (err0 < err1) { best_err = cvt_x86(err0); best_q = q1 - 1; }
        out[k >> 3] += (u32)best_q << (4 * (k & 7));
        total += (float)best_err;
    }
    idx0 = out[0];
    idx1 = out[1];
    return total;
}
ITW_HD void scalar_solve(float (&ep)[2], const float* a, int bits, u32 idx0, u32 idx1)
{
    const float top = (float)((1 << bits) - 1);
    float atb1 = 0.0f, sq1 = 0.0f, sqq = 0.0f, sum = 0.0f;
    for (int k = 0; k < 16; k++) {
        float q = (float)(((k < 8 ? idx0 : idx1) >> (4 * (k & 7))) & 15u);
        float x = (float)cvt_x86(top - q);
        sq1 += q;
        sqq += q * q;
        sum += a[k];
        atb1 += x * a[k];
    }
    float atb2 = top * sum - atb1;
    float cxx = 16.0f * sq(top) - (2.0f * top) * sq1 + sqq;
    float cyy = sqq;
    float cxy = top * sq1 - sqq;
    float det = cxx * cyy - cxy * cxy;
    float scale = top / det;
    ep[0] = clamp_sse((atb1 * cyy - atb2 * cxy) * scale, 0.0f, 255.0f);
    ep[1] = clamp_sse((atb2 * cxx - atb1 * cxy) * scale, 0.0f, 255.0f);
    if (fabsf(det) < 0.001f) {
        ep[0] = sum / 16.0f;
        ep[1] = ep[0];
    }
}

// ---- chain: one mode 4/5 candidate (rotation, index swap); K:1565-1621, :1879-1939 ----
ITW_HD void bc7_chain_mode45(Bc7Warp& W, const Bc7Params& P, int slot, int role, int mode, int rotation, int swap)
{
    const float* src = W.px[slot];
    int bits = 2, abits = (mode == 4) ? 3 : 2;
    const int aepbits = (mode == 4) ? 6 : 8;
    if (swap == 1) { bits = 3; abits = 2; }

    // rotated copy of the colour planes: the rotated-in plane is alpha, or 255 for RGB profiles
    float px[48];
    for (int k = 0; k < 16; k++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float v = src[16 * c + k];
            if (c == rotation) v = (P.channels == 4) ? src[48 + k] : 255.0f;
            px[16 * c + k] = v;
        }
    }
    float ep[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ep[i] = 0.0f;
    fit_segment(ep, px, 0xFFFF, 3, true);
    bc7_quantise_pair(q, ep, mode, 3);
    u32 i0, i1;
    float err = assign_indices(i0, i1, px, bits, ep, 0u, 3);
    for (int it = 0; it < P.refine[mode]; it++) {
        solve_endpoints(ep, px, bits, i0, i1, 0xFFFF, 3);
        bc7_quantise_pair(q, ep, mode, 3);
        err = assign_indices(i0, i1, px, bits, ep, 0u, 3);
    }
    // the channel that was rotated out (always the ORIGINAL plane `rotation`; K:1608)
    const float* a = src + 16 * rotation;
    float aep[2] = {255.0f, 0.0f};
    for (int k = 0; k < 16; k++) { aep[0] = min_sse(aep[0], a[k]); aep[1] = max_sse(aep[1], a[k]); }
    int aq[2];
    u32 a0, a1;
    scalar_quantise(aq, aep, aepbits);
    float aerr = scalar_assign(a0, a1, a, abits, aep);
    for (int it = 0; it < P.rch; it++) {
        scalar_solve(aep, a, abits, a0, a1);
        scalar_quantise(aq, aep, aepbits);
        aerr = scalar_assign(a0, a1, a, abits, aep);
    }
    err += aerr;
    W.res_err[slot][role] = err;

    // layout; K:1879-1939
    const int epbits = (mode == 4) ? 5 : 7;
    const int cbits = 2, sbits = (mode == 4) ? 3 : 2;          // widths of the first / second index set
    if (!swap) {
        orient_single(q, 4, i0, i1, cbits);
        orient_single(aq, 1, a0, a1, sbits);
    } else {                                                    // the two index sets trade places
        u32 t0 = i0, t1 = i1;
        i0 = a0; i1 = a1;
        a0 = t0; a1 = t1;
        orient_single(aq, 1, i0, i1, cbits);
        orient_single(q, 4, a0, a1, sbits);
    }
    BitSink s;
    s.reset();
    s.put(mode + 1, 1u << mode);
    s.put(2, (u32)((rotation + 1) & 3));
    if (mode == 4) s.put(1, (u32)swap);
#pragma unroll
    for (int c = 0; c < 3; c++) { s.put(epbits, (u32)q[c]); s.put(epbits, (u32)q[4 + c]); }
    s.put(aepbits, (u32)aq[0]);
    s.put(aepbits, (u32)aq[1]);
    put_indices(s, i0, i1, cbits, 0, -1, -1);
    put_indices(s, a0, a1, sbits, 0, -1, -1);
    u32* out = W.res_code[slot][role];
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}

// ---- chain: mode 6; K:1657-1689, :1941-1964 ----
ITW_HD void bc7_chain_mode6(Bc7Warp& W, const Bc7Params& P, int slot, int role)
{
    const float* px = W.px[slot];
    const int channels = P.channels;
    float ep[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ep[i] = 0.0f;
    fit_segment(ep, px, 0xFFFF, channels, true);
    if (channels == 3) ep[3] = ep[7] = 255.0f;
    bc7_quantise_pair(q, ep, 6, channels);
    u32 i0, i1;
    float err = assign_indices(i0, i1, px, 4, ep, 0u, channels);
    for (int it = 0; it < P.refine[6]; it++) {
        solve_endpoints(ep, px, 4, i0, i1, 0xFFFF, channels);
        bc7_quantise_pair(q, ep, 6, channels);
        err = assign_indices(i0, i1, px, 4, ep, 0u, channels);
    }
    W.res_err[slot][role] = err;
    orient_single(q, 4, i0, i1, 4);
    BitSink s;
    s.reset();
    s.put(7, 64u);
#pragma unroll
    for (int c = 0; c < 4; c++) { s.put(7, (u32)(q[c] >> 1)); s.put(7, (u32)(q[4 + c] >> 1)); }
    s.put(1, (u32)(q[0] & 1));
    s.put(1, (u32)(q[4] & 1));
    put_indices(s, i0, i1, 4, 0, -1, -1);
    u32* out = W.res_code[slot][role];
    out[0] = s.w0; out[1] = s.w1; out[2] = s.w2; out[3] = s.w3;
}

// =============================================================================================
// Warp program: per-lane phase functions.  A phase reads what earlier phases wrote to W and
// writes disjoint locations; the caller separates phases with a warp barrier.
// =============================================================================================
// texels of `nvalid` consecutive blocks starting at first_block -> W.px
ITW_HD void bc7_phase_load(int lane, Bc7Warp& W, const SurfaceView& s, long long first_block, int nvalid)
{
    const int bw = s.width >> 2;
    for (int t = lane; t < nvalid * 16; t += 32) {
        const int slot = t >> 4, k = t & 15;
        const long long id = first_block + slot;
        const int by = (int)(id / bw), bx = (int)(id - (long long)by * bw);
        const uint8_t* p = s.ptr + (size_t)(by * 4 + (k >> 2)) * (size_t)s.stride + (size_t)(bx * 4 + (k & 3)) * 4;
        float* px = W.px[slot];
        px[k] = (float)p[0];
        px[16 + k] = (float)p[1];
        px[32 + k] = (float)p[2];
        px[48 + k] = (float)p[3];
    }
    if (lane == 0) W.nvalid = nvalid;
}
// candidates of mode slot m (m = 0,1 need no ranking; m = 2,3,4 need bc7_phase_rank first)
ITW_HD void bc7_phase_candidates(int lane, Bc7Warp& W, const Bc7Params& P, int m)
{
    const int count = bc7_slot_count(P, m);
    const int mode = bc7_slot_mode(m);
    for (int t = lane; t < W.nvalid * count; t += 32) {
        const int slot = t / count, n = t - slot * count;
        int q[24];
        u32 i0, i1;
        W.cand_err[slot][m][n] = bc7_eval_partitioned(W.px[slot], mode, bc7_slot_shape(W, slot, m, n), q, i0, i1);
    }
}
// split-bound keys of the 64 two-subset shapes; set 0 = RGB (modes 1,3), set 1 = profile channels (mode 7)
ITW_HD bool bc7_needs_keys(const Bc7Params& P, int set)
{
    if (!P.sel[1]) return false;
    return set == 0 ? !(P.t1 == 0 && P.t3 == 0) : (P.t7 != 0);
}
ITW_HD void bc7_phase_keys(int lane, Bc7Warp& W, const Bc7Params& P, int set)
{
    const int channels = (set == 0) ? 3 : P.channels;
    for (int t = lane; t < W.nvalid * 64; t += 32) {
        const int slot = t >> 6, shape = t & 63;
        float full[15];
        masked_moments(full, W.px[slot], 0xFFFF, channels);
        W.keys[slot][set][shape] = split_bound_key(W.px[slot], shape, full, channels);
    }
}
ITW_HD void bc7_phase_rank(int lane, Bc7Warp& W, int set)
{
    for (int t = lane; t < W.nvalid * 64; t += 32) {
        const int slot = t >> 6, i = t & 63;
        W.order[slot][set][rank_of(W.keys[slot][set], 64, i)] = W.keys[slot][set][i];
    }
}
// first minimum of each mode slot's candidate list; K:1320 (strict <)
ITW_HD void bc7_phase_winners(int lane, Bc7Warp& W, const Bc7Params& P)
{
    for (int t = lane; t < W.nvalid * 5; t += 32) {
        const int slot = t / 5, m = t - slot * 5;
        const int count = bc7_slot_count(P, m);
        int best = -1;
        float best_err = inf_f();
        for (int n = 0; n < count; n++) {
            float e = W.cand_err[slot][m][n];
            if (e < best_err) { best_err = e; best = n; }
        }
        W.win_pos[slot][m] = best;
    }
}
ITW_HD void bc7_phase_chain_partitioned(int lane, Bc7Warp& W, const Bc7Params& P)
{
    for (int t = lane; t < W.nvalid * 5; t += 32) bc7_chain_partitioned(W, P, t / 5, t % 5);
}
ITW_HD int bc7_rotations(const Bc7Params& P) { return P.sel[2] ? maxi(P.channels - P.ch0, 0) : 0; }
ITW_HD void bc7_phase_chain_mode45(int lane, Bc7Warp& W, const Bc7Params& P)
{
    const int nrot = bc7_rotations(P), per = 3 * nrot;        // 2*nrot mode-4 roles then nrot mode-5 roles
    for (int t = lane; t < W.nvalid * per; t += 32) {
        const int slot = t / per, r = t - slot * per;
        if (r < 2 * nrot) bc7_chain_mode45(W, P, slot, 5 + r, 4, P.ch0 + (r >> 1), r & 1);
        else              bc7_chain_mode45(W, P, slot, 5 + r, 5, P.ch0 + (r - 2 * nrot), 0);
    }
}
ITW_HD void bc7_phase_chain_mode6(int lane, Bc7Warp& W, const Bc7Params& P)
{
    if (!P.sel[3]) return;
    const int role = 5 + 3 * bc7_rotations(P);
    for (int t = lane; t < W.nvalid; t += 32) bc7_chain_mode6(W, P, t, role);
}
// first strict minimum over the roles in the reference's order, then the 16-byte store; K:2027
ITW_HD void bc7_phase_store(int lane, Bc7Warp& W, const Bc7Params& P, uint8_t* dst, long long first_block)
{
    const int nroles = 5 + 3 * bc7_rotations(P) + (P.sel[3] ? 1 : 0);
    for (int t = lane; t < W.nvalid; t += 32) {
        float best_err = inf_f();
        u32 code[4] = {0u, 0u, 0u, 0u};
        for (int r = 0; r < nroles; r++) {
            float e = W.res_err[t][r];
            if (e < best_err) {
                best_err = e;
#pragma unroll
                for (int i = 0; i < 4; i++) code[i] = W.res_code[t][r][i];
            }
        }
        u32* out = reinterpret_cast<u32*>(dst + (size_t)(first_block + t) * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = code[i];
    }
}

// The whole program for one batch, as a list of (phase, barrier) pairs.  SYNC is __syncwarp() on
// the device; the CPU emulation runs each phase for lanes 0..31 in turn.
#define ITW_BC7_PROGRAM(PHASE)                                                         \
    PHASE(bc7_phase_load(lane, W, surf, first_block, nvalid));                         \
    PHASE(bc7_phase_candidates(lane, W, P, 0));                                        \
    PHASE(bc7_phase_candidates(lane, W, P, 1));                                        \
    if (bc7_needs_keys(P, 0)) {                                                        \
        PHASE(bc7_phase_keys(lane, W, P, 0));                                          \
        PHASE(bc7_phase_rank(lane, W, 0));                                             \
        PHASE(bc7_phase_candidates(lane, W, P, 2));                                    \
        PHASE(bc7_phase_candidates(lane, W, P, 3));                                    \
    }                                                                                  \
    if (bc7_needs_keys(P, 1)) {                                                        \
        PHASE(bc7_phase_keys(lane, W, P, 1));                                          \
        PHASE(bc7_phase_rank(lane, W, 1));                                             \
        PHASE(bc7_phase_candidates(lane, W, P, 4));                                    \
    }                                                                                  \
    PHASE(bc7_phase_winners(lane, W, P));                                              \
    PHASE(bc7_phase_chain_partitioned(lane, W, P));                                    \
    PHASE(bc7_phase_chain_mode45(lane, W, P));                                         \
    PHASE(bc7_phase_chain_mode6(lane, W, P));                                          \
    PHASE(bc7_phase_store(lane, W, P, dst, first_block));

#if defined(__CUDACC__)
constexpr int kBc7WarpsPerCta = 4;

__global__ void __launch_bounds__(kBc7WarpsPerCta * 32)
bc7_kernel(SurfaceView surf, uint8_t* __restrict__ dst, Bc7Params P, long long nblocks)
{
    __shared__ Bc7Warp warps[kBc7WarpsPerCta];
    Bc7Warp& W = warps[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const long long nbatches = (nblocks + kBc7Slots - 1) / kBc7Slots;
    const long long warp0 = (long long)blockIdx.x * kBc7WarpsPerCta + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kBc7WarpsPerCta;
    for (long long batch = warp0; batch < nbatches; batch += nwarps) {
        const long long first_block = batch * kBc7Slots;
        const int nvalid = (int)((nblocks - first_block < kBc7Slots) ? (nblocks - first_block) : kBc7Slots);
#define ITW_PHASE_DEVICE(call) call; __syncwarp()
        ITW_BC7_PROGRAM(ITW_PHASE_DEVICE)
#undef ITW_PHASE_DEVICE
    }
}
#endif

}  // namespace itw
