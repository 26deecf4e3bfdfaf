// itw_params.h -- host-side conversion of the reference's settings structs (ispc_texcomp.h:27-50)
// into the flat int structs the kernels take, plus the profile tables (ispc_texcomp.cpp:20-410).
// Plain C++ (no CUDA): shared by the product's host API and by the test-only CPU emulation.
#pragma once
#include "../../include/itw_bcn.h"
#include "bc6h.cuh"
#include "bc7.cuh"

namespace itw {

inline Bc7Params bc7_params_from(const bc7_enc_settings& s)
{
    Bc7Params p;
    for (int i = 0; i < 4; i++) p.sel[i] = s.mode_selection[i] ? 1 : 0;
    for (int i = 0; i < 8; i++) p.refine[i] = s.refineIterations[i];
    p.skip2 = s.skip_mode2 ? 1 : 0;
    p.t1 = s.fastSkipTreshold_mode1;
    p.t3 = s.fastSkipTreshold_mode3;
    p.t7 = s.fastSkipTreshold_mode7;
    p.ch0 = s.mode45_channel0;
    p.rch = s.refineIterations_channel;
    p.channels = s.channels;
    return p;
}
// The reference indexes 64-entry candidate lists with these counts and has undefined behaviour
// outside the ranges below; the kernels reject such settings instead (itw_get_last_error).
inline const char* bc7_params_check(const Bc7Params& p)
{
    if (p.channels != 3 && p.channels != 4) return "bc7_enc_settings.channels must be 3 or 4";
    if (p.t1 < 0 || p.t1 > 64 || p.t3 < 0 || p.t3 > 64 || p.t7 < 0 || p.t7 > 64)
        return "bc7_enc_settings.fastSkipTreshold_mode* must be in [0,64]";
    if (p.ch0 < 0 || p.ch0 > 4) return "bc7_enc_settings.mode45_channel0 must be in [0,4]";
    for (int i = 0; i < 8; i++)
        if (i != 7 || p.t7 != 0)          // refineIterations[7] is left unset by the RGB profiles
            if (p.refine[i] < 0 || p.refine[i] > 64) return "bc7_enc_settings.refineIterations out of range";
    if (p.rch < 0 || p.rch > 64) return "bc7_enc_settings.refineIterations_channel out of range";
    return nullptr;
}
inline Bc6Params bc6_params_from(const bc6h_enc_settings& s)
{
    Bc6Params p;
    p.slow_mode = s.slow_mode ? 1 : 0;
    p.fast_mode = s.fast_mode ? 1 : 0;
    p.refine_1p = s.refineIterations_1p;
    p.refine_2p = s.refineIterations_2p;
    p.fast_skip = s.fastSkipTreshold;
    return p;
}
inline const char* bc6_params_check(const Bc6Params& p)
{
    if (p.fast_skip < 0 || p.fast_skip > 32) return "bc6h_enc_settings.fastSkipTreshold must be in [0,32]";
    if (p.refine_1p < 0 || p.refine_1p > 64 || p.refine_2p < 0 || p.refine_2p > 64)
        return "bc6h_enc_settings.refineIterations_* out of range";
    return nullptr;
}

// ---- profiles; ispc_texcomp.cpp:20-365.  One row per GetProfile_* in header order. ----
// {channels, sel0..3, skip2, t1, t3, t7, refine0..6, refine7 (-1: the RGB profiles leave it unset),
//  mode45_channel0, refineIterations_channel}
struct Bc7ProfileRow { signed char v[20]; };
static const Bc7ProfileRow kBc7ProfileRows[10] = {
    {{3, 0, 0, 0, 1, 1, 3, 1, 0, 2, 2, 2, 1, 2, 2, 1, -1, 0, 0, 0}},     // ultrafast        :20-50
    {{3, 0, 1, 0, 1, 1, 3, 1, 0, 2, 2, 2, 1, 2, 2, 1, -1, 0, 0, 0}},     // veryfast         :52-82
    {{3, 0, 1, 0, 1, 1, 12, 4, 0, 2, 2, 2, 1, 2, 2, 2, -1, 0, 0, 0}},    // fast             :84-120
    {{3, 1, 1, 1, 1, 1, 12, 8, 0, 2, 2, 2, 2, 2, 2, 2, -1, 0, 2, 0}},    // basic            :122-154
    {{3, 1, 1, 1, 1, 0, 64, 64, 0, 4, 4, 4, 4, 4, 4, 4, -1, 0, 4, 0}},   // slow             :156-189
    {{4, 0, 0, 1, 1, 1, 0, 0, 4, 2, 1, 2, 1, 1, 1, 2, 2, 3, 1, 0}},      // alpha_ultrafast  :191-224
    {{4, 0, 1, 1, 1, 1, 0, 0, 4, 2, 1, 2, 1, 2, 2, 2, 2, 3, 2, 0}},      // alpha_veryfast   :226-259
    {{4, 0, 1, 1, 1, 1, 4, 4, 8, 2, 1, 2, 1, 2, 2, 2, 2, 3, 2, 0}},      // alpha_fast       :261-294
    {{4, 1, 1, 1, 1, 1, 12, 8, 8, 2, 2, 2, 2, 2, 2, 2, 2, 0, 2, 0}},     // alpha_basic      :296-329
    {{4, 1, 1, 1, 1, 0, 64, 64, 64, 4, 4, 4, 4, 4, 4, 4, 4, 0, 4, 0}},   // alpha_slow       :331-365
};
inline void bc7_fill_profile(bc7_enc_settings* s, int row)
{
    const signed char* v = kBc7ProfileRows[row].v;
    s->channels = v[0];
    for (int i = 0; i < 4; i++) s->mode_selection[i] = v[1 + i] != 0;
    s->skip_mode2 = v[5] != 0;
    s->fastSkipTreshold_mode1 = v[6];
    s->fastSkipTreshold_mode3 = v[7];
    s->fastSkipTreshold_mode7 = v[8];
    for (int i = 0; i < 8; i++)
        if (v[9 + i] >= 0) s->refineIterations[i] = v[9 + i];
    s->mode45_channel0 = v[17];
    s->refineIterations_channel = v[18];
}
// {slow_mode, fast_mode, fastSkipTreshold, refine_1p, refine_2p}; ispc_texcomp.cpp:367-410
static const signed char kBc6ProfileRows[5][5] = {
    {0, 1, 0, 0, 0}, {0, 1, 2, 0, 1}, {0, 0, 4, 2, 2}, {1, 0, 10, 2, 2}, {1, 0, 32, 2, 2},
};
inline void bc6_fill_profile(bc6h_enc_settings* s, int row)
{
    const signed char* v = kBc6ProfileRows[row];
    s->slow_mode = v[0] != 0;
    s->fast_mode = v[1] != 0;
    s->fastSkipTreshold = v[2];
    s->refineIterations_1p = v[3];
    s->refineIterations_2p = v[4];
}

}  // namespace itw
