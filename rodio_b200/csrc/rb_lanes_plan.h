// rb_lanes_plan.h — host-side arithmetic of the lane-per-stream kernel's plan (plain C++: shared by rb_lanes.cu and
// by the CPU emulator under tests/emu/ so that both fill lanes::Row / lanes::Args identically).
#pragma once
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "rb_lanes_core.h"

namespace lanes {

// Outputs [0, n) of a converter run over L frames whose left frame is <= L-2, i.e. that interpolate:
// n = ceil((L-1) * to / from), clipped to the stream's length (src/conversions/sample_rate.rs:157-199: the last
// frame is emitted raw, once).
inline uint64_t n_interp(uint64_t L, uint32_t from, uint32_t to, uint64_t out_len) {
    if (L < 2) return 0;
    const uint64_t n = ((L - 1) * (uint64_t)to + from - 1) / from;
    return n < out_len ? n : out_len;
}

// b1 == +-2*b0 and b2 == b0 (low_pass / high_pass of src/source/blt.rs:504-541: b0 = b1/2 resp. b1 = -(2*b0),
// b2 = b0, all three divided by the same a0 -- halving and doubling commute with a correctly rounded division).
// b0 must be far enough from the subnormals that b0*x is normal for every x the classified inputs produce.
inline bool ff2_coeffs(float b0, float b1, float b2, float* ffk) {
    uint32_t u0, u2;
    std::memcpy(&u0, &b0, 4), std::memcpy(&u2, &b2, 4);
    if (u0 != u2) return false;
    const float a0 = b0 < 0 ? -b0 : b0;
    if (!(a0 >= 9.3132257e-10f && a0 <= 1.0737418e9f)) return false;   // [2^-30, 2^30]
    if (b1 == 2.0f * b0) *ffk = 2.0f;
    else if (b1 == -2.0f * b0) *ffk = -2.0f;
    else return false;
    return true;
}

// A gain in front of the conversion scales what the exact-reciprocal class was measured on.  For |g| in [2^-6, 2^6] the scaled
// taps are 0 or in [2^-76, 2^66): their difference is 0 or at least one ulp of the smaller tap (>= 2^-99), times num >= 1, and
// below 2^67 * 2^20 -- inside [2^-100, 2^100) where the reciprocal step is exact (rb_lanes_core.h).  Any other gain makes the
// class run the guarded tile (GUARD: every quotient checked), or, above the mixer's rate, sends the stream to the slow tiles.  (Zero -- Player::set_volume(0.0) -- makes every tap zero: exact.)
inline bool pre_gain_keeps_class(float g) {
    const float a = g < 0 ? -g : g;
    return a == 0.0f || (a >= 0.015625f && a <= 64.0f);   // a muted source (all taps zero) is exact as well; false for NaN
}

// A class above the mixer's rate, up to twice, runs on the DOWN instantiation (fast tiles); beyond that on the slow tiles.
inline bool ratio_runs_down(uint32_t from, uint32_t to) { return from > to && (uint64_t)from <= 2ull * to; }

inline void fill_ratio(Args& a, uint32_t from, uint32_t to, uint32_t channels) {
    a.from = from, a.to = to;
    const uint64_t tf = TILE / channels;   // frames per tile
    a.q8 = (uint32_t)((tf * from) / to);
    a.r8 = (uint32_t)((tf * from) % to);
    a.den_f = (float)to;
    a.rcp_den = 1.0f / a.den_f;
    a.from_f = (float)from;
    a.adv_q = from / to, a.rem_f = (float)(from % to);
    a.neg1 = -1.0f;
}

// Streams of one kernel launch share a reduced rate pair.  A batch / session with several pairs (44.1 kHz and 48 kHz
// sources in one mixer) is served class by class: stable partition by (from, to) in order of first appearance; every
// class gets its own launch over its own rows and its own partial rows, k_sum_groups adds all partial rows in order.
// Mono sources in a stereo mixer form classes of their own (the CI = 1, CO = 2 instantiation).
// `ch` (source channels per stream) may be NULL when all streams have the mixer's channel count.
inline std::vector<std::vector<uint32_t>> classes_by_ratio(const uint32_t* from, const uint32_t* to, const uint32_t* ch, uint32_t n) {
    struct Key { uint32_t from, to, ch; };
    std::vector<Key> keys;
    std::vector<std::vector<uint32_t>> out;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t c = ch ? ch[i] : 0;
        size_t k = 0;
        while (k < keys.size() && !(keys[k].from == from[i] && keys[k].to == to[i] && keys[k].ch == c)) k++;
        if (k == keys.size()) keys.push_back({from[i], to[i], c}), out.emplace_back();
        out[k].push_back(i);
    }
    return out;
}

inline uint64_t round_up_tile(uint64_t n) { return (n + TILE - 1) / TILE * TILE; }   // n in floats (frames * channels)

// Input classification for the exact-reciprocal division: every non-zero |x| inside [2^-70, 2^60].
inline bool sample_in_class(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0x7fffffffu;
    return u == 0 || (u >= 0x1c800000u && u < 0x5d800000u);   // 2^-70 = 0x1c800000, 2^60 = 0x5d800000
}

}  // namespace lanes
