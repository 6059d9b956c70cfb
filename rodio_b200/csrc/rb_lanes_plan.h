// rb_lanes_plan.h — host-side arithmetic of the lane-per-stream kernel's plan (plain C++: shared by rb_lanes.cu and
// by the CPU emulator under tests/emu/ so that both fill lanes::Row / lanes::Args identically).
#pragma once
#include <cmath>
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

// ---- the lane-pair kernel (rb_duo_core.h) ----
// Two rows may share a lane when they are in phase on the mixer timeline: (o0 - mix_start) congruent modulo 4 * to -- the
// same numerator and the same offset of the left frame inside its 16-byte quad at every timeline frame.
inline bool duo_in_phase(const Row& a, const Row& b, uint32_t to) {
    const uint64_t m = 4ull * to;
    const uint64_t pa = (a.o0 % m + m - a.mix_start % m) % m, pb = (b.o0 % m + m - b.mix_start % m) % m;
    return pa == pb;
}
// A class goes to the lane-pair kernel when every lane's two rows are in phase (rows 2l, 2l + 1 of every group of 64).
inline bool duo_compatible(const Row* rows, size_t n, uint32_t to) {
    for (size_t i = 0; i + 1 < n; i += 2)
        if (rows[i].out_len && rows[i + 1].out_len && !duo_in_phase(rows[i], rows[i + 1], to)) return false;
    return true;
}

// ---- the time-parallel biquad plan (RB_BIQUAD_TIME_PARALLEL) ----
// Largest pole magnitude and noise gain (sum of squares of the impulse response of 1 / (1 + a1 z^-1 + a2 z^-2)) of a filter.
// The f32 recurrence of the reference (src/source/blt.rs:558-560) carries rounding noise of about 2.1e-7 * sqrt(gain) * peak
// (measured against an f64 run for 200 Hz ... 5 kHz, white noise; up to twice that for sines).  Two f32 trajectories of the
// same filter on the same input that start from different states differ by at most that much until they merge bit for
// bit, which they do within a few hundred samples above ~700 Hz and practically never at 200 Hz (tools/microbench/
// biquad_merge.cpp).  The plan is therefore taken only when 3.5e-7 * sqrt(gain) <= 7e-6 (gain <= 400: a cut-off of >= ~700 Hz
// at q = 0.5), leaving the north-star tolerance of 1e-5 * peak a margin; other filters keep the exact serial path.
inline bool tp_filter_ok(float a1, float a2, double* radius, double* gain) {
    const double A1 = a1, A2 = a2;
    const double disc = A1 * A1 - 4.0 * A2;
    double r;
    if (disc < 0) r = A2 > 0 ? std::sqrt(A2) : 2.0;
    else {
        const double s = std::sqrt(disc), r1 = (-A1 + s) / 2, r2 = (-A1 - s) / 2;
        r = (r1 < 0 ? -r1 : r1) > (r2 < 0 ? -r2 : r2) ? (r1 < 0 ? -r1 : r1) : (r2 < 0 ? -r2 : r2);
    }
    *radius = r, *gain = 1e300;
    if (!(r < 0.985)) return false;
    double h1 = 0, h2 = 0, g = 0;                 // h[n] = delta[n] - a1 h[n-1] - a2 h[n-2]
    for (int n = 0; n < 200000; n++) {
        const double h = (n == 0 ? 1.0 : 0.0) - A1 * h1 - A2 * h2;
        g += h * h, h2 = h1, h1 = h;
        if (n > 16 && h * h + h2 * h2 < 1e-24 * g) break;
    }
    *gain = g;
    return g <= 400.0;
}
// Warm-up in front of a segment: 60 time constants of the slowest pole -- the zero-state start has decayed far below an ulp
// long before, the length is chosen so that most segments have MERGED with the exact trajectory bit for bit when they
// begin to count (measured: 3 of 4 segments at 1 kHz; the rest stay within the rounding noise above).
inline uint32_t tp_warmup(double radius) {
    const double w = 60.0 / (1.0 - radius);
    uint32_t W = w > 16000.0 ? 16000u : (uint32_t)w + 1u;
    return (W + TILE - 1) / TILE * TILE;
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
