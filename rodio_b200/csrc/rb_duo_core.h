// rb_duo_core.h — the lane-PAIR fused kernel: k_fused_lanes' warp program (rb_lanes_core.h) with TWO mono streams per lane
// and every floating-point step issued once for both as a packed f32x2 instruction (sm_100: FADD2 / FMUL2 / FFMA2).
//
// Why.  ncu on k_fused_lanes at 65 536 streams (profiles/r2_ncu_k_fused_lanes_65536.json): 30 warp instructions per sample,
// 81 % issue utilisation on average and 94 % on the busiest sub-partition -- the kernel is bound by issue slots, not by HBM
// (0.46 of the roofline) and not by the FMA pipe (45 %).  17 of the 30 are rounded float operations that the reference's
// arithmetic dictates one by one (src/math.rs:24-26, src/source/blt.rs:558-560): they cannot be removed, but two streams
// can share the instruction.  A packed instruction occupies the FMA pipe for two cycles and ONE issue slot
// (tools/microbench/fp32_throughput.cu), so per pair of samples the lane spends
//     interpolation 6 + feed-forward 3 + recurrence 4 + gain 1  packed,  1 scalar add of the two halves,
//     index step 7 (one numerator for both: the two streams of a lane are in phase, see below),  mixer tree 4,  ring 1.5
// = 28.6 slots per TWO samples against 30 per sample.
//
// What a lane carries: rows 64*g + 2*l and 64*g + 2*l + 1 of the class (insertion order), the even one in the low halves.
// The two share the numerator and the ring cursor, so they must be IN PHASE whenever both are inside a fast run:
// (o0 - mix_start) of the two rows congruent modulo 4 * to (same numerator, same offset of the left frame inside its
// 16-byte quad).  The host checks that for every lane (rb_lanes_plan.h duo_compatible) and otherwise keeps the class on
// k_fused_lanes; batches whose streams start together (every bench configuration, the time-parallel plan) qualify.
// A half that is idle (not started, finished, or no row at all) walks a ring of zeros like an idle lane of k_fused_lanes
// and contributes +0.
//
// Mixer sum: lane value = even row + odd row, then the fixed 32-lane tree of k_fused_lanes (reduce_tile): one partial row
// per group of 64 streams.  Same tolerance class as the other fused kernels (<= 1e-5 * peak against the sequential sum of
// src/mixer.rs:185-198), bit-exact against the oracle's streams added in this documented order (tests).
//
// Group spans (Args::spans, optional): group g adds into partial row `slot` and only stores tiles at or behind `store_lo`
// -- the time-parallel plan (rb_lanes_batch.cu) runs segment k of every stream from a warm-up point in front of the
// segment with zero filter state; what it computes in front of `store_lo` is never stored, and groups of different
// segments share partial rows because their stored ranges are disjoint.
//
// Shapes served: mono f32 sources below the mixer's rate, whole streams (no streaming state), [biquad] [one gain] -- the
// bench chain.  Everything else stays on k_fused_lanes.  Compiled twice like rb_lanes_core.h (nvcc / CPU emulator).
#pragma once
#include <cstdint>

#include "rb_lanes_core.h"

namespace duo {

using lanes::Args;
using lanes::Row;
using lanes::GroupSpan;
using lanes::TILE;
using lanes::RUN_CAP;
using lanes::MIN_RUN_TILES;

// Ring slots per stream.  4 (default): chunk c-1 (draining), c, c+1, c+2 (in flight) -- two chunks of look-ahead.  3: ONE chunk
// of look-ahead and 68 instead of 84 words per stream (12 instead of 10 warps per SM).  Measured: 3 slots are slower at every
// batch size (0.526 against 0.544 of the roofline at 65 536 streams, 0.41 against 0.44 on the time-parallel plan): one chunk
// of look-ahead does not cover the HBM latency under load.
//
// Bank conflicts.  Every lane reads its tap at the same ring offset (the streams of a warp are in phase) and the rings start on
// 16-byte boundaries (cp.async), so the 32 addresses of a tap load share their residue modulo 4 words: only 8 of the 32 banks can
// be hit -- a four-way conflict whatever the stride or rotation of the rings (a rotation by l / 8 quads was tried: the same banks,
// more instructions, 4-10 % slower).  64 of the 93 shared-memory wavefronts of a tile are these loads.
#ifndef RB_DUO_SLOTS
#define RB_DUO_SLOTS 4
#endif
#ifndef RB_DUO_SCALAR_B
#define RB_DUO_SCALAR_B 0   // A/B knobs of the fast run, see the pipeline below
#endif
#ifndef RB_DUO_ORDER
#define RB_DUO_ORDER 0
#endif
static_assert(RB_DUO_SLOTS == 3 || RB_DUO_SLOTS == 4, "3 (one chunk ahead) or 4 (two chunks ahead)");
constexpr int NSLOT = RB_DUO_SLOTS;
constexpr bool ONE_AHEAD = RB_DUO_SLOTS == 3;
constexpr int CHW = lanes::CHUNK;           // words (= mono frames) per chunk
constexpr int CHF = CHW;
constexpr int RING = CHW * NSLOT;
constexpr int MIRROR = CHW;
constexpr int RS = RING + MIRROR + 4;       // 84 words per stream
constexpr int QPC = CHW / 4;                // 16-byte quads per chunk and stream
constexpr int SPI = 32 / QPC;               // streams served by one cp.async warp instruction (8)
constexpr int NCOPY = 64 / SPI;             // cp.async instructions per chunk of the 64 streams
constexpr int HALF_WORDS = 32 * RS;         // ring of the odd row lies this many words behind the ring of the even row
constexpr int WARP_WORDS = 64 * RS;         // 17.0 KB per warp (3 slots)
constexpr int TF = TILE;                    // frames per tile (mono)
static_assert((RS / 4) % 2 == 1 && MIRROR >= TILE && CHF >= 2 * TF, "ring geometry");

// ROLE 0: one warp does everything.  ROLE 1 / 2: the group is served by a PAIR of warps of one CTA (k_fused_duo_split):
// warp A (1) owns the rings and produces the interpolated samples of a tile, warp B (2) takes them from `handoff`
// ([2][TILE][32] packed pairs in shared memory, one CTA barrier per tile) and runs the filter, the gain and the mixer tree.
// Both warps walk the same sequence of runs (each derives it from the rows), B alone computes the slow tiles.  Twice the
// warps per stream, half the instructions per warp: the single-warp form is bound by the latency of its warps
// (profiles/README.md: 0.33 instructions per clock and warp at 1.7 warps per sub-partition).
template <bool HASB, bool FF2, int NPOST, int ROLE = 0>
SIMT_FN void warp_main(const Args& a, uint32_t group, float* ring_warp, simt::f2* handoff = nullptr) {
    using simt::f2;
    constexpr bool DO_A = ROLE != 2, DO_B = ROLE != 1;
    const uint32_t ln = simt::lane();
    Row row[2];
    bool has[2], safe[2], live[2];
    uint64_t ms[2], end[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t r = group * 64u + 2u * ln + (uint32_t)h;
        has[h] = r < a.n_rows;
        if (has[h]) {
            row[h] = a.rows[r];
        } else {
            row[h].in = a.zeros, row[h].L = 0, row[h].out_len = 0, row[h].mix_start = 0, row[h].n_int = 0, row[h].o0 = 0, row[h].i0 = 0, row[h].state = nullptr;
            row[h].b0 = row[h].b1 = row[h].b2 = row[h].a1 = row[h].a2 = row[h].ffk = row[h].post = row[h].pre = row[h].mid = row[h].ga = row[h].gb = 0.0f;
            row[h].flags = 0, row[h].f0 = 0;
        }
        ms[h] = row[h].mix_start, end[h] = row[h].mix_start + row[h].out_len;
        safe[h] = has[h] && !(row[h].flags & (lanes::ROW_UNSAFE | lanes::ROW_FORCE_SLOW)) && !(a.unsafe && a.unsafe[r]);
        live[h] = has[h] && row[h].out_len != 0;
    }
    const uint64_t lo_l = (live[0] ? ms[0] : ~0ull) < (live[1] ? ms[1] : ~0ull) ? (live[0] ? ms[0] : ~0ull) : (live[1] ? ms[1] : ~0ull);
    const uint64_t hi_l = (live[0] ? end[0] : 0ull) > (live[1] ? end[1] : 0ull) ? (live[0] ? end[0] : 0ull) : (live[1] ? end[1] : 0ull);
    const uint64_t t_lo = simt::reduce_min64(lo_l), t_hi = simt::reduce_max64(hi_l);
    if (t_lo >= t_hi) return;
    uint64_t st_lo = 0;
    uint32_t slot = group;
    if (a.spans) st_lo = a.spans[group].store_lo, slot = a.spans[group].slot;
    const uint64_t t_end = (t_hi + TF - 1) / TF * TF;
    uint64_t t = t_lo / TF * TF;
    float* const prow = a.partial + (uint64_t)slot * a.pstride;
    const uint32_t from = a.from, to = a.to;
    const f2 NDEN = simt::pack2(-a.den_f, -a.den_f), RCP = simt::pack2(a.rcp_den, a.rcp_den);
    const f2 NEG1 = simt::pack2(a.neg1, a.neg1);
    const float from_f = a.from_f, den = a.den_f;
    const f2 B0 = simt::pack2(row[0].b0, row[1].b0), B1 = simt::pack2(row[0].b1, row[1].b1), B2 = simt::pack2(row[0].b2, row[1].b2);
    const f2 A1 = simt::pack2(row[0].a1, row[1].a1), A2 = simt::pack2(row[0].a2, row[1].a2);
    const f2 FFK = simt::pack2(row[0].ffk, row[1].ffk), POST = simt::pack2(row[0].post, row[1].post);
    const f2 ONE = simt::pack2(-a.neg1, -a.neg1);   // +1.0 as a run-time value, see rb_simt.h on contraction
    // canonical filter state per half: x[n-1], x[n-2], y[n-1], y[n-2] (zero at the start of a row: whole streams only)
    float xh1[2] = {0.f, 0.f}, xh2[2] = {0.f, 0.f}, y1[2] = {0.f, 0.f}, y2[2] = {0.f, 0.f};
    const uint32_t cq = ln % QPC;                      // the quad of a chunk this lane copies ...
    const uint32_t cr = ln / QPC;                      // ... for stream cr + SPI * j of the group's 64, j < NCOPY:
    const uint32_t chalf = cr & 1u;                    //     half cr & 1 of lane (cr >> 1) + 4 j

    while (t < t_end) {
        // ---- how long does every half stay "interior or idle" from t on? ----
        uint32_t d = RUN_CAP;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t dh;
            if (!has[h] || t >= end[h]) {
                dh = RUN_CAP;
                xh1[h] = xh2[h] = y1[h] = y2[h] = 0.f;   // a finished stream contributes nothing, its filter stops
            } else if (t < ms[h]) {
                const uint64_t g = (ms[h] - t) / TF * TF;
                dh = g > RUN_CAP ? RUN_CAP : (uint32_t)g;
            } else {
                const uint64_t o = t - ms[h];
                dh = 0;
                if (safe[h] && o + TF <= row[h].n_int) {
                    const uint64_t g = (row[h].n_int - o) / TF * TF;
                    dh = g > RUN_CAP ? RUN_CAP : (uint32_t)g;
                }
            }
            d = dh < d ? dh : d;
        }
        const uint64_t left = t_end - t;
        const uint32_t cap = left > RUN_CAP ? RUN_CAP : (uint32_t)left;
        const uint32_t run = simt::reduce_min(d < cap ? d : cap);

        if (run >= (uint32_t)(MIN_RUN_TILES * TF)) {
            // =================================== FAST RUN ===================================
            uint32_t num = 0, k0 = 0;
            bool any = false;
            const float* src[2] = {a.zeros, a.zeros};
            uint32_t maxq[2] = {QPC - 1, QPC - 1};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (has[h] && t >= ms[h] && t < end[h]) {
                    const uint64_t prod = (row[h].o0 + (t - ms[h])) * (uint64_t)from;
                    const uint64_t ia = prod / to;
                    const uint64_t i = ia - row[h].i0;
                    const uint64_t ibase = i & ~3ull;
                    // the same for both halves when both are active (the host's duo_compatible)
                    simt::emu_assert(!any || (num == (uint32_t)(prod - ia * to) && k0 == (uint32_t)(i - ibase)), "the two rows of a lane are out of phase");
                    num = (uint32_t)(prod - ia * to);
                    k0 = (uint32_t)(i - ibase);
                    any = true;
                    src[h] = row[h].in + ibase;
                    const uint64_t mq = ((row[h].L - 1) >> 2) - (ibase >> 2);
                    maxq[h] = mq > 0x7fffffffull ? 0x7fffffffu : (uint32_t)mq;
                }
            }
            uint64_t sq[NCOPY];
            uint32_t mqc[NCOPY];
#pragma unroll
            for (int j = 0; j < NCOPY; j++) {
                const uint32_t sl = (cr >> 1) + 4u * (uint32_t)j;
                const uint64_t s0 = simt::shfl_idx64((uint64_t)(uintptr_t)src[0], sl), s1 = simt::shfl_idx64((uint64_t)(uintptr_t)src[1], sl);
                const uint32_t m0 = simt::shfl_idx(maxq[0], sl), m1 = simt::shfl_idx(maxq[1], sl);
                sq[j] = chalf ? s1 : s0, mqc[j] = chalf ? m1 : m0;
            }
            auto issue = [&](uint32_t c) {
                const uint32_t sl = c % NSLOT;
                const uint32_t want = c * QPC + cq;
#pragma unroll
                for (int j = 0; j < NCOPY; j++) {
                    const uint32_t off = want < mqc[j] ? want : mqc[j];
                    const float* s = (const float*)(uintptr_t)sq[j] + 4ull * off;
                    float* dst = ring_warp + chalf * HALF_WORDS + ((cr >> 1) + 4u * (uint32_t)j) * RS + sl * CHW + cq * 4;
                    simt::cp16(dst, s);
                    if (sl == 0 && (int)(cq * 4) < MIRROR) simt::cp16(dst + RING, s);
                }
                simt::cp_commit();
            };
            if (DO_A) {
                issue(0);
                issue(1);
                simt::cp_wait<1>();
                simt::syncwarp();
                if (!ONE_AHEAD) issue(2);
            }
            uint32_t c_ready = 1;   // chunks [0, c_ready) are readable; c_ready (and, two ahead, c_ready + 1) are in flight
            float* const ringl = ring_warp + ln * RS;
            const simt::sptr ring_end = simt::sptr_of(ringl + RING);
            simt::sptr p = simt::sptr_of(ringl + k0);
            f2 X0 = simt::pack2(0.f, 0.f), X1 = X0;
            if (DO_A) {
                X0 = simt::pack2(simt::lds(p), simt::lds(simt::sptr_add(p, HALF_WORDS)));
                X1 = simt::pack2(simt::lds(simt::sptr_add(p, 1)), simt::lds(simt::sptr_add(p, HALF_WORDS + 1)));
            }
            p = simt::sptr_add(p, 2);
            float nf = simt::u2f(num);
            uint32_t kb = 5, kbn = to - 1;
            f2 XH1 = simt::pack2(xh1[0], xh1[1]), XH2 = simt::pack2(xh2[0], xh2[1]);
            f2 Y1 = simt::pack2(y1[0], y1[1]), Y2 = simt::pack2(y2[0], y2[1]);
            f2 P1 = XH1, P2 = XH2;
            if (HASB && FF2) P1 = simt::mul2(B0, XH1), P2 = simt::mul2(B0, XH2);

            // The run is a software pipeline of three stages, one tile apart, so that a warp always has three independent
            // instruction streams to issue from (measured on the first version, one stage after the other: 0.3 instructions per
            // clock for a lone warp -- the tap loads, the interpolation, the recurrence and the five shuffle levels of the mixer
            // sum each waited for the one before):
            //   A(k + 1)  ring bookkeeping, index steps, tap loads, interpolation, feed-forward half  -> T' (tt per step)
            //   B(k)      the recurrence on T, gain, the two halves added                               -> v
            //   C(k - 1)  the transposing tree over the 32 lanes (lanes::reduce_tile, stage by stage), store
            auto refill = [&]() {
                kb += a.q8, kbn += a.r8;
                if (kbn >= to) kbn -= to, kb += 1;
                if ((kb - 1) / CHF >= c_ready) {
                    if (ONE_AHEAD) {
                        simt::cp_wait<0>();      // chunk c_ready has landed
                        simt::syncwarp();        // ... for every lane, and nobody reads chunk c_ready - 2 any more
                        issue(c_ready + 1);      // into the slot of chunk c_ready - 2
                    } else {
                        simt::cp_wait<1>();      // chunk c_ready has landed (c_ready + 1 may still be in flight)
                        simt::syncwarp();        // ... for every lane, and nobody reads chunk c_ready - 2 any more
                        issue(c_ready + 2);
                    }
                    c_ready += 1;
                    simt::emu_count(2, 1);
                }
                if (simt::sptr_ge(p, ring_end)) p = simt::sptr_add(p, -RING);
            };
            auto step_a = [&](f2& out) {
                // src/math.rs:24-26: first + (second - first) * num / den, the division as an exact reciprocal step
                const f2 m = simt::mul2(simt::sub2(X1, X0), simt::pack2(nf, nf));
                const f2 q0 = simt::mul2(m, RCP);
                const f2 q = simt::fma2(simt::fma2(q0, NDEN, m), RCP, q0);
                const f2 x = simt::add2(X0, q);
                simt::lerp_advance2<HALF_WORDS>(nf, X0, X1, p, from_f, den);
                out = x;
                if (HASB && ROLE == 0) {
                    if (FF2) {
                        const f2 pz = simt::mul2(B0, x);
                        out = simt::fma2(P2, ONE, simt::fma2(P1, FFK, pz));   // (.. ) + p2 with one rounding; p2 stays a multiplicand (rb_simt.h)
                        P2 = P1, P1 = pz;
                    } else {
                        // (b0*x + b1*x1) + b2*x2: every product rounded, then added (fma(p, 1, s) = p + s, one rounding)
                        const f2 s01 = simt::fma2(simt::mul2(B1, XH1), ONE, simt::mul2(B0, x));
                        out = simt::fma2(simt::mul2(B2, XH2), ONE, s01);
                    }
                    XH2 = XH1, XH1 = x;
                }
            };
            auto step_b = [&](f2 tt, float& v) {
                if (HASB && ROLE == 2) {   // the feed-forward half runs here when the warps are split: tt arrives as x
                    const f2 x = tt;
                    if (FF2) {
                        const f2 pz = simt::mul2(B0, x);
                        tt = simt::fma2(P2, ONE, simt::fma2(P1, FFK, pz));
                        P2 = P1, P1 = pz;
                    } else {
                        const f2 s01 = simt::fma2(simt::mul2(B1, XH1), ONE, simt::mul2(B0, x));
                        tt = simt::fma2(simt::mul2(B2, XH2), ONE, s01);
                    }
                    XH2 = XH1, XH1 = x;
                }
                f2 y = tt;
#if RB_DUO_SCALAR_B
                if (HASB) {   // A/B knob: the recurrence as two scalar chains (4-cycle dependent issue, 1-cycle pipe) instead of one packed one
                    const float yl = lanes::fb(simt::lo2(A1), simt::lo2(A2), simt::lo2(tt), simt::lo2(Y1), simt::lo2(Y2), simt::lo2(NEG1));
                    const float yh = lanes::fb(simt::hi2(A1), simt::hi2(A2), simt::hi2(tt), simt::hi2(Y1), simt::hi2(Y2), simt::hi2(NEG1));
                    y = simt::pack2(yl, yh);
                    Y2 = Y1, Y1 = y;
                }
#else
                if (HASB) {
                    // (t - a1*y1) - a2*y2, src/source/blt.rs:558-560
                    y = simt::fma2(simt::mul2(A2, Y2), NEG1, simt::fma2(simt::mul2(A1, Y1), NEG1, tt));
                    Y2 = Y1, Y1 = y;
                }
#endif
                const f2 val = NPOST ? simt::mul2(y, POST) : y;
                v = simt::fadd(simt::lo2(val), simt::hi2(val));
            };
            // lanes::reduce_tile in five stages (the same operations in the same order)
            const bool rb4 = ln & 16u, rb3 = ln & 8u, rb2 = ln & 4u;
            float rw[4], rz[2], rs = 0.f;
            auto stage_c = [&](int stage, const float (&vv)[TILE], uint64_t pos) {
                if (stage == 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++) rw[j] = simt::fadd(rb4 ? vv[j + 4] : vv[j], simt::shfl_xor(rb4 ? vv[j] : vv[j + 4], 16));
                } else if (stage == 1) {
#pragma unroll
                    for (int j = 0; j < 2; j++) rz[j] = simt::fadd(rb3 ? rw[j + 2] : rw[j], simt::shfl_xor(rb3 ? rw[j] : rw[j + 2], 8));
                } else if (stage == 2) {
                    rs = simt::fadd(rb2 ? rz[1] : rz[0], simt::shfl_xor(rb2 ? rz[0] : rz[1], 4));
                } else if (stage == 3) {
                    rs = simt::fadd(rs, simt::shfl_xor(rs, 1));
                } else if (stage == 4) {
                    rs = simt::fadd(rs, simt::shfl_xor(rs, 2));
                    rs = simt::fadd(rs, 0.0f);
                } else if (stage == 5) {
                    if (pos >= st_lo && (ln & 3u) == 0) prow[pos + (ln >> 2)] = rs;
                }
            };
            auto all_c = [&](const float (&vv)[TILE], uint64_t pos) {
#pragma unroll
                for (int st = 0; st < 6; st++) stage_c(st, vv, pos);
            };
            const uint32_t n_tiles = run / TF;     // >= MIN_RUN_TILES
            if (ROLE == 1) {
                // ---- warp A of a split pair: produce tile k into handoff[k & 1], one barrier per tile, one at the end of the run
                // (B may still be reading the last tile when the next run starts to write) ----
                for (uint32_t k = 0; k < n_tiles; k++) {
                    f2 X[TILE];
                    refill();
#pragma unroll
                    for (int f = 0; f < TF; f++) step_a(X[f]);
                    f2* dst = handoff + (k & 1u) * (TILE * 32) + ln;
#pragma unroll
                    for (int f = 0; f < TF; f++) simt::sts2(dst + f * 32, X[f]);
                    simt::cta_sync();
                }
                simt::cta_sync();
            } else if (ROLE == 2) {
                // ---- warp B: consume tile j (feed-forward, recurrence, gain), the mixer tree of tile j - 1 interleaved ----
                float v[TILE], vp[TILE];
                for (uint32_t j = 0; j < n_tiles; j++) {
                    simt::cta_sync();      // tile j has been written
                    f2 X[TILE];
                    const f2* srcx = handoff + (j & 1u) * (TILE * 32) + ln;
#pragma unroll
                    for (int f = 0; f < TF; f++) X[f] = simt::lds2(srcx + f * 32);
                    const uint64_t pos = t + (uint64_t)(j - 1) * TF;
                    if (j == 0) {
#pragma unroll
                        for (int f = 0; f < TF; f++) step_b(X[f], vp[f]);
                    } else {
#pragma unroll
                        for (int f = 0; f < TF; f++) {
                            step_b(X[f], v[f]);
                            stage_c(f == 0 ? 0 : f == 2 ? 1 : f == 4 ? 2 : f == 5 ? 3 : f == 6 ? 4 : f == 7 ? 5 : -1, vp, pos);
                        }
#pragma unroll
                        for (int f = 0; f < TF; f++) vp[f] = v[f];
                    }
                }
                all_c(vp, t + (uint64_t)(n_tiles - 1) * TF);
                simt::cta_sync();
            } else {
            f2 T[TILE], Tn[TILE];
            float v[TILE], vp[TILE];
            refill();
#pragma unroll
            for (int f = 0; f < TF; f++) step_a(T[f]);                       // A(0)
            refill();
#pragma unroll
            for (int f = 0; f < TF; f++) step_a(Tn[f]), step_b(T[f], vp[f]);   // A(1), B(0)
#pragma unroll
            for (int f = 0; f < TF; f++) T[f] = Tn[f];
            // steady state, two tiles per trip so that the tile buffers alternate instead of being copied
            auto body = [&](const f2 (&tin)[TILE], f2 (&tout)[TILE], const float (&vred)[TILE], float (&vout)[TILE], uint64_t pos) {
                refill();
#pragma unroll
                for (int f = 0; f < TF; f++) {
#if RB_DUO_ORDER == 1
                    step_b(tin[f], vout[f]);
                    step_a(tout[f]);
#else
                    step_a(tout[f]);
                    step_b(tin[f], vout[f]);
#endif
                    stage_c(f == 0 ? 0 : f == 2 ? 1 : f == 4 ? 2 : f == 5 ? 3 : f == 6 ? 4 : f == 7 ? 5 : -1, vred, pos);
                }
            };
            uint32_t k = 2;                                                  // A(k), B(k - 1), C(k - 2)
            for (; k + 1 < n_tiles; k += 2) {
                body(T, Tn, vp, v, t + (uint64_t)(k - 2) * TF);
                body(Tn, T, v, vp, t + (uint64_t)(k - 1) * TF);
            }
            if (k < n_tiles) {
                body(T, Tn, vp, v, t + (uint64_t)(k - 2) * TF);
#pragma unroll
                for (int f = 0; f < TF; f++) T[f] = Tn[f], vp[f] = v[f];
            }
#pragma unroll
            for (int f = 0; f < TF; f++) step_b(T[f], v[f]);                 // B(n - 1)
            all_c(vp, t + (uint64_t)(n_tiles - 2) * TF);                     // C(n - 2)
            all_c(v, t + (uint64_t)(n_tiles - 1) * TF);                      // C(n - 1)
            }
            if (DO_A) {
                simt::cp_wait<0>();
                simt::syncwarp();
            }
            simt::emu_count(0, run / TF);
            xh1[0] = simt::lo2(XH1), xh1[1] = simt::hi2(XH1), xh2[0] = simt::lo2(XH2), xh2[1] = simt::hi2(XH2);
            y1[0] = simt::lo2(Y1), y1[1] = simt::hi2(Y1), y2[0] = simt::lo2(Y2), y2[1] = simt::hi2(Y2);
            t += run;
        } else {
            // =================================== SLOW TILE: per-half closed form ===================================
            if (!DO_B) {     // warp A of a split pair has no part in it
                t += TF;
                continue;
            }
            float v[TILE];
#pragma unroll 1
            for (int f = 0; f < TF; f++) {
                const uint64_t tt = t + (uint64_t)f;
                float val[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    val[h] = 0.0f;
                    const bool on = has[h] && tt >= ms[h] && tt < end[h];
                    if (on) {
                        const uint64_t prod = (row[h].o0 + (tt - ms[h])) * (uint64_t)from;
                        const uint64_t ia = prod / to;
                        const uint32_t num = (uint32_t)(prod - ia * to);
                        const uint64_t i = ia - row[h].i0;
                        const float xa = simt::ldg(row[h].in + i);
                        float x = xa;
                        if (i + 1 < row[h].L) {
                            const float xb = simt::ldg(row[h].in + i + 1);
                            x = simt::fadd(xa, simt::fdiv(simt::fmul(simt::fsub(xb, xa), simt::u2f(num)), a.den_f));
                        }
                        float y = x;
                        if (HASB) {
                            const float ff = simt::fadd(simt::fadd(simt::fmul(row[h].b0, x), simt::fmul(row[h].b1, xh1[h])), simt::fmul(row[h].b2, xh2[h]));
                            y = lanes::fb(row[h].a1, row[h].a2, ff, y1[h], y2[h], a.neg1);
                            xh2[h] = xh1[h], xh1[h] = x, y2[h] = y1[h], y1[h] = y;
                        }
                        val[h] = NPOST ? simt::fmul(y, row[h].post) : y;
                    } else if (has[h] && tt >= end[h]) {
                        xh1[h] = xh2[h] = y1[h] = y2[h] = 0.f;
                    }
                }
                v[f] = simt::fadd(val[0], val[1]);
            }
            if (t >= st_lo) {
                const float s = lanes::reduce_tile(v, ln);
                if ((ln & 3u) == 0) prow[t + (ln >> 2)] = s;
            }
            simt::emu_count(1, 1);
            t += TF;
        }
    }
}

}  // namespace duo
