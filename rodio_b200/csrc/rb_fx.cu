// rb_fx.cu — k_fused_fx: the effect chain of BASELINE cfg4 in ONE launch, no intermediates in HBM:
//     [Spatial / ChannelVolume] -> [reverb] -> [automatic_gain_control] -> [limit] -> mixer sum      (at least one of AGC / limiter)
// for f32 sources that already have the mixer's rate and channel count (references:
// src/source/channel_volume.rs:71-88, src/source/spatial.rs:48-69 (the two volumes are computed on the host),
// src/source/mod.rs:628-634 + mix.rs:43-53 + delay.rs:68-75 (reverb = x + delayed, amplified x), src/source/agc.rs:397-504,
// src/source/limit.rs:854-988).
//
// Where the time goes.  The AGC is three recurrences over the INTERLEAVED sample sequence of a stream (agc.rs keeps one
// state for all channels): the peak follower, the running sum of the last 8192 squares, and the gain smoother with its clamp.
// Each is a dependent chain of 2-5 rounded operations per sample, i.e. 9-22 cycles per sample whatever the batch, and a
// stream of one second of stereo audio with a 50 ms echo is 100 800 such steps: 1.1 ms for the longest chain.  Everything
// else (channel volumes, echo, |x|, squares, sqrt and the three IEEE divisions of the desired gain, the final product, the
// mixer sum) is parallel in time.  The general path ran this as seven launches with f32 intermediates in HBM (4.33 ms for
// 512 stereo streams x 1 s, the two chain passes 1.7 and 1.9 ms one after the other).  Here the three chains run CONCURRENTLY
// as warp roles of one CTA, a tile apart, fed and drained through shared memory by worker warps:
//
//     tile it      workers   front:  x (+ the echo tap x[n-D], and both again 8192 samples earlier for the square that
//                                    leaves the RMS window) -> e, |e|, e^2, old e^2                  [global -> shared]
//     tile it-1    warp P    peak[n]  = |e| > peak ? |e| : peak*release + |e|*(1-release)             (3 dependent operations)
//                  warp S    sum[n]   = (sum - old) + new                                              (2)
//     tile it-2    workers   desired  = max(min(target/sqrt(sum/8192), min(target/peak, max_gain)), floor)   (parallel)
//     tile it-3    warp G    gain[n]  = clamp(gain*k + desired*(1-k), 0.1, max_gain), k = desired > gain ? attack : release
//                                       both candidates computed, one selected: FMUL, FADD, SEL, FMNMX, FMNMX on the chain
//     tile it-4    workers   y = e * gain, summed over the CTA's streams in insertion order -> one partial row per CTA
//   with a limiter behind (or instead of) the AGC the last stage splits in three:
//     tile it-4    workers   y, and the gain computer's dB for it (log2: within 2 ulp of glibc's, the limiter's 1e-5 * peak class)
//     tile it-5    warp L    per channel: integrator = max(dB, release-smoothed), peak = attack-smoothed; max over the channels
//     tile it-6    workers   y * 2^(-max_peak * 0.05 * log2(10)), summed
// (eight worker warps, two per stream; the gain chain has a sub-partition to itself: sharing one with the other two chains cost it a
// third of its pace -- 19 instructions per sample from three warps that each want a slot every 4.3 cycles.)
//
// A CTA owns FX_R = 4 consecutive streams (lane = stream in the chain warps), so that 512 streams spread over 128 SMs: the
// chains are latency-bound and gain nothing from sharing an SM, the parallel stages need the SMs.  The 8192-entry ring of
// squares is not stored: the square that leaves the window is recomputed from the input, bit-identical (as k_agc_pipe does).
// Mixer sum: sequential over the 4 streams of a CTA from +0.0, the partial rows added in CTA order (rb_batch_mix_group = 4):
// the tolerance class of the other fused kernels; per-stream samples are bit-exact.  RB_MIX_EXACT_ORDER: k_fused_fx<C, true> hands the
// running sum of every tile from CTA to CTA (one sequential sum over all streams: the reference's, bit for bit; streams that join late
// keep the general path).
#include <algorithm>
#include <cstring>
#include <vector>

#include "rb_dsp.cuh"
#include "rb_fused.h"

using namespace rbd;

namespace {

constexpr int FX_R = 4;                    // streams per CTA
constexpr int FX_T = 256;                  // samples per stream and tile
constexpr int FX_TS = FX_T + 4;            // padded row
constexpr int FX_SLOTS = 7;                // tiles in flight: front | P,S | desired | G | out, or (with a limiter) y, gain computer | L | out
constexpr int FX_ARR = FX_R * FX_TS;       // one array of a tile
constexpr int FX_SLOT = 4 * FX_ARR;        // e | v -> peak | sq -> sum -> desired -> gain | old sq
constexpr size_t FX_SMEM = (size_t)FX_SLOTS * FX_SLOT * sizeof(float);
constexpr int FX_THREADS = 15 * 32;        // sub-partition 3: warp 3 = the gain chain, alone (its 5 dependent operations per sample set the pace);
                                           // sub-partition 2: warps 2, 6 = peak and sum chains beside two workers; 0 and 1: three workers each
constexpr uint32_t RMS_WINDOW = 8192;

struct FxRow {
    const float* in;
    uint64_t n_in;        // samples
    uint64_t n_out;       // samples (n_in + echo delay)
    uint64_t mix_start;   // samples
    uint64_t delay;       // echo delay in samples (has_echo)
    float amp;            // echo amplitude
    float vol[2];         // channel volumes (has_cv)
    float target, max_gain, floor, attack, release;                 // AGC (has_agc)
    float l_thr, l_knee, l_ik8, l_attack, l_release;                // limiter (has_lim), limit.rs:94-130
};

struct FxArgs {
    const FxRow* rows;
    uint32_t n_rows;
    uint32_t channels;    // 1 or 2 (source == mixer)
    uint32_t has_cv, has_echo, has_agc, has_lim;
    float* partial;       // [n_ctas][mix_len]  (chain: [n_ctas][mix_len] of (value, tag) pairs)
    uint64_t mix_len;
    // RB_MIX_EXACT_ORDER (k_fused_fx<C, true>): the running sum of every tile is handed from CTA to CTA like in k_fused_hot
    // (rb_fused.cu, "Chain mode"): tickets order the chain, (value, tag) pairs are the hand-over, the last CTA writes `out`
    float* out;
    uint32_t* ticket;
    uint32_t epoch, pad_;
};

__device__ __forceinline__ uint2 fx_ld_pair(const uint2* p) {      // one 64-bit relaxed access: single-copy atomic
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}
__device__ __forceinline__ void fx_st_pair(uint2* p, float v, uint32_t tag) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}

// e[n] of one stream: channel volume, then the echo (mix.rs:47-52: both / whichever exists)
template <int C>
__device__ __forceinline__ float fx_cv(const FxRow& r, bool has_cv, uint64_t m) {
    // ChannelVolume over C input channels (channel_volume.rs:75-88): mono = ((0 + s0) + s1) / C, out[ch] = mono * vol[ch]
    if (!has_cv) return __ldg(r.in + m);
    // (x / 1 == x and x / 2 == x * 0.5 for every x, subnormal results included: both are one rounding of the same real number)
    if (C == 1) return mul(add(0.0f, __ldg(r.in + m)), r.vol[0]);
    const float2 fr = __ldg(reinterpret_cast<const float2*>(r.in) + (m >> 1));
    return mul(mul(add(add(0.0f, fr.x), fr.y), 0.5f), (m & 1) ? r.vol[1] : r.vol[0]);
}
template <int C, bool INTERIOR = false>   // INTERIOR: the sample and its echo tap both exist (no range checks)
__device__ __forceinline__ float fx_e(const FxRow& r, bool has_cv, bool has_echo, uint64_t n) {
    if (!has_echo) return fx_cv<C>(r, has_cv, n);
    if (INTERIOR) return add(fx_cv<C>(r, has_cv, n), mul(fx_cv<C>(r, has_cv, n - r.delay), r.amp));
    const float s2 = n < r.delay ? 0.0f : mul(fx_cv<C>(r, has_cv, n - r.delay), r.amp);   // Delay emits literal 0.0 first
    return n < r.n_in ? add(fx_cv<C>(r, has_cv, n), s2) : s2;
}

template <int C, bool CHAIN>
__global__ void __launch_bounds__(FX_THREADS) k_fused_fx(FxArgs a) {
    extern __shared__ __align__(16) float fx_sm[];   // [FX_SLOTS][4][FX_R][FX_TS]
    __shared__ FxRow s_rows[FX_R];
    __shared__ uint64_t s_max_n;
    __shared__ uint32_t s_cta;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (CHAIN) {   // the CTA's place in the chain is the order in which the CTAs start
        if (threadIdx.x == 0) s_cta = atomicAdd(a.ticket, 1u) - a.epoch * gridDim.x;
        __syncthreads();
    }
    const uint32_t cta = CHAIN ? s_cta : blockIdx.x;
    const uint32_t s0 = cta * FX_R;
    const uint32_t cnt = min((uint32_t)FX_R, a.n_rows - s0);
    if (threadIdx.x < FX_R) {
        FxRow r;
        if (threadIdx.x < cnt) r = a.rows[s0 + threadIdx.x];
        else memset(&r, 0, sizeof(r));
        s_rows[threadIdx.x] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t m = 0;
        for (uint32_t i = 0; i < cnt; i++) m = max(m, s_rows[i].n_out);
        s_max_n = CHAIN ? a.mix_len : m;      // chain: every CTA hands the running sum on for every tile of the timeline (mix_start = 0)
    }
    __syncthreads();
    const uint32_t n_tiles = (uint32_t)((s_max_n + FX_T - 1) / FX_T);
    const bool has_cv = a.has_cv != 0, has_echo = a.has_echo != 0, has_agc = a.has_agc != 0, has_lim = a.has_lim != 0;
    const uint32_t out_lag = has_lim ? 6u : 4u;   // tiles between the front and the mixer sum
    // warps 0,4,8,12 | 1,5,9 | 10 are the eight workers; 2 = peak, 6 = sum, 14 = limiter envelopes (one sub-partition), 3 = gain;
    // 7, 11, 13 only keep the barrier count
    const int worker = warp == 0 ? 0 : warp == 4 ? 1 : warp == 8 ? 2 : warp == 1 ? 3 : warp == 5 ? 4 : warp == 9 ? 5 : warp == 10 ? 6 : warp == 12 ? 7 : -1;
    const bool chain_p = warp == 2, chain_s = warp == 6, chain_g = warp == 3, chain_l = warp == 14;
    const uint64_t mix_start = s_rows[0].mix_start;      // equal for the CTA's streams (planner)
    float* const prow = a.partial + (uint64_t)cta * a.mix_len + mix_start;
    const uint2* carry_row = nullptr;      // chain: the row of (value, tag) pairs the CTA in front wrote
    uint2* tagged_row = nullptr;           // chain: this CTA's row (the last CTA writes a.out instead)
    if (CHAIN) {
        uint2* rows2 = reinterpret_cast<uint2*>(a.partial);
        if (cta + 1 != gridDim.x) tagged_row = rows2 + (uint64_t)cta * a.mix_len;
        if (cta > 0) carry_row = rows2 + (uint64_t)(cta - 1) * a.mix_len;
    }
    const int chain_out = !CHAIN ? -1 : warp == 7 ? 0 : warp == 11 ? 1 : warp == 13 ? 2 : -1;   // the mixer sum moves to the idle warps

    // chain state, lane = stream
    float peak = 0.0f, sum = 0.0f, gain = 1.0f;
    const FxRow& my = s_rows[lane < FX_R ? lane : 0];
    const uint64_t my_n = (lane < cnt) ? my.n_out : 0;
    const float attack = my.attack, release = my.release, max_gain = my.max_gain;
    const float oma = sub(1.0f, attack), omr = sub(1.0f, release);
    const float l_att = my.l_attack, l_rel = my.l_release, l_oma = sub(1.0f, my.l_attack), l_omr = sub(1.0f, my.l_release);
    float l_int[2] = {0.0f, 0.0f}, l_pk[2] = {0.0f, 0.0f};   // limiter: integrator and peak per channel (limit.rs:903-916)

    for (uint32_t it = 0; it < n_tiles + out_lag; it++) {
        if (worker >= 0) {
            // ---- front, tile `it`: stream worker / 2, half tile worker % 2, 4 consecutive samples per lane ----
            const uint32_t ws = (uint32_t)worker >> 1, wo = ((uint32_t)worker & 1u) * (FX_T / 2) + 4 * lane;
            if (it < n_tiles && ws < cnt) {
                const FxRow& r = s_rows[ws];
                float* base = fx_sm + (it % FX_SLOTS) * FX_SLOT + ws * FX_TS + wo;
                const uint64_t n0 = (uint64_t)it * FX_T + wo;
                float e[4], v[4], q[4], o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) e[j] = v[j] = q[j] = o[j] = 0.0f;
                if (n0 < r.n_out) {
                    const bool interior = n0 + 4 <= r.n_in && n0 >= r.delay + RMS_WINDOW;   // every tap of every sample exists
                    if (!has_agc) {
#pragma unroll 1
                        for (int j = 0; j < 4; j++)
                            if (n0 + j < r.n_out) e[j] = fx_e<C>(r, has_cv, has_echo, n0 + j);
                    } else if (interior) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            e[j] = fx_e<C, true>(r, has_cv, has_echo, n0 + j);
                            v[j] = fabsf(e[j]);
                            q[j] = mul(v[j], v[j]);
                            const float w = fabsf(fx_e<C, true>(r, has_cv, has_echo, n0 + j - RMS_WINDOW));
                            o[j] = mul(w, w);
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 4; j++) {
                            const uint64_t n = n0 + j;
                            if (n < r.n_out) {
                                e[j] = fx_e<C>(r, has_cv, has_echo, n);
                                v[j] = fabsf(e[j]);
                                q[j] = mul(v[j], v[j]);
                                if (n >= RMS_WINDOW) {
                                    const float w = fabsf(fx_e<C>(r, has_cv, has_echo, n - RMS_WINDOW));
                                    o[j] = mul(w, w);
                                }
                            }
                        }
                    }
                }
                *reinterpret_cast<float4*>(base) = make_float4(e[0], e[1], e[2], e[3]);
                *reinterpret_cast<float4*>(base + FX_ARR) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(base + 2 * FX_ARR) = make_float4(q[0], q[1], q[2], q[3]);
                *reinterpret_cast<float4*>(base + 3 * FX_ARR) = make_float4(o[0], o[1], o[2], o[3]);
            }
            // ---- desired gain, tile `it - 2` (agc.rs:413-431, :466-470): sum -> desired in place ----
            if (has_agc && it >= 2 && it - 2 < n_tiles && ws < cnt) {
                const FxRow& r = s_rows[ws];
                float* base = fx_sm + ((it - 2) % FX_SLOTS) * FX_SLOT + ws * FX_TS + wo;
                const float4 p0 = *reinterpret_cast<const float4*>(base + FX_ARR);
                const float4 u0 = *reinterpret_cast<const float4*>(base + 2 * FX_ARR);
                const float pk[4] = {p0.x, p0.y, p0.z, p0.w};
                const float sm[4] = {u0.x, u0.y, u0.z, u0.w};
                const float target = r.target, mg = r.max_gain, fl = r.floor;
                float d[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float rms = __fsqrt_rn(mul(sm[j], 1.0f / 8192.0f));     // sum / 8192: a power of two, the product is the quotient
                    const float rms_gain = (rms > 0.0f) ? divf(target, rms) : mg;
                    const float peak_gain = (pk[j] > 0.0f) ? fminf(divf(target, pk[j]), mg) : mg;
                    d[j] = fmaxf(fminf(rms_gain, peak_gain), fl);
                }
                *reinterpret_cast<float4*>(base + 2 * FX_ARR) = make_float4(d[0], d[1], d[2], d[3]);
            }
            // ---- with a limiter, tile `it - 4`: y = e * gain stays in the tile, the gain computer's dB beside it (limit.rs:854-873) ----
            if (has_lim && it >= 4 && it - 4 < n_tiles && ws < cnt) {
                const FxRow& r = s_rows[ws];
                float* base = fx_sm + ((it - 4) % FX_SLOTS) * FX_SLOT + ws * FX_TS + wo;
                float4 ev = *reinterpret_cast<const float4*>(base);
                if (has_agc) {
                    const float4 gv = *reinterpret_cast<const float4*>(base + 2 * FX_ARR);
                    ev.x = mul(ev.x, gv.x), ev.y = mul(ev.y, gv.y), ev.z = mul(ev.z, gv.z), ev.w = mul(ev.w, gv.w);
                    *reinterpret_cast<float4*>(base) = ev;
                }
                *reinterpret_cast<float4*>(base + FX_ARR) = make_float4(limiter_db(ev.x, r.l_thr, r.l_knee, r.l_ik8), limiter_db(ev.y, r.l_thr, r.l_knee, r.l_ik8),
                                                                        limiter_db(ev.z, r.l_thr, r.l_knee, r.l_ik8), limiter_db(ev.w, r.l_thr, r.l_knee, r.l_ik8));
            }
            // ---- out, tile `it - out_lag`: the last factor, summed over the CTA's streams; one position per worker lane ----
            if (!CHAIN && it >= out_lag) {
                const float* base = fx_sm + ((it - out_lag) % FX_SLOTS) * FX_SLOT;
                const uint32_t pos = 32 * (uint32_t)worker + lane;
                const uint64_t n = (uint64_t)(it - out_lag) * FX_T + pos;
                float acc = 0.0f;
                bool any = false;
#pragma unroll
                for (uint32_t s = 0; s < FX_R; s++)
                    if (n < s_rows[s].n_out) {   // absent rows: n_out = 0
                        float y = base[s * FX_TS + pos];
                        if (has_lim) y = mul(y, db_to_linear(-base[FX_ARR + s * FX_TS + pos]));                 // limit.rs:927-988
                        else if (has_agc) y = mul(y, base[2 * FX_ARR + s * FX_TS + pos]);
                        acc = add(acc, y), any = true;
                    }
                if (any && mix_start + n < a.mix_len) prow[n] = acc;
            }
        } else if (CHAIN && chain_out >= 0) {
            // ---- chain mode, out, tile `it - out_lag`: start from the running sum of the CTA in front, add this CTA's streams in
            // insertion order, hand the sum on (three otherwise idle warps, 96 lanes over the 256 positions of a tile) ----
            if (it >= out_lag) {
                const float* base = fx_sm + ((it - out_lag) % FX_SLOTS) * FX_SLOT;
                const uint32_t tag = a.epoch * (n_tiles + out_lag + 1u) + (it - out_lag) + 1u;
                for (uint32_t pos = 32 * (uint32_t)chain_out + lane; pos < (uint32_t)FX_T; pos += 96) {
                    const uint64_t n = (uint64_t)(it - out_lag) * FX_T + pos;
                    if (n >= a.mix_len) continue;
                    float acc = 0.0f;
                    if (carry_row) {
                        uint2 w;
                        uint32_t spins = 0;
                        while ((w = fx_ld_pair(carry_row + n)).y != tag)
                            if (++spins > (1u << 22)) __trap();      // the CTA in front never arrived: fail loudly
                        acc = __uint_as_float(w.x);
                    }
#pragma unroll
                    for (uint32_t s = 0; s < FX_R; s++)
                        if (n < s_rows[s].n_out) {
                            float y = base[s * FX_TS + pos];
                            if (has_lim) y = mul(y, db_to_linear(-base[FX_ARR + s * FX_TS + pos]));
                            else if (has_agc) y = mul(y, base[2 * FX_ARR + s * FX_TS + pos]);
                            acc = add(acc, y);
                        }
                    if (tagged_row) fx_st_pair(tagged_row + n, acc, tag);
                    else a.out[n] = acc;
                }
            }
        } else if (has_agc && it >= 1 && it - 1 < n_tiles && (chain_p || chain_s)) {
            // ---- chains P and S, tile `it - 1`, lane = stream ----
            if (lane < cnt) {
                const uint64_t nb = (uint64_t)(it - 1) * FX_T;
                const int c4 = (int)((min((uint64_t)FX_T, my_n > nb ? my_n - nb : 0) + 3) / 4);
                float* slot = fx_sm + ((it - 1) % FX_SLOTS) * FX_SLOT + lane * FX_TS;
                if (chain_p) {
                    float4* pv = reinterpret_cast<float4*>(slot + FX_ARR);
                    float4 cur = c4 > 0 ? pv[0] : make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int k = 0; k < c4; k++) {
                        const float4 nxt = k + 1 < c4 ? pv[k + 1] : cur;
                        const float av[4] = {cur.x, cur.y, cur.z, cur.w};
                        float r1[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float v = av[j];
                            const float decayed = add(mul(peak, release), mul(v, omr));   // agc.rs:397-408
                            peak = (v > peak) ? v : decayed;
                            r1[j] = peak;
                        }
                        pv[k] = make_float4(r1[0], r1[1], r1[2], r1[3]);
                        cur = nxt;
                    }
                } else {
                    float4* pq = reinterpret_cast<float4*>(slot + 2 * FX_ARR);
                    const float4* po = reinterpret_cast<const float4*>(slot + 3 * FX_ARR);
                    float4 cq = c4 > 0 ? pq[0] : make_float4(0.f, 0.f, 0.f, 0.f), co = c4 > 0 ? po[0] : cq;
                    for (int k = 0; k < c4; k++) {
                        const float4 nq = k + 1 < c4 ? pq[k + 1] : cq, no = k + 1 < c4 ? po[k + 1] : co;
                        float4 o;
                        sum = add(sub(sum, co.x), cq.x), o.x = sum;                          // agc.rs:157
                        sum = add(sub(sum, co.y), cq.y), o.y = sum;
                        sum = add(sub(sum, co.z), cq.z), o.z = sum;
                        sum = add(sub(sum, co.w), cq.w), o.w = sum;
                        pq[k] = o;
                        cq = nq, co = no;
                    }
                }
            }
        } else if (has_agc && chain_g && it >= 3 && it - 3 < n_tiles) {
            // ---- chain G, tile `it - 3` (agc.rs:474-491): desired -> gain in place ----
            if (lane < cnt) {
                const uint64_t nb = (uint64_t)(it - 3) * FX_T;
                const int c4 = (int)((min((uint64_t)FX_T, my_n > nb ? my_n - nb : 0) + 3) / 4);
                float4* pd = reinterpret_cast<float4*>(fx_sm + ((it - 3) % FX_SLOTS) * FX_SLOT + 2 * FX_ARR + lane * FX_TS);
                float4 cur = c4 > 0 ? pd[0] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < c4; k++) {
                    const float4 nxt = k + 1 < c4 ? pd[k + 1] : cur;
                    const float dv[4] = {cur.x, cur.y, cur.z, cur.w};
                    float r1[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float d = dv[j];
                        const float da = mul(d, oma), dr = mul(d, omr);            // off the chain
                        const float ga = add(mul(gain, attack), da), gr = add(mul(gain, release), dr);
                        const float g = d > gain ? ga : gr;
                        gain = fminf(fmaxf(g, 0.1f), max_gain);
                        r1[j] = gain;
                    }
                    pd[k] = make_float4(r1[0], r1[1], r1[2], r1[3]);
                    cur = nxt;
                }
            }
        } else if (has_lim && chain_l && it >= 5 && it - 5 < n_tiles) {
            // ---- chain L, tile `it - 5`: per channel integrator (release) and peak (attack) envelopes, the channels coupled by
            // max(peaks) AFTER the current channel's update (limit.rs:903-916, :946-960): gain-computer dB -> max_peak in place ----
            if (lane < cnt) {
                const uint64_t nb = (uint64_t)(it - 5) * FX_T;
                const int c4 = (int)((min((uint64_t)FX_T, my_n > nb ? my_n - nb : 0) + 3) / 4);
                float4* pd = reinterpret_cast<float4*>(fx_sm + ((it - 5) % FX_SLOTS) * FX_SLOT + FX_ARR + lane * FX_TS);
                float4 cur = c4 > 0 ? pd[0] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < c4; k++) {
                    const float4 nxt = k + 1 < c4 ? pd[k + 1] : cur;
                    const float dv[4] = {cur.x, cur.y, cur.z, cur.w};
                    float r1[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        constexpr int CM = C - 1;
                        const int c = j & CM;                                   // tiles and groups of four start on channel 0
                        const float ldb = dv[j];
                        const float in_c = fmaxf(ldb, add(mul(l_rel, l_int[c]), mul(l_omr, ldb)));
                        l_int[c] = in_c;
                        l_pk[c] = add(mul(l_att, l_pk[c]), mul(l_oma, in_c));
                        r1[j] = C == 1 ? l_pk[0] : fmaxf(l_pk[0], l_pk[1]);
                    }
                    pd[k] = make_float4(r1[0], r1[1], r1[2], r1[3]);
                    cur = nxt;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_fx_sum_partials(const float* __restrict__ partial, uint32_t n_ctas, uint64_t mix_len,
                                                         float* __restrict__ out) {
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < mix_len; m += (uint64_t)gridDim.x * blockDim.x) {
        float acc = partial[m];
        for (uint32_t c = 1; c < n_ctas; c++) acc = add(acc, partial[(uint64_t)c * mix_len + m]);
        out[m] = acc;
    }
}

}  // namespace

struct rb_fx_plan {
    FxArgs args{};
    FxRow* d_rows = nullptr;
    float* d_partial = nullptr;
    float* d_out = nullptr;
    uint32_t n_ctas = 0;
    bool chain = false;          // RB_MIX_EXACT_ORDER: k_fused_fx<C, true>, no k_fx_sum_partials
    uint32_t* d_ticket = nullptr;
    uint32_t epoch = 0;
};

// Shape: every stream f32 with the mixer's channel count (1 or 2) and no conversion, nodes = [CHANVOL]? [ECHO]? [AGC]? [LIMIT]? with
// at least one of the last two, the same
// node kinds for every stream; the streams of a CTA (groups of FX_R in insertion order) share their mix_start.
cudaError_t rb_fx_try_create(const rb_fused_stream* streams, size_t n_streams, uint16_t mixer_channels, float* d_out, uint64_t mix_len,
                             uint32_t flags, cudaStream_t st, rb_fx_plan** out) {
    *out = nullptr;
    if (n_streams == 0 || mix_len == 0 || (mixer_channels != 1 && mixer_channels != 2)) return cudaSuccess;
    const bool exact = (flags & RB_MIX_EXACT_ORDER) != 0;
    std::vector<FxRow> rows(n_streams);
    uint32_t has_cv = 0, has_echo = 0, has_agc = 0, has_lim = 0;
    for (size_t i = 0; i < n_streams; i++) {
        const rb_fused_stream& s = streams[i];
        if (s.fmt != RB_FMT_F32 || s.c_in != mixer_channels || s.n_nodes == 0 || s.n_nodes > 4) return cudaSuccess;
        if (s.n_in % mixer_channels || (reinterpret_cast<uintptr_t>(s.in) & 7u)) return cudaSuccess;
        FxRow& r = rows[i];
        memset(&r, 0, sizeof(r));
        r.in = (const float*)s.in, r.n_in = s.n_in, r.n_out = s.out_len, r.mix_start = s.mix_start;
        uint32_t cv = 0, echo = 0, agc = 0, lim = 0, k = 0;
        auto node = [&](uint32_t j) -> const rb_node_dev& {
            return *reinterpret_cast<const rb_node_dev*>(reinterpret_cast<const char*>(s.nodes) + (size_t)j * s.node_stride);
        };
        if (k < s.n_nodes && node(k).kind == RB_N_CHANVOL) {
            const rb_node_dev& nd = node(k++);
            if (nd.c_in != mixer_channels || nd.c_out != mixer_channels) return cudaSuccess;
            r.vol[0] = nd.p.cv.vol[0], r.vol[1] = mixer_channels == 2 ? nd.p.cv.vol[1] : 0.0f, cv = 1;
        }
        if (k < s.n_nodes && node(k).kind == RB_N_ECHO) {
            const rb_node_dev& nd = node(k++);
            r.delay = nd.p.echo.delay, r.amp = nd.p.echo.amplitude, echo = 1;
        }
        if (k < s.n_nodes && node(k).kind == RB_N_AGC) {
            const rb_node_dev& nd = node(k++);
            r.target = nd.p.agc.target, r.max_gain = nd.p.agc.max_gain, r.floor = nd.p.agc.floor, r.attack = nd.p.agc.attack, r.release = nd.p.agc.release;
            agc = 1;
        }
        if (k < s.n_nodes && node(k).kind == RB_N_LIMIT) {
            const rb_node_dev& nd = node(k++);
            r.l_thr = nd.p.lim.threshold, r.l_knee = nd.p.lim.knee, r.l_ik8 = nd.p.lim.inv_knee_8, r.l_attack = nd.p.lim.attack, r.l_release = nd.p.lim.release;
            lim = 1;
        }
        if (!(agc || lim) || k != s.n_nodes) return cudaSuccess;
        if (r.n_out != r.n_in + (echo ? r.delay : 0)) return cudaSuccess;
        if (i == 0) has_cv = cv, has_echo = echo, has_agc = agc, has_lim = lim;
        else if (cv != has_cv || echo != has_echo || agc != has_agc || lim != has_lim) return cudaSuccess;
        if (i % FX_R && r.mix_start != rows[i - 1].mix_start) return cudaSuccess;
        if (exact && r.mix_start != 0) return cudaSuccess;      // the chain walks one common timeline: late joiners keep the general path
    }
    auto p = new rb_fx_plan;
    p->n_ctas = (uint32_t)((n_streams + FX_R - 1) / FX_R);
    p->d_out = d_out;
    p->chain = exact && p->n_ctas > 1;
    if (exact && !p->chain) {      // one CTA: its partial row is the sequential sum already; keep the plain launch
    }
    const size_t partial_bytes = (size_t)p->n_ctas * mix_len * (p->chain ? sizeof(uint2) : sizeof(float));
    cudaError_t e = cudaMalloc(&p->d_rows, n_streams * sizeof(FxRow));
    if (e == cudaSuccess && p->chain) e = cudaMalloc(&p->d_ticket, sizeof(uint32_t));
    if (e == cudaSuccess && p->chain) e = cudaMemsetAsync(p->d_ticket, 0, sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, partial_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_partial, 0, partial_bytes, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rows, rows.data(), n_streams * sizeof(FxRow), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused_fx<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused_fx<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused_fx<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FX_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fused_fx<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FX_SMEM);
    if (e != cudaSuccess) {
        rb_fx_destroy(p);
        return e;
    }
    p->args.rows = p->d_rows, p->args.n_rows = (uint32_t)n_streams, p->args.channels = mixer_channels;
    p->args.has_cv = has_cv, p->args.has_echo = has_echo, p->args.has_agc = has_agc, p->args.has_lim = has_lim, p->args.partial = p->d_partial, p->args.mix_len = mix_len;
    p->args.out = d_out, p->args.ticket = p->d_ticket;
    *out = p;
    return cudaSuccess;
}

cudaError_t rb_fx_run(rb_fx_plan* p, cudaStream_t st) {
    if (p->chain) {
        p->args.epoch = p->epoch++;      // tickets and tags of this render
        if (p->args.channels == 2) k_fused_fx<2, true><<<p->n_ctas, FX_THREADS, FX_SMEM, st>>>(p->args);
        else k_fused_fx<1, true><<<p->n_ctas, FX_THREADS, FX_SMEM, st>>>(p->args);
        return cudaGetLastError();
    }
    if (p->args.channels == 2) k_fused_fx<2, false><<<p->n_ctas, FX_THREADS, FX_SMEM, st>>>(p->args);
    else k_fused_fx<1, false><<<p->n_ctas, FX_THREADS, FX_SMEM, st>>>(p->args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    uint64_t blocks = (p->args.mix_len + 255) / 256;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    k_fx_sum_partials<<<(uint32_t)blocks, 256, 0, st>>>(p->d_partial, p->n_ctas, p->args.mix_len, p->d_out);
    return cudaGetLastError();
}

void rb_fx_destroy(rb_fx_plan* p) {
    if (!p) return;
    cudaFree(p->d_rows);
    cudaFree(p->d_partial);
    cudaFree(p->d_ticket);
    delete p;
}
bool rb_fx_chain(const rb_fx_plan* p) { return p && p->chain; }
