// rb_fused_rows.h — host-visible part of the fused planner: the per-stream record the fused kernels read (FusedRow), the
// parser that recognises the fused chain family, and the hand-over of a batch to the lane-per-stream kernel.
// No device syntax: included by rb_fused.cu (product) and by the host-emulated library of the CPU suite
// (tests/emu/hostemu.cpp), so that rb_batch_create -> parse -> lane plan runs on the CPU as it runs in front of the GPU.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rb_fused.h"
#include "rb_lanes.h"

namespace {

constexpr int TT = 256;            // mixer-timeline samples per tile
constexpr int MAX_GAINS = 4;

struct FusedRow {                  // one stream, device side
    const void* in;
    uint64_t n_in;
    uint64_t out_len;              // samples on the mixer timeline
    uint64_t mix_start;
    uint32_t fmt, c_in;
    uint32_t has_uniform;
    uint32_t mode;                 // row-wide fast mode, see ROW_*
    rb_uniform_params uni;
    uint32_t q32, r32;             // divmod(32 * from, to): index advance for a lane stride of 32 output frames
    uint32_t qT, rT;               // divmod(TT * from, to): index advance from one full tile to the next
    float den_f, rcp_den;          // (f32)to and RN(1 / (f32)to)
    float pre[MAX_GAINS];          // gains applied to raw input samples (before interpolation)
    float mid[MAX_GAINS];          // gains between the uniform conversion and the biquad
    float post[MAX_GAINS];         // gains after the biquad
    float b0, b1, b2, a1, a2;
    uint32_t pad_[1];
};
enum : uint32_t {
    ROW_GENERIC = 0,   // exact closed form per sample (span chunks, trailing partial frames, huge ratios)
    ROW_DIRECT = 1,    // no conversion at all: out[o] = in[o]
    ROW_PASS = 2,      // same rate, channel map only
    ROW_LERP = 3       // linear interpolation on the reduced grid, one segment, whole frames
};

}  // namespace

static const rb_node_dev& node_at(const rb_fused_stream& s, uint32_t i) {
    return *reinterpret_cast<const rb_node_dev*>(reinterpret_cast<const char*>(s.nodes) + (size_t)i * s.node_stride);
}

// Parse one stream into a FusedRow; returns false when its chain is outside the fused shape.
// front = 1: the biquad sits in front of the conversion (`source.low_pass(f)` handed to the mixer): pre = gains in front of the
// filter, mid = gains between the filter and the conversion, post = gains behind the conversion.
static bool parse_row(const rb_fused_stream& s, uint16_t mixer_ch, FusedRow& r, uint32_t& n_pre, uint32_t& n_mid,
                      uint32_t& n_post, uint32_t& has_uniform, uint32_t& has_biquad, uint32_t& front) {
    memset(&r, 0, sizeof(r));
    front = 0;
    r.in = s.in, r.n_in = s.n_in, r.out_len = s.out_len, r.mix_start = s.mix_start;
    r.fmt = s.fmt, r.c_in = s.c_in;
    n_pre = n_mid = n_post = has_uniform = has_biquad = 0;
    uint32_t cur_c = s.c_in;
    for (uint32_t i = 0; i < s.n_nodes; i++) {
        const rb_node_dev& nd = node_at(s, i);
        switch (nd.kind) {
            case RB_N_CONVERT:
                if (i != 0) return false;
                break;   // the format is applied at load time
            case RB_N_AMPLIFY:
                if (has_biquad) {
                    if (n_post >= MAX_GAINS) return false;
                    r.post[n_post++] = nd.p.amp.factor;
                } else if (has_uniform) {
                    if (n_mid >= MAX_GAINS) return false;
                    r.mid[n_mid++] = nd.p.amp.factor;
                } else {
                    if (n_pre >= MAX_GAINS) return false;
                    r.pre[n_pre++] = nd.p.amp.factor;
                }
                break;
            case RB_N_UNIFORM:
                if (has_uniform) return false;
                if (has_biquad) {   // what was collected behind the filter sits between it and the conversion
                    front = 1;
                    for (uint32_t g = 0; g < n_post; g++) r.mid[g] = r.post[g], r.post[g] = 0.0f;
                    n_mid = n_post, n_post = 0;
                }
                has_uniform = 1;
                r.uni = nd.p.uni;
                if (nd.c_in != s.c_in) return false;
                cur_c = nd.c_out;
                break;
            case RB_N_BIQUAD:
                if (has_biquad) return false;
                has_biquad = 1;
                r.b0 = nd.p.blt.b0, r.b1 = nd.p.blt.b1, r.b2 = nd.p.blt.b2, r.a1 = nd.p.blt.a1, r.a2 = nd.p.blt.a2;
                break;
            default: return false;
        }
    }
    if (cur_c != mixer_ch) return false;
    if (!has_uniform) {
        // gains before a (missing) uniform were collected as `pre`; keep that, n_mid stays 0
    }
    r.has_uniform = has_uniform;
    r.mode = ROW_GENERIC;
    if (!has_uniform) {
        r.mode = ROW_DIRECT;
        // the HOT kernel walks same-rate rows as a 1:1 "ratio"
        r.uni.from = r.uni.to = 1, r.uni.tail.L = s.n_in / (s.c_in ? s.c_in : 1);
        r.q32 = 32, r.r32 = 0, r.qT = TT, r.rT = 0, r.den_f = 1.0f, r.rcp_den = 1.0f;
    } else {
        const rb_uniform_params& u = r.uni;
        r.den_f = (float)u.to;
        r.rcp_den = 1.0f / r.den_f;
        if (u.from == u.to) {
            if (u.tail.p == 0 && (u.chunk_samples == 0 || u.chunk_samples % s.c_in == 0)) {
                r.mode = ROW_PASS;
                r.uni.tail.L = r.n_in / (s.c_in ? s.c_in : 1);   // chunks are irrelevant for a whole-frame pass-through
                r.q32 = 32, r.r32 = 0, r.qT = TT, r.rT = 0;
            }
        } else if (u.chunk_samples == 0 && u.tail.p == 0 && u.from <= (1u << 20) && u.to <= (1u << 20)) {
            r.mode = ROW_LERP;
            r.q32 = (uint32_t)((32ull * u.from) / u.to);
            r.r32 = (uint32_t)((32ull * u.from) % u.to);
            r.qT = (uint32_t)(((uint64_t)TT * u.from) / u.to);
            r.rT = (uint32_t)(((uint64_t)TT * u.from) % u.to);
        }
    }
    return true;
}

// Parse every stream of a batch.  The fused kernels want one chain shape for the whole batch (gain counts, uniform?, biquad?).
// One difference is harmless: a conversion that is the identity (source already in the mixer's format) is dropped by the
// planner, so such a row has no uniform node where its neighbours have one -- `mixed_u` then tells that only kernels
// which treat rows individually (the lane kernel) may take the batch.  Returns false when the batch is outside the family.
static bool fused_parse_rows(const rb_fused_stream* streams, size_t n_streams, uint16_t mixer_channels, std::vector<FusedRow>& rows,
                             uint32_t& n_pre, uint32_t& n_mid, uint32_t& n_post, uint32_t& has_u, uint32_t& has_b, bool& mixed_u,
                             uint32_t& front) {
    n_pre = n_mid = n_post = has_u = has_b = front = 0;
    mixed_u = false;
    struct Shape { uint32_t p, m, q, u, b, f; };
    std::vector<Shape> sh(n_streams);
    for (size_t i = 0; i < n_streams; i++) {
        Shape& k = sh[i];
        if (!parse_row(streams[i], mixer_channels, rows[i], k.p, k.m, k.q, k.u, k.b, k.f)) return false;
        if (k.f && !front) front = 1, n_mid = k.m;   // n_mid: how many gains the batch has between filter and conversion
    }
    const uint32_t m_front = n_mid;
    for (size_t i = 0; i < n_streams; i++) {
        Shape& k = sh[i];
        // a filtered row without a conversion (it is in the mixer's format already) beside rows whose filter sits in front of
        // theirs: its filter is "in front" just as well -- the first gains behind it file as `mid` like theirs, the rest as `post`
        if (front && !k.f && !k.u && k.b && k.q >= m_front) {
            FusedRow& r = rows[i];
            for (uint32_t g = 0; g < m_front; g++) r.mid[g] = r.post[g];
            for (uint32_t g = m_front; g < k.q; g++) r.post[g - m_front] = r.post[g];
            for (uint32_t g = k.q - m_front; g < k.q; g++) r.post[g] = 0.0f;
            k.m = m_front, k.q -= m_front, k.f = 1;
        }
        if (i == 0) n_pre = k.p, n_mid = k.m, n_post = k.q, has_u = k.u, has_b = k.b;
        else if (k.p != n_pre || k.m != n_mid || k.q != n_post || k.b != has_b) return false;
        else if (k.u != has_u) mixed_u = true, has_u = 1;
        if (k.f != front) return false;
    }
    // with gains in front of the biquad the two row kinds file them differently (pre / mid): not the same shape after all
    if (mixed_u && !front && (n_pre || n_mid)) return false;
    if (front) mixed_u = true;   // only the lane kernel knows a filter in front of the conversion
    return true;
}

// The lane kernels take the batch when RB_FUSED_LANES / RB_FUSED_DUO ask for them, or from 128 streams per SM on (18 944 on a
// B200), where they are the faster kernels anyway (measured, profiles/README.md round 2: k_fused_duo 1.96 / 2.14 / 3.25 ms at
// 16 384 / 32 768 / 65 536 streams x 1 s against 1.76 / 3.51 / 7.0 ms for k_fused_hot).  *lanes stays NULL when the shape is not
// the kernels'.
static cudaError_t fused_lanes_hook(const std::vector<FusedRow>& rows, size_t n_streams, uint16_t mixer_channels, bool all_f32,
                                    uint32_t n_pre, uint32_t n_mid, uint32_t n_post, uint32_t has_u, uint32_t has_b, uint32_t front,
                                    uint32_t flags, int sm_count, float* d_out, uint64_t mix_len, cudaStream_t st, rb_lanes_plan** lanes) {
    *lanes = nullptr;
    // A filter in front of the conversion has no other fused kernel: the alternative is the general path (one kernel per adapter,
    // intermediates in HBM: 24.5 ms against 0.85 ms on cfg3), so the lane kernel takes such a batch as soon as every SM gets a warp.
    const size_t sms = (size_t)(sm_count > 0 ? sm_count : 148);
    const bool want_lanes = (flags & (RB_FUSED_LANES | RB_FUSED_DUO)) || n_streams >= 128 * sms || (front && n_streams >= 32 * sms);
    // RB_BIQUAD_TIME_PARALLEL: the lane kernels serve the batch cut into timeline segments when it qualifies (rb_lanes_batch.cu);
    // when it does not, the flag changes nothing
    bool want_tp = (flags & RB_BIQUAD_TIME_PARALLEL) && has_b && !front && mixer_channels == 1;
    // a chain without a filter carries nothing from sample to sample: cut into timeline segments it is still the serial run bit
    // for bit, and a batch of a few thousand mono streams fills the machine that way (measured against k_fused_hot<1, false>)
    // (k_lerp_mix, parallel over the timeline: from 64 streams on, and for RB_MIX_EXACT_ORDER, whose order it keeps)
    // RB_FUSED_DUO names the lane-pair kernel itself: k_lerp_mix stays out, and the segment plan starts where it was measured
    // Measured (profiles/README.md R2.3): at 4096 streams x 2 s the segment rows on k_fused_duo take 0.402 ms, k_lerp_mix 0.422 ms;
    // at 512 streams k_lerp_mix 0.147 ms against 0.20 ms for the generic fused kernel -> k_lerp_mix below 1024 streams and for
    // RB_MIX_EXACT_ORDER (any size), the segment plan from 1024 streams on
    const bool name_duo = (flags & RB_FUSED_DUO) != 0;
    size_t seg_min = name_duo ? 1024 : 64;
    const bool lerpmix_small = !name_duo && (n_streams < 1024 || getenv("RB_LERPMIX"));   // RB_LERPMIX=1: A/B runs at any size
    if (const char* e = getenv("RB_SEGMENTS_FROM")) seg_min = (size_t)atoll(e);
    const bool exact_order = (flags & RB_MIX_EXACT_ORDER) != 0;
    if (!has_b && has_u && !front && mixer_channels == 1 && (n_streams >= seg_min || exact_order) && !(flags & RB_FUSED_LANES)) want_tp = true;
    if (!((want_lanes || want_tp) && (mixer_channels == 1 || mixer_channels == 2) && all_f32 && (has_u || has_b) && n_pre <= 1)) return cudaSuccess;
    // f32 streams with the mixer's channel count (or mono in a stereo mixer), each at or below the mixer's rate (classes per
    // rate pair), at most one gain in front of the conversion (source.amplify(v) handed to the mixer), optional biquad, at most
    // one gain directly in front of the sum.
    const uint32_t C = mixer_channels;
    // front: [gain] filter [gain] conversion [gain]
    const bool shape = front ? (has_b && n_mid <= 1 && n_post <= 1) : (has_b ? (n_mid == 0 && n_post <= 1) : (n_mid + n_post <= 1));
    if (!shape || mix_len % C != 0) return cudaSuccess;
    std::vector<rb_lanes_stream> ls(n_streams);
    for (size_t i = 0; i < n_streams; i++) {
        const FusedRow& r = rows[i];
        // every stream interpolates upwards on its own reduced grid (several rate pairs are served class by class),
        // or is at the mixer's rate already (UniformSourceIterator hands it through / there is no conversion in the chain)
        // ... or downwards by at most a factor of two (fast tiles of their own, untimed so far: only on request)
        const bool lerp_up = r.mode == ROW_LERP && (r.uni.from < r.uni.to ||
                                                    ((flags & RB_FUSED_LANES) && !front && r.uni.from > r.uni.to &&
                                                     (uint64_t)r.uni.from <= 2ull * r.uni.to));   // = lanes::ratio_runs_down
        const bool pass = r.mode == ROW_PASS || r.mode == ROW_DIRECT;
        // the stream has the mixer's channels, or is mono in a stereo mixer (repeated on both channels, channels.rs:57-85)
        if (!((lerp_up || pass) && (r.c_in == C || (r.c_in == 1 && C == 2)) && r.out_len % C == 0 && r.mix_start % C == 0 &&
              r.n_in % r.c_in == 0))
            return cudaSuccess;
        rb_lanes_stream& l = ls[i];
        l.channels = r.c_in;
        l.in = (const float*)r.in, l.n_frames = r.uni.tail.L, l.out_len = r.out_len / C, l.mix_start = r.mix_start / C;
        l.from = pass ? 1u : r.uni.from, l.to = pass ? 1u : r.uni.to;
        l.b0 = r.b0, l.b1 = r.b1, l.b2 = r.b2, l.a1 = r.a1, l.a2 = r.a2;
        l.post = n_post ? r.post[0] : (n_mid && !front ? r.mid[0] : 1.0f);
        l.pre = n_pre ? r.pre[0] : 1.0f;
        l.mid = front && n_mid ? r.mid[0] : 1.0f;
    }
    cudaError_t e = rb_lanes_try_create(ls.data(), n_streams, C, has_b != 0, front ? n_post != 0 : (n_mid + n_post) != 0, n_pre != 0, front != 0, d_out,
                                        mix_len / C, sm_count, st, lanes,
                                        (want_tp ? LANES_TIME_PARALLEL : 0u) | ((flags & RB_FUSED_LANES) ? LANES_NO_DUO : 0u) | (exact_order ? LANES_ONE_GROUP : 0u) | ((name_duo || !(lerpmix_small || exact_order)) ? LANES_NO_LERPMIX : 0u));   // RB_FUSED_LANES names k_fused_lanes itself
    if (e == cudaSuccess && *lanes && !want_lanes && rb_lanes_kind(*lanes) != 4 && rb_lanes_kind(*lanes) != 6) {   // asked for the time-parallel plan only, and it was not built
        rb_lanes_destroy(*lanes);
        *lanes = nullptr;
    }
    return e;
}
