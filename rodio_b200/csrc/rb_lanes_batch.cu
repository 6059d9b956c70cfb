// rb_lanes_batch.cu — host-side plan of the lane kernels for whole batches: streams partitioned into classes (rate pair x
// source channels), one launch per class over its rows and partial rows, inputs classified once per upload.
//   * a class of mono sources below the mixer's rate whose neighbours are in phase goes to the lane-PAIR kernel
//     (k_fused_duo, rb_duo_core.h: 64 streams per warp, packed arithmetic), every other class to k_fused_lanes;
//   * RB_BIQUAD_TIME_PARALLEL (mode bit LANES_TIME_PARALLEL): the mixer timeline is cut into segments, every stream
//     contributes one row per segment that starts a warm-up in front of the segment from zero filter state, and the rows of
//     one segment form warps of their own -- a 4096-stream batch then runs as 60 000 rows.  Taken only for filters that
//     pass lanes::tp_filter_ok (the result stays within the tolerance of the exact path); otherwise the exact plan is built.
//     A chain WITHOUT a filter has no state at all: its segments need no warm-up and reproduce the serial run bit for bit,
//     so the fused planner asks for this plan by itself (rb_fused_rows.h) when the batch alone cannot fill the machine.
// No device syntax in this file: besides nvcc (product) it is compiled as plain C++ against tests/emu/mockcuda by the CPU
// suite, with the launchers of rb_lanes.cu replaced by the SIMT emulator (tests/test_session_hostemu.py).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rb_lanes.h"
#include "rb_lanes_plan.h"

struct rb_lanes_plan {
    struct Class {
        lanes::Args args{};
        bool ff2 = false;
        bool guard = false;   // some row's gain in front is outside the range of the unguarded tile
        bool duo = false;     // served by k_fused_duo (groups of 64 rows)
        uint32_t ch_in = 1;
    };
    std::vector<Class> classes;      // one launch per reduced rate pair
    lanes::Row* d_rows = nullptr;    // class after class
    float* d_partial = nullptr;      // [n_partial_rows][pstride]
    float* d_zeros = nullptr;
    float* d_out = nullptr;
    uint8_t* d_row_channels = nullptr;   // [n_rows], class order: interleaved channels of every stream
    lanes::GroupSpan* d_spans = nullptr; // time-parallel plan: one record per group
    lanes::Row* d_stream_rows = nullptr; // time-parallel plan: one row per STREAM, what the classification runs on ...
    uint32_t* d_row_stream = nullptr;    // ... and the stream of every plan row (~0u: padding)
    uint32_t n_rows = 0, n_partial_rows = 0, n_streams = 0, channels = 1;   // channels: of the mixer
    uint64_t pstride = 0, mix_len = 0;
    bool has_biquad = false, has_post = false, has_pre = false, front = false;
    bool classified = false;
    bool time_parallel = false;
    uint32_t tp_segments = 0, tp_warmup = 0, tp_seg_len = 0;
    // filter-free chains: the time-parallel resample -> gain -> mix kernel (k_lerp_mix) serves the batch
    bool lerpmix = false;
    rb_lerpmix_args lm{};
    rb_lerpmix_row* d_lm_rows = nullptr;
};

namespace {

void fill_row(lanes::Row& r, const rb_lanes_stream& s, bool has_biquad, bool has_post, bool has_pre, bool front) {
    memset(&r, 0, sizeof(r));
    r.in = s.in, r.L = s.n_frames, r.out_len = s.out_len, r.mix_start = s.mix_start;
    r.n_int = lanes::n_interp(r.L, s.from, s.to, r.out_len);
    r.b0 = s.b0, r.b1 = s.b1, r.b2 = s.b2, r.a1 = s.a1, r.a2 = s.a2;
    r.post = has_post ? s.post : 1.0f;
    r.pre = has_pre ? s.pre : 1.0f;
    r.mid = front ? s.mid : 1.0f;
    r.flags = lanes::ROW_UNSAFE;   // until classified
    (void)has_biquad;
}

// The time-parallel plan of one class (mono, below the mixer's rate, whole streams that are pairwise in phase, every filter
// within the accuracy gate).  Returns false when the class does not qualify.
// Without a filter (has_biquad false) nothing is carried from sample to sample: the segments need no warm-up and every output
// is the bit the serial run produces -- the plan is then simply the way a few thousand streams fill the machine.
bool plan_time_parallel(const rb_lanes_stream* streams, const std::vector<uint32_t>& cls, uint64_t mix_len, int sm_count, bool has_biquad, bool has_post,
                        std::vector<lanes::Row>& rows, std::vector<uint32_t>& row_stream, std::vector<lanes::GroupSpan>& spans,
                        uint32_t* n_slots, uint32_t* warmup, uint32_t* seg_len, bool* ff2) {
    const uint32_t S = (uint32_t)cls.size();
    if (S == 0) return false;
    double rmax = 0.0;
    for (uint32_t i : cls) {
        double r = 0, g = 0;
        if (!has_biquad) break;
        if (!lanes::tp_filter_ok(streams[i].a1, streams[i].a2, &r, &g)) return false;
        rmax = std::max(rmax, r);
    }
    const uint32_t to = streams[cls[0]].to;
    for (size_t k = 0; k + 1 < cls.size(); k++)   // every stream in phase with its neighbour: any two may share a lane
        if ((streams[cls[k]].mix_start % (4ull * to)) != (streams[cls[k + 1]].mix_start % (4ull * to))) return false;
    const uint32_t W = has_biquad ? lanes::tp_warmup(rmax) : 0u;
    // segments: enough rows to fill the machine (7 groups of 64 rows per SM), but never shorter than 8 warm-ups
    // (2048 frames without a filter: a run has to be worth priming the rings for)
    uint64_t warps_per_sm = 7;       // groups of 64 rows per SM (10 fit: one wave with room to spare); measured best among 7 ... 24
    if (const char* e = getenv("RB_TP_WARPS_PER_SM")) warps_per_sm = std::max<uint64_t>(1, (uint64_t)atoll(e));
    const uint64_t want_rows = 64ull * (uint64_t)(sm_count > 0 ? sm_count : 148) * warps_per_sm;
    uint64_t K = (want_rows + S - 1) / S;
    const uint64_t max_k = std::max<uint64_t>(1, mix_len / (has_biquad ? 8ull * W : 2048ull));
    K = std::max<uint64_t>(1, std::min<uint64_t>(K, max_k));
    if (const char* e = getenv("RB_TP_SEGMENTS")) K = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)atoll(e), std::max<uint64_t>(1, mix_len / lanes::TILE)));
    if (K < 2) return false;   // nothing to gain
    const uint64_t L = lanes::round_up_tile((mix_len + K - 1) / K);
    *warmup = W, *seg_len = (uint32_t)L, *n_slots = (S + 63) / 64;
    *ff2 = true;
    for (uint64_t lo = 0; lo < mix_len; lo += L) {
        const uint64_t hi = std::min(mix_len, lo + L);
        const uint64_t from_t = lo > W ? lo - W : 0;
        const size_t first = rows.size();
        for (uint32_t i : cls) {
            const rb_lanes_stream& s = streams[i];
            const uint64_t s_end = s.mix_start + s.out_len;
            if (s.out_len == 0 || s.mix_start >= hi || s_end <= lo) continue;      // not active inside the segment
            lanes::Row r;
            fill_row(r, s, has_biquad, has_post, false, false);
            const uint64_t n_int_stream = r.n_int;
            r.mix_start = std::max(s.mix_start, from_t);
            r.o0 = r.mix_start - s.mix_start;
            r.out_len = std::min(s_end, hi) - r.mix_start;
            r.n_int = n_int_stream > r.o0 ? std::min<uint64_t>(r.out_len, n_int_stream - r.o0) : 0;
            float k = 0.0f;
            if (has_biquad && lanes::ff2_coeffs(r.b0, r.b1, r.b2, &k)) r.ffk = k;
            else *ff2 = false;
            rows.push_back(r);
            row_stream.push_back(i);
        }
        while ((rows.size() - first) % 64) {       // groups never straddle segments
            lanes::Row r;
            memset(&r, 0, sizeof(r));
            rows.push_back(r);
            row_stream.push_back(~0u);
        }
        for (size_t g = 0; g < (rows.size() - first) / 64; g++) spans.push_back(lanes::GroupSpan{lo, (uint32_t)g, 0u});
    }
    return true;
}

}  // namespace

cudaError_t rb_lanes_try_create(const rb_lanes_stream* streams, size_t n_streams, uint32_t channels, bool has_biquad, bool has_post,
                                bool has_pre, bool front, float* d_out, uint64_t mix_len, int sm_count, cudaStream_t st, rb_lanes_plan** out,
                                uint32_t mode) {
    *out = nullptr;
    if (n_streams == 0 || n_streams > 0x7fffffffull || mix_len == 0 || (channels != 1 && channels != 2)) return cudaSuccess;
    std::vector<uint32_t> from(n_streams), to(n_streams), chs(n_streams);
    for (size_t i = 0; i < n_streams; i++) {
        from[i] = streams[i].from, to[i] = streams[i].to, chs[i] = streams[i].channels;
        // at or below the mixer's rate, or above it by at most a factor of two (the DOWN tiles)
        if (!(from[i] <= to[i] || lanes::ratio_runs_down(from[i], to[i])) || from[i] == 0 || from[i] > (1u << 20) || to[i] > (1u << 20)) return cudaSuccess;
        if (!(chs[i] == channels || (chs[i] == 1 && channels == 2))) return cudaSuccess;
        if (reinterpret_cast<uintptr_t>(streams[i].in) & 15u) return cudaSuccess;
    }
    const auto classes = lanes::classes_by_ratio(from.data(), to.data(), chs.data(), (uint32_t)n_streams);
    const bool no_duo = (mode & LANES_NO_DUO) || getenv("RB_NO_DUO") != nullptr;
    const bool duo_shape = channels == 1 && !has_pre && !front && !no_duo;
    auto p = new rb_lanes_plan;
    p->has_biquad = has_biquad, p->has_post = has_post, p->has_pre = has_pre, p->front = front, p->d_out = d_out, p->channels = channels;
    p->n_streams = (uint32_t)n_streams, p->mix_len = mix_len, p->pstride = lanes::round_up_tile(mix_len * channels);
    std::vector<lanes::Row> rows;
    rows.reserve(n_streams);
    std::vector<uint8_t> row_channels;
    std::vector<uint32_t> row_stream;
    std::vector<lanes::GroupSpan> spans;
    std::vector<size_t> first_row, first_slot;
    uint32_t n_slots_total = 0;

    // ---- chains without a filter: nothing is carried from sample to sample, the whole batch is parallel over the timeline ----
    if ((mode & LANES_TIME_PARALLEL) && !has_biquad && channels == 1 && !has_pre && !front && classes.size() == 1 && from[0] < to[0] &&
        mix_len < (1ull << 31) && !(mode & LANES_NO_LERPMIX) && !getenv("RB_NO_LERPMIX")) {
        const uint32_t T = to[0];
        uint64_t origin = ~0ull;
        bool in_phase = true;
        for (size_t i = 0; i < n_streams; i++)
            if (streams[i].out_len) origin = std::min(origin, streams[i].mix_start);
        for (size_t i = 0; i < n_streams && in_phase; i++)
            in_phase = streams[i].out_len == 0 || (streams[i].mix_start - origin) % T == 0;
        if (in_phase && origin != ~0ull) {
            std::vector<rb_lerpmix_row> lr(n_streams);
            for (size_t i = 0; i < n_streams; i++) {
                const rb_lanes_stream& s = streams[i];
                lanes::Row r;
                fill_row(r, s, false, has_post, false, false);
                rows.push_back(r);
                row_channels.push_back(1);
                const uint64_t shift = s.out_len ? (s.mix_start - origin) / T * from[0] : 0;
                lr[i].p = s.in - shift;
                lr[i].lo = (uint32_t)s.mix_start, lr[i].hi_int = (uint32_t)(s.mix_start + r.n_int), lr[i].hi = (uint32_t)(s.mix_start + s.out_len);
                lr[i].post = has_post ? s.post : 1.0f, lr[i].row = (uint32_t)i, lr[i].pad_ = 0;
            }
            // stream groups: a group is summed sequentially in insertion order by one CTA per tile of 1024 frames.  64 streams per
            // group keep the partial rows at 3 % of the input bytes; fewer (down to 16) when that leaves the machine short of CTAs
            const uint64_t tiles = (mix_len + 1023) / 1024;
            const uint64_t two_waves = 2ull * (uint64_t)(sm_count > 0 ? sm_count : 148);
            uint64_t groups = std::max<uint64_t>((two_waves + tiles - 1) / tiles, (n_streams + 63) / 64);
            groups = std::max<uint64_t>(1, std::min<uint64_t>(groups, (n_streams + 15) / 16));
            if (mode & LANES_ONE_GROUP) groups = 1;
            const uint32_t per = (uint32_t)((n_streams + groups - 1) / groups);
            groups = (n_streams + per - 1) / per;
            p->lerpmix = true;
            p->n_rows = (uint32_t)n_streams, p->n_partial_rows = groups > 1 ? (uint32_t)groups : 0;
            cudaError_t e = cudaMalloc(&p->d_rows, n_streams * sizeof(lanes::Row));
            if (e == cudaSuccess) e = cudaMalloc(&p->d_lm_rows, n_streams * sizeof(rb_lerpmix_row));
            if (e == cudaSuccess) e = cudaMalloc(&p->d_row_channels, n_streams);
            if (e == cudaSuccess && groups > 1) e = cudaMalloc(&p->d_partial, (size_t)groups * p->pstride * sizeof(float));
            if (e == cudaSuccess && groups > 1) e = cudaMemsetAsync(p->d_partial, 0, (size_t)groups * p->pstride * sizeof(float), st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rows, rows.data(), n_streams * sizeof(lanes::Row), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_lm_rows, lr.data(), n_streams * sizeof(rb_lerpmix_row), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_row_channels, row_channels.data(), n_streams, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                rb_lanes_destroy(p);
                return e;
            }
            rb_lerpmix_args& a = p->lm;
            a.rows = p->d_lm_rows, a.lane_rows = p->d_rows, a.n_rows = (uint32_t)n_streams, a.rows_per_group = per, a.n_groups = (uint32_t)groups;
            a.from = from[0], a.to = T, a.origin = origin, a.den_f = (float)T, a.rcp_den = 1.0f / (float)T, a.mix_len = mix_len;
            a.out = groups > 1 ? p->d_partial : d_out, a.pstride = p->pstride, a.has_post = has_post ? 1u : 0u;
            *out = p;
            return cudaSuccess;
        }
    }
    // ---- the time-parallel plan: one class of mono sources, a filter, nothing in front of the conversion ----
    if ((mode & LANES_TIME_PARALLEL) && duo_shape && classes.size() == 1 && from[0] < to[0]) {
        uint32_t n_slots = 0;
        bool ff2 = false;
        const bool force_post = true;   // (y * 1.0 is exact: the plan always runs the instantiation with a gain)
        std::vector<rb_lanes_stream> ls(streams, streams + n_streams);
        if (!has_post)
            for (auto& s : ls) s.post = 1.0f;
        if (plan_time_parallel(ls.data(), classes[0], mix_len, sm_count, has_biquad, force_post, rows, row_stream, spans, &n_slots, &p->tp_warmup, &p->tp_seg_len, &ff2)) {
            rb_lanes_plan::Class c;
            c.ch_in = 1, c.duo = true, c.ff2 = ff2;
            c.args.n_rows = (uint32_t)rows.size(), c.args.n_groups = (uint32_t)(rows.size() / 64);
            lanes::fill_ratio(c.args, from[0], to[0], 1);
            c.args.mix_len = mix_len, c.args.pstride = p->pstride;
            first_row.push_back(0), first_slot.push_back(0);
            n_slots_total = n_slots;
            p->classes.push_back(c);
            p->time_parallel = true, p->has_post = true;
            p->tp_segments = (uint32_t)((mix_len + p->tp_seg_len - 1) / p->tp_seg_len);
            row_channels.assign(n_streams, 1);   // of the per-stream rows the classification runs on
        }
    }
    if (!p->time_parallel) {
        for (const auto& cls : classes) {
            rb_lanes_plan::Class c;
            c.ch_in = chs[cls[0]];
            lanes::Args& a = c.args;
            a.n_rows = (uint32_t)cls.size();
            lanes::fill_ratio(a, from[cls[0]], to[cls[0]], channels);
            a.mix_len = mix_len, a.pstride = p->pstride;
            c.ff2 = has_biquad;
            first_row.push_back(rows.size());
            for (uint32_t i : cls) {
                const rb_lanes_stream& s = streams[i];
                lanes::Row r;
                fill_row(r, s, has_biquad, has_post, has_pre, front);
                if (has_pre && !front && !lanes::pre_gain_keeps_class(r.pre)) r.flags |= lanes::ROW_FORCE_SLOW, c.guard = true;
                float k = 0.0f;
                if (has_biquad && lanes::ff2_coeffs(r.b0, r.b1, r.b2, &k)) r.ffk = k;
                else c.ff2 = false;
                rows.push_back(r);
                row_channels.push_back((uint8_t)s.channels);
            }
            c.duo = duo_shape && c.ch_in == 1 && a.from < a.to && lanes::duo_compatible(rows.data() + first_row.back(), cls.size(), a.to);
            a.n_groups = c.duo ? (a.n_rows + 63) / 64 : (a.n_rows + 31) / 32;
            first_slot.push_back(n_slots_total);
            n_slots_total += a.n_groups;
            p->classes.push_back(c);
        }
    }
    p->n_rows = (uint32_t)rows.size(), p->n_partial_rows = n_slots_total;
    const size_t partial_bytes = (size_t)n_slots_total * p->pstride * sizeof(float);
    cudaError_t e = cudaMalloc(&p->d_rows, rows.size() * sizeof(lanes::Row));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, partial_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_zeros, 256);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_row_channels, row_channels.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_row_channels, row_channels.data(), row_channels.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_partial, 0, partial_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_zeros, 0, 256, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rows, rows.data(), rows.size() * sizeof(lanes::Row), cudaMemcpyHostToDevice, st);
    std::vector<lanes::Row> stream_rows;
    if (e == cudaSuccess && p->time_parallel) {
        stream_rows.resize(n_streams);
        for (size_t i = 0; i < n_streams; i++) {
            memset(&stream_rows[i], 0, sizeof(lanes::Row));
            stream_rows[i].in = streams[i].in, stream_rows[i].L = streams[i].n_frames, stream_rows[i].flags = lanes::ROW_UNSAFE;
        }
        e = cudaMalloc(&p->d_stream_rows, n_streams * sizeof(lanes::Row));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_row_stream, row_stream.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_spans, spans.size() * sizeof(lanes::GroupSpan));
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_stream_rows, stream_rows.data(), n_streams * sizeof(lanes::Row), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_row_stream, row_stream.data(), row_stream.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_spans, spans.data(), spans.size() * sizeof(lanes::GroupSpan), cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        rb_lanes_destroy(p);
        return e;
    }
    for (size_t k = 0; k < p->classes.size(); k++) {
        lanes::Args& a = p->classes[k].args;
        a.rows = p->d_rows + first_row[k], a.partial = p->d_partial + first_slot[k] * p->pstride, a.zeros = p->d_zeros;
        a.spans = p->time_parallel ? p->d_spans : nullptr;
    }
    *out = p;
    return cudaSuccess;
}

void rb_lanes_inputs_changed(rb_lanes_plan* p) {
    if (p) p->classified = false;
}

cudaError_t rb_lanes_run(rb_lanes_plan* p, cudaStream_t st) {
    if (p->lerpmix) {
        if (!p->classified) {
            cudaError_t e = rb_lanes_launch_classify(p->d_rows, p->n_rows, p->d_row_channels, st);
            if (e != cudaSuccess) return e;
            p->classified = true;
        }
        cudaError_t e = rb_lerpmix_launch(p->lm, st);
        if (e != cudaSuccess || p->lm.n_groups == 1) return e;
        return rb_lanes_launch_sum(p->d_partial, p->lm.n_groups, p->pstride, p->mix_len, p->d_out, st);
    }
    if (!p->classified) {
        cudaError_t e;
        if (p->time_parallel) {
            e = rb_lanes_launch_classify(p->d_stream_rows, p->n_streams, p->d_row_channels, st);
            if (e == cudaSuccess) e = rb_lanes_spread_flags(p->d_rows, p->n_rows, p->d_row_stream, p->d_stream_rows, st);
        } else {
            e = rb_lanes_launch_classify(p->d_rows, p->n_rows, p->d_row_channels, st);
        }
        if (e != cudaSuccess) return e;
        p->classified = true;
    }
    for (const auto& c : p->classes) {
        cudaError_t e = c.duo ? rb_duo_launch_kernel(c.args, p->has_biquad, c.ff2, p->has_post, st)
                              : rb_lanes_launch_kernel(c.args, c.ch_in, p->channels, p->has_biquad, c.ff2 && !p->front, p->has_post, p->has_pre, p->front, c.guard, st);
        if (e != cudaSuccess) return e;
    }
    return rb_lanes_launch_sum(p->d_partial, p->n_partial_rows, p->pstride, p->mix_len * p->channels, p->d_out, st);
}

uint32_t rb_lanes_launch_count(const rb_lanes_plan* p) { return p->lerpmix ? (p->lm.n_groups > 1 ? 2u : 1u) : (uint32_t)p->classes.size() + 1u; }
// 3: every class on the lane-pair kernel, 4: the time-parallel plan, 2: k_fused_lanes (alone or beside pair classes)
int rb_lanes_kind(const rb_lanes_plan* p) {
    if (p->lerpmix) return 6;
    if (p->time_parallel) return 4;
    bool all = !p->classes.empty();
    for (const auto& c : p->classes) all = all && c.duo;
    return all ? 3 : 2;
}
uint32_t rb_lanes_mix_group(const rb_lanes_plan* p) { return p->lerpmix ? (p->lm.n_groups > 1 ? p->lm.rows_per_group : 0u) : rb_lanes_kind(p) >= 3 ? 64u : 32u; }
void rb_lanes_tp_geometry(const rb_lanes_plan* p, uint32_t* segments, uint32_t* seg_len, uint32_t* warmup) {
    *segments = p->tp_segments, *seg_len = p->tp_seg_len, *warmup = p->tp_warmup;
}

void rb_lanes_destroy(rb_lanes_plan* p) {
    if (!p) return;
    cudaFree(p->d_rows);
    cudaFree(p->d_partial);
    cudaFree(p->d_zeros);
    cudaFree(p->d_row_channels);
    cudaFree(p->d_spans);
    cudaFree(p->d_stream_rows);
    cudaFree(p->d_row_stream);
    cudaFree(p->d_lm_rows);
    delete p;
}
