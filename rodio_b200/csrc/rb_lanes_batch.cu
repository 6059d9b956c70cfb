// rb_lanes_batch.cu — host-side plan of the lane-per-stream kernel for whole batches: streams partitioned into classes (rate
// pair x source channels), one launch per class over its rows and partial rows, inputs classified once per upload.
// No device syntax in this file: besides nvcc (product) it is compiled as plain C++ against tests/emu/mockcuda by the CPU
// suite, with the launchers of rb_lanes.cu replaced by the SIMT emulator (tests/test_session_hostemu.py).
#include <cstring>
#include <vector>

#include "rb_lanes.h"
#include "rb_lanes_plan.h"

struct rb_lanes_plan {
    struct Class {
        lanes::Args args{};
        bool ff2 = false;
        bool guard = false;   // some row's gain in front is outside the range of the unguarded tile
        uint32_t ch_in = 1;
    };
    std::vector<Class> classes;      // one launch per reduced rate pair
    lanes::Row* d_rows = nullptr;    // class after class
    float* d_partial = nullptr;      // [n_groups_total][pstride]
    float* d_zeros = nullptr;
    float* d_out = nullptr;
    uint8_t* d_row_channels = nullptr;   // [n_rows], class order: interleaved channels of every stream
    uint32_t n_rows = 0, n_groups_total = 0, channels = 1;   // channels: of the mixer
    uint64_t pstride = 0, mix_len = 0;
    bool has_biquad = false, has_post = false, has_pre = false, front = false;
    bool classified = false;
};

cudaError_t rb_lanes_try_create(const rb_lanes_stream* streams, size_t n_streams, uint32_t channels, bool has_biquad, bool has_post,
                                bool has_pre, bool front, float* d_out, uint64_t mix_len, int sm_count, cudaStream_t st, rb_lanes_plan** out) {
    (void)sm_count;
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
    auto p = new rb_lanes_plan;
    p->has_biquad = has_biquad, p->has_post = has_post, p->has_pre = has_pre, p->front = front, p->d_out = d_out, p->channels = channels;
    p->n_rows = (uint32_t)n_streams, p->mix_len = mix_len, p->pstride = lanes::round_up_tile(mix_len * channels);
    std::vector<lanes::Row> rows;
    rows.reserve(n_streams);
    std::vector<uint8_t> row_channels;
    std::vector<size_t> first_row;
    for (const auto& cls : classes) {
        rb_lanes_plan::Class c;
        c.ch_in = chs[cls[0]];
        lanes::Args& a = c.args;
        a.n_rows = (uint32_t)cls.size(), a.n_groups = (a.n_rows + 31) / 32;
        lanes::fill_ratio(a, from[cls[0]], to[cls[0]], channels);
        a.mix_len = mix_len, a.pstride = p->pstride;
        c.ff2 = has_biquad;
        first_row.push_back(rows.size());
        for (uint32_t i : cls) {
            const rb_lanes_stream& s = streams[i];
            lanes::Row r;
            memset(&r, 0, sizeof(r));
            r.in = s.in, r.L = s.n_frames, r.out_len = s.out_len, r.mix_start = s.mix_start;
            r.n_int = lanes::n_interp(r.L, s.from, s.to, r.out_len);
            r.b0 = s.b0, r.b1 = s.b1, r.b2 = s.b2, r.a1 = s.a1, r.a2 = s.a2;
            r.post = has_post ? s.post : 1.0f;
            r.pre = has_pre ? s.pre : 1.0f;
            r.mid = front ? s.mid : 1.0f;
            r.flags = lanes::ROW_UNSAFE;   // until classified
            if (has_pre && !front && !lanes::pre_gain_keeps_class(r.pre)) r.flags |= lanes::ROW_FORCE_SLOW, c.guard = true;
            float k = 0.0f;
            if (has_biquad && lanes::ff2_coeffs(r.b0, r.b1, r.b2, &k)) r.ffk = k;
            else c.ff2 = false;
            rows.push_back(r);
            row_channels.push_back((uint8_t)s.channels);
        }
        p->n_groups_total += a.n_groups;
        p->classes.push_back(c);
    }
    const size_t partial_bytes = (size_t)p->n_groups_total * p->pstride * sizeof(float);
    cudaError_t e = cudaMalloc(&p->d_rows, n_streams * sizeof(lanes::Row));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, partial_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_zeros, 256);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_row_channels, n_streams);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_row_channels, row_channels.data(), n_streams, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_partial, 0, partial_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_zeros, 0, 256, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rows, rows.data(), n_streams * sizeof(lanes::Row), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        rb_lanes_destroy(p);
        return e;
    }
    uint32_t g0 = 0;
    for (size_t k = 0; k < p->classes.size(); k++) {
        lanes::Args& a = p->classes[k].args;
        a.rows = p->d_rows + first_row[k], a.partial = p->d_partial + (size_t)g0 * p->pstride, a.zeros = p->d_zeros;
        g0 += a.n_groups;
    }
    *out = p;
    return cudaSuccess;
}

void rb_lanes_inputs_changed(rb_lanes_plan* p) {
    if (p) p->classified = false;
}

cudaError_t rb_lanes_run(rb_lanes_plan* p, cudaStream_t st) {
    if (!p->classified) {
        cudaError_t e = rb_lanes_launch_classify(p->d_rows, p->n_rows, p->d_row_channels, st);
        if (e != cudaSuccess) return e;
        p->classified = true;
    }
    for (const auto& c : p->classes) {
        cudaError_t e = rb_lanes_launch_kernel(c.args, c.ch_in, p->channels, p->has_biquad, c.ff2 && !p->front, p->has_post, p->has_pre, p->front, c.guard, st);
        if (e != cudaSuccess) return e;
    }
    return rb_lanes_launch_sum(p->d_partial, p->n_groups_total, p->pstride, p->mix_len * p->channels, p->d_out, st);
}

uint32_t rb_lanes_launch_count(const rb_lanes_plan* p) { return (uint32_t)p->classes.size() + 1u; }

void rb_lanes_destroy(rb_lanes_plan* p) {
    if (!p) return;
    cudaFree(p->d_rows);
    cudaFree(p->d_partial);
    cudaFree(p->d_zeros);
    cudaFree(p->d_row_channels);
    delete p;
}
