// librodio_b200_hostemu.so (test infrastructure): the library's real host code -- rodio_b200/csrc/rb_api.cu, compiled as C++
// against tests/emu/mockcuda/cuda_runtime.h -- with the lane kernel's launchers replaced by the SIMT emulator.  It exports the
// C ABI of include/rodio_b200.h, so the Python mirror (rodio_b200.Session ...) drives it unchanged: the session bookkeeping of
// rb_api.cu (class order, packed pushes, FIFO compaction, state blobs) runs on the CPU exactly as it runs in front of the GPU.
// Only the streaming sessions are served: batches need the kernels of rb_kernels.cu / rb_fused.cu and fail loudly here.
#include <cstdio>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "warp_variants.h"
#include "../../rodio_b200/csrc/rb_fused.h"
#include "../../rodio_b200/csrc/rb_p2p.h"
#include "../../rodio_b200/csrc/rb_fused_rows.h"
#include "../../rodio_b200/csrc/rb_lanes.h"
#include "../../rodio_b200/csrc/rb_lanes_plan.h"
#include "../../rodio_b200/csrc/rb_duo_core.h"

// ---- allocation registry of the mock runtime ----
static std::map<void*, size_t> g_allocs;
static std::mutex g_alloc_mutex;
extern "C" void mock_cuda_register(void* p, size_t n) {
    std::lock_guard<std::mutex> l(g_alloc_mutex);
    g_allocs[p] = n;
}
extern "C" void mock_cuda_unregister(void* p) {
    std::lock_guard<std::mutex> l(g_alloc_mutex);
    g_allocs.erase(p);
}

// ---- the lane kernel on the SIMT emulator (warp_variants.cpp) ----
cudaError_t rb_lanes_launch_kernel(const lanes::Args& a, uint32_t ch_in, uint32_t ch_out, bool has_biquad, bool ff2, bool has_post,
                                   bool has_pre, bool front, bool guard, cudaStream_t) {
    if (a.mix_len == 0 || a.n_groups == 0) return cudaSuccess;
    if (!((ch_in == 1 || ch_in == 2) && (ch_in == ch_out || (ch_in == 1 && ch_out == 2)))) return cudaErrorInvalidValue;
    simt::WarpEmu warp;
    {
        std::lock_guard<std::mutex> l(g_alloc_mutex);
        for (auto& kv : g_allocs) warp.readable.push_back({(const char*)kv.first, (const char*)kv.first + kv.second});
    }
    const float nan = std::numeric_limits<float>::quiet_NaN();
    constexpr int MAX_RS = lanes::Geo<2, 4>::RS;   // the largest ring of any variant
    std::vector<float> ring_store(32 * MAX_RS + 4, nan);
    float* ring = ring_store.data();
    while ((uintptr_t)ring & 15) ring++;
    for (uint32_t g = 0; g < a.n_groups; g++) {
        for (int i = 0; i < 32 * MAX_RS; i++) ring[i] = nan;
        emu_run_group(ch_in, ch_out, a, g, &warp, ring, has_biquad, ff2, has_post, has_pre, front, guard);
    }
    return cudaSuccess;
}

// ---- the lane-pair kernel (rb_duo_core.h) on the same emulator ----
template <bool HASB, bool FF2, int NPOST>
static void duo_run(const lanes::Args& a, uint32_t g, simt::WarpEmu* w, float* ring) {
    simt::run_warp(w, [&] { duo::warp_main<HASB, FF2, NPOST>(a, g, ring); });
}
cudaError_t rb_duo_launch_kernel(const lanes::Args& a, bool has_biquad, bool ff2, bool has_post, cudaStream_t) {
    if (a.mix_len == 0 || a.n_groups == 0) return cudaSuccess;
    simt::WarpEmu warp;
    {
        std::lock_guard<std::mutex> l(g_alloc_mutex);
        for (auto& kv : g_allocs) warp.readable.push_back({(const char*)kv.first, (const char*)kv.first + kv.second});
    }
    const float nan = std::numeric_limits<float>::quiet_NaN();
    std::vector<float> ring_store(duo::WARP_WORDS + 4, nan);
    float* ring = ring_store.data();
    while ((uintptr_t)ring & 15) ring++;
    for (uint32_t g = 0; g < a.n_groups; g++) {
        for (int i = 0; i < duo::WARP_WORDS; i++) ring[i] = nan;
        if (has_biquad) {
            if (ff2) has_post ? duo_run<true, true, 1>(a, g, &warp, ring) : duo_run<true, true, 0>(a, g, &warp, ring);
            else has_post ? duo_run<true, false, 1>(a, g, &warp, ring) : duo_run<true, false, 0>(a, g, &warp, ring);
        } else {
            has_post ? duo_run<false, false, 1>(a, g, &warp, ring) : duo_run<false, false, 0>(a, g, &warp, ring);
        }
    }
    return cudaSuccess;
}
cudaError_t rb_lanes_spread_flags(lanes::Row* rows, uint32_t n_rows, const uint32_t* row_stream, const lanes::Row* stream_rows, cudaStream_t) {
    for (uint32_t r = 0; r < n_rows; r++) {
        const uint32_t bad = row_stream[r] == ~0u ? 0u : (stream_rows[row_stream[r]].flags & lanes::ROW_UNSAFE);
        rows[r].flags = (rows[r].flags & ~lanes::ROW_UNSAFE) | bad;
    }
    return cudaSuccess;
}

// k_lerp_mix on the host: the same sums in the same order (plain loops: the kernel has no warp-level structure to emulate)
cudaError_t rb_lerpmix_launch(const rb_lerpmix_args& a, cudaStream_t) {
    for (uint32_t g = 0; g < a.n_groups; g++) {
        float* out = a.out + (uint64_t)g * a.pstride;
        const uint32_t r_lo = g * a.rows_per_group, r_hi = r_lo + a.rows_per_group < a.n_rows ? r_lo + a.rows_per_group : a.n_rows;
        for (uint64_t n = 0; n < a.mix_len; n++) {
            const uint64_t prod = (n - a.origin) * (uint64_t)a.from, i = prod / a.to;
            const float numf = (float)(uint32_t)(prod - i * a.to);
            float acc = 0.0f;
            bool any = false;
            for (uint32_t r = r_lo; r < r_hi; r++) {
                const rb_lerpmix_row& row = a.rows[r];
                if (n < row.lo || n >= row.hi) continue;
                const float xa = row.p[(uint32_t)i];
                float x = xa;
                if (n < row.hi_int) x = xa + ((row.p[(uint32_t)i + 1] - xa) * numf) / a.den_f;
                if (a.has_post) x = x * row.post;
                acc = acc + x, any = true;
            }
            if (any || a.n_groups == 1) out[n] = acc;
        }
    }
    return cudaSuccess;
}

cudaError_t rb_lanes_launch_sum(const float* partial, uint32_t n_groups, uint64_t pstride, uint64_t n_floats, float* out, cudaStream_t) {
    for (uint64_t m = 0; m < n_floats; m++) {
        float acc = 0.0f;
        for (uint32_t g = 0; g < n_groups; g++) acc = acc + partial[(uint64_t)g * pstride + m];
        out[m] = acc;
    }
    return cudaSuccess;
}

cudaError_t rb_lanes_classify_range(const float* p, uint64_t n, uint32_t* flag, cudaStream_t) {
    for (uint64_t i = 0; i < n; i++)
        if (!lanes::sample_in_class(p[i])) *flag = 1u;
    return cudaSuccess;
}

cudaError_t rb_lanes_launch_classify(lanes::Row* rows, uint32_t n_rows, const uint8_t* channels, cudaStream_t) {
    for (uint32_t r = 0; r < n_rows; r++) {
        bool ok = true;
        for (uint64_t i = 0; i < rows[r].L * channels[r] && ok; i++) ok = lanes::sample_in_class(rows[r].in[i]);
        rows[r].flags = ok ? 0u : lanes::ROW_UNSAFE;
    }
    return cudaSuccess;
}

cudaError_t rb_lanes_fifo_append(const float* staging, const uint64_t* offset, const uint32_t* count, const uint32_t* fill, float* fifo,
                                 uint64_t stride, uint32_t* flags, uint32_t n_streams, cudaStream_t) {
    for (uint32_t r = 0; r < n_streams; r++)
        for (uint32_t i = 0; i < count[r]; i++) {
            const float v = staging[offset[r] + i];
            if (!lanes::sample_in_class(v)) flags[r] = 1u;
            fifo[(uint64_t)r * stride + fill[r] + i] = v;
        }
    return cudaSuccess;
}

cudaError_t rb_lanes_fifo_compact(const float* src, float* dst, uint64_t stride, const uint32_t* drop, const uint32_t* keep,
                                  uint32_t n_streams, cudaStream_t) {
    for (uint32_t r = 0; r < n_streams; r++)
        for (uint32_t i = 0; i < keep[r]; i++) dst[(uint64_t)r * stride + i] = src[(uint64_t)r * stride + drop[r] + i];
    return cudaSuccess;
}

// ---- everything else the host code links against: not available without the real kernels ----
// The fused planner as far as the CPU can follow it: the product's own parser and hand-over to the lane kernel
// (rb_fused_rows.h); batches the lane kernel does not take have no kernels here and fall to the (absent) general path.
struct rb_fused_plan {
    rb_lanes_plan* lanes = nullptr;
};
cudaError_t rb_fused_try_create(const rb_fused_stream* streams, size_t n_streams, uint16_t mixer_channels, float* d_out, uint64_t mix_len,
                                uint32_t flags, int sm_count, cudaStream_t st, rb_fused_plan** out) {
    *out = nullptr;
    if (n_streams == 0 || mix_len == 0 || (flags & RB_MIX_EXACT_ORDER)) return cudaSuccess;
    std::vector<FusedRow> rows(n_streams);
    uint32_t n_pre = 0, n_mid = 0, n_post = 0, has_u = 0, has_b = 0, front = 0;
    bool mixed_u = false, all_f32 = true;
    if (!fused_parse_rows(streams, n_streams, mixer_channels, rows, n_pre, n_mid, n_post, has_u, has_b, mixed_u, front)) return cudaSuccess;
    for (size_t i = 0; i < n_streams; i++) all_f32 = all_f32 && streams[i].fmt == RB_FMT_F32;
    rb_lanes_plan* lanes = nullptr;
    cudaError_t e = fused_lanes_hook(rows, n_streams, mixer_channels, all_f32, n_pre, n_mid, n_post, has_u, has_b, front, flags, sm_count, d_out,
                                     mix_len, st, &lanes);
    if (e != cudaSuccess || !lanes) return e;
    *out = new rb_fused_plan{lanes};
    return cudaSuccess;
}
cudaError_t rb_fused_run(rb_fused_plan* p, cudaStream_t st, bool) { return rb_lanes_run(p->lanes, st); }
bool rb_fused_partial_rows(const rb_fused_plan*, const float**, uint32_t*, uint64_t*) { return false; }
bool rb_fx_chain(const rb_fx_plan*) { return false; }
void rb_fused_destroy(rb_fused_plan* p) {
    if (p) rb_lanes_destroy(p->lanes), delete p;
}
uint32_t rb_fused_launch_count(const rb_fused_plan* p) { return rb_lanes_launch_count(p->lanes); }
void rb_fused_inputs_changed(rb_fused_plan* p) {
    if (p) rb_lanes_inputs_changed(p->lanes);
}
int rb_fused_kind(const rb_fused_plan* p) { return rb_lanes_kind(p->lanes); }
uint32_t rb_fused_mix_group(const rb_fused_plan* p) { return rb_lanes_mix_group(p->lanes); }
cudaError_t rb_launch_nodes(uint32_t, const rb_node_dev*, uint32_t, uint64_t, uint32_t, cudaStream_t) { return cudaErrorInvalidValue; }
// the NVLink peer-memory exchange (rb_p2p.cu) has no host emulation: the communicator stays on NCCL here
cudaError_t rb_p2p_create_rank(int, int, int, cudaStream_t, uint64_t, const rb_p2p_allgather&, rb_p2p** out, std::string* why) {
    *out = nullptr;
    if (why) *why = "host emulation";
    return cudaErrorInvalidValue;
}
cudaError_t rb_p2p_create_local(int n, const int*, const cudaStream_t*, uint64_t, rb_p2p** out, std::string* why) {
    for (int i = 0; i < n; i++) out[i] = nullptr;
    if (why) *why = "host emulation";
    return cudaErrorInvalidValue;
}
uint64_t rb_p2p_capacity(const rb_p2p*) { return 0; }
cudaError_t rb_p2p_allreduce(rb_p2p*, float*, uint64_t, const float*, uint32_t, uint64_t, cudaStream_t) { return cudaErrorInvalidValue; }
void rb_p2p_destroy(rb_p2p*) {}
cudaError_t rb_launch_mix(const rb_mix_src*, uint32_t, float*, uint64_t, cudaStream_t, float*, uint32_t) { return cudaErrorInvalidValue; }
cudaError_t rb_launch_convert(const void*, uint32_t, void*, uint32_t, uint64_t, cudaStream_t) { return cudaErrorInvalidValue; }


// ---- test entry: the batch plan of the lane kernel (rb_lanes_batch.cu: classes, row order, partial-row offsets) ----
// pcm[r] holds n_frames[r] * ch_in[r] floats; out_len / mix_start / mix_len in frames; out receives mix_len * channels floats.
extern "C" int hostemu_lanes_batch(const float* const* pcm, const uint64_t* n_frames, const uint64_t* out_len, const uint64_t* mix_start,
                                   const float* coefs, const float* post, const float* pre, uint32_t n, uint32_t channels,
                                   const uint32_t* ch_in, const uint32_t* from, const uint32_t* to, uint64_t mix_len, int hasb, int npost,
                                   int npre, float* out, uint32_t* n_launches) {
    std::vector<float*> d_in(n, nullptr);
    std::vector<rb_lanes_stream> st(n);
    for (uint32_t r = 0; r < n; r++) {
        if (cudaMalloc(&d_in[r], (n_frames[r] * ch_in[r] + 4) * sizeof(float)) != cudaSuccess) return 1;
        std::memcpy(d_in[r], pcm[r], n_frames[r] * ch_in[r] * sizeof(float));
        rb_lanes_stream& s = st[r];
        s.in = d_in[r], s.n_frames = n_frames[r], s.out_len = out_len[r], s.mix_start = mix_start[r];
        s.from = from[r], s.to = to[r], s.channels = ch_in[r];
        const float* c = coefs + 5 * r;
        s.b0 = c[0], s.b1 = c[1], s.b2 = c[2], s.a1 = c[3], s.a2 = c[4], s.post = post[r], s.pre = pre[r], s.mid = 1.0f;
    }
    float* d_out = nullptr;
    if (cudaMalloc(&d_out, (mix_len * channels + 8) * sizeof(float)) != cudaSuccess) return 1;
    rb_lanes_plan* plan = nullptr;
    int rc = 0;
    if (rb_lanes_try_create(st.data(), n, channels, hasb != 0, npost != 0, npre != 0, false, d_out, mix_len, 148, nullptr, &plan) != cudaSuccess || !plan) rc = 2;
    if (!rc && (rb_lanes_run(plan, nullptr) != cudaSuccess || rb_lanes_run(plan, nullptr) != cudaSuccess)) rc = 3;   // twice: idempotent
    if (!rc) std::memcpy(out, d_out, mix_len * channels * sizeof(float)), *n_launches = rb_lanes_launch_count(plan);
    rb_lanes_destroy(plan);
    cudaFree(d_out);
    for (float* p : d_in) cudaFree(p);
    return rc;
}
