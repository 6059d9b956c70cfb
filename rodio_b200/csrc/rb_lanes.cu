// rb_lanes.cu — k_fused_lanes: the lane-per-stream fused kernel for large uniform batches (the warp program lives in
// rb_lanes_core.h, shared with the CPU emulator of tests/emu/), its input classification, the ordered sum of the
// per-warp partial rows, and the host-side plan (one launch per class of streams: rate pair x source channels).
// Chosen by the fused planner for very large batches or on RB_FUSED_LANES (include/rodio_b200.h); also renders the blocks
// of the streaming sessions (rb_session_* in rb_api.cu).
#include <algorithm>
#include <vector>

#include "rb_lanes.h"
#include "rb_lanes_plan.h"

namespace {

constexpr int LANES_WARPS = 2;                      // warps per CTA: every warp is independent, small CTAs pack the SM
constexpr int LANES_THREADS = 32 * LANES_WARPS;
template <int C>
constexpr size_t lanes_smem_bytes() { return (size_t)LANES_WARPS * 32 * lanes::Geo<C>::RS * sizeof(float); }   // 21.0 / 41.0 KB

template <int CI, int CO, bool HASB, bool FF2, int NPOST, bool PASS>
__global__ void __launch_bounds__(LANES_THREADS) k_fused_lanes(lanes::Args a) {
    extern __shared__ __align__(16) float lanes_smem[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x * LANES_WARPS + warp;
    if (group >= a.n_groups) return;   // whole warps leave: the warp program only synchronises within a warp
    lanes::warp_main<CI, CO, HASB, FF2, NPOST, PASS>(a, group, lanes_smem + (size_t)warp * 32 * lanes::Geo<CI>::RS);
}

// One CTA per stream: does every non-zero |x| lie inside [2^-70, 2^60]?  (rb_lanes_core.h, "Exact division".)
// channels[r]: interleaved channels of stream r.
__global__ void __launch_bounds__(256) k_classify_inputs(lanes::Row* rows, uint32_t n_rows, const uint8_t* __restrict__ channels) {
    const uint32_t r = blockIdx.x;
    if (r >= n_rows) return;
    const float* __restrict__ x = rows[r].in;
    const uint64_t L = rows[r].L * channels[r];   // floats
    bool bad = false;
    const uint64_t n4 = L / 4;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    auto out_of_class = [](float v) {
        const uint32_t u = __float_as_uint(v) & 0x7fffffffu;
        return u != 0u && (u - 0x1c800000u) >= (0x5d800000u - 0x1c800000u);
    };
    for (uint64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = __ldg(x4 + i);
        bad |= out_of_class(v.x) | out_of_class(v.y) | out_of_class(v.z) | out_of_class(v.w);
    }
    for (uint64_t i = n4 * 4 + threadIdx.x; i < L; i += blockDim.x) bad |= out_of_class(__ldg(x + i));
    const int any = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) rows[r].flags = any ? lanes::ROW_UNSAFE : 0u;
}

// out[m] = +0.0 + partial[0][m] + partial[1][m] + ...  (warp order = insertion order of the streams)
__global__ void __launch_bounds__(256) k_sum_groups(const float* __restrict__ partial, uint32_t n_groups, uint64_t pstride,
                                                    uint64_t mix_len, float* __restrict__ out) {
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < mix_len; m += (uint64_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (uint32_t g = 0; g < n_groups; g++) acc = __fadd_rn(acc, partial[(uint64_t)g * pstride + m]);
        out[m] = acc;
    }
}

// FIFO append + sticky classification, one CTA per stream
__global__ void __launch_bounds__(128) k_fifo_append(const float* __restrict__ staging, const uint64_t* __restrict__ offset,
                                                     const uint32_t* __restrict__ count, const uint32_t* __restrict__ fill,
                                                     float* __restrict__ fifo, uint64_t stride, uint32_t* __restrict__ flags) {
    const uint32_t r = blockIdx.x, n = count[r];
    if (n == 0) return;
    const float* __restrict__ src = staging + offset[r];
    float* __restrict__ dst = fifo + (uint64_t)r * stride + fill[r];
    bool bad = false;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = src[i];
        const uint32_t u = __float_as_uint(v) & 0x7fffffffu;
        bad |= u != 0u && (u - 0x1c800000u) >= (0x5d800000u - 0x1c800000u);
        dst[i] = v;
    }
    if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) flags[r] = 1u;
}

__global__ void __launch_bounds__(128) k_fifo_compact(const float* __restrict__ src, float* __restrict__ dst, uint64_t stride,
                                                      const uint32_t* __restrict__ drop, const uint32_t* __restrict__ keep) {
    const uint32_t r = blockIdx.x, n = keep[r];
    const float* __restrict__ s = src + (uint64_t)r * stride + drop[r];   // drop is a multiple of 4: both sides 16-byte aligned
    float* __restrict__ d = dst + (uint64_t)r * stride;
    const uint32_t n4 = n / 4;
    for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x)
        reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    for (uint32_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

__global__ void __launch_bounds__(256) k_classify_range(const float* __restrict__ x, uint64_t n, uint32_t* __restrict__ flag) {
    bool bad = false;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t u = __float_as_uint(x[i]) & 0x7fffffffu;
        bad |= u != 0u && (u - 0x1c800000u) >= (0x5d800000u - 0x1c800000u);
    }
    if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) *flag = 1u;
}

template <int CI, int CO, bool HASB, bool FF2, int NPOST, bool PASS>
static void launch_lanes(const lanes::Args& a, cudaStream_t st) {
    const uint32_t n_ctas = (a.n_groups + LANES_WARPS - 1) / LANES_WARPS;
    k_fused_lanes<CI, CO, HASB, FF2, NPOST, PASS><<<n_ctas, LANES_THREADS, lanes_smem_bytes<CI>(), st>>>(a);   // < 48 KB: no opt-in needed
}
template <int CI, int CO, bool PASS>
static void launch_lanes_c(const lanes::Args& a, bool has_biquad, bool ff2, bool has_post, cudaStream_t st) {
    if (has_biquad) {
        if (ff2) has_post ? launch_lanes<CI, CO, true, true, 1, PASS>(a, st) : launch_lanes<CI, CO, true, true, 0, PASS>(a, st);
        else has_post ? launch_lanes<CI, CO, true, false, 1, PASS>(a, st) : launch_lanes<CI, CO, true, false, 0, PASS>(a, st);
    } else {
        has_post ? launch_lanes<CI, CO, false, false, 1, PASS>(a, st) : launch_lanes<CI, CO, false, false, 0, PASS>(a, st);
    }
}

}  // namespace

cudaError_t rb_lanes_launch_kernel(const lanes::Args& a, uint32_t ch_in, uint32_t ch_out, bool has_biquad, bool ff2, bool has_post,
                                   cudaStream_t st) {
    if (a.mix_len == 0 || a.n_groups == 0) return cudaSuccess;
    if (!((ch_in == 1 || ch_in == 2) && (ch_in == ch_out || (ch_in == 1 && ch_out == 2)))) return cudaErrorInvalidValue;
    const bool pass = a.from == a.to;   // sources at the mixer's rate: taps used raw
    if (ch_in == 2) pass ? launch_lanes_c<2, 2, true>(a, has_biquad, ff2, has_post, st) : launch_lanes_c<2, 2, false>(a, has_biquad, ff2, has_post, st);
    else if (ch_out == 2) pass ? launch_lanes_c<1, 2, true>(a, has_biquad, ff2, has_post, st) : launch_lanes_c<1, 2, false>(a, has_biquad, ff2, has_post, st);
    else pass ? launch_lanes_c<1, 1, true>(a, has_biquad, ff2, has_post, st) : launch_lanes_c<1, 1, false>(a, has_biquad, ff2, has_post, st);
    return cudaGetLastError();
}

cudaError_t rb_lanes_launch_sum(const float* d_partial, uint32_t n_groups, uint64_t pstride, uint64_t n_floats, float* d_out,
                                cudaStream_t st) {
    if (n_floats == 0) return cudaSuccess;
    uint64_t blocks = (n_floats + 255) / 256;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    k_sum_groups<<<(uint32_t)blocks, 256, 0, st>>>(d_partial, n_groups, pstride, n_floats, d_out);
    return cudaGetLastError();
}

cudaError_t rb_lanes_fifo_append(const float* d_staging, const uint64_t* d_offset, const uint32_t* d_count, const uint32_t* d_fill,
                                 float* d_fifo, uint64_t stride, uint32_t* d_flags, uint32_t n_streams, cudaStream_t st) {
    if (n_streams == 0) return cudaSuccess;
    k_fifo_append<<<n_streams, 128, 0, st>>>(d_staging, d_offset, d_count, d_fill, d_fifo, stride, d_flags);
    return cudaGetLastError();
}

cudaError_t rb_lanes_classify_range(const float* d_ptr, uint64_t n, uint32_t* d_flag, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_classify_range<<<1, 256, 0, st>>>(d_ptr, n, d_flag);
    return cudaGetLastError();
}

cudaError_t rb_lanes_fifo_compact(const float* d_src, float* d_dst, uint64_t stride, const uint32_t* d_drop, const uint32_t* d_keep,
                                  uint32_t n_streams, cudaStream_t st) {
    if (n_streams == 0) return cudaSuccess;
    k_fifo_compact<<<n_streams, 128, 0, st>>>(d_src, d_dst, stride, d_drop, d_keep);
    return cudaGetLastError();
}

struct rb_lanes_plan {
    struct Class {
        lanes::Args args{};
        bool ff2 = false;
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
    bool has_biquad = false, has_post = false;
    bool classified = false;
};

cudaError_t rb_lanes_try_create(const rb_lanes_stream* streams, size_t n_streams, uint32_t channels, bool has_biquad, bool has_post,
                                float* d_out, uint64_t mix_len, int sm_count, cudaStream_t st, rb_lanes_plan** out) {
    (void)sm_count;
    *out = nullptr;
    if (n_streams == 0 || n_streams > 0x7fffffffull || mix_len == 0 || (channels != 1 && channels != 2)) return cudaSuccess;
    std::vector<uint32_t> from(n_streams), to(n_streams), chs(n_streams);
    for (size_t i = 0; i < n_streams; i++) {
        from[i] = streams[i].from, to[i] = streams[i].to, chs[i] = streams[i].channels;
        if (!(from[i] <= to[i]) || from[i] == 0 || to[i] > (1u << 20)) return cudaSuccess;
        if (!(chs[i] == channels || (chs[i] == 1 && channels == 2))) return cudaSuccess;
        if (reinterpret_cast<uintptr_t>(streams[i].in) & 15u) return cudaSuccess;
    }
    const auto classes = lanes::classes_by_ratio(from.data(), to.data(), chs.data(), (uint32_t)n_streams);
    auto p = new rb_lanes_plan;
    p->has_biquad = has_biquad, p->has_post = has_post, p->d_out = d_out, p->channels = channels;
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
            r.flags = lanes::ROW_UNSAFE;   // until classified
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
        k_classify_inputs<<<p->n_rows, 256, 0, st>>>(p->d_rows, p->n_rows, p->d_row_channels);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        p->classified = true;
    }
    for (const auto& c : p->classes) {
        cudaError_t e = rb_lanes_launch_kernel(c.args, c.ch_in, p->channels, p->has_biquad, c.ff2, p->has_post, st);
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
