// rb_lanes.cu — k_fused_lanes: the lane-per-stream fused kernel for large uniform batches (the warp program lives in
// rb_lanes_core.h, shared with the CPU emulator of tests/emu/), its input classification, the ordered sum of the
// per-warp partial rows, and the host-side plan.  Opt-in through RB_FUSED_LANES (include/rodio_b200.h).
#include <algorithm>
#include <vector>

#include "rb_lanes.h"
#include "rb_lanes_plan.h"

namespace {

constexpr int LANES_WARPS = 2;                      // warps per CTA: every warp is independent, small CTAs pack the SM
constexpr int LANES_THREADS = 32 * LANES_WARPS;
template <int C>
constexpr size_t lanes_smem_bytes() { return (size_t)LANES_WARPS * 32 * lanes::Geo<C>::RS * sizeof(float); }   // 21.0 / 41.0 KB

template <int C, bool HASB, bool FF2, int NPOST>
__global__ void __launch_bounds__(LANES_THREADS) k_fused_lanes(lanes::Args a) {
    extern __shared__ __align__(16) float lanes_smem[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x * LANES_WARPS + warp;
    if (group >= a.n_groups) return;   // whole warps leave: the warp program only synchronises within a warp
    lanes::warp_main<C, HASB, FF2, NPOST>(a, group, lanes_smem + (size_t)warp * 32 * lanes::Geo<C>::RS);
}

// One CTA per stream: does every non-zero |x| lie inside [2^-70, 2^60]?  (rb_lanes_core.h, "Exact division".)
__global__ void __launch_bounds__(256) k_classify_inputs(lanes::Row* rows, uint32_t n_rows, uint32_t channels) {
    const uint32_t r = blockIdx.x;
    if (r >= n_rows) return;
    const float* __restrict__ x = rows[r].in;
    const uint64_t L = rows[r].L * channels;   // floats
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

template <int C, bool HASB, bool FF2, int NPOST>
static void launch_lanes(const lanes::Args& a, cudaStream_t st) {
    const uint32_t n_ctas = (a.n_groups + LANES_WARPS - 1) / LANES_WARPS;
    k_fused_lanes<C, HASB, FF2, NPOST><<<n_ctas, LANES_THREADS, lanes_smem_bytes<C>(), st>>>(a);   // < 48 KB: no opt-in needed
}
template <int C>
static void launch_lanes_c(const lanes::Args& a, bool has_biquad, bool ff2, bool has_post, cudaStream_t st) {
    if (has_biquad) {
        if (ff2) has_post ? launch_lanes<C, true, true, 1>(a, st) : launch_lanes<C, true, true, 0>(a, st);
        else has_post ? launch_lanes<C, true, false, 1>(a, st) : launch_lanes<C, true, false, 0>(a, st);
    } else {
        has_post ? launch_lanes<C, false, false, 1>(a, st) : launch_lanes<C, false, false, 0>(a, st);
    }
}

}  // namespace

static cudaError_t launch_lanes_any(const lanes::Args& a, uint32_t channels, bool has_biquad, bool ff2, bool has_post, cudaStream_t st) {
    if (channels == 2) launch_lanes_c<2>(a, has_biquad, ff2, has_post, st);
    else launch_lanes_c<1>(a, has_biquad, ff2, has_post, st);
    return cudaGetLastError();
}

static cudaError_t launch_sum_groups(const lanes::Args& a, uint32_t channels, float* d_out, cudaStream_t st) {
    const uint64_t n = a.mix_len * channels;   // floats
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    k_sum_groups<<<(uint32_t)blocks, 256, 0, st>>>(a.partial, a.n_groups, a.pstride, n, d_out);
    return cudaGetLastError();
}

cudaError_t rb_lanes_launch_block(const lanes::Args& a, uint32_t channels, bool has_biquad, bool ff2, bool has_post, float* d_out,
                                  cudaStream_t st) {
    if (a.mix_len == 0 || a.n_groups == 0 || (channels != 1 && channels != 2)) return cudaSuccess;
    cudaError_t e = launch_lanes_any(a, channels, has_biquad, ff2, has_post, st);
    if (e != cudaSuccess) return e;
    return launch_sum_groups(a, channels, d_out, st);
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
    lanes::Args args{};
    lanes::Row* d_rows = nullptr;
    float* d_partial = nullptr;
    float* d_zeros = nullptr;
    float* d_out = nullptr;
    bool has_biquad = false, ff2 = false, has_post = false;
    bool classified = false;
    uint32_t channels = 1;
};

cudaError_t rb_lanes_try_create(const rb_lanes_stream* streams, size_t n_streams, uint32_t channels, uint32_t from, uint32_t to,
                                bool has_biquad, bool has_post, float* d_out, uint64_t mix_len, int sm_count, cudaStream_t st,
                                rb_lanes_plan** out) {
    (void)sm_count;
    *out = nullptr;
    if (n_streams == 0 || mix_len == 0 || !(from < to) || to > (1u << 20) || (channels != 1 && channels != 2)) return cudaSuccess;
    std::vector<lanes::Row> rows(n_streams);
    bool ff2 = has_biquad;
    for (size_t i = 0; i < n_streams; i++) {
        const rb_lanes_stream& s = streams[i];
        if (reinterpret_cast<uintptr_t>(s.in) & 15u) return cudaSuccess;
        lanes::Row& r = rows[i];
        memset(&r, 0, sizeof(r));
        r.in = s.in, r.L = s.n_frames, r.out_len = s.out_len, r.mix_start = s.mix_start;
        r.n_int = lanes::n_interp(r.L, from, to, r.out_len);
        r.b0 = s.b0, r.b1 = s.b1, r.b2 = s.b2, r.a1 = s.a1, r.a2 = s.a2;
        r.post = has_post ? s.post : 1.0f;
        r.flags = lanes::ROW_UNSAFE;   // until classified
        float k = 0.0f;
        if (has_biquad && lanes::ff2_coeffs(r.b0, r.b1, r.b2, &k)) r.ffk = k;
        else ff2 = false;
    }
    auto p = new rb_lanes_plan;
    p->has_biquad = has_biquad, p->ff2 = ff2, p->has_post = has_post, p->d_out = d_out, p->channels = channels;
    lanes::Args& a = p->args;
    a.n_rows = (uint32_t)n_streams, a.n_groups = (uint32_t)((n_streams + 31) / 32);
    lanes::fill_ratio(a, from, to, channels);
    a.mix_len = mix_len, a.pstride = lanes::round_up_tile(mix_len * channels);
    cudaError_t e = cudaMalloc(&p->d_rows, n_streams * sizeof(lanes::Row));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, (size_t)a.n_groups * a.pstride * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_zeros, 256);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_partial, 0, (size_t)a.n_groups * a.pstride * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->d_zeros, 0, 256, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_rows, rows.data(), n_streams * sizeof(lanes::Row), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        rb_lanes_destroy(p);
        return e;
    }
    a.rows = p->d_rows, a.partial = p->d_partial, a.zeros = p->d_zeros;
    *out = p;
    return cudaSuccess;
}

void rb_lanes_inputs_changed(rb_lanes_plan* p) {
    if (p) p->classified = false;
}

cudaError_t rb_lanes_run(rb_lanes_plan* p, cudaStream_t st) {
    const lanes::Args& a = p->args;
    if (!p->classified) {
        k_classify_inputs<<<a.n_rows, 256, 0, st>>>(p->d_rows, a.n_rows, p->channels);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        p->classified = true;
    }
    cudaError_t e = launch_lanes_any(a, p->channels, p->has_biquad, p->ff2, p->has_post, st);
    if (e != cudaSuccess) return e;
    return launch_sum_groups(a, p->channels, p->d_out, st);
}

uint32_t rb_lanes_launch_count(const rb_lanes_plan*) { return 2u; }

void rb_lanes_destroy(rb_lanes_plan* p) {
    if (!p) return;
    cudaFree(p->d_rows);
    cudaFree(p->d_partial);
    cudaFree(p->d_zeros);
    delete p;
}
