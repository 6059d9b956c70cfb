// Dependent-issue latency of the FP32 operations on the biquad recurrence chain, one warp per SM sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false chain_latency.cu -o chain_latency && ./chain_latency
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 4096

template <int MODE, bool REGCOEF = false>
__global__ void k(float* out, long long* cycles, float a1, float a2, float neg1, int active_lanes) {
    __shared__ float row[32 * 33];
    if (REGCOEF) a1 += out[threadIdx.x + 96], a2 += out[threadIdx.x + 128], neg1 += out[threadIdx.x + 160];   // per-lane registers, like the kernel
    float y1 = out[threadIdx.x], y2 = out[threadIdx.x + 32];
    float t = out[threadIdx.x + 64];
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) row[i] = t;
    __syncthreads();
    if ((int)(threadIdx.x & 31) >= active_lanes) return;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_ITER; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            float y;
            if (MODE == 0) y = __fsub_rn(__fsub_rn(t, __fmul_rn(a1, y1)), __fmul_rn(a2, y2));            // FMUL, FADD, FADD
            if (MODE == 1) y = __fmaf_rn(__fmul_rn(a2, y2), neg1, __fmaf_rn(__fmul_rn(a1, y1), neg1, t)); // FMUL, FFMA, FFMA
            if (MODE == 2) y = __fmaf_rn(y1, a1, t);                                                        // one FFMA per step
            if (MODE == 3) y = __fadd_rn(y1, t);                                                            // one FADD per step
            if (MODE == 4) y = __fmul_rn(y1, a1);                                                           // one FMUL per step
            if (MODE == 5) y = __fmaf_rn(y1, 0.999f, t);                                                    // FFMA, immediate operand
            if (MODE == 6) y = __fadd_rn(__fmul_rn(y1, a1), t);                                             // FMUL -> FADD
            if (MODE == 7) {                                                                                // chain + LDS/STS like the kernel
                float x = row[(i + u) & 31];
                y = __fmaf_rn(__fmul_rn(a2, y2), neg1, __fmaf_rn(__fmul_rn(a1, y1), neg1, x));
                row[(i + u) & 31] = y;
            }
            y2 = y1, y1 = y;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = y1 + y2;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE, bool REGCOEF = false>
void run(const char* name, int ops, int threads = 32, int active = 32) {
    float* d;
    long long* c;
    cudaMalloc(&d, 4096 * 4);
    cudaMemset(d, 0, 4096 * 4);
    cudaMalloc(&c, 8);
    for (int r = 0; r < 2; r++) k<MODE, REGCOEF><<<1, threads>>>(d, c, 0.5f, 0.25f, -1.0f, active);
    long long h = 0;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads=%4d active=%2d  %.2f cycles/step  (%d dependent ops -> %.2f each)\n", name, threads, active,
           (double)h / N_ITER, ops, (double)h / N_ITER / ops);
    cudaFree(d), cudaFree(c);
}


// ---- the recurrence warp's real loop shapes: t comes from shared memory (16-byte loads), y goes back ----
#define FBX(t_, y1_, y2_) __fmaf_rn(__fmul_rn(a2, y2_), neg1, __fmaf_rn(__fmul_rn(a1, y1_), neg1, t_))
#define FB4(v_, o_)                 \
    o_.x = FBX(v_.x, y1, y2);       \
    o_.y = FBX(v_.y, o_.x, y1);     \
    o_.z = FBX(v_.z, o_.y, o_.x);   \
    o_.w = FBX(v_.w, o_.z, o_.y);   \
    y2 = o_.z, y1 = o_.w;
template <int SHAPE>
__global__ void krow(float* out, long long* cycles, float a1, float a2, float neg1, int n_tiles) {
    extern __shared__ __align__(16) float rows[];   // [32][260]
    const int lane = threadIdx.x & 31;
    a1 += out[lane + 96], a2 += out[lane + 128], neg1 += out[lane + 160];
    for (int i = threadIdx.x; i < 32 * 260; i += blockDim.x) rows[i] = 0.001f * (i & 15);
    __syncthreads();
    if (lane >= 28 || threadIdx.x >= 32) return;
    float y1 = 0.f, y2 = 0.f;
    float4* p0 = reinterpret_cast<float4*>(rows + lane * 260 + 4);
    long long t0 = clock64();
    for (int tile = 0; tile < n_tiles; tile++) {
        float4* p4 = p0;
        if (SHAPE == 0) {           // kernel shape: 16 samples per step, loads one step ahead, stores interleaved
            float4 c0 = p4[0], c1 = p4[1], c2 = p4[2], c3 = p4[3];
            for (int n4 = 64; n4 >= 8; n4 -= 8) {
                const float4 d0 = p4[4], d1 = p4[5], d2 = p4[6], d3 = p4[7];
                float4 o;
                FB4(c0, o) p4[0] = o; FB4(c1, o) p4[1] = o; FB4(c2, o) p4[2] = o; FB4(c3, o) p4[3] = o;
                if (n4 > 8) c0 = p4[8], c1 = p4[9], c2 = p4[10], c3 = p4[11];
                FB4(d0, o) p4[4] = o; FB4(d1, o) p4[5] = o; FB4(d2, o) p4[6] = o; FB4(d3, o) p4[7] = o;
                p4 += 8;
            }
        } else if (SHAPE == 1) {    // loads only (results folded into the state so they stay alive)
            float4 c0 = p4[0], c1 = p4[1], c2 = p4[2], c3 = p4[3];
            for (int n4 = 64; n4 >= 8; n4 -= 8) {
                const float4 d0 = p4[4], d1 = p4[5], d2 = p4[6], d3 = p4[7];
                float4 o;
                FB4(c0, o) FB4(c1, o) FB4(c2, o) FB4(c3, o)
                if (n4 > 8) c0 = p4[8], c1 = p4[9], c2 = p4[10], c3 = p4[11];
                FB4(d0, o) FB4(d1, o) FB4(d2, o) FB4(d3, o)
                p4 += 8;
            }
        } else if (SHAPE == 2) {    // stores only
            float4 c0 = make_float4(y1, y2, a1, a2);
            for (int n4 = 64; n4 >= 8; n4 -= 8) {
                float4 o;
                FB4(c0, o) p4[0] = o; FB4(c0, o) p4[1] = o; FB4(c0, o) p4[2] = o; FB4(c0, o) p4[3] = o;
                FB4(c0, o) p4[4] = o; FB4(c0, o) p4[5] = o; FB4(c0, o) p4[6] = o; FB4(c0, o) p4[7] = o;
                p4 += 8;
            }
        } else if (SHAPE == 3) {    // 32 samples per step: eight loads up front, eight stores at the end
            for (int n4 = 64; n4 >= 8; n4 -= 8) {
                float4 v[8], o[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = p4[j];
#pragma unroll
                for (int j = 0; j < 8; j++) { FB4(v[j], o[j]) }
#pragma unroll
                for (int j = 0; j < 8; j++) p4[j] = o[j];
                p4 += 8;
            }
        } else if (SHAPE == 4) {    // scalar loads / stores
            float* p = reinterpret_cast<float*>(p4);
#pragma unroll 8
            for (int n = 0; n < 256; n++) {
                const float y = FBX(p[n], y1, y2);
                y2 = y1, y1 = y;
                p[n] = y;
            }
        }
    }
    long long t1 = clock64();
    out[lane] = y1 + y2;
    if (lane == 0) cycles[0] = t1 - t0;
}
template <int SHAPE>
void runrow(const char* name, int threads = 32) {
    float* d;
    long long* c;
    cudaMalloc(&d, 4096 * 4);
    cudaMemset(d, 0, 4096 * 4);
    cudaMalloc(&c, 8);
    const int n_tiles = 64;
    for (int r = 0; r < 2; r++) krow<SHAPE><<<1, threads, 32 * 260 * 4>>>(d, c, 0.5f, 0.25f, -1.0f, n_tiles);
    long long h = 0;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %.2f cycles/sample\n", name, (double)h / (n_tiles * 256.0));
    cudaFree(d), cudaFree(c);
}

int main() {
    run<0>("FMUL -> FADD -> FADD", 3);
    run<1>("FMUL -> FFMA(-1) -> FFMA(-1)", 3);
    run<2>("FFMA (3 reg)", 1);
    run<3>("FADD", 1);
    run<4>("FMUL", 1);
    run<5>("FFMA (immediate)", 1);
    run<6>("FMUL -> FADD", 2);
    run<7>("chain + LDS/STS", 3);
    run<1, true>("FMUL -> FFMA -> FFMA, coefficients in registers", 3);
    run<0, true>("FMUL -> FADD -> FADD, coefficients in registers", 3);
    run<7, true>("chain + LDS/STS, coefficients in registers", 3);
    run<1>("FMUL -> FFMA -> FFMA, 16 active lanes", 3, 32, 16);
    run<1>("FMUL -> FFMA -> FFMA, 1 active lane", 3, 32, 1);
    run<1>("same, 4 warps (one per sub-partition)", 3, 128, 32);
    run<1>("same, 8 warps (two per sub-partition)", 3, 256, 32);
    run<0>("FMUL -> FADD -> FADD, 8 warps", 3, 256, 32);
    runrow<0>("row loop as in k_fused_hot (16/step, look-ahead, stores)");
    runrow<1>("  loads only");
    runrow<2>("  stores only");
    runrow<3>("  32/step: loads up front, stores at the end");
    runrow<4>("  scalar loads and stores");
    return 0;
}
