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
    return 0;
}
