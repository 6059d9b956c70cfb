// Issue throughput of scalar and packed (f32x2) FP32 instructions on one SM, and dependent-issue latency of the packed forms.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false fp32_throughput.cu -o fp32_throughput && ./fp32_throughput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(uint64_t v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// MODE: 0 FFMA 1 FMUL 2 FADD 3 FFMA2 4 FMUL2 5 FADD2 6 FFMA+FADD alternating (same pipe?) 7 FMUL + LOP3 (fma + alu pipe)
// 8 FFMA2 + FADD2 alternating  9 FSEL/select  10 SHFL.BFLY  11 FFMA2 dependent chain  12 FMUL2->FADD2->FADD2 chain (the packed recurrence)
// 13 FFMA immediate form
template <int MODE>
__global__ void k(float* out, long long* cycles) {
    float s[ILP];
    uint64_t p[ILP];
    const float c0 = out[threadIdx.x & 31], c1 = out[32 + (threadIdx.x & 31)];
    const uint64_t pc0 = pk(c0, c1), pc1 = pk(c1, c0);
#pragma unroll
    for (int j = 0; j < ILP; j++) { s[j] = out[64 + j + threadIdx.x]; p[j] = pk(s[j], c0); }
    uint32_t m = __float_as_uint(c0);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (MODE == 0) s[j] = __fmaf_rn(s[j], c0, c1);
            if (MODE == 1) s[j] = __fmul_rn(s[j], c0);
            if (MODE == 2) s[j] = __fadd_rn(s[j], c0);
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(pc0), "l"(pc1));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(pc0));
            if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(pc0));
            if (MODE == 6) s[j] = (j & 1) ? __fmaf_rn(s[j], c0, c1) : __fadd_rn(s[j], c0);
            if (MODE == 7) { if (j & 1) s[j] = __fmul_rn(s[j], c0); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(reinterpret_cast<uint32_t&>(s[j])) : "r"(m), "r"(i)); }
            if (MODE == 8) { if (j & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(pc0), "l"(pc1)); else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(pc0)); }
            if (MODE == 9) asm volatile("{.reg .pred q; setp.gt.f32 q, %0, %1; selp.f32 %0, %1, %2, q;}" : "+f"(s[j]) : "f"(c0), "f"(c1));
            if (MODE == 10) s[j] = __shfl_xor_sync(0xffffffffu, s[j], 1 + (j & 3));
            if (MODE == 13) s[j] = __fmaf_rn(s[j], 0.999f, c1);
        }
        if (MODE == 11) {
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[0]) : "l"(pc0), "l"(pc1));
        }
        if (MODE == 12) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {   // y = (t - a1*y1) - a2*y2 on both halves
                uint64_t q, r, y;
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(p[0]), "l"(pc0));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p[1]), "l"(pc1));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(p[2]), "l"(q));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(y), "l"(r));
                p[1] = p[0], p[0] = y;
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < ILP; j++) acc += s[j] + lo(p[j]);
    out[threadIdx.x] = acc + __uint_as_float(m);
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int flops_per_instr) {
    float* d;
    long long* c;
    cudaMalloc(&d, 8192 * 4);
    cudaMemset(d, 0, 8192 * 4);
    cudaMalloc(&c, 8);
    printf("%-46s", name);
    for (int warps : {1, 4, 8, 16, 32}) {
        for (int r = 0; r < 2; r++) k<MODE><<<1, warps * 32>>>(d, c);
        long long h = 0;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double instr = (double)ITERS * ILP * warps;
        printf("  w%-2d %.3f i/clk", warps, instr / h);
    }
    printf("   (x%d lane-flop per instr)\n", flops_per_instr);
    cudaFree(d), cudaFree(c);
}

int main() {
    printf("warp instructions per clock on ONE SM (4 sub-partitions; 4.0 = every issue slot)\n");
    run<0>("FFMA 3-reg", 1);
    run<13>("FFMA immediate", 1);
    run<1>("FMUL", 1);
    run<2>("FADD", 1);
    run<3>("FFMA2 (fma.rn.f32x2)", 2);
    run<4>("FMUL2", 2);
    run<5>("FADD2", 2);
    run<6>("FFMA / FADD alternating", 1);
    run<7>("FMUL / LOP3 alternating", 1);
    run<8>("FFMA2 / FADD2 alternating", 2);
    run<9>("FSETP+SELP pairs (count = pairs)", 1);
    run<10>("SHFL.BFLY", 1);
    run<11>("FFMA2 dependent chain (w1: 1/latency)", 2);
    run<12>("FMUL2,FMUL2,FADD2,FADD2 packed recurrence (8 steps = ILP)", 2);
    return 0;
}
