// rb_simt.h — the handful of warp-level primitives the lane-per-stream kernel (rb_lanes_core.h) is written
// against, with two implementations:
//   * device (nvcc, default): thin wrappers over the CUDA intrinsics / PTX they name;
//   * RB_SIMT_EMULATE (g++, tests/emu/): every lane is a host thread, collectives meet at a std::barrier,
//     cp.async is a queue of pending 16-byte copies that only land at the matching wait_group and whose
//     destination is poisoned with NaN the moment the copy is issued -- so a read that the device code could
//     only get right by luck (too short a look-ahead, a slot refilled while still in use) reads NaN in the
//     emulator.  The emulator is test infrastructure: it lets the CPU suite run the kernel's index / ring / run
//     logic bit for bit against the oracle without a GPU.  The product never links it.
// Every float operation is an explicitly rounded single operation on both sides (no contraction: the device TU is
// built --fmad=false, the emulator -ffp-contract=off).
#pragma once
#include <cstdint>

#if defined(RB_SIMT_EMULATE)
#include <barrier>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <utility>
#include <vector>
#define SIMT_FN inline
#else
#include <cuda_runtime.h>
#define SIMT_FN __device__ __forceinline__
#endif

namespace simt {

#if defined(RB_SIMT_EMULATE)
// ------------------------------------------------------------------------------------------ emulator
struct WarpEmu;
struct LaneEmu {
    WarpEmu* w = nullptr;
    uint32_t lane = 0;
    struct Copy { float* dst; const float* src; };
    std::vector<Copy> open;                 // copies issued since the last commit
    std::deque<std::vector<Copy>> groups;   // committed, not yet landed
    uint64_t n_instr_hint = 0;
};
struct WarpEmu {
    uint32_t xu[32];
    uint64_t xl[32];
    float xf[32];
    // address ranges cp16 / ldg may read (the harness registers every input row with its 16-byte tail pad)
    std::vector<std::pair<const char*, const char*>> readable;
    bool ok_read(const void* p, size_t n) const {
        for (auto& r : readable)
            if ((const char*)p >= r.first && (const char*)p + n <= r.second) return true;
        return false;
    }
#if defined(__x86_64__)
    // The 32 lanes are fibers of ONE host thread (run_warp below): a collective is a counter, and a lane that has to wait
    // hands the processor to the scheduler, which resumes the others in an order that changes from pass to pass.
    uint32_t epoch = 0, arrived = 0, order_seed = 12345u;
    void* main_sp = nullptr;
    void* sp[32] = {};
    bool done[32] = {};
    LaneEmu lanes[32];
    std::function<void()> body;
#else
    std::barrier<> bar{32};
#endif
};

#if defined(__x86_64__)
// Context switch between fibers: callee-saved registers on the old stack, stack pointers exchanged (System V x86-64).
extern "C" void rb_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak rb_fiber_switch
.type rb_fiber_switch,@function
rb_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size rb_fiber_switch,.-rb_fiber_switch
)");
inline thread_local LaneEmu* g_cur = nullptr;
inline LaneEmu& cur() { return *g_cur; }
#else
inline thread_local LaneEmu g_lane;
inline LaneEmu& cur() { return g_lane; }
#endif

[[noreturn]] inline void emu_fail(const char* what) {
    std::fprintf(stderr, "simt emulator: %s (lane %u)\n", what, cur().lane);
    std::abort();
}
SIMT_FN uint32_t lane() { return cur().lane; }
// coverage counters of the emulator (0: fast tiles, 1: slow tiles, 2: ring refills), counted by lane 0
inline uint64_t g_emu_count[4] = {0, 0, 0, 0};
SIMT_FN void emu_count(int which, uint64_t n) {
    if (cur().lane == 0) g_emu_count[which] += n;
}
SIMT_FN void emu_assert(bool ok, const char* what) {
    if (!ok) emu_fail(what);
}
SIMT_FN float fmul(float a, float b) { return a * b; }
SIMT_FN float fadd(float a, float b) { return a + b; }
SIMT_FN float fsub(float a, float b) { return a - b; }
SIMT_FN float fdiv(float a, float b) { return a / b; }
SIMT_FN float fdiv_cold(float a, float b) { return a / b; }   // the same division, kept out of line on the device
SIMT_FN float ffma(float a, float b, float c) { return std::fmaf(a, b, c); }
SIMT_FN float u2f(uint32_t v) { return (float)v; }
SIMT_FN float ldg(const float* p) {
    if (!cur().w->ok_read(p, 4)) emu_fail("ldg outside the registered input rows");
    return *p;
}
#if defined(__x86_64__)
SIMT_FN void sync() {
    LaneEmu* me = g_cur;
    WarpEmu* w = me->w;
    const uint32_t e = w->epoch;
    if (++w->arrived == 32) {   // the last lane to arrive releases the others and runs on
        w->arrived = 0, w->epoch = e + 1;
        return;
    }
    while (w->epoch == e) {
        rb_fiber_switch(&w->sp[me->lane], w->main_sp);
        g_cur = me;
    }
}
[[noreturn]] inline void fiber_entry() {
    LaneEmu* me = g_cur;
    WarpEmu* w = me->w;
    w->body();
    w->done[me->lane] = true;
    rb_fiber_switch(&w->sp[me->lane], w->main_sp);
    std::abort();   // a finished lane is never resumed
}
// Run `body` as the 32 lanes of warp `w` (every lane executes it; simt::lane() tells them apart).
template <class F>
void run_warp(WarpEmu* w, F&& body) {
    constexpr size_t STACK = 256 * 1024;
    static thread_local std::vector<char> stacks(32 * STACK + 64);
    w->body = body, w->epoch = 0, w->arrived = 0;
    for (uint32_t l = 0; l < 32; l++) {
        w->lanes[l] = LaneEmu{};
        w->lanes[l].w = w, w->lanes[l].lane = l, w->done[l] = false;
        uintptr_t top = ((uintptr_t)stacks.data() + (l + 1) * STACK) & ~(uintptr_t)15;
        void** q = (void**)top;
        *--q = nullptr;                 // the frame fiber_entry "returns" to (never used); entry sees rsp % 16 == 8
        *--q = (void*)&fiber_entry;     // popped by the `ret` of the first switch
        for (int i = 0; i < 6; i++) *--q = nullptr;   // rbp rbx r12 r13 r14 r15
        w->sp[l] = q;
    }
    uint32_t left = 32;
    while (left) {
        // a different lane order every pass: what lockstep execution would hide (a lane touching another lane's ring slot
        // on the wrong side of a syncwarp) must not be hidden by a fixed order either
        w->order_seed = w->order_seed * 1664525u + 1013904223u;
        const uint32_t first = w->order_seed >> 27, step = ((w->order_seed >> 20) & 30u) | 1u;   // odd step: a permutation of 0..31
        const uint32_t before = left, epoch = w->epoch;
        for (uint32_t k = 0; k < 32; k++) {
            const uint32_t l = (first + k * step) & 31u;
            if (w->done[l]) continue;
            g_cur = &w->lanes[l];
            rb_fiber_switch(&w->main_sp, w->sp[l]);
            if (w->done[l]) left--;
        }
        if (left && left == before && w->epoch == epoch) {   // every live lane waits, and nobody is left to release them
            std::fprintf(stderr, "simt emulator: collective reached by %u of %u live lanes\n", w->arrived, left);
            std::abort();
        }
    }
    g_cur = nullptr;
}
#else
SIMT_FN void sync() { cur().w->bar.arrive_and_wait(); }
template <class F>
void run_warp(WarpEmu* w, F&& body) {
    std::vector<std::thread> th;
    for (uint32_t l = 0; l < 32; l++)
        th.emplace_back([&, l] {
            g_lane = LaneEmu{};
            g_lane.w = w, g_lane.lane = l;
            body();
        });
    for (auto& t : th) t.join();
}
#endif
SIMT_FN void syncwarp() { sync(); }
SIMT_FN float shfl_xor(float v, int m) {
    WarpEmu* w = cur().w;
    w->xf[cur().lane] = v;
    sync();
    const float r = w->xf[cur().lane ^ (uint32_t)m];
    sync();
    return r;
}
SIMT_FN uint32_t shfl_idx(uint32_t v, uint32_t src) {
    WarpEmu* w = cur().w;
    w->xu[cur().lane] = v;
    sync();
    const uint32_t r = w->xu[src & 31u];
    sync();
    return r;
}
SIMT_FN uint64_t shfl_idx64(uint64_t v, uint32_t src) {
    WarpEmu* w = cur().w;
    w->xl[cur().lane] = v;
    sync();
    const uint64_t r = w->xl[src & 31u];
    sync();
    return r;
}
SIMT_FN uint32_t reduce_min(uint32_t v) {
    WarpEmu* w = cur().w;
    w->xu[cur().lane] = v;
    sync();
    uint32_t r = w->xu[0];
    for (int i = 1; i < 32; i++) r = w->xu[i] < r ? w->xu[i] : r;
    sync();
    return r;
}
SIMT_FN uint64_t reduce_min64(uint64_t v) {
    WarpEmu* w = cur().w;
    w->xl[cur().lane] = v;
    sync();
    uint64_t r = w->xl[0];
    for (int i = 1; i < 32; i++) r = w->xl[i] < r ? w->xl[i] : r;
    sync();
    return r;
}
SIMT_FN uint64_t reduce_max64(uint64_t v) {
    WarpEmu* w = cur().w;
    w->xl[cur().lane] = v;
    sync();
    uint64_t r = w->xl[0];
    for (int i = 1; i < 32; i++) r = w->xl[i] > r ? w->xl[i] : r;
    sync();
    return r;
}
// Shared-memory cursor of a lane (the device keeps a 32-bit shared-window address).
using sptr = const float*;
SIMT_FN sptr sptr_of(const float* p) { return p; }
SIMT_FN sptr sptr_add(sptr p, int words) { return p + words; }
SIMT_FN bool sptr_ge(sptr a, sptr b) { return a >= b; }
SIMT_FN float lds(sptr p) { return *p; }
// One index step of the resampler for a frame of C channels: numerator += from (mod den); on a carry the right taps
// become the left ones and the next ring frame is fetched.  (Six / eight instructions on the device, see below.)
// PRE: the fetched frame is multiplied by `gpre` (an Amplify in front of the conversion, one rounding like Amplify::next).
template <int C, bool PRE = false>
SIMT_FN void lerp_advance(float& nf, float (&x0)[C], float (&x1)[C], sptr& p, float from_f, float den, float gpre = 1.0f) {
    const float nf2 = nf + from_f;
    if (nf2 >= den) {
        nf = nf2 - den;
        for (int c = 0; c < C; c++) x0[c] = x1[c], x1[c] = PRE ? p[c] * gpre : p[c];
        p += C;
    } else {
        nf = nf2;
    }
}
// m is zero or 2^-100 <= |m| < 2^100: the range in which q0 = m*r, q = fma(fma(-q0, den, m), r, q0) is the correctly
// rounded m / den (rb_lanes_core.h, "Exact division")
SIMT_FN bool in_exact_quotient_class(float m) {
    uint32_t u;
    std::memcpy(&u, &m, 4);
    u &= 0x7fffffffu;
    return u == 0u || (u - 0x0d800000u) < 0x64000000u;
}
// ---- packed pairs (two independent f32 values per lane; the device uses the f32x2 instructions of sm_100) ----
struct f2 { float lo, hi; };
SIMT_FN f2 pack2(float lo, float hi) { return f2{lo, hi}; }
SIMT_FN float lo2(f2 v) { return v.lo; }
SIMT_FN float hi2(f2 v) { return v.hi; }
SIMT_FN f2 add2(f2 a, f2 b) { return f2{a.lo + b.lo, a.hi + b.hi}; }
SIMT_FN f2 sub2(f2 a, f2 b) { return f2{a.lo - b.lo, a.hi - b.hi}; }
SIMT_FN f2 mul2(f2 a, f2 b) { return f2{a.lo * b.lo, a.hi * b.hi}; }
SIMT_FN f2 fma2(f2 a, f2 b, f2 c) { return f2{std::fmaf(a.lo, b.lo, c.lo), std::fmaf(a.hi, b.hi, c.hi)}; }
// split warps (k_fused_duo_split) exist on the device only: the emulator runs one warp at a time
SIMT_FN void cta_sync() { emu_fail("cta_sync: the split-warp variant is not emulated"); }
SIMT_FN void sts2(f2* p, f2 v) { *p = v; }
SIMT_FN f2 lds2(const f2* p) { return *p; }
// One index step of the resampler for the TWO streams of a lane that share their phase: numerator += from (mod den); on a
// carry the right taps become the left ones and the next ring frame of either stream is fetched (ring B lies `B_OFF` words
// behind ring A).  The numerator is one scalar for both (the packed multiply takes it as a broadcast operand).
template <int B_OFF>
SIMT_FN void lerp_advance2(float& nf, f2& x0, f2& x1, sptr& p, float from_f, float den) {
    const float t = nf + from_f;
    if (t >= den) {
        nf = t - den;
        x0 = x1, x1 = f2{p[0], p[B_OFF]};
        p += 1;
    } else {
        nf = t;
    }
}
// 16-byte asynchronous copy global -> shared (cp.async.cg.shared.global): lands at the matching cp_wait.
SIMT_FN void cp16(float* smem_dst, const float* gsrc) {
    if (((uintptr_t)smem_dst & 15) || ((uintptr_t)gsrc & 15)) emu_fail("cp16: operands must be 16-byte aligned");
    if (!cur().w->ok_read(gsrc, 16)) emu_fail("cp16 source outside the registered input rows");
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (int i = 0; i < 4; i++) smem_dst[i] = nan;   // the engine may overwrite the slot any time from now on
    cur().open.push_back({smem_dst, gsrc});
}
SIMT_FN void cp_commit() {
    cur().groups.push_back(std::move(cur().open));
    cur().open.clear();
}
template <int N>
SIMT_FN void cp_wait() {
    while ((int)cur().groups.size() > N) {
        for (auto& c : cur().groups.front()) std::memcpy(c.dst, c.src, 16);
        cur().groups.pop_front();
    }
}
#else
// ------------------------------------------------------------------------------------------ device
SIMT_FN uint32_t lane() { return threadIdx.x & 31u; }
SIMT_FN void emu_count(int, uint64_t) {}
SIMT_FN void emu_assert(bool, const char*) {}
SIMT_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
SIMT_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
SIMT_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
SIMT_FN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// the same division for paths that are taken once in a blue moon: one copy per kernel instead of one per call site
static __device__ __noinline__ float fdiv_cold(float a, float b) { return __fdiv_rn(a, b); }
SIMT_FN float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SIMT_FN float u2f(uint32_t v) { return __uint2float_rn(v); }
SIMT_FN float ldg(const float* p) { return __ldg(p); }
SIMT_FN void syncwarp() { __syncwarp(); }
SIMT_FN float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
SIMT_FN uint32_t shfl_idx(uint32_t v, uint32_t src) { return __shfl_sync(0xffffffffu, v, (int)src); }
SIMT_FN uint64_t shfl_idx64(uint64_t v, uint32_t src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, (int)src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), (int)src);
    return ((uint64_t)hi << 32) | lo;
}
SIMT_FN uint32_t reduce_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
SIMT_FN uint64_t reduce_min64(uint64_t v) {
    // high words first, then the low words of the lanes that hold the minimal high word
    const uint32_t hi = __reduce_min_sync(0xffffffffu, (uint32_t)(v >> 32));
    const uint32_t lo = __reduce_min_sync(0xffffffffu, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0xffffffffu);
    return ((uint64_t)hi << 32) | lo;
}
SIMT_FN uint64_t reduce_max64(uint64_t v) {
    const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(v >> 32));
    const uint32_t lo = __reduce_max_sync(0xffffffffu, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0u);
    return ((uint64_t)hi << 32) | lo;
}
using sptr = uint32_t;   // address in the shared window
SIMT_FN sptr sptr_of(const float* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SIMT_FN sptr sptr_add(sptr p, int words) { return p + (uint32_t)(words * 4); }
SIMT_FN bool sptr_ge(sptr a, sptr b) { return a >= b; }
SIMT_FN float lds(sptr p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(p) : "memory");
    return v;
}
// FADD, FSETP, then predicated instructions: numerator wrap, tap move(s), tap load (one LDS / LDS.64), cursor increment.
// Written in PTX because nvcc otherwise keeps a second cursor and copies the taps through temporaries (3 extra moves
// per step).
template <int C, bool PRE = false>
SIMT_FN void lerp_advance(float& nf, float (&x0)[C], float (&x1)[C], sptr& p, float from_f, float den, float gpre = 1.0f) {
    static_assert(C == 1 || C == 2, "mono or stereo");
    if constexpr (C == 1 && !PRE) {
        asm volatile(
            "{\n"
            ".reg .pred c;\n"
            ".reg .f32 t;\n"
            "add.rn.f32 t, %0, %4;\n"
            "setp.ge.f32 c, t, %5;\n"
            "@c sub.rn.f32 t, t, %5;\n"
            "mov.f32 %0, t;\n"
            "@c mov.f32 %1, %2;\n"
            "@c ld.shared.f32 %2, [%3];\n"
            "@c add.u32 %3, %3, 4;\n"
            "}\n"
            : "+f"(nf), "+f"(x0[0]), "+f"(x1[0]), "+r"(p)
            : "f"(from_f), "f"(den)
            : "memory");
    } else if constexpr (C == 1 && PRE) {
        asm volatile(
            "{\n"
            ".reg .pred c;\n"
            ".reg .f32 t;\n"
            "add.rn.f32 t, %0, %4;\n"
            "setp.ge.f32 c, t, %5;\n"
            "@c sub.rn.f32 t, t, %5;\n"
            "mov.f32 %0, t;\n"
            "@c mov.f32 %1, %2;\n"
            "@c ld.shared.f32 %2, [%3];\n"
            "@c mul.rn.f32 %2, %2, %6;\n"          // the gain in front of the conversion, applied once to the fetched frame
            "@c add.u32 %3, %3, 4;\n"
            "}\n"
            : "+f"(nf), "+f"(x0[0]), "+f"(x1[0]), "+r"(p)
            : "f"(from_f), "f"(den), "f"(gpre)
            : "memory");
    } else if constexpr (C == 2 && !PRE) {
        asm volatile(
            "{\n"
            ".reg .pred c;\n"
            ".reg .f32 t;\n"
            "add.rn.f32 t, %0, %6;\n"
            "setp.ge.f32 c, t, %7;\n"
            "@c sub.rn.f32 t, t, %7;\n"
            "mov.f32 %0, t;\n"
            "@c mov.f32 %1, %3;\n"
            "@c mov.f32 %2, %4;\n"
            "@c ld.shared.v2.f32 {%3, %4}, [%5];\n"
            "@c add.u32 %5, %5, 8;\n"
            "}\n"
            : "+f"(nf), "+f"(x0[0]), "+f"(x0[1]), "+f"(x1[0]), "+f"(x1[1]), "+r"(p)
            : "f"(from_f), "f"(den)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred c;\n"
            ".reg .f32 t;\n"
            "add.rn.f32 t, %0, %6;\n"
            "setp.ge.f32 c, t, %7;\n"
            "@c sub.rn.f32 t, t, %7;\n"
            "mov.f32 %0, t;\n"
            "@c mov.f32 %1, %3;\n"
            "@c mov.f32 %2, %4;\n"
            "@c ld.shared.v2.f32 {%3, %4}, [%5];\n"
            "@c mul.rn.f32 %3, %3, %8;\n"
            "@c mul.rn.f32 %4, %4, %8;\n"
            "@c add.u32 %5, %5, 8;\n"
            "}\n"
            : "+f"(nf), "+f"(x0[0]), "+f"(x0[1]), "+f"(x1[0]), "+f"(x1[1]), "+r"(p)
            : "f"(from_f), "f"(den), "f"(gpre)
            : "memory");
    }
}
// ---- packed pairs: add/mul/fma.rn.f32x2 (SASS FADD2 / FMUL2 / FFMA2): one issue slot for two lanes' worth of arithmetic.
// NOTE ptxas (12.9) contracts a mul.rn.f32x2 whose only use is an add/sub.rn.f32x2 into FFMA2 (it never does that to the
// scalar forms): code that needs the rounded product must not feed it to add2 / sub2 directly -- the kernels use
// fma2(p, +-1, t) with the constant in a register, which keeps both roundings (tools/sass_loop_count.py checks the counts).
using f2 = unsigned long long;
SIMT_FN f2 pack2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
SIMT_FN float lo2(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
SIMT_FN float hi2(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
SIMT_FN f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMT_FN f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMT_FN f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMT_FN f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
SIMT_FN void cta_sync() { __syncthreads(); }
SIMT_FN void sts2(f2* p, f2 v) { *p = v; }    // STS.64 / LDS.64 of a packed pair
SIMT_FN f2 lds2(const f2* p) { return *p; }
// FADD, FSETP, then predicated: numerator wrap, tap moves, two tap loads (ring A at p, ring B B_OFF words behind it: an
// immediate offset), cursor increment.  The taps are handed over as separate halves so that the loads land in place.
template <int B_OFF>
SIMT_FN void lerp_advance2(float& nf, f2& x0, f2& x1, sptr& p, float from_f, float den) {
    asm volatile(
        "{\n"
        ".reg .pred c;\n"
        ".reg .f32 t, a, b, u, v;\n"
        "add.rn.f32 t, %0, %4;\n"
        "setp.ge.f32 c, t, %5;\n"
        "@c sub.rn.f32 t, t, %5;\n"
        "mov.f32 %0, t;\n"
        "mov.b64 {u, v}, %1;\n"
        "mov.b64 {a, b}, %2;\n"
        "@c mov.f32 u, a;\n"
        "@c mov.f32 v, b;\n"
        "@c ld.shared.f32 a, [%3];\n"
        "@c ld.shared.f32 b, [%3 + %6];\n"
        "mov.b64 %1, {u, v};\n"
        "mov.b64 %2, {a, b};\n"
        "@c add.u32 %3, %3, 4;\n"
        "}\n"
        : "+f"(nf), "+l"(x0), "+l"(x1), "+r"(p)
        : "f"(from_f), "f"(den), "n"(B_OFF * 4)
        : "memory");
}
SIMT_FN bool in_exact_quotient_class(float m) {
    const uint32_t u = __float_as_uint(m) & 0x7fffffffu;
    return u == 0u || (u - 0x0d800000u) < 0x64000000u;
}
SIMT_FN void cp16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
SIMT_FN void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SIMT_FN void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

}  // namespace simt
