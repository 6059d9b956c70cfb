// rb_dsp.cuh — the arithmetic of rodio's per-sample adapters as device functions.
// Every float op is an explicitly rounded intrinsic (__fmul_rn/__fadd_rn/...), which nvcc never
// contracts into an FMA: rustc does not fuse a*b+c, and bit-parity with the reference depends on it.
// (The translation unit is additionally compiled with --fmad=false.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rb_internal.h"

namespace rbd {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divf(float a, float b) { return __fdiv_rn(a, b); }

// src/math.rs:24-26 : first + (second - first) * numerator as f32 / denominator as f32
__device__ __forceinline__ float lerp(float first, float second, uint32_t numerator, uint32_t denominator) {
    float d = sub(second, first);
    float m = mul(d, __uint2float_rn(numerator));
    float q = divf(m, __uint2float_rn(denominator));
    return add(first, q);
}
// Same with the two u32->f32 casts hoisted by the caller.
__device__ __forceinline__ float lerp_f(float first, float second, float num_f, float den_f) {
    return add(first, divf(mul(sub(second, first), num_f), den_f));
}

// src/source/blt.rs:558-560 : b0*x + b1*x1 + b2*x2 - a1*y1 - a2*y2, strictly left to right
__device__ __forceinline__ float biquad(float b0, float b1, float b2, float a1, float a2, float x, float x1, float x2,
                                        float y1, float y2) {
    float r = mul(b0, x);
    r = add(r, mul(b1, x1));
    r = add(r, mul(b2, x2));
    r = sub(r, mul(a1, y1));
    r = sub(r, mul(a2, y2));
    return r;
}
// Feed-forward half (no dependence on y): ((b0*x + b1*x1) + b2*x2)
__device__ __forceinline__ float biquad_ff(float b0, float b1, float b2, float x, float x1, float x2) {
    return add(add(mul(b0, x), mul(b1, x1)), mul(b2, x2));
}
// Recurrent half: (t - a1*y1) - a2*y2
__device__ __forceinline__ float biquad_fb(float a1, float a2, float t, float y1, float y2) {
    return sub(sub(t, mul(a1, y1)), mul(a2, y2));
}
// The same value with the two subtractions written as fma(p, -1, t) = RN(t - p): p * -1 is exact, so each step
// still rounds exactly once, but all three dependent operations now run on the FMA pipe.  An FMUL feeding an FADD
// crosses pipes (5 instead of 4 cycles each way): 18 cycles per sample for the serial chain instead of 12.
__device__ __forceinline__ float biquad_fb_chain(float a1, float a2, float t, float y1, float y2, float neg1) {
    return __fmaf_rn(mul(a2, y2), neg1, __fmaf_rn(mul(a1, y1), neg1, t));
}

// src/conversions/channels.rs:57-85 as a pure map: which input channel feeds output channel j
// (-1 = literal 0.0).
__device__ __forceinline__ int chan_map(uint32_t j, uint32_t c_in) {
    if (j < c_in) return (int)j;
    if (j == 1 && c_in == 1) return 0;
    return -1;
}

// Closed form of UniformSourceIterator (src/source/uniform.rs:50-97) =
// ChannelCountConverter(SampleRateConverter(Take(input))) re-bootstrapped per span chunk
// (src/conversions/sample_rate.rs:131-201, src/conversions/channels.rs:57-85).
// For output sample `o`: where the value comes from.  Exact integer math (the planner rejects
// from*to >= 2^32, so the reference's u32 products never wrap either).
struct UniformTap {
    uint64_t i0;     // flat input index of the left sample
    uint32_t num;    // interpolation numerator (0..to-1)
    uint32_t kind;   // 0: literal 0.0 (zero-filled channel), 1: x[i0] raw, 2: lerp(x[i0], x[i0 + c_in], num, to)
};
__device__ __forceinline__ UniformTap uniform_tap(const rb_uniform_params& u, uint32_t c_in, uint32_t c_out, uint64_t o) {
    UniformTap t;
    t.i0 = 0, t.num = 0, t.kind = 0;
    uint64_t base = 0;
    const rb_uniform_seg* seg = &u.tail;
    if (u.chunk_samples) {
        uint64_t k = u.full.out_samples ? o / u.full.out_samples : u.n_full_chunks;
        if (k >= u.n_full_chunks) k = u.n_full_chunks;
        else seg = &u.full;
        o -= k * u.full.out_samples;
        base = k * u.chunk_samples;
    }
    uint64_t fo = o / c_out;
    uint32_t j = (uint32_t)(o - fo * c_out);
    int c = chan_map(j, c_in);
    if (c < 0) return t;
    uint64_t q = fo * c_in + (uint32_t)c;            // position in the converter's flat output
    if (u.from == u.to) {
        t.i0 = base + q, t.kind = 1;
        return t;
    }
    uint64_t n;
    uint32_t ch;
    uint64_t full = seg->full_out_frames * c_in;
    if (q < full) {
        n = q / c_in;
        ch = (uint32_t)(q - n * c_in);
    } else {
        uint64_t r = q - full;
        uint64_t k = r / seg->p;
        n = seg->full_out_frames + k;
        ch = (uint32_t)(r - k * seg->p);
    }
    uint64_t frames_ch = seg->L + (ch < seg->p ? 1u : 0u);
    uint64_t prod = n * (uint64_t)u.from;
    uint64_t i = prod / u.to;
    t.num = (uint32_t)(prod - i * u.to);
    t.i0 = base + i * c_in + ch;
    t.kind = (i + 1 < frames_ch) ? 2u : 1u;
    return t;
}
__device__ __forceinline__ float uniform_sample(const float* __restrict__ x, const rb_uniform_params& u, uint32_t c_in,
                                                uint32_t c_out, float den_f, uint64_t o) {
    UniformTap t = uniform_tap(u, c_in, c_out, o);
    if (t.kind == 0) return 0.0f;
    float a = x[t.i0];
    if (t.kind == 1) return a;
    return lerp_f(a, x[t.i0 + c_in], __uint2float_rn(t.num), den_f);
}

// src/source/limit.rs:854-873 gain computer (dB of reduction wanted for this sample)
__device__ __forceinline__ float limiter_db(float sample, float threshold, float knee_width, float inv_knee_8) {
    const float LOG10_2 = 0.301029995663981195213738894724493027f;
    const float MIN_POSITIVE = 1.17549435e-38f;
    float lin = add(fabsf(sample), MIN_POSITIVE);
    float db = mul(mul(log2f(lin), LOG10_2), 20.0f);     // src/math.rs:87-90
    float bias_db = sub(db, threshold);
    float knee_boundary_db = mul(bias_db, 2.0f);
    if (knee_boundary_db < -knee_width) return 0.0f;
    if (fabsf(knee_boundary_db) <= knee_width) {
        float x = add(knee_boundary_db, knee_width);
        return mul(mul(x, x), inv_knee_8);
    }
    return bias_db;
}
// src/math.rs:52-56 : 2^(dB * 0.05 * LOG2_10)
__device__ __forceinline__ float db_to_linear(float decibels) {
    const float LOG2_10 = 3.32192809488736234787031942948939018f;
    return exp2f(mul(mul(decibels, 0.05f), LOG2_10));
}

// Rust float->int `as` casts: truncate, saturate, NaN -> 0 (what dasp_sample's conversions rely on).
__device__ __forceinline__ int32_t f32_as_i32(float v) { return __float2int_rz(v); }   // cvt.rzi.s32.f32 saturates, NaN->0
__device__ __forceinline__ int16_t f32_as_i16(float v) {
    int32_t i = __float2int_rz(v);
    i = max(-32768, min(32767, i));
    return (int16_t)i;
}
__device__ __forceinline__ int8_t f32_as_i8(float v) {
    int32_t i = __float2int_rz(v);
    i = max(-128, min(127, i));
    return (int8_t)i;
}

// dasp_sample 0.11.0 conv: X -> f32 (call site src/conversions/sample.rs:42-44)
__device__ __forceinline__ float load_as_f32(const void* p, uint32_t fmt, uint64_t i) {
    switch (fmt) {
        case RB_FMT_F32: return ((const float*)p)[i];
        case RB_FMT_I16: return divf((float)((const int16_t*)p)[i], 32768.0f);
        case RB_FMT_U16: return divf((float)((int32_t)((const uint16_t*)p)[i] - 32768), 32768.0f);
        case RB_FMT_I8: return divf((float)((const int8_t*)p)[i], 128.0f);
        case RB_FMT_U8: return divf((float)((int32_t)((const uint8_t*)p)[i] - 128), 128.0f);
        case RB_FMT_I32: return divf(__int2float_rn(((const int32_t*)p)[i]), 2147483648.0f);
        default: return divf(__int2float_rn(((const int32_t*)p)[i]), 8388608.0f);  // RB_FMT_I24_IN_I32
    }
}
// f32 -> X
__device__ __forceinline__ void store_from_f32(void* p, uint32_t fmt, uint64_t i, float s) {
    switch (fmt) {
        case RB_FMT_F32: ((float*)p)[i] = s; break;
        case RB_FMT_I16: ((int16_t*)p)[i] = f32_as_i16(mul(s, 32768.0f)); break;
        case RB_FMT_U16: ((uint16_t*)p)[i] = (uint16_t)((int32_t)f32_as_i16(mul(s, 32768.0f)) + 32768); break;
        case RB_FMT_I8: ((int8_t*)p)[i] = f32_as_i8(mul(s, 128.0f)); break;
        case RB_FMT_U8: ((uint8_t*)p)[i] = (uint8_t)((int32_t)f32_as_i8(mul(s, 128.0f)) + 128); break;
        case RB_FMT_I32: ((int32_t*)p)[i] = f32_as_i32(mul(s, 2147483648.0f)); break;
        default: ((int32_t*)p)[i] = f32_as_i32(mul(s, 8388608.0f)); break;
    }
}

}  // namespace rbd
