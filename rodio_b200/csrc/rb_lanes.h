// rb_lanes.h — interface between the fused planner (rb_fused.cu) and the lane-per-stream kernel (rb_lanes.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

// One stream of a batch that fits the lane kernel's shape:
//   f32 mono source -> UniformSourceIterator(1 ch, to) with reduced from < to -> [biquad] -> [one gain] -> mixer(1 ch)
struct rb_lanes_stream {
    const float* in;       // device, 16-byte aligned, 16-byte tail pad
    uint64_t n_frames;
    uint64_t out_len;
    uint64_t mix_start;
    float b0, b1, b2, a1, a2;
    float post;
};

struct rb_lanes_plan;

// *out stays NULL when the shape is not covered.  `d_out` is the mixer output ([mix_len] f32).
cudaError_t rb_lanes_try_create(const rb_lanes_stream* streams, size_t n_streams, uint32_t from, uint32_t to, bool has_biquad,
                                bool has_post, float* d_out, uint64_t mix_len, int sm_count, cudaStream_t st,
                                rb_lanes_plan** out);
// Inputs were (re)written: classify them again before the next render.
void rb_lanes_inputs_changed(rb_lanes_plan* p);
cudaError_t rb_lanes_run(rb_lanes_plan* p, cudaStream_t st);
uint32_t rb_lanes_launch_count(const rb_lanes_plan* p);
void rb_lanes_destroy(rb_lanes_plan* p);
