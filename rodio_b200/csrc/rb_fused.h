// rb_fused.h — interface between the planner (rb_api.cu) and the fused fast path (rb_fused.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "rb_internal.h"

// One mixer input as the fused planner sees it: the raw input region plus the planned node list
// (rb_node_dev embedded at offset 0 of records `node_stride` bytes apart; src/dst not yet filled).
struct rb_fused_stream {
    const void* in;
    uint32_t fmt;
    uint32_t n_nodes;
    const rb_node_dev* nodes;
    size_t node_stride;
    uint64_t out_len;
    uint64_t mix_start;
    uint64_t n_in;
    uint32_t c_in;
};

struct rb_fused_plan;

// Sets *out to a plan when every stream of the batch has a chain shape the fused kernels cover,
// else leaves it NULL (the general per-adapter path then serves the batch).
cudaError_t rb_fused_try_create(const rb_fused_stream* streams, size_t n_streams, uint16_t mixer_channels, float* d_out,
                                uint64_t mix_len, uint32_t flags, int sm_count, cudaStream_t st, rb_fused_plan** out);
cudaError_t rb_fused_run(rb_fused_plan* plan, cudaStream_t st, bool skip_final_sum = false);
// The per-CTA partial rows of a plan whose last launch is k_sum_partials (rows added in order give the mix): the cross-GPU sum can
// then take the place of that launch (rb_p2p.h).  false: the plan has another shape (large-batch kernels, effect chain, one CTA, chain).
bool rb_fused_partial_rows(const rb_fused_plan* plan, const float** partial, uint32_t* n_rows, uint64_t* pstride);
void rb_fused_destroy(rb_fused_plan* plan);
uint32_t rb_fused_launch_count(const rb_fused_plan* plan);
// The input PCM of the batch was (re)written: plans that keep per-stream facts about it refresh them at the next run.
void rb_fused_inputs_changed(rb_fused_plan* plan);
// Which kernel family serves the plan: 0 = k_fused_biquad / k_fused_nobiquad, 1 = k_fused_hot, 2 = k_fused_lanes,
// 3 = k_fused_duo, 4 = the time-parallel plan on k_fused_duo, 5 = k_fused_fx (effect chain).
int rb_fused_kind(const rb_fused_plan* plan);
// streams per partial sum of the mixer (rows per CTA; 32 for the lane kernel)
uint32_t rb_fused_mix_group(const rb_fused_plan* plan);

// ---- the effect-chain kernel (rb_fx.cu): [Spatial / ChannelVolume] -> [reverb] -> automatic_gain_control -> mix in one launch ----
struct rb_fx_plan;
// *out stays NULL when the batch has another shape (see rb_fx.cu).
cudaError_t rb_fx_try_create(const rb_fused_stream* streams, size_t n_streams, uint16_t mixer_channels, float* d_out, uint64_t mix_len,
                             uint32_t flags, cudaStream_t st, rb_fx_plan** out);
cudaError_t rb_fx_run(rb_fx_plan* plan, cudaStream_t st);
void rb_fx_destroy(rb_fx_plan* plan);
bool rb_fx_chain(const rb_fx_plan* plan);   // RB_MIX_EXACT_ORDER: the running sum handed from CTA to CTA (one launch, one sequential sum)
