// rb_p2p.h -- the cross-shard mixer sum as ONE kernel over NVLink peer memory ("one-shot all-reduce fused after the mix
// kernel, since the message is far below the bandwidth-bound regime").  Every rank owns a mailbox in its HBM that all peers can
// write (cudaIpc between processes, peer access inside one process).  k_mix_exchange on rank r
//   1. forms its shard's mix -- for the fused kernel's per-CTA partial rows it adds them up itself (k_sum_partials is not launched) --
//   2. pushes it as (value, tag) pairs into the mailbox of every peer with 8-byte stores over NVLink (posted writes, one-way latency),
//   3. polls its OWN mailbox until every peer's pair carries this render's tag and adds the shards in RANK ORDER from +0.0:
//      deterministic, identical on every rank (NCCL's all-reduce order is neither).
// No flag, no fence, no barrier: the tag is the flag, two mailbox buffers alternate (a rank that starts render e + 1 has seen every
// peer's render-e data, so every peer has finished reading render e - 1: the buffer it overwrites).  src/mixer.rs:185-198 is the sum
// being distributed; the result is within the tolerance class of the sharded sum (<= 1e-5 * peak), as with NCCL.
// Plain C++ interface: rb_api.cu (which also compiles against the mock runtime of tests/emu) only sees these declarations.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>

#include <cuda_runtime.h>

struct rb_p2p;
constexpr int RB_P2P_MAX_RANKS = 16;

// all-gather of `bytes` bytes per rank between DEVICE buffers, in rank order, asynchronous on `st` (rb_api.cu: ncclAllGather)
using rb_p2p_allgather = std::function<cudaError_t(const void* send_dev, void* recv_dev, size_t bytes, cudaStream_t st)>;

// One process per GPU: collective over the communicator's ranks.  cap = floats of mix the mailbox takes.  On failure `why` says what
// (peers on another node, IPC disabled ...) and the caller keeps NCCL.
cudaError_t rb_p2p_create_rank(int n_ranks, int rank, int device, cudaStream_t st, uint64_t cap, const rb_p2p_allgather& allgather,
                               rb_p2p** out, std::string* why);
// One process, several GPUs: out[i] for device i of `devices`.
cudaError_t rb_p2p_create_local(int n, const int* devices, const cudaStream_t* streams, uint64_t cap, rb_p2p** out, std::string* why);
uint64_t rb_p2p_capacity(const rb_p2p* p);
// d_out[0, mix_len) <- sum over ranks.  partial != nullptr: this rank's mix is the sum of n_rows rows of `pstride` floats (in row order).
cudaError_t rb_p2p_allreduce(rb_p2p* p, float* d_out, uint64_t mix_len, const float* partial, uint32_t n_rows, uint64_t pstride, cudaStream_t st);
void rb_p2p_destroy(rb_p2p* p);
