// rb_p2p.cu -- see rb_p2p.h.  Product code: never includes anything under oracle/.
#include "rb_p2p.h"

#include <cstring>
#include <vector>

namespace {

struct ExchangeArgs {
    unsigned long long* peer[RB_P2P_MAX_RANKS];   // slot of THIS rank inside the current buffer of every rank's mailbox ([cap] pairs)
    const unsigned long long* recv;                // current buffer of the own mailbox: [n_ranks][cap] pairs
    uint64_t cap;
    uint32_t n_ranks, rank, tag;
};

// (value, tag) as ONE naturally aligned 64-bit access: single-copy atomic, also across NVLink
__device__ __forceinline__ void st_pair_sys(unsigned long long* p, float v, uint32_t tag) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_pair_sys(const unsigned long long* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}

// One thread per TWO timeline positions.  FUSE: the shard's value is the ordered sum of the fused kernel's per-CTA partial rows.
template <bool FUSE>
__global__ void __launch_bounds__(256) k_mix_exchange(const float* __restrict__ partial, uint32_t n_rows, uint64_t pstride,
                                                      float* __restrict__ d_out, uint64_t mix_len, ExchangeArgs a) {
    const uint64_t m = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (m >= mix_len) return;
    const bool two = m + 1 < mix_len;
    float own0, own1 = 0.0f;
    if (FUSE) {
        own0 = partial[m];
        if (two) own1 = partial[m + 1];
        for (uint32_t c = 1; c < n_rows; c++) {
            const float* row = partial + (uint64_t)c * pstride + m;
            own0 = __fadd_rn(own0, row[0]);
            if (two) own1 = __fadd_rn(own1, row[1]);
        }
    } else {
        own0 = d_out[m];
        if (two) own1 = d_out[m + 1];
    }
    // push: posted 8-byte stores over NVLink into every peer's mailbox
    for (uint32_t q = 0; q < a.n_ranks; q++) {
        if (q == a.rank) continue;
        st_pair_sys(a.peer[q] + m, own0, a.tag);
        if (two) st_pair_sys(a.peer[q] + m + 1, own1, a.tag);
    }
    // collect: the own mailbox, in rank order from +0.0
    float acc0 = 0.0f, acc1 = 0.0f;
    for (uint32_t q = 0; q < a.n_ranks; q++) {
        float v0 = own0, v1 = own1;
        if (q != a.rank) {
            const unsigned long long* src = a.recv + (uint64_t)q * a.cap + m;
            unsigned long long w0, w1 = 0;
            uint32_t spins = 0;
            while (true) {
                w0 = ld_pair_sys(src);
                if (two) w1 = ld_pair_sys(src + 1);
                if ((uint32_t)(w0 >> 32) == a.tag && (!two || (uint32_t)(w1 >> 32) == a.tag)) break;
                if (++spins > (1u << 24)) __trap();     // a peer never arrived: fail loudly instead of hanging the device
            }
            v0 = __uint_as_float((uint32_t)w0), v1 = __uint_as_float((uint32_t)w1);
        }
        acc0 = __fadd_rn(acc0, v0), acc1 = __fadd_rn(acc1, v1);
    }
    d_out[m] = acc0;
    if (two) d_out[m + 1] = acc1;
}

}  // namespace

struct rb_p2p {
    int n_ranks = 0, rank = 0, device = 0;
    uint64_t cap = 0;
    unsigned long long* mailbox = nullptr;                      // [2][n_ranks][cap] pairs, this rank's
    unsigned long long* peer_mailbox[RB_P2P_MAX_RANKS] = {};    // every rank's mailbox as mapped here (own: mailbox)
    bool opened[RB_P2P_MAX_RANKS] = {};                         // mapped through cudaIpcOpenMemHandle
    uint32_t epoch = 0;
};

uint64_t rb_p2p_capacity(const rb_p2p* p) { return p ? p->cap : 0; }

void rb_p2p_destroy(rb_p2p* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int q = 0; q < p->n_ranks; q++)
        if (p->opened[q] && p->peer_mailbox[q]) cudaIpcCloseMemHandle(p->peer_mailbox[q]);
    cudaFree(p->mailbox);
    delete p;
}

static cudaError_t alloc_mailbox(rb_p2p* p, cudaStream_t st) {
    const size_t bytes = (size_t)2 * p->n_ranks * p->cap * sizeof(unsigned long long);
    cudaError_t e = cudaMalloc(&p->mailbox, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(p->mailbox, 0, bytes, st);      // tag 0: never written
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);                      // ... before any peer learns the address
    return e;
}

cudaError_t rb_p2p_create_rank(int n_ranks, int rank, int device, cudaStream_t st, uint64_t cap, const rb_p2p_allgather& allgather,
                               rb_p2p** out, std::string* why) {
    *out = nullptr;
    if (n_ranks < 2 || n_ranks > RB_P2P_MAX_RANKS || cap == 0) {
        if (why) *why = "rank count outside 2..16";
        return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    rb_p2p* p = new rb_p2p;
    p->n_ranks = n_ranks, p->rank = rank, p->device = device, p->cap = (cap + 1) & ~1ull;
    auto fail = [&](cudaError_t err, const char* what) {
        if (why) *why = std::string(what) + ": " + cudaGetErrorString(err);
        cudaGetLastError();
        rb_p2p_destroy(p);
        return err;
    };
    e = alloc_mailbox(p, st);
    if (e != cudaSuccess) return fail(e, "mailbox allocation");
    // every rank's IPC handle, through the communicator (which doubles as the barrier behind the memset)
    struct Slot {
        cudaIpcMemHandle_t h;
        int device;
        int ok;
    };
    Slot mine{};
    mine.device = device;
    mine.ok = cudaIpcGetMemHandle(&mine.h, p->mailbox) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    Slot *d_send = nullptr, *d_recv = nullptr;
    e = cudaMalloc(&d_send, sizeof(Slot));
    if (e == cudaSuccess) e = cudaMalloc(&d_recv, sizeof(Slot) * n_ranks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_send, &mine, sizeof(Slot), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = allgather(d_send, d_recv, sizeof(Slot), st);
    std::vector<Slot> all(n_ranks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(all.data(), d_recv, sizeof(Slot) * n_ranks, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_send), cudaFree(d_recv);
    if (e != cudaSuccess) return fail(e, "handle exchange");
    // Every rank sees the same table, so every rank takes the same decision (a rank that cannot export keeps everybody on NCCL).
    bool all_ok = true;
    for (int q = 0; q < n_ranks; q++) all_ok = all_ok && all[q].ok;
    if (!all_ok) return fail(cudaErrorNotSupported, "a rank could not export its mailbox (cudaIpcGetMemHandle)");
    int mapped_ok = 1;
    cudaError_t map_err = cudaSuccess;
    for (int q = 0; q < n_ranks && mapped_ok; q++) {
        if (q == rank) {
            p->peer_mailbox[q] = p->mailbox;
            continue;
        }
        void* ptr = nullptr;
        map_err = cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess);
        if (map_err != cudaSuccess) mapped_ok = 0, cudaGetLastError();
        else p->peer_mailbox[q] = (unsigned long long*)ptr, p->opened[q] = true;
    }
    // ... and the same for the mapping: one more exchange, so that nobody pushes into a peer that stays on NCCL
    int *d_ok = nullptr, *d_oks = nullptr;
    e = cudaMalloc(&d_ok, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_oks, sizeof(int) * n_ranks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ok, &mapped_ok, sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = allgather(d_ok, d_oks, sizeof(int), st);
    std::vector<int> oks(n_ranks, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(oks.data(), d_oks, sizeof(int) * n_ranks, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_ok), cudaFree(d_oks);
    if (e != cudaSuccess) return fail(e, "mapping exchange");
    for (int q = 0; q < n_ranks; q++)
        if (!oks[q]) return fail(map_err != cudaSuccess ? map_err : cudaErrorNotSupported, "a rank could not map a peer's mailbox (cudaIpcOpenMemHandle)");
    *out = p;
    return cudaSuccess;
}

cudaError_t rb_p2p_create_local(int n, const int* devices, const cudaStream_t* streams, uint64_t cap, rb_p2p** out, std::string* why) {
    for (int i = 0; i < n; i++) out[i] = nullptr;
    if (n < 2 || n > RB_P2P_MAX_RANKS || cap == 0) {
        if (why) *why = "rank count outside 2..16";
        return cudaErrorInvalidValue;
    }
    auto fail = [&](cudaError_t err, const char* what) {
        if (why) *why = std::string(what) + ": " + cudaGetErrorString(err);
        cudaGetLastError();
        for (int i = 0; i < n; i++) rb_p2p_destroy(out[i]), out[i] = nullptr;
        return err;
    };
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (e != cudaSuccess || !can) return fail(e != cudaSuccess ? e : cudaErrorNotSupported, "no peer access between two of the GPUs");
        }
    for (int i = 0; i < n; i++) {
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e != cudaSuccess) return fail(e, "cudaSetDevice");
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(), e = cudaSuccess;
            if (e != cudaSuccess) return fail(e, "cudaDeviceEnablePeerAccess");
        }
        rb_p2p* p = new rb_p2p;
        p->n_ranks = n, p->rank = i, p->device = devices[i], p->cap = (cap + 1) & ~1ull;
        out[i] = p;
        e = alloc_mailbox(p, streams[i]);
        if (e != cudaSuccess) return fail(e, "mailbox allocation");
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i]->peer_mailbox[j] = out[j]->mailbox;
    return cudaSuccess;
}

cudaError_t rb_p2p_allreduce(rb_p2p* p, float* d_out, uint64_t mix_len, const float* partial, uint32_t n_rows, uint64_t pstride, cudaStream_t st) {
    if (!p || mix_len > p->cap) return cudaErrorInvalidValue;
    if (mix_len == 0) return cudaSuccess;
    ExchangeArgs a{};
    const uint32_t buf = p->epoch & 1u;
    const uint64_t buf_pairs = (uint64_t)p->n_ranks * p->cap;
    for (int q = 0; q < p->n_ranks; q++) a.peer[q] = p->peer_mailbox[q] + buf * buf_pairs + (uint64_t)p->rank * p->cap;
    a.recv = p->mailbox + buf * buf_pairs;
    a.cap = p->cap, a.n_ranks = (uint32_t)p->n_ranks, a.rank = (uint32_t)p->rank;
    a.tag = p->epoch % 0xFFFFFFFFu + 1u;   // never 0 (0 = never written); equal on every rank: renders are collective
    p->epoch++;
    const uint64_t threads = (mix_len + 1) / 2;
    const uint32_t blocks = (uint32_t)((threads + 255) / 256);
    if (partial) k_mix_exchange<true><<<blocks, 256, 0, st>>>(partial, n_rows, pstride, d_out, mix_len, a);
    else k_mix_exchange<false><<<blocks, 256, 0, st>>>(nullptr, 0, 0, d_out, mix_len, a);
    return cudaGetLastError();
}
