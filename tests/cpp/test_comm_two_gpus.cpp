// The cross-shard mixer sum through the C ABI alone (include/rodio_b200.h rb_comm_*): one process drives N GPUs (default 2),
// each renders the partial mix of its shard of a cfg3-shaped batch, rb_batch_render_mix_allreduce sums the shards -- by
// k_mix_exchange over NVLink peer memory, and once more by NCCL (RB_COMM_NCCL_ONLY=1) -- and every GPU must then hold (within the
// fused kernels' tolerance, 1e-5 * peak) the mix a single GPU renders from all sources.  On the peer-memory path the result is also
// the shards added in RANK ORDER from +0.0, bit for bit, and identical on every GPU.
// Exit code 0 = passed, 77 = fewer GPUs than ranks (skipped).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rodio_b200.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        rb_status s_ = (call);                                                             \
        if (s_ != RB_OK) {                                                                 \
            std::fprintf(stderr, "%s failed: %s (%s)\n", #call, rb_status_string(s_), rb_last_error()); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static rb_effect fx(uint32_t kind, uint32_t u0, uint32_t u1, float f0) {
    rb_effect e{};
    e.kind = kind, e.u32[0] = u0, e.u32[1] = u1, e.f32[0] = f0;
    return e;
}

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 2;
    const size_t S = 300, frames = 4000;
    std::vector<rb_context*> ctxs(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; g++)
        if (rb_context_create(g, &ctxs[g]) != RB_OK) {
            std::printf("skipped: %d GPUs wanted, context %d failed (%s)\n", n_gpus, g, rb_last_error());
            return 77;
        }
    // sources: 44.1 kHz mono noise -> uniform(1, 48000) -> low_pass(300) -> amplify(0.9)
    std::vector<std::vector<float>> pcm(S, std::vector<float>(frames));
    uint32_t lcg = 12345u;
    for (auto& p : pcm)
        for (auto& v : p) lcg = lcg * 1664525u + 1013904223u, v = (float)((int32_t)lcg) / 2147483648.0f;
    rb_effect chain[3] = {fx(RB_FX_UNIFORM, 1, 48000, 0.f), fx(RB_FX_LOW_PASS, 300, 0, 0.5f), fx(RB_FX_AMPLIFY, 0, 0, 0.9f)};
    auto desc = [&](size_t) {
        rb_stream_desc d{};
        d.sample_rate = 44100, d.channels = 1, d.format = RB_FMT_F32, d.n_samples = frames, d.span_len = 0, d.n_effects = 3, d.effects = chain, d.mix_start = 0;
        return d;
    };
    auto build = [&](rb_context* ctx, size_t lo, size_t hi, rb_batch** out) -> rb_status {
        std::vector<rb_stream_desc> ds;
        for (size_t i = lo; i < hi; i++) ds.push_back(desc(i));
        rb_status s = rb_batch_create(ctx, 1, 48000, ds.data(), ds.size(), 0, out);
        for (size_t i = lo; i < hi && s == RB_OK; i++) s = rb_batch_upload(*out, i - lo, pcm[i].data(), frames);
        return s;
    };
    // the whole batch on GPU 0
    rb_batch* whole = nullptr;
    CHECK(build(ctxs[0], 0, S, &whole));
    uint64_t mix_len = 0, got = 0;
    CHECK(rb_batch_mix_len(whole, &mix_len));
    std::vector<float> ref(mix_len);
    CHECK(rb_batch_render_mix(whole, ref.data(), mix_len, &got));
    // sharded over the GPUs
    std::vector<rb_batch*> shard(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; g++) CHECK(build(ctxs[g], S * g / n_gpus, S * (g + 1) / n_gpus, &shard[g]));
    // every shard's own mix, added in rank order from +0.0 on the host: what the peer-memory exchange must give, bit for bit
    std::vector<float> ordered(mix_len, 0.0f);
    for (int g = 0; g < n_gpus; g++) {
        std::vector<float> own(mix_len);
        CHECK(rb_batch_render_mix(shard[g], own.data(), mix_len, &got));
        for (uint64_t i = 0; i < mix_len; i++) ordered[i] = ordered[i] + own[i];
    }
    float peak = 0.f;
    for (float v : ref) peak = std::fmax(peak, std::fabs(v));
    for (int pass = 0; pass < 2; pass++) {           // 0: whatever the communicator picks (peer memory on an NVLink box), 1: NCCL
        if (pass == 1) setenv("RB_COMM_NCCL_ONLY", "1", 1);
        rb_comm* comm = nullptr;
        CHECK(rb_comm_init_all(ctxs.data(), n_gpus, &comm));
        for (int rep = 0; rep < 3; rep++) CHECK(rb_batch_render_mix_allreduce(shard.data(), n_gpus, comm));
        char how[400];
        CHECK(rb_comm_transport(comm, how, sizeof how));
        std::printf("pass %d transport: %s\n", pass, how);
        const bool p2p = how[0] == 'p';
        if (pass == 1 && p2p) return std::fprintf(stderr, "RB_COMM_NCCL_ONLY=1 was ignored\n"), 1;
        std::vector<float> first;
        for (int g = 0; g < n_gpus; g++) {
            std::vector<float> out(mix_len);
            CHECK(rb_batch_read_mix(shard[g], 0, out.data(), mix_len, &got));
            if (got != mix_len) return std::fprintf(stderr, "short read on GPU %d\n", g), 1;
            float err = 0.f;
            uint64_t bits_differ = 0;
            for (uint64_t i = 0; i < mix_len; i++) {
                err = std::fmax(err, std::fabs(out[i] - ref[i]));
                bits_differ += std::memcmp(&out[i], &ordered[i], 4) != 0;
            }
            std::printf("GPU %d: max |sharded - single| = %.3e (peak %.3e), %llu samples differ from the rank-ordered sum\n", g, err, peak,
                        (unsigned long long)bits_differ);
            if (!(err <= 1e-5f * peak)) return std::fprintf(stderr, "all-reduced mix differs on GPU %d\n", g), 1;
            if (p2p && bits_differ) return std::fprintf(stderr, "peer-memory exchange: not the rank-ordered sum on GPU %d\n", g), 1;
            if (g == 0) first = out;
            else if (p2p && std::memcmp(first.data(), out.data(), mix_len * 4) != 0) return std::fprintf(stderr, "GPUs disagree\n"), 1;
        }
        CHECK(rb_comm_destroy(comm));
    }
    unsetenv("RB_COMM_NCCL_ONLY");
    for (auto* b : shard) rb_batch_destroy(b);
    rb_batch_destroy(whole);
    for (auto* c : ctxs) rb_context_destroy(c);
    std::printf("all communicator tests passed (%d GPUs)\n", n_gpus);
    return 0;
}
