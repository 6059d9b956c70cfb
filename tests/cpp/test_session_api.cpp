// Streaming through the C++ mirror (include/rodio_b200.hpp, rodio::mixer::LiveMixer over rb_session_*): sources pushed in
// uneven blocks and pulled sample by sample give exactly the samples of the whole-stream mixer (rodio::mixer::mixer with
// the same chains), whatever the split.  Runs on the GPU box (tests/test_cpp_mirror.py, gated like the other session tests).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "rodio_b200.hpp"

using namespace rodio;

static int failures = 0;
#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            failures++;                                                 \
        }                                                               \
    } while (0)

int main() {
    std::mt19937 g(7);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    for (uint16_t ch : {(uint16_t)1, (uint16_t)2}) {
        const size_t n_src = 5;
        std::vector<std::vector<Sample>> pcm(n_src);
        std::vector<Source> chains, whole;
        for (size_t i = 0; i < n_src; i++) {
            pcm[i].resize((3000 + 517 * i) * ch);
            for (auto& v : pcm[i]) v = U(g);
            chains.push_back(TestSource({}, ch, 44100).uniform(ch, 48000).low_pass(300).amplify(0.9f));
            whole.push_back(TestSource(pcm[i], ch, 44100).uniform(ch, 48000).low_pass(300).amplify(0.9f));
        }
        // the whole-stream render through the lane kernel is the reference for the split (same summation tree)
        std::vector<rb_stream_desc> descs;
        for (auto& w : whole) descs.push_back(w.desc(0));
        rb_batch* b = nullptr;
        check(rb_batch_create(Context::get(), ch, 48000, descs.data(), descs.size(), RB_FUSED_LANES, &b), "rb_batch_create");
        int family = -1;
        check(rb_batch_kernel_family(b, &family), "rb_batch_kernel_family");
        CHECK(family == 2);
        for (size_t i = 0; i < n_src; i++) check(rb_batch_upload(b, i, pcm[i].data(), pcm[i].size()), "rb_batch_upload");
        uint64_t n = 0, w = 0;
        check(rb_batch_mix_len(b, &n), "rb_batch_mix_len");
        std::vector<Sample> want(n);
        check(rb_batch_render_mix(b, want.data(), n, &w), "rb_batch_render_mix");
        rb_batch_destroy(b);

        mixer::LiveMixer live(chains, ch, 48000, 4096, 500);
        std::vector<size_t> at(n_src, 0);
        std::vector<Sample> got;
        while (!live.ended()) {
            for (size_t i = 0; i < n_src; i++) {
                const size_t left = pcm[i].size() / ch - at[i];
                if (!left) continue;
                const size_t k = std::min<size_t>(left, 1 + g() % 700);
                std::vector<Sample> blk(pcm[i].begin() + at[i] * ch, pcm[i].begin() + (at[i] + k) * ch);
                at[i] += k;
                live.push(i, blk, at[i] * ch == pcm[i].size());
            }
            while (auto s = live.next()) got.push_back(*s);
        }
        CHECK(got.size() == want.size());
        size_t bad = 0;
        for (size_t i = 0; i < got.size() && i < want.size(); i++) bad += std::memcmp(&got[i], &want[i], 4) != 0;
        CHECK(bad == 0);
        std::printf("%u ch: %zu mixer samples streamed, %zu differ from the whole-stream render\n", (unsigned)ch, got.size(), bad);
    }
    if (failures) return 1;
    std::printf("all session API tests passed\n");
    return 0;
}
