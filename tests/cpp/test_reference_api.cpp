// The reference's own unit tests for this path, restated against the C++ mirror (include/rodio_b200.hpp):
//   src/mixer.rs:208-341, src/conversions/channels.rs:114-143, src/conversions/sample_rate.rs:356-387,
//   src/source/channel_volume.rs:135-166 (via Spatial-free ChannelVolume is Python-only), player.rs:454-470.
// Runs on the GPU through the C ABI; exits non-zero on the first failed expectation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rodio_b200.hpp"

using namespace rodio;
using V = std::vector<float>;
using O = std::optional<float>;

static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::fprintf(stderr, "FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                     \
        }                                                                   \
    } while (0)

static bool eq(O a, O b) { return a.has_value() == b.has_value() && (!a || *a == *b); }

static void mixer_basic() {   // mixer.rs:214-237
    auto [tx, rx] = mixer::mixer(1, 48000);
    tx.add(SamplesBuffer(1, 48000, {10.0f, -10.0f, 10.0f, -10.0f}));
    tx.add(SamplesBuffer(1, 48000, {5.0f, 5.0f, 5.0f, 5.0f}));
    EXPECT(rx.channels() == 1 && rx.sample_rate() == 48000);
    EXPECT(eq(rx.next(), 15.0f)); EXPECT(eq(rx.next(), -5.0f)); EXPECT(eq(rx.next(), 15.0f));
    EXPECT(eq(rx.next(), -5.0f)); EXPECT(eq(rx.next(), std::nullopt));
}
static void mixer_channels_conv() {   // mixer.rs:240-267
    auto [tx, rx] = mixer::mixer(2, 48000);
    tx.add(SamplesBuffer(1, 48000, {10.0f, -10.0f, 10.0f, -10.0f}));
    tx.add(SamplesBuffer(1, 48000, {5.0f, 5.0f, 5.0f, 5.0f}));
    float want[] = {15, 15, -5, -5, 15, 15, -5, -5};
    for (float w : want) EXPECT(eq(rx.next(), w));
    EXPECT(eq(rx.next(), std::nullopt));
}
static void mixer_rate_conv() {   // mixer.rs:270-296
    auto [tx, rx] = mixer::mixer(1, 96000);
    tx.add(SamplesBuffer(1, 48000, {10.0f, -10.0f, 10.0f, -10.0f}));
    tx.add(SamplesBuffer(1, 48000, {5.0f, 5.0f, 5.0f, 5.0f}));
    float want[] = {15, 5, -5, 5, 15, 5, -5};
    for (float w : want) EXPECT(eq(rx.next(), w));
    EXPECT(eq(rx.next(), std::nullopt));
}
static void mixer_start_afterwards() {   // mixer.rs:299-328
    auto [tx, rx] = mixer::mixer(1, 48000);
    tx.add(SamplesBuffer(1, 48000, {10.0f, -10.0f, 10.0f, -10.0f}));
    EXPECT(eq(rx.next(), 10.0f)); EXPECT(eq(rx.next(), -10.0f));
    tx.add(SamplesBuffer(1, 48000, {5.0f, 5.0f, 6.0f, 6.0f, 7.0f, 7.0f, 7.0f}));
    EXPECT(eq(rx.next(), 15.0f)); EXPECT(eq(rx.next(), -5.0f)); EXPECT(eq(rx.next(), 6.0f)); EXPECT(eq(rx.next(), 6.0f));
    tx.add(SamplesBuffer(1, 48000, {2.0f}));
    EXPECT(eq(rx.next(), 9.0f)); EXPECT(eq(rx.next(), 7.0f)); EXPECT(eq(rx.next(), 7.0f));
    EXPECT(eq(rx.next(), std::nullopt));
}
static void mixer_phase() {   // mixer.rs:331-341
    auto [tx, rx] = mixer::mixer(2, 48000);
    tx.add(SamplesBuffer(2, 48000, {10.0f, -10.0f, 10.0f, -10.0f}));
    EXPECT(eq(rx.next(), 10.0f));
    tx.add(SamplesBuffer(2, 48000, {5.0f, -5.0f, 6.0f, -6.0f}));
    EXPECT(eq(rx.next(), -10.0f));   // not yet mixed (out of phase)
    EXPECT(eq(rx.next(), 15.0f));    // mixing starts
}
static void channels() {   // channels.rs:114-143
    using conversions::ChannelCountConverter;
    EXPECT((ChannelCountConverter({1, 2, 3, 4, 5, 6}, 3, 2) == V{1, 2, 4, 5}));
    EXPECT((ChannelCountConverter({1, 2, 3, 4, 5, 6, 7, 8}, 4, 1) == V{1, 5}));
    EXPECT((ChannelCountConverter({1, 2, 3, 4}, 1, 2) == V{1, 1, 2, 2, 3, 3, 4, 4}));
    EXPECT((ChannelCountConverter({1, 2}, 1, 4) == V{1, 1, 0, 0, 2, 2, 0, 0}));
    EXPECT((ChannelCountConverter({1, 2, 3, 4}, 2, 4) == V{1, 2, 0, 0, 3, 4, 0, 0}));
}
static V trunc(V v) { for (float& x : v) x = std::trunc(x); return v; }
static void sample_rate() {   // sample_rate.rs:356-387
    using conversions::SampleRateConverter;
    EXPECT((trunc(SampleRateConverter({2, 16, 4, 18, 6, 20, 8, 22}, 2000, 3000, 2)) == V{2, 16, 3, 17, 4, 18, 6, 20, 7, 21, 8, 22}));
    EXPECT((trunc(SampleRateConverter({1, 14}, 1000, 7000, 1)) == V{1, 2, 4, 6, 8, 10, 12, 14}));
    V in;
    for (int i = 0; i < 17; i++) in.push_back((float)i);
    EXPECT((SampleRateConverter(in, 12000, 2400, 1) == V{0, 5, 10, 15}));
}
static void amplify_is_volume() {   // player.rs:454-470
    V x{0.1f, -0.4f, 0.7f, 1.0f};
    V y = SamplesBuffer(1, 44100, x).amplify(0.5f).collect();
    EXPECT(y.size() == 4);
    for (size_t i = 0; i < y.size(); i++) EXPECT(y[i] == x[i] * 0.5f);
}
static void errors() {
    bool threw = false;
    try { mixer::mixer(0, 48000); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw);
    auto [tx, rx] = mixer::mixer(1, 48000);
    threw = false;
    try { rx.try_seek(Duration(0)); } catch (const Error& e) { threw = e.status == RB_ERR_NOT_SUPPORTED_SEEK; }
    EXPECT(threw);
    EXPECT(SamplesBuffer(2, 44100, {0, 0}).speed(0.9f).sample_rate() == 39690);   // speed.rs:130-133
}

static V one_to(int n) {   // crossfade.rs:40-43 dummy_source
    V v;
    for (int i = 1; i <= n; i++) v.push_back((float)i);
    return v;
}
static void crossfade_with_self() {   // crossfade.rs:45-63
    const Duration d = std::chrono::seconds(5) + std::chrono::nanoseconds(1);
    V y = SamplesBuffer(1, 1, one_to(10)).take_crossfade_with(SamplesBuffer(1, 1, one_to(10)), d).collect();
    EXPECT(y.size() == 5);
    for (size_t i = 0; i < y.size() && i < 5; i++) EXPECT(std::fabs(y[i] - (float)(i + 1)) < 1e-6f);
}
static void crossfade_against_silence() {   // crossfade.rs:65-80 (source2 = Zero: endless silence)
    const Duration d = std::chrono::seconds(5) + std::chrono::nanoseconds(1);
    V y = SamplesBuffer(1, 1, one_to(10)).take_crossfade_with(TestSource(V(64, 0.0f), 1, 1), d).collect();
    const float want[5] = {1.0f, 2.0f * 0.8f, 3.0f * 0.6f, 4.0f * 0.4f, 5.0f * 0.2f};
    EXPECT(y.size() == 5);
    for (size_t i = 0; i < y.size() && i < 5; i++) EXPECT(std::fabs(y[i] - want[i]) < 1e-6f);
}
static void from_iter_basic() {   // from_iter.rs:129-157: the format changes between the two buffers
    Source seq = Source::from_iter({SamplesBuffer(1, 48000, {10.0f, -10.0f, 10.0f, -10.0f}), SamplesBuffer(2, 96000, {5.0f, 5.0f, 5.0f, 5.0f})});
    EXPECT(seq.channels() == 1 && seq.sample_rate() == 48000);
    V y = seq.collect();
    const V want{10.0f, -10.0f, 10.0f, -10.0f, 5.0f, 5.0f, 5.0f, 5.0f};
    EXPECT(y == want);
}
static void signal_generators() {   // signal_generator.rs:181-238
    V sq = Source::signal_generator(2000, 500.0f, RB_SIGNAL_SQUARE, 8).collect();
    EXPECT((sq == V{1.0f, 1.0f, -1.0f, -1.0f, 1.0f, 1.0f, -1.0f, -1.0f}));
    V saw = Source::signal_generator(200, 50.0f, RB_SIGNAL_SAWTOOTH, 7).collect();
    EXPECT((saw == V{0.0f, 0.5f, -1.0f, -0.5f, 0.0f, 0.5f, -1.0f}));
    V sine = Source::signal_generator(1000, 100.0f, RB_SIGNAL_SINE, 7).collect();
    const float want[7] = {0.0f, 0.58778524f, 0.95105654f, 0.95105654f, 0.58778524f, 0.0f, -0.58778554f};
    EXPECT(sine.size() == 7);
    for (size_t i = 0; i < sine.size() && i < 7; i++) EXPECT(std::fabs(sine[i] - want[i]) < 1e-4f);   // TEST_EPSILON
}

int main() {
    try {
        mixer_basic(); mixer_channels_conv(); mixer_rate_conv(); mixer_start_afterwards(); mixer_phase();
        channels(); sample_rate(); amplify_is_volume(); errors();
        crossfade_with_self(); crossfade_against_silence(); from_iter_basic(); signal_generators();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    if (failures) return 1;
    std::puts("all reference API tests passed");
    return 0;
}
