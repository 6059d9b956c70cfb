// rodio_oracle.hpp — CPU restatement of rodio's per-sample DSP path (TEST INFRASTRUCTURE).
//
// This file is the parity oracle for rodio_b200.  It is NOT product code: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may build,
// load or call it.  The product library (rodio_b200/csrc) never includes this header.
//
// The reference (RustAudio/rodio @ 1f927962) is Rust; no Rust toolchain exists in the build
// image, so the reference cannot be compiled or run (oracle/_ref does not exist).  Every class
// below restates one reference iterator with the SAME control flow (pull model, one next() per
// sample, virtual dispatch where rodio uses Box<dyn Source>) and the SAME f32 operation order,
// citing the file:line it follows.  Build flags must keep it faithful:
//     g++ -O2 -ffp-contract=off -fno-fast-math     (rustc never contracts a*b+c into an FMA)
//
// Pinning status (see DESIGN.md "Oracle"):
//   pinned by the reference's own golden vectors (tests/test_oracle_golden.py): SampleRateConverter,
//     ChannelCountConverter, MixerSource, ChannelVolume, lerp, db<->linear tables, SignalGenerator,
//     limiter behavioural bands (tests/limit.rs).
//   restated only (the reference has no result-pinning tests: blt.rs:563-568 and agc.rs:600-605 are
//     empty test modules): BltFilter, AutomaticGainControl, Spatial, reverb/Delay/Mix, Amplify, Speed,
//     UniformSourceIterator.
//   PARITY UNPINNED: integer<->float sample conversion — the arithmetic lives in the un-vendored crate
//     dasp_sample 0.11.0 (Cargo.lock:317-319) and no reference test pins its values.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <numeric>
#include <optional>
#include <utility>
#include <vector>

namespace rodio_oracle {

using Sample = float;  // src/common.rs:36,:48 (default feature set: f32)

// ---------------------------------------------------------------------------------------------
// src/math.rs
// ---------------------------------------------------------------------------------------------
constexpr float PI_F = 3.14159265358979323846264338327950288f;        // std::f32::consts::PI
constexpr float TAU_F = 6.28318530717958647692528676655900577f;       // std::f32::consts::TAU
constexpr float LOG2_10_F = 3.32192809488736234787031942948939018f;   // std::f32::consts::LOG2_10
constexpr float LOG10_2_F = 0.301029995663981195213738894724493027f;  // std::f32::consts::LOG10_2

// src/math.rs:24-26 : first + (second - first) * numerator as f32 / denominator as f32
inline float lerp(float first, float second, uint32_t numerator, uint32_t denominator) {
    float d = second - first;
    float m = d * (float)numerator;
    float q = m / (float)denominator;
    return first + q;
}
// src/math.rs:52-56 : f32::powf(2.0, decibels * 0.05 * LOG2_10)
inline float db_to_linear(float decibels) { return powf(2.0f, (decibels * 0.05f) * LOG2_10_F); }
// src/math.rs:87-90 : linear.log2() * LOG10_2 * 20.0
inline float linear_to_db(float linear) { return (log2f(linear) * LOG10_2_F) * 20.0f; }
// std::time::Duration::as_secs_f32 : (secs as f32) + (nanos as f32) / 1e9
inline float duration_secs_f32(uint64_t ns) {
    uint64_t secs = ns / 1000000000ull;
    uint32_t nanos = (uint32_t)(ns % 1000000000ull);
    return (float)secs + (float)nanos / 1000000000.0f;
}
// src/math.rs:111-113 : exp(-1.0 / (duration_secs * sample_rate as f32))
inline float duration_to_coefficient(uint64_t ns, uint32_t sample_rate) {
    return expf(-1.0f / (duration_secs_f32(ns) * (float)sample_rate));
}

// Rust `as u32` from f32: truncating, saturating, NaN -> 0.
inline uint32_t f32_as_u32(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
// Rust f32::clamp(min,max): NaN passes through (src: core::f32::clamp).
inline float clampf(float x, float lo, float hi) {
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}

// ---------------------------------------------------------------------------------------------
// trait Source (src/source/mod.rs:179-218)
// ---------------------------------------------------------------------------------------------
struct Source {
    virtual ~Source() = default;
    virtual std::optional<Sample> next() = 0;
    virtual std::optional<size_t> current_span_len() const = 0;
    virtual uint16_t channels() const = 0;
    virtual uint32_t sample_rate() const = 0;
    virtual std::unique_ptr<Source> clone() const = 0;  // Rust `Clone` where the adapter derives it
};
using Src = std::unique_ptr<Source>;

// src/buffer.rs:40-60,:76-82,:127-131  (SamplesBuffer)
struct SamplesBuffer : Source {
    std::shared_ptr<const std::vector<Sample>> data;
    size_t pos = 0;
    uint16_t ch;
    uint32_t rate;
    SamplesBuffer(uint16_t c, uint32_t r, std::vector<Sample> d)
        : data(std::make_shared<const std::vector<Sample>>(std::move(d))), ch(c), rate(r) {}
    SamplesBuffer(uint16_t c, uint32_t r, std::shared_ptr<const std::vector<Sample>> d)
        : data(std::move(d)), ch(c), rate(r) {}
    std::optional<Sample> next() override {
        if (pos >= data->size()) return std::nullopt;
        return (*data)[pos++];
    }
    std::optional<size_t> current_span_len() const override {
        if (pos >= data->size()) return (size_t)0;
        return data->size();
    }
    uint16_t channels() const override { return ch; }
    uint32_t sample_rate() const override { return rate; }
    Src clone() const override {
        auto p = std::make_unique<SamplesBuffer>(ch, rate, data);
        p->pos = pos;
        return p;
    }
};

// A Vec-backed source with a caller-chosen span report:
//   span = nullopt : benches/shared.rs:7-50 TestSource (span-less "forever")
//   span = n       : src/source/mod.rs:865-929 test_utils::TestSource (always Some(total len))
struct VecSource : Source {
    std::shared_ptr<const std::vector<Sample>> data;
    size_t pos = 0;
    uint16_t ch;
    uint32_t rate;
    std::optional<size_t> span;
    VecSource(uint16_t c, uint32_t r, std::shared_ptr<const std::vector<Sample>> d, std::optional<size_t> s)
        : data(std::move(d)), ch(c), rate(r), span(s) {}
    std::optional<Sample> next() override {
        if (pos >= data->size()) {
            pos++;
            return std::nullopt;
        }
        return (*data)[pos++];
    }
    std::optional<size_t> current_span_len() const override { return span; }
    uint16_t channels() const override { return ch; }
    uint32_t sample_rate() const override { return rate; }
    Src clone() const override {
        auto p = std::make_unique<VecSource>(ch, rate, data, span);
        p->pos = pos;
        return p;
    }
};

// src/source/signal_generator.rs:51-53,:107-135 (Function::Sine) and src/source/sine.rs:23-27
struct SignalGenerator : Source {
    enum Fn { Sine, Triangle, Square, Sawtooth } fn;
    uint32_t rate;
    float phase_step, phase = 0.0f, period;
    SignalGenerator(uint32_t sample_rate, float frequency, Fn f) : fn(f), rate(sample_rate) {
        period = (float)sample_rate / frequency;  // :118
        phase_step = 1.0f / period;               // :119
    }
    static float rem_euclid1(float x) {  // f32::rem_euclid(1.0)
        float r = fmodf(x, 1.0f);
        if (r < 0.0f) r += 1.0f;
        return r;
    }
    std::optional<Sample> next() override {
        float v;
        switch (fn) {
            case Sine: v = sinf(TAU_F * phase); break;                                         // :51-53
            case Triangle: v = 4.0f * fabsf(phase - floorf(phase + 0.5f)) - 1.0f; break;       // :55-57
            case Square: v = (fmodf(phase, 1.0f) < 0.5f) ? 1.0f : -1.0f; break;                // :59-65
            default: v = 2.0f * (phase - floorf(phase + 0.5f)); break;                         // :67-69
        }
        phase = rem_euclid1(phase + phase_step);  // :133
        return v;
    }
    std::optional<size_t> current_span_len() const override { return std::nullopt; }
    uint16_t channels() const override { return 1; }
    uint32_t sample_rate() const override { return rate; }
    Src clone() const override { return std::make_unique<SignalGenerator>(*this); }
};
inline Src sine_wave(float freq) {  // SineWave::new, 48 kHz mono (src/common.rs:10)
    return std::make_unique<SignalGenerator>(48000u, freq, SignalGenerator::Sine);
}

// Iterator::take(n) on a Source (used by tests: `.take(2600)`), keeps metadata.
struct TakeN : Source {
    Src in;
    size_t n;
    TakeN(Src i, size_t k) : in(std::move(i)), n(k) {}
    std::optional<Sample> next() override {
        if (n == 0) return std::nullopt;
        n--;
        return in->next();
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override { return std::make_unique<TakeN>(in->clone(), n); }
};

// src/source/amplify.rs:63-65
struct Amplify : Source {
    Src in;
    float factor;
    Amplify(Src i, float f) : in(std::move(i)), factor(f) {}
    std::optional<Sample> next() override {
        auto v = in->next();
        if (!v) return std::nullopt;
        return *v * factor;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override { return std::make_unique<Amplify>(in->clone(), factor); }
};

// src/source/speed.rs:103-105,:130-133
struct Speed : Source {
    Src in;
    float factor;
    Speed(Src i, float f) : in(std::move(i)), factor(f) {}
    std::optional<Sample> next() override { return in->next(); }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override {
        float r = (float)in->sample_rate() * factor;
        r = fmaxf(r, 1.0f);  // f32::max
        return f32_as_u32(r);
    }
    Src clone() const override { return std::make_unique<Speed>(in->clone(), factor); }
};

// src/source/blt.rs — RBJ biquad, Direct Form I, per-channel state.  The SpanTracker branch
// (blt.rs:122-138) only fires when rate/channels change at a span boundary; every source in this
// oracle has stable parameters, so it is never taken and is not restated.
struct BltCoeffs {
    float b0, b1, b2, a1, a2;
};
inline BltCoeffs blt_low_pass(uint32_t freq, float q, uint32_t fs) {  // blt.rs:504-522
    float w0 = ((2.0f * PI_F) * (float)freq) / (float)fs;
    float alpha = sinf(w0) / (2.0f * q);
    float b1 = 1.0f - cosf(w0);
    float b0 = b1 / 2.0f;
    float b2 = b0;
    float a0 = 1.0f + alpha;
    float a1 = -2.0f * cosf(w0);
    float a2 = 1.0f - alpha;
    return {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}
inline BltCoeffs blt_high_pass(uint32_t freq, float q, uint32_t fs) {  // blt.rs:523-541
    float w0 = ((2.0f * PI_F) * (float)freq) / (float)fs;
    float cos_w0 = cosf(w0);
    float alpha = sinf(w0) / (2.0f * q);
    float b0 = (1.0f + cos_w0) / 2.0f;
    float b1 = -1.0f - cos_w0;
    float b2 = b0;
    float a0 = 1.0f + alpha;
    float a1 = -2.0f * cos_w0;
    float a2 = 1.0f - alpha;
    return {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}
// src/source/span.rs:33-101: SpanTracker (span-counting mode while `cached_span_len` is Some, parameter comparison on every sample
// while it is None -- the state a tracker starts in, span.rs:56-63)
struct SpanTracker {
    size_t samples_counted = 0;
    std::optional<size_t> cached_span_len;
    uint32_t last_sample_rate;
    uint16_t last_channels;
    SpanTracker(uint32_t rate, uint16_t ch) : last_sample_rate(rate), last_channels(ch) {}
    struct Detection {
        bool at_span_boundary, parameters_changed;
    };
    Detection advance(const Source& source) {  // span.rs:66-101
        if (samples_counted != SIZE_MAX) samples_counted++;   // saturating_add(1)
        std::optional<size_t> input_span_len = source.current_span_len();
        bool parameters_changed = false, at_span_boundary = false;
        if (input_span_len) {
            std::optional<bool> known_boundary;
            if (cached_span_len) known_boundary = samples_counted >= *cached_span_len;
            if (!known_boundary || *known_boundary) {
                uint16_t cur_ch = source.channels();
                uint32_t cur_rate = source.sample_rate();
                parameters_changed = cur_ch != last_channels || cur_rate != last_sample_rate;
                last_channels = cur_ch, last_sample_rate = cur_rate;
            }
            at_span_boundary = known_boundary ? *known_boundary : parameters_changed;
        }
        if (at_span_boundary) samples_counted = 0, cached_span_len = input_span_len;
        return {at_span_boundary, parameters_changed};
    }
};

// src/source/blt.rs:43-55 (constructor), :111-139 (next: the sample first, then the span tracker; a parameter change recomputes the
// coefficients for the samples that follow -- `current_channels != channels()` at :128 compares a value with itself, so the state
// layout chosen at construction, blt.rs:247-283, is never rebuilt), :397-492 (per-sample step)
struct BltFilter : Source {
    Src in;
    bool high;
    uint32_t freq;
    float q;
    BltCoeffs k;
    std::vector<float> x1, x2, y1, y2;
    size_t position = 0;
    SpanTracker span;
    BltFilter(Src i, bool hp, uint32_t f, float qq)
        : in(std::move(i)), high(hp), freq(f), q(qq), span(in->sample_rate(), in->channels()) {
        uint32_t fs = in->sample_rate();  // blt.rs:195-199
        k = hp ? blt_high_pass(f, qq, fs) : blt_low_pass(f, qq, fs);
        size_t n = in->channels();
        x1.assign(n, 0.0f), x2.assign(n, 0.0f), y1.assign(n, 0.0f), y2.assign(n, 0.0f);
    }
    std::optional<Sample> next() override {  // blt.rs:397-410 / :431-451 / :472-492
        auto s = in->next();
        if (!s) return std::nullopt;
        size_t c = position;
        position = (position + 1) % x1.size();
        float x = *s;
        // blt.rs:558-560, strictly left to right
        float r = k.b0 * x;
        r = r + k.b1 * x1[c];
        r = r + k.b2 * x2[c];
        r = r - k.a1 * y1[c];
        r = r - k.a2 * y2[c];
        y2[c] = y1[c];
        x2[c] = x1[c];
        y1[c] = r;
        x1[c] = x;
        SpanTracker::Detection d = span.advance(*in);   // blt.rs:122-137
        if (d.at_span_boundary && d.parameters_changed) {
            uint32_t fs = in->sample_rate();
            k = high ? blt_high_pass(freq, q, fs) : blt_low_pass(freq, q, fs);   // recreate_applier
        }
        return r;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<BltFilter>(in->clone(), high, freq, q);
        p->x1 = x1, p->x2 = x2, p->y1 = y1, p->y2 = y2, p->position = position, p->k = k, p->span = span;
        return p;
    }
};

// src/source/pausable.rs:8-21,:31-50,:85-97 with the control calls of a Player scripted: set_paused(true) is observed when `at`
// samples have been pulled from the input, set_paused(false) after `frames` whole frames of silence.
struct Pausable : Source {
    Src in;
    std::optional<uint16_t> paused_channels;
    uint16_t remaining_paused_samples = 0;
    uint64_t at, frames, pulled = 0, silent_frames = 0;
    bool done = false;
    Pausable(Src i, uint64_t at_sample, uint64_t n_frames) : in(std::move(i)), at(at_sample), frames(n_frames) {}
    void set_paused(bool paused) {  // :42-50
        if (!paused_channels && paused) paused_channels = in->channels();
        else if (paused_channels && !paused) paused_channels.reset();
    }
    std::optional<Sample> next() override {  // :85-97
        // the script (what Player's periodic access does between two calls)
        if (!done && frames && pulled == at && !paused_channels && silent_frames == 0) set_paused(true);
        if (paused_channels && remaining_paused_samples == 0 && silent_frames == frames) set_paused(false), done = true;
        if (remaining_paused_samples > 0) {
            remaining_paused_samples--;
            return 0.0f;
        }
        if (paused_channels) {
            remaining_paused_samples = (uint16_t)(*paused_channels - 1);
            silent_frames++;
            return 0.0f;
        }
        auto v = in->next();
        if (v) pulled++;
        return v;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<Pausable>(in->clone(), at, frames);
        p->paused_channels = paused_channels, p->remaining_paused_samples = remaining_paused_samples;
        p->pulled = pulled, p->silent_frames = silent_frames, p->done = done;
        return p;
    }
};

// src/source/from_iter.rs:16-127: the sources of an iterator played one after the other; the format may change from one to the next.
// current_span_len(): the current source's while it is not exhausted, None otherwise (:83-91); channels() / sample_rate(): the
// current source's, also when it is exhausted -- the next one is only fetched inside next() (:48-63).
struct FromIter : Source {
    std::vector<Src> rest;   // in order
    size_t next_idx = 0;
    Src current;
    explicit FromIter(std::vector<Src> srcs) : rest(std::move(srcs)) {
        if (!rest.empty()) current = std::move(rest[0]), next_idx = 1;
    }
    std::optional<Sample> next() override {
        while (true) {
            if (current) {
                auto v = current->next();
                if (v) return v;
            }
            if (next_idx < rest.size()) current = std::move(rest[next_idx++]);
            else return std::nullopt;
        }
    }
    std::optional<size_t> current_span_len() const override {
        if (current) {
            auto s = current->current_span_len();
            if (!(s && *s == 0)) return s;   // !is_exhausted() (src/source/mod.rs:204-206)
        }
        return std::nullopt;
    }
    uint16_t channels() const override { return current ? current->channels() : (uint16_t)2; }
    uint32_t sample_rate() const override { return current ? current->sample_rate() : 48000u; }
    Src clone() const override {
        std::vector<Src> v;
        if (current) v.push_back(current->clone());
        for (size_t i = next_idx; i < rest.size(); i++) v.push_back(rest[i]->clone());
        return std::make_unique<FromIter>(std::move(v));
    }
};

// src/source/delay.rs:8-16,:68-75
inline size_t delay_remaining_samples(uint64_t ns, uint32_t rate, uint16_t channels) {
    unsigned __int128 s = (unsigned __int128)ns * channels * rate / 1000000000ull;
    return (size_t)s;
}
struct Delay : Source {
    Src in;
    size_t remaining;
    Delay(Src i, uint64_t ns) : in(std::move(i)) {
        remaining = delay_remaining_samples(ns, in->sample_rate(), in->channels());
    }
    Delay(Src i, size_t rem, int) : in(std::move(i)), remaining(rem) {}
    std::optional<Sample> next() override {
        if (remaining >= 1) {
            remaining--;
            return 0.0f;
        }
        return in->next();
    }
    std::optional<size_t> current_span_len() const override {
        auto v = in->current_span_len();
        if (!v) return std::nullopt;
        return *v + remaining;
    }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override { return std::make_unique<Delay>(in->clone(), remaining, 0); }
};

// src/source/distortion.rs:66-72
struct Distortion : Source {
    Src in;
    float gain, threshold;
    Distortion(Src i, float g, float t) : in(std::move(i)), gain(g), threshold(t) {}
    std::optional<Sample> next() override {
        auto v = in->next();
        if (!v) return std::nullopt;
        return clampf(*v * gain, -threshold, threshold);
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override { return std::make_unique<Distortion>(in->clone(), gain, threshold); }
};

// src/source/linear_ramp.rs:9-34,:79-104 (fade_in = ramp(d, 0, 1, false), fade_out = ramp(d, 1, 0, true))
struct LinearGainRamp : Source {
    Src in;
    uint64_t elapsed_ns = 0, total_ns;
    float start_gain, end_gain;
    bool clamp_end;
    uint64_t sample_idx = 0;
    LinearGainRamp(Src i, uint64_t d, float s, float e, bool c)
        : in(std::move(i)), total_ns(d), start_gain(s), end_gain(e), clamp_end(c) {}
    std::optional<Sample> next() override {
        float factor;
        if (elapsed_ns >= total_ns) {
            factor = clamp_end ? end_gain : 1.0f;
        } else {
            sample_idx += 1;
            float p = duration_secs_f32(elapsed_ns) / duration_secs_f32(total_ns);
            factor = start_gain * (1.0f - p) + end_gain * p;
        }
        if (sample_idx % in->channels() == 0) elapsed_ns += 1000000000ull / in->sample_rate();   // :94-100
        auto v = in->next();
        if (!v) return std::nullopt;
        return *v * factor;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<LinearGainRamp>(in->clone(), total_ns, start_gain, end_gain, clamp_end);
        p->elapsed_ns = elapsed_ns, p->sample_idx = sample_idx;
        return p;
    }
};

// src/source/take.rs:9-26,:34-41,:107-148,:180-196
struct TakeDuration : Source {
    Src in;
    uint64_t remaining_ns, requested_ns, dps_ns;
    bool fadeout;
    size_t samples_in_current_frame = 0, silence_samples_remaining = 0;
    TakeDuration(Src i, uint64_t d, bool f) : in(std::move(i)), remaining_ns(d), requested_ns(d), fadeout(f) {
        dps_ns = 1000000000ull / ((uint64_t)in->sample_rate() * in->channels());   // :65-69
    }
    std::optional<Sample> next() override {
        while (true) {
            if (silence_samples_remaining > 0) {
                silence_samples_remaining -= 1;
                return 0.0f;
            }
            if (remaining_ns < dps_ns) {
                silence_samples_remaining = samples_in_current_frame > 0 ? in->channels() - samples_in_current_frame : 0;
                if (silence_samples_remaining > 0) {
                    samples_in_current_frame = 0;
                    continue;
                }
                return std::nullopt;
            }
            auto s = in->next();
            if (!s) return std::nullopt;
            samples_in_current_frame = (samples_in_current_frame + 1) % in->channels();
            float sample = *s;
            if (fadeout) {   // :34-41  sample * remaining / total
                float remaining = (float)(remaining_ns / 1000000ull);
                float total = (float)(requested_ns / 1000000ull);
                sample = sample * remaining / total;
            }
            remaining_ns -= dps_ns;
            return sample;
        }
    }
    std::optional<size_t> current_span_len() const override {   // :180-196
        if (dps_ns == 0 || remaining_ns == 0) return (size_t)0;
        size_t remaining_samples = (size_t)(remaining_ns / dps_ns);
        auto v = in->current_span_len();
        if (v && *v < remaining_samples) return v;
        return remaining_samples;
    }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<TakeDuration>(in->clone(), requested_ns, fadeout);
        p->remaining_ns = remaining_ns, p->samples_in_current_frame = samples_in_current_frame;
        p->silence_samples_remaining = silence_samples_remaining;
        return p;
    }
};

// src/conversions/channels.rs:57-85
struct ChannelCountConverter {
    Source* input_src = nullptr;  // set by owner
    uint16_t from, to;
    std::optional<Sample> sample_repeat;
    uint16_t next_output_sample_pos = 0;
    ChannelCountConverter(uint16_t f, uint16_t t) : from(f), to(t) {}
    template <class In>
    std::optional<Sample> next(In& input) {
        std::optional<Sample> result;
        uint16_t x = next_output_sample_pos;
        if (x == 0) {
            auto value = input.next();
            sample_repeat = value;
            result = value;
        } else if (x < from) {
            result = input.next();
        } else if (x == 1) {
            result = sample_repeat;
        } else {
            result = 0.0f;
        }
        if (result) next_output_sample_pos += 1;
        if (next_output_sample_pos == to) {
            next_output_sample_pos = 0;
            if (from > to) {
                for (uint16_t i = to; i < from; i++) input.next();  // discarding extra input
            }
        }
        return result;
    }
};

// src/conversions/sample_rate.rs:52-201.  u32 arithmetic wraps exactly like a Rust release build.
template <class In>
struct SampleRateConverter {
    In input;
    uint32_t from, to;
    uint16_t channels;
    std::vector<Sample> current_span, next_frame;
    uint32_t current_span_pos_in_chunk = 0;
    uint32_t next_output_span_pos_in_chunk = 0;
    std::deque<Sample> output_buffer;

    SampleRateConverter(In in, uint32_t from_rate, uint32_t to_rate, uint16_t ch)
        : input(std::move(in)), channels(ch) {
        if (from_rate != to_rate) {  // :58-71 eager read of two frames
            for (uint16_t i = 0; i < ch; i++) {
                auto v = input.next();
                if (!v) break;
                current_span.push_back(*v);
            }
            for (uint16_t i = 0; i < ch; i++) {
                auto v = input.next();
                if (!v) break;
                next_frame.push_back(*v);
            }
        }
        uint32_t g = std::gcd(from_rate, to_rate);  // :74 Ratio::new(to, from).into_raw()
        from = from_rate / g;
        to = to_rate / g;
    }
    void next_input_span() {  // :110-122
        current_span_pos_in_chunk += 1;
        std::swap(current_span, next_frame);
        next_frame.clear();
        for (uint16_t i = 0; i < channels; i++) {
            auto v = input.next();
            if (v)
                next_frame.push_back(*v);
            else
                break;
        }
    }
    std::optional<Sample> next() {  // :131-201
        if (from == to) return input.next();
        if (!output_buffer.empty()) {
            Sample s = output_buffer.front();
            output_buffer.pop_front();
            return s;
        }
        if (next_output_span_pos_in_chunk == to) {
            next_output_span_pos_in_chunk = 0;
            next_input_span();
            while (current_span_pos_in_chunk != from) next_input_span();
            current_span_pos_in_chunk = 0;
        } else {
            uint32_t req_left_sample = (uint32_t)(from * next_output_span_pos_in_chunk) / to % from;
            while (current_span_pos_in_chunk != req_left_sample) next_input_span();
        }
        std::optional<Sample> result;
        uint32_t numerator = (uint32_t)(from * next_output_span_pos_in_chunk) % to;
        size_t n = std::min(current_span.size(), next_frame.size());  // zip
        for (size_t off = 0; off < n; off++) {
            Sample sample = lerp(current_span[off], next_frame[off], numerator, to);
            if (off == 0)
                result = sample;
            else
                output_buffer.push_back(sample);
        }
        next_output_span_pos_in_chunk += 1;
        if (result) return result;
        // draining `current_span`
        if (current_span.empty()) return std::nullopt;
        Sample r = current_span[0];
        for (size_t i = 1; i < current_span.size(); i++) output_buffer.push_back(current_span[i]);
        current_span.clear();
        return r;
    }
};

// src/source/uniform.rs:33-97,:148-197
struct UniformSourceIterator : Source {
    struct Take {  // uniform.rs:148-178
        Source* iter;
        std::optional<size_t> n;
        std::optional<Sample> next() {
            if (n) {
                if (*n != 0) {
                    *n -= 1;
                    return iter->next();
                }
                return std::nullopt;
            }
            return iter->next();
        }
    };
    struct Inner {
        SampleRateConverter<Take> src;
        ChannelCountConverter ccc;
        Inner(Take t, uint32_t from_rate, uint32_t to_rate, uint16_t from_ch, uint16_t to_ch)
            : src(t, from_rate, to_rate, from_ch), ccc(from_ch, to_ch) {}
        std::optional<Sample> next() { return ccc.next(src); }
    };
    Src input;  // both `pending` and the iterator buried in `inner`
    std::unique_ptr<Inner> inner;
    uint16_t target_channels;
    uint32_t target_sample_rate;
    UniformSourceIterator(Src in, uint16_t ch, uint32_t rate)
        : input(std::move(in)), target_channels(ch), target_sample_rate(rate) {}
    std::unique_ptr<Inner> bootstrap() {  // uniform.rs:50-68
        std::optional<size_t> span_len = input->current_span_len();
        if (span_len) span_len = std::min<size_t>(*span_len, 32768);
        uint16_t from_channels = input->channels();
        uint32_t from_sample_rate = input->sample_rate();
        Take t{input.get(), span_len};
        return std::make_unique<Inner>(t, from_sample_rate, target_sample_rate, from_channels, target_channels);
    }
    std::optional<Sample> next() override {  // uniform.rs:78-97
        if (inner) {
            auto v = inner->next();
            if (v) return v;
        }
        inner = bootstrap();
        return inner->next();
    }
    std::optional<size_t> current_span_len() const override { return std::nullopt; }
    uint16_t channels() const override { return target_channels; }
    uint32_t sample_rate() const override { return target_sample_rate; }
    Src clone() const override {
        // only ever cloned before the first next() in this oracle (reverb clones its input first)
        return std::make_unique<UniformSourceIterator>(input->clone(), target_channels, target_sample_rate);
    }
};

// src/source/mix.rs:10-53
struct Mix : Source {
    UniformSourceIterator input1, input2;
    Mix(Src a, Src b, uint16_t ch, uint32_t rate) : input1(std::move(a), ch, rate), input2(std::move(b), ch, rate) {}
    static std::unique_ptr<Mix> make(Src a, Src b) {
        uint16_t ch = a->channels();
        uint32_t rate = a->sample_rate();
        return std::make_unique<Mix>(std::move(a), std::move(b), ch, rate);
    }
    std::optional<Sample> next() override {
        auto s1 = input1.next();
        auto s2 = input2.next();
        if (s1 && s2) return *s1 + *s2;
        if (s1) return s1;
        if (s2) return s2;
        return std::nullopt;
    }
    std::optional<size_t> current_span_len() const override { return std::nullopt; }  // min(None, None)
    uint16_t channels() const override { return input1.channels(); }
    uint32_t sample_rate() const override { return input1.sample_rate(); }
    Src clone() const override {
        return std::make_unique<Mix>(input1.input->clone(), input2.input->clone(), input1.target_channels,
                                     input1.target_sample_rate);
    }
};
// src/source/mod.rs:628-634 : let echo = self.clone().amplify(amplitude).delay(duration); self.mix(echo)
inline Src reverb(Src self, uint64_t duration_ns, float amplitude) {
    Src echo = std::make_unique<Delay>(std::make_unique<Amplify>(self->clone(), amplitude), duration_ns);
    return Mix::make(std::move(self), std::move(echo));
}

// src/source/agc.rs:133-171 (CircularBuffer), :183-236 (ctor), :397-504 (per-sample).
struct AutomaticGainControl : Source {
    static constexpr size_t RMS_WINDOW_SIZE = 8192;
    Src in;
    float target_level, floor_ = 0.0f, absolute_max_gain;
    float current_gain = 1.0f, attack_coeff, release_coeff, peak_level = 0.0f;
    std::vector<float> buffer;
    float sum = 0.0f;
    size_t index = 0;
    bool is_enabled = true;
    uint64_t attack_time_ns, release_time_ns;
    SpanTracker span;
    AutomaticGainControl(Src i, float target, uint64_t attack_ns, uint64_t release_ns, float max_gain, float floor_v)
        : in(std::move(i)), target_level(target), floor_(floor_v), absolute_max_gain(max_gain),
          buffer(RMS_WINDOW_SIZE, 0.0f), span(in->sample_rate(), in->channels()) {
        // src/source/mod.rs:432-433 : times limited to 10 s
        uint64_t ten = 10ull * 1000000000ull;
        attack_ns = std::min(attack_ns, ten);
        release_ns = std::min(release_ns, ten);
        attack_time_ns = attack_ns, release_time_ns = release_ns;
        attack_coeff = duration_to_coefficient(attack_ns, in->sample_rate());
        release_coeff = duration_to_coefficient(release_ns, in->sample_rate());
    }
    float push(float value) {  // agc.rs:154-163
        float old_value = buffer[index];
        sum = sum - old_value + value;
        buffer[index] = value;
        index = (index + 1) & (RMS_WINDOW_SIZE - 1);
        return old_value;
    }
    float process_sample(float sample) {  // agc.rs:433-504
        float sample_value = fabsf(sample);
        // update_peak_level :397-408
        float coeff = (sample_value > peak_level) ? 0.0f : release_coeff;
        peak_level = peak_level * coeff + sample_value * (1.0f - coeff);
        // update_rms :413-418
        float squared_sample = sample_value * sample_value;
        push(squared_sample);
        float rms = sqrtf(sum / (float)RMS_WINDOW_SIZE);
        float rms_gain = (rms > 0.0f) ? target_level / rms : absolute_max_gain;
        // calculate_peak_gain :424-431
        float peak_gain = (peak_level > 0.0f) ? fminf(target_level / peak_level, absolute_max_gain) : absolute_max_gain;
        float desired_gain = fmaxf(fminf(rms_gain, peak_gain), floor_);
        float attack_speed = (desired_gain > current_gain) ? attack_coeff : release_coeff;
        current_gain = current_gain * attack_speed + desired_gain * (1.0f - attack_speed);
        current_gain = clampf(current_gain, 0.1f, absolute_max_gain);
        return sample * current_gain;
    }
    std::optional<Sample> next() override {  // agc.rs:524-557: the tracker BEFORE the sample is pulled
        SpanTracker::Detection d = span.advance(*in);
        if (d.at_span_boundary && d.parameters_changed) {   // :527-548: coefficients for the new rate, everything else from scratch
            uint32_t rate = in->sample_rate();
            attack_coeff = duration_to_coefficient(attack_time_ns, rate);
            release_coeff = duration_to_coefficient(release_time_ns, rate);
            buffer.assign(RMS_WINDOW_SIZE, 0.0f), sum = 0.0f, index = 0;   // CircularBuffer::new()
            peak_level = 0.0f;
            current_gain = 1.0f;
        }
        auto s = in->next();
        if (!s) return std::nullopt;
        return is_enabled ? process_sample(*s) : *s;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<AutomaticGainControl>(in->clone(), target_level, 0, 0, absolute_max_gain, floor_);
        p->attack_coeff = attack_coeff, p->release_coeff = release_coeff, p->current_gain = current_gain;
        p->peak_level = peak_level, p->buffer = buffer, p->sum = sum, p->index = index, p->is_enabled = is_enabled;
        p->attack_time_ns = attack_time_ns, p->release_time_ns = release_time_ns, p->span = span;
        return p;
    }
};

// src/source/limit.rs:94-130 (ctor), :854-873 (gain computer), :903-916 (envelope), :927-988 (variants)
struct Limit : Source {
    Src in;
    float threshold, knee_width, inv_knee_8, attack, release;
    std::vector<float> integrators, peaks;
    size_t position = 0;
    SpanTracker span;
    Limit(Src i, float thr, float knee, uint64_t attack_ns, uint64_t release_ns)
        : in(std::move(i)), threshold(thr), knee_width(knee), span(in->sample_rate(), in->channels()) {
        attack = duration_to_coefficient(attack_ns, in->sample_rate());
        release = duration_to_coefficient(release_ns, in->sample_rate());
        inv_knee_8 = 1.0f / (8.0f * knee_width);  // :877
        size_t n = in->channels();
        integrators.assign(n, 0.0f), peaks.assign(n, 0.0f);
    }
    static constexpr float MIN_POSITIVE = 1.17549435e-38f;  // f32::MIN_POSITIVE
    float process_sample(float sample) const {              // :854-873
        float bias_db = linear_to_db(fabsf(sample) + MIN_POSITIVE) - threshold;
        float knee_boundary_db = bias_db * 2.0f;
        if (knee_boundary_db < -knee_width) return 0.0f;
        if (fabsf(knee_boundary_db) <= knee_width) {
            float x = knee_boundary_db + knee_width;
            return x * x * inv_knee_8;
        }
        return bias_db;
    }
    std::optional<Sample> next() override {
        auto s = in->next();
        if (!s) return std::nullopt;
        float sample = *s;
        size_t n = integrators.size();
        size_t c = position;
        position = (position + 1) % n;
        float limiter_db = process_sample(sample);  // :906
        integrators[c] = fmaxf(limiter_db, release * integrators[c] + (1.0f - release) * limiter_db);  // :909-912
        peaks[c] = attack * peaks[c] + (1.0f - attack) * integrators[c];                                 // :913
        float max_peak;
        if (n == 1)
            max_peak = peaks[0];  // :927-934
        else if (n == 2)
            max_peak = fmaxf(peaks[0], peaks[1]);  // :946-960
        else {
            max_peak = 0.0f;  // :971-988 fold(0.0, max)
            for (float p : peaks) max_peak = fmaxf(max_peak, p);
        }
        const float out = sample * db_to_linear(-max_peak);
        // limit.rs:651-697: the tracker behind the sample; another channel count rebuilds the per-channel state (`base`, with the
        // coefficients of the rate at construction, is kept)
        SpanTracker::Detection d = span.advance(*in);
        if (d.at_span_boundary && d.parameters_changed) {
            size_t nc = in->channels();
            if (nc != integrators.size()) integrators.assign(nc, 0.0f), peaks.assign(nc, 0.0f), position = 0;
        }
        return out;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return in->channels(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<Limit>(in->clone(), threshold, knee_width, 0, 0);
        p->attack = attack, p->release = release, p->integrators = integrators, p->peaks = peaks, p->position = position;
        p->span = span;
        return p;
    }
};

// src/source/channel_volume.rs:30-38,:71-88
struct ChannelVolume : Source {
    Src in;
    std::vector<float> channel_volumes;
    size_t current_channel;
    std::optional<Sample> current_sample;
    ChannelVolume(Src i, std::vector<float> v) : in(std::move(i)), channel_volumes(std::move(v)) {
        current_channel = channel_volumes.size();
    }
    std::optional<Sample> next() override {
        if (current_channel >= channel_volumes.size()) {
            current_channel = 0;
            current_sample = std::nullopt;
            uint16_t c = in->channels();
            for (uint16_t i = 0; i < c; i++) {
                auto s = in->next();
                if (!s) return std::nullopt;  // `?`
                current_sample = (current_sample ? *current_sample : 0.0f) + *s;
            }
            if (current_sample) current_sample = *current_sample / (float)in->channels();
        }
        std::optional<Sample> result;
        if (current_sample) result = *current_sample * channel_volumes[current_channel];
        current_channel += 1;
        return result;
    }
    std::optional<size_t> current_span_len() const override { return in->current_span_len(); }
    uint16_t channels() const override { return (uint16_t)channel_volumes.size(); }
    uint32_t sample_rate() const override { return in->sample_rate(); }
    Src clone() const override {
        auto p = std::make_unique<ChannelVolume>(in->clone(), channel_volumes);
        p->current_channel = current_channel, p->current_sample = current_sample;
        return p;
    }
};
// src/source/spatial.rs:19-24,:48-69
inline float dist_sq(const float a[3], const float b[3]) {
    float s = 0.0f;  // Iterator::sum::<f32>() folds from 0.0
    for (int i = 0; i < 3; i++) s = s + (a[i] - b[i]) * (a[i] - b[i]);
    return s;
}
inline void spatial_volumes(const float emitter[3], const float left[3], const float right[3], float out[2]) {
    float left_dist_sq = dist_sq(left, emitter);
    float right_dist_sq = dist_sq(right, emitter);
    float max_diff = sqrtf(dist_sq(left, right));
    float left_dist = sqrtf(left_dist_sq);
    float right_dist = sqrtf(right_dist_sq);
    float left_diff_modifier = fminf(((left_dist - right_dist) / max_diff + 1.0f) / 4.0f + 0.5f, 1.0f);
    float right_diff_modifier = fminf(((right_dist - left_dist) / max_diff + 1.0f) / 4.0f + 0.5f, 1.0f);
    float left_dist_modifier = fminf(1.0f / left_dist_sq, 1.0f);
    float right_dist_modifier = fminf(1.0f / right_dist_sq, 1.0f);
    out[0] = left_diff_modifier * left_dist_modifier;
    out[1] = right_diff_modifier * right_dist_modifier;
}
inline Src spatial(Src in, const float emitter[3], const float left[3], const float right[3]) {
    float v[2];
    spatial_volumes(emitter, left, right, v);
    return std::make_unique<ChannelVolume>(std::move(in), std::vector<float>{v[0], v[1]});
}

// src/mixer.rs:25-43,:58-66,:120-136,:175-198.  `add` queues into `pending` exactly like the channel.
struct MixerSource : Source {
    std::vector<Src> current_sources, still_pending, pending_rx;
    uint16_t ch;
    uint32_t rate;
    uint16_t current_channel = 0;
    MixerSource(uint16_t c, uint32_t r) : ch(c), rate(r) {}
    void add(Src source) {  // Mixer::add, mixer.rs:58-66
        pending_rx.push_back(std::make_unique<UniformSourceIterator>(std::move(source), ch, rate));
    }
    std::optional<Sample> next() override {
        // start_pending_sources :175-183
        for (auto& s : pending_rx) still_pending.push_back(std::move(s));
        pending_rx.clear();
        if (current_channel == 0) {
            for (auto& s : still_pending) current_sources.push_back(std::move(s));
            still_pending.clear();
        }
        // sum_current_sources :185-198
        float sum = 0.0f;
        size_t w = 0;
        for (size_t i = 0; i < current_sources.size(); i++) {
            auto v = current_sources[i]->next();  // one virtual call per source per sample
            if (v) {
                sum += *v;
                if (w != i) current_sources[w] = std::move(current_sources[i]);
                w++;
            }
        }
        current_sources.resize(w);
        current_channel += 1;
        if (current_channel >= ch) current_channel = 0;
        if (current_sources.empty()) return std::nullopt;
        return sum;
    }
    std::optional<size_t> current_span_len() const override { return std::nullopt; }
    uint16_t channels() const override { return ch; }
    uint32_t sample_rate() const override { return rate; }
    Src clone() const override { return nullptr; }
};

// ---------------------------------------------------------------------------------------------
// Statically dispatched variants (CPU baseline only).  rustc monomorphises an adapter chain such as
// Amplify<BltFilter<UniformSourceIterator<TestSource>>> into one inlined next(); only the mixer holds
// Box<dyn Source> (mixer.rs:72).  Timing the fully virtual classes above would understate the
// reference, so the baseline builds the bench chain from these templates: same arithmetic, same
// control flow, static dispatch inside a source, one virtual call per source per sample at the mixer.
// tests/test_oracle_golden.py checks they are bit-identical to the virtual classes.
// ---------------------------------------------------------------------------------------------
struct VecT {   // benches/shared.rs TestSource (span-less)
    const Sample* p;
    size_t n, pos = 0;
    uint16_t ch;
    uint32_t rate;
    inline std::optional<Sample> next() {
        if (pos >= n) return std::nullopt;
        return p[pos++];
    }
    std::optional<size_t> current_span_len() const { return std::nullopt; }
    uint16_t channels() const { return ch; }
    uint32_t sample_rate() const { return rate; }
};
template <class In>
struct AmplifyT {
    In in;
    float factor;
    inline std::optional<Sample> next() {
        auto v = in.next();
        if (!v) return std::nullopt;
        return *v * factor;
    }
    std::optional<size_t> current_span_len() const { return in.current_span_len(); }
    uint16_t channels() const { return in.channels(); }
    uint32_t sample_rate() const { return in.sample_rate(); }
};
template <class In>
struct BltT {   // mono / stereo / multi share one body here; state selection is position % channels
    In in;
    BltCoeffs k;
    std::vector<float> x1, x2, y1, y2;
    size_t position = 0;
    BltT(In i, bool hp, uint32_t f, float q) : in(std::move(i)) {
        k = hp ? blt_high_pass(f, q, in.sample_rate()) : blt_low_pass(f, q, in.sample_rate());
        size_t n = in.channels();
        x1.assign(n, 0.0f), x2.assign(n, 0.0f), y1.assign(n, 0.0f), y2.assign(n, 0.0f);
    }
    inline std::optional<Sample> next() {
        auto s = in.next();
        if (!s) return std::nullopt;
        size_t c = position;
        position = position + 1 == x1.size() ? 0 : position + 1;
        float x = *s;
        float r = k.b0 * x;
        r = r + k.b1 * x1[c];
        r = r + k.b2 * x2[c];
        r = r - k.a1 * y1[c];
        r = r - k.a2 * y2[c];
        y2[c] = y1[c], x2[c] = x1[c], y1[c] = r, x1[c] = x;
        return r;
    }
    std::optional<size_t> current_span_len() const { return in.current_span_len(); }
    uint16_t channels() const { return in.channels(); }
    uint32_t sample_rate() const { return in.sample_rate(); }
};
template <class In>
struct UniformT {   // src/source/uniform.rs:33-97 with a concrete input type
    struct Take {
        In* iter;
        std::optional<size_t> n;
        inline std::optional<Sample> next() {
            if (n) {
                if (*n != 0) {
                    *n -= 1;
                    return iter->next();
                }
                return std::nullopt;
            }
            return iter->next();
        }
    };
    struct Inner {
        SampleRateConverter<Take> src;
        ChannelCountConverter ccc;
        Inner(Take t, uint32_t fr, uint32_t tr, uint16_t fc, uint16_t tc) : src(t, fr, tr, fc), ccc(fc, tc) {}
    };
    In input;
    std::optional<Inner> inner;
    uint16_t target_channels;
    uint32_t target_sample_rate;
    UniformT(In in, uint16_t ch, uint32_t rate) : input(std::move(in)), target_channels(ch), target_sample_rate(rate) {}
    UniformT(UniformT&& o) noexcept
        : input(std::move(o.input)), target_channels(o.target_channels), target_sample_rate(o.target_sample_rate) {}
    void bootstrap() {
        std::optional<size_t> span_len = input.current_span_len();
        if (span_len) span_len = std::min<size_t>(*span_len, 32768);
        inner.emplace(Take{&input, span_len}, input.sample_rate(), target_sample_rate, input.channels(), target_channels);
    }
    inline std::optional<Sample> next() {
        if (inner) {
            auto v = inner->ccc.next(inner->src);
            if (v) return v;
        }
        bootstrap();
        return inner->ccc.next(inner->src);
    }
    std::optional<size_t> current_span_len() const { return std::nullopt; }
    uint16_t channels() const { return target_channels; }
    uint32_t sample_rate() const { return target_sample_rate; }
};
// Box<dyn Source>: the one virtual boundary the mixer sees.
template <class T>
struct Boxed : Source {
    T t;
    explicit Boxed(T v) : t(std::move(v)) {}
    std::optional<Sample> next() override { return t.next(); }
    std::optional<size_t> current_span_len() const override { return t.current_span_len(); }
    uint16_t channels() const override { return t.channels(); }
    uint32_t sample_rate() const override { return t.sample_rate(); }
    Src clone() const override { return nullptr; }
};
// MixerSource over already-uniform boxed sources (Mixer::add's wrap is part of the typed chain).
struct MixerSourceBoxed {
    std::vector<Src> current_sources;
    uint16_t ch;
    uint16_t current_channel = 0;
    explicit MixerSourceBoxed(uint16_t c) : ch(c) {}
    inline std::optional<Sample> next() {
        float sum = 0.0f;
        size_t w = 0;
        for (size_t i = 0; i < current_sources.size(); i++) {
            auto v = current_sources[i]->next();
            if (v) {
                sum += *v;
                if (w != i) current_sources[w] = std::move(current_sources[i]);
                w++;
            }
        }
        current_sources.resize(w);
        current_channel += 1;
        if (current_channel >= ch) current_channel = 0;
        if (current_sources.empty()) return std::nullopt;
        return sum;
    }
};

// dasp_sample 0.11.0 `conv` (un-vendored; PARITY UNPINNED).  Call site src/conversions/sample.rs:42-44.
inline int32_t f32_as_i32_sat(float v) {  // Rust `as i32`
    if (!(v == v)) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}
inline int16_t f32_as_i16_sat(float v) {
    if (!(v == v)) return 0;
    if (v <= -32768.0f) return INT16_MIN;
    if (v >= 32767.0f) return INT16_MAX;
    return (int16_t)v;
}
inline int8_t f32_as_i8_sat(float v) {
    if (!(v == v)) return 0;
    if (v <= -128.0f) return INT8_MIN;
    if (v >= 127.0f) return INT8_MAX;
    return (int8_t)v;
}
inline float i16_to_f32(int16_t s) { return (float)s / 32768.0f; }
inline float i8_to_f32(int8_t s) { return (float)s / 128.0f; }
inline float i32_to_f32(int32_t s) { return (float)s / 2147483648.0f; }
inline float i24_to_f32(int32_t s) { return (float)s / 8388608.0f; }
inline float u16_to_f32(uint16_t s) { return i16_to_f32((int16_t)((int32_t)s - 32768)); }
inline float u8_to_f32(uint8_t s) { return i8_to_f32((int8_t)((int32_t)s - 128)); }
inline int16_t f32_to_i16(float s) { return f32_as_i16_sat(s * 32768.0f); }
inline int8_t f32_to_i8(float s) { return f32_as_i8_sat(s * 128.0f); }
inline int32_t f32_to_i32(float s) { return f32_as_i32_sat(s * 2147483648.0f); }
inline int32_t f32_to_i24(float s) {
    int32_t v = f32_as_i32_sat(s * 8388608.0f);
    return v;
}
inline uint16_t f32_to_u16(float s) { return (uint16_t)((int32_t)f32_to_i16(s) + 32768); }
inline uint8_t f32_to_u8(float s) { return (uint8_t)((int32_t)f32_to_i8(s) + 128); }

}  // namespace rodio_oracle
