// rodio_oracle_capi.cpp — extern "C" surface of the CPU oracle (TEST INFRASTRUCTURE, see
// rodio_oracle.hpp).  Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg.  Never linked into or called from the product library.
#include "rodio_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace rodio_oracle;

extern "C" {

// Same field layout as rb_effect / rb_stream_desc in include/rodio_b200.h, declared independently.
struct ro_effect {
    uint32_t kind;
    uint32_t u32[3];
    float f32[12];
    uint64_t ns[2];
};
struct ro_stream {
    uint32_t sample_rate;
    uint16_t channels;
    uint16_t format;  // must be 0 (f32): convert with ro_convert first
    uint64_t n_samples;
    uint32_t span_len;  // 0 = None
    uint32_t n_effects;
    const ro_effect* effects;
    uint64_t mix_start;
    const float* pcm;
};
enum { FX_AMPLIFY = 1, FX_SPEED, FX_LOW_PASS, FX_HIGH_PASS, FX_REVERB, FX_AGC, FX_LIMIT, FX_SPATIAL,
       FX_CHANNEL_VOLUME, FX_UNIFORM, FX_DELAY, FX_DISTORTION, FX_LINEAR_RAMP, FX_TAKE_DURATION, FX_SIGNAL, FX_MIX, FX_APPEND, FX_PAUSE };

float ro_lerp(float a, float b, uint32_t num, uint32_t den) { return lerp(a, b, num, den); }
float ro_db_to_linear(float d) { return db_to_linear(d); }
float ro_linear_to_db(float l) { return linear_to_db(l); }
float ro_duration_to_coefficient(uint64_t ns, uint32_t rate) { return duration_to_coefficient(ns, rate); }
uint32_t ro_speed_sample_rate(uint32_t rate, float factor) {
    return f32_as_u32(fmaxf((float)rate * factor, 1.0f));
}
uint64_t ro_delay_samples(uint64_t ns, uint32_t rate, uint16_t ch) { return delay_remaining_samples(ns, rate, ch); }
void ro_spatial_volumes(const float* e, const float* l, const float* r, float* out) { spatial_volumes(e, l, r, out); }
void ro_blt_coeffs(int high, uint32_t freq, float q, uint32_t fs, float* out5) {
    BltCoeffs k = high ? blt_high_pass(freq, q, fs) : blt_low_pass(freq, q, fs);
    out5[0] = k.b0, out5[1] = k.b1, out5[2] = k.b2, out5[3] = k.a1, out5[4] = k.a2;
}

static Src build(const ro_stream& s);
static Src make_input(const ro_stream& s) {
    auto data = std::make_shared<const std::vector<Sample>>(s.pcm, s.pcm + s.n_samples);
    if (s.span_len != 0 && (uint64_t)s.span_len == s.n_samples)
        return std::make_unique<SamplesBuffer>(s.channels, s.sample_rate, data);  // Some(len) / Some(0) at end
    std::optional<size_t> span;
    if (s.span_len) span = (size_t)s.span_len;
    return std::make_unique<VecSource>(s.channels, s.sample_rate, data, span);
}
static Src apply_effects(Src src, const ro_effect* fx, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        const ro_effect& e = fx[i];
        switch (e.kind) {
            case FX_AMPLIFY: src = std::make_unique<Amplify>(std::move(src), e.f32[0]); break;
            case FX_SPEED: src = std::make_unique<Speed>(std::move(src), e.f32[0]); break;
            case FX_LOW_PASS: src = std::make_unique<BltFilter>(std::move(src), false, e.u32[0], e.f32[0]); break;
            case FX_HIGH_PASS: src = std::make_unique<BltFilter>(std::move(src), true, e.u32[0], e.f32[0]); break;
            case FX_REVERB: src = reverb(std::move(src), e.ns[0], e.f32[0]); break;
            case FX_AGC:
                src = std::make_unique<AutomaticGainControl>(std::move(src), e.f32[0], e.ns[0], e.ns[1], e.f32[1], e.f32[2]);
                break;
            case FX_LIMIT: src = std::make_unique<Limit>(std::move(src), e.f32[0], e.f32[1], e.ns[0], e.ns[1]); break;
            case FX_SPATIAL: src = spatial(std::move(src), e.f32, e.f32 + 3, e.f32 + 6); break;
            case FX_CHANNEL_VOLUME:
                src = std::make_unique<ChannelVolume>(std::move(src), std::vector<float>(e.f32, e.f32 + e.u32[0]));
                break;
            case FX_UNIFORM: src = std::make_unique<UniformSourceIterator>(std::move(src), (uint16_t)e.u32[0], e.u32[1]); break;
            case FX_DELAY: src = std::make_unique<Delay>(std::move(src), e.ns[0]); break;
            case FX_DISTORTION: src = std::make_unique<Distortion>(std::move(src), e.f32[0], e.f32[1]); break;
            case FX_LINEAR_RAMP:
                src = std::make_unique<LinearGainRamp>(std::move(src), e.ns[0], e.f32[0], e.f32[1], e.u32[0] != 0);
                break;
            case FX_TAKE_DURATION: src = std::make_unique<TakeDuration>(std::move(src), e.ns[0], e.u32[0] != 0); break;
            case FX_PAUSE: src = std::make_unique<Pausable>(std::move(src), e.ns[0], e.ns[1]); break;
            case FX_MIX: {   // Source::mix(other): the second input is another ro_stream, its address in u32[1] (low) / u32[2] (high)
                const ro_stream* o = (const ro_stream*)(uintptr_t)(((uint64_t)e.u32[2] << 32) | (uint64_t)e.u32[1]);
                Src other = build(*o);
                if (!other) return nullptr;
                src = Mix::make(std::move(src), std::move(other));
                break;
            }
            default: return nullptr;
        }
    }
    return src;
}
static Src build(const ro_stream& s) {
    // effects[0] == FX_SIGNAL: the stream's source is SignalGenerator::new(rate, f32[0], u32[0]).take(ns[0]) instead of PCM
    if (s.n_effects && s.effects[0].kind == FX_SIGNAL) {
        const ro_effect& e = s.effects[0];
        Src gen = std::make_unique<TakeN>(std::make_unique<SignalGenerator>(s.sample_rate, e.f32[0], (SignalGenerator::Fn)e.u32[0]), (size_t)e.ns[0]);
        return apply_effects(std::move(gen), s.effects + 1, s.n_effects - 1);
    }
    // leading FX_APPEND adapters: source::from_iter([this buffer, the appended buffers ...]) (src/source/from_iter.rs:16-27)
    uint32_t n_app = 0;
    while (n_app < s.n_effects && s.effects[n_app].kind == FX_APPEND) n_app++;
    if (n_app) {
        std::vector<Src> parts;
        parts.push_back(make_input(s));
        for (uint32_t i = 0; i < n_app; i++) {
            const ro_effect& e = s.effects[i];
            const ro_stream* o = (const ro_stream*)(uintptr_t)(((uint64_t)e.u32[2] << 32) | (uint64_t)e.u32[1]);
            parts.push_back(make_input(*o));
        }
        return apply_effects(std::make_unique<FromIter>(std::move(parts)), s.effects + n_app, s.n_effects - n_app);
    }
    return apply_effects(make_input(s), s.effects, s.n_effects);
}

static int drain(Source& src, float* out, uint64_t cap, uint64_t* n_out) {
    uint64_t n = 0;
    while (true) {
        auto v = src.next();
        if (!v) break;
        if (n < cap) out[n] = *v;
        n++;
    }
    *n_out = n;
    return n <= cap ? 0 : 8;
}

// SampleRateConverter::new(input, from, to, channels).collect()
int ro_sample_rate_converter(const float* in, uint64_t n, uint32_t from, uint32_t to, uint16_t ch, float* out,
                             uint64_t cap, uint64_t* n_out) {
    struct VecIt {
        const float* p;
        uint64_t n, i = 0;
        std::optional<Sample> next() {
            if (i >= n) return std::nullopt;
            return p[i++];
        }
    };
    SampleRateConverter<VecIt> c(VecIt{in, n}, from, to, ch);
    uint64_t k = 0;
    while (true) {
        auto v = c.next();
        if (!v) break;
        if (k < cap) out[k] = *v;
        k++;
    }
    *n_out = k;
    return k <= cap ? 0 : 8;
}
// ChannelCountConverter::new(input, from, to).collect()
int ro_channel_count_converter(const float* in, uint64_t n, uint16_t from, uint16_t to, float* out, uint64_t cap,
                               uint64_t* n_out) {
    struct VecIt {
        const float* p;
        uint64_t n, i = 0;
        std::optional<Sample> next() {
            if (i >= n) return std::nullopt;
            return p[i++];
        }
    } it{in, n};
    ChannelCountConverter c(from, to);
    uint64_t k = 0;
    while (true) {
        auto v = c.next(it);
        if (!v) break;
        if (k < cap) out[k] = *v;
        k++;
    }
    *n_out = k;
    return k <= cap ? 0 : 8;
}
// SignalGenerator::new(rate, freq, fn).take(n)
void ro_signal(int fn, uint32_t rate, float freq, uint64_t n, float* out) {
    SignalGenerator g(rate, freq, (SignalGenerator::Fn)fn);
    for (uint64_t i = 0; i < n; i++) out[i] = *g.next();
}
// The stream's own chain drained (no mixer wrap); reports final channels / rate.
int ro_chain(const ro_stream* s, float* out, uint64_t cap, uint64_t* n_out, uint16_t* ch_out, uint32_t* rate_out) {
    Src src = build(*s);
    if (!src) return 1;
    if (ch_out) *ch_out = src->channels();
    if (rate_out) *rate_out = src->sample_rate();
    return drain(*src, out, cap, n_out);
}
// What MixerSource pulls from this source: UniformSourceIterator::new(chain, mixer_ch, mixer_rate) drained.
int ro_chain_uniform(const ro_stream* s, uint16_t mixer_ch, uint32_t mixer_rate, float* out, uint64_t cap,
                     uint64_t* n_out) {
    Src src = build(*s);
    if (!src) return 1;
    UniformSourceIterator u(std::move(src), mixer_ch, mixer_rate);
    return drain(u, out, cap, n_out);
}
// mixer(ch, rate); sources added when the output position reaches mix_start (array order breaks ties).
// A position where the reference would yield None while later adds are still scheduled is written as 0.0
// and counted in *gaps (the block renderer has no way to say None mid-block).
int ro_mixer(const ro_stream* streams, uint64_t n_streams, uint16_t ch, uint32_t rate, float* out, uint64_t cap,
             uint64_t* n_out, uint64_t* gaps) {
    MixerSource mx(ch, rate);
    std::vector<uint64_t> order(n_streams);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint64_t a, uint64_t b) { return streams[a].mix_start < streams[b].mix_start; });
    uint64_t next_add = 0, pos = 0, g = 0;
    while (true) {
        while (next_add < n_streams && streams[order[next_add]].mix_start <= pos) {
            Src s = build(streams[order[next_add]]);
            if (!s) return 1;
            mx.add(std::move(s));
            next_add++;
        }
        auto v = mx.next();
        if (!v) {
            if (next_add >= n_streams && mx.still_pending.empty()) break;
            if (pos < cap) out[pos] = 0.0f;
            g++;
        } else if (pos < cap)
            out[pos] = *v;
        pos++;
    }
    // trailing gap samples are not output
    *n_out = pos;
    if (gaps) *gaps = g;
    return pos <= cap ? 0 : 8;
}
// CPU baseline: the same mixer drain with the streams sharded over n_threads host threads (contiguous
// index ranges, one MixerSource per shard), partial mixes added in shard order.  mix_start must be 0.
// Returns wall seconds of the drain (inputs already in RAM) in *seconds.
int ro_mixer_mt(const ro_stream* streams, uint64_t n_streams, uint16_t ch, uint32_t rate, int n_threads, float* out,
                uint64_t cap, uint64_t* n_out, double* seconds) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_streams) n_threads = (int)std::max<uint64_t>(1, n_streams);
    std::vector<std::vector<float>> partial(n_threads);
    std::vector<int> rc(n_threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        uint64_t lo = n_streams * t / n_threads, hi = n_streams * (t + 1) / n_threads;
        MixerSource mx(ch, rate);
        for (uint64_t i = lo; i < hi; i++) {
            Src s = build(streams[i]);
            if (!s) {
                rc[t] = 1;
                return;
            }
            mx.add(std::move(s));
        }
        auto& p = partial[t];
        while (true) {
            auto v = mx.next();
            if (!v) break;
            p.push_back(*v);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    uint64_t n = 0;
    for (auto& p : partial) n = std::max<uint64_t>(n, p.size());
    for (uint64_t i = 0; i < n && i < cap; i++) {
        float acc = 0.0f;
        for (auto& p : partial)
            if (i < p.size()) acc += p[i];
        out[i] = acc;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int r : rc)
        if (r) return r;
    *n_out = n;
    return n <= cap ? 0 : 8;
}

// Statically dispatched build of the bench chain shape
//   [UniformSourceIterator(ch, rate)] -> low_pass/high_pass -> amplify   (then Mixer::add's own uniform wrap)
// Returns nullptr when the stream does not have exactly that shape.
static Src build_static_cfg3(const ro_stream& s, uint16_t mix_ch, uint32_t mix_rate) {
    if (s.span_len != 0 || s.n_effects != 3) return nullptr;
    const ro_effect* e = s.effects;
    if (e[0].kind != FX_UNIFORM || (e[1].kind != FX_LOW_PASS && e[1].kind != FX_HIGH_PASS) || e[2].kind != FX_AMPLIFY)
        return nullptr;
    using U1 = UniformT<VecT>;
    using B = BltT<U1>;
    using A = AmplifyT<B>;
    using U2 = UniformT<A>;
    VecT v{s.pcm, (size_t)s.n_samples, 0, s.channels, s.sample_rate};
    U1 u1(v, (uint16_t)e[0].u32[0], e[0].u32[1]);
    B b(std::move(u1), e[1].kind == FX_HIGH_PASS, e[1].u32[0], e[1].f32[0]);
    A a{std::move(b), e[2].f32[0]};
    return std::make_unique<Boxed<U2>>(U2(std::move(a), mix_ch, mix_rate));
}

// CPU baseline proper: like ro_mixer_mt, but every source is the monomorphised chain rustc would emit
// (static dispatch inside a source, one virtual call per source per sample at the mixer).
// Falls back to the virtual classes for streams of another shape.  mix_start must be 0.
int ro_mixer_mt_static(const ro_stream* streams, uint64_t n_streams, uint16_t ch, uint32_t rate, int n_threads,
                       float* out, uint64_t cap, uint64_t* n_out, double* seconds) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_streams) n_threads = (int)std::max<uint64_t>(1, n_streams);
    std::vector<std::vector<float>> partial(n_threads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        uint64_t lo = n_streams * t / n_threads, hi = n_streams * (t + 1) / n_threads;
        MixerSourceBoxed mx(ch);
        for (uint64_t i = lo; i < hi; i++) {
            Src s = build_static_cfg3(streams[i], ch, rate);
            if (!s) s = std::make_unique<UniformSourceIterator>(build(streams[i]), ch, rate);
            mx.current_sources.push_back(std::move(s));
        }
        auto& p = partial[t];
        p.reserve(1 << 16);
        while (true) {
            auto v = mx.next();
            if (!v) break;
            p.push_back(*v);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    uint64_t n = 0;
    for (auto& p : partial) n = std::max<uint64_t>(n, p.size());
    for (uint64_t i = 0; i < n && i < cap; i++) {
        float acc = 0.0f;
        for (auto& p : partial)
            if (i < p.size()) acc += p[i];
        out[i] = acc;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    *n_out = n;
    return n <= cap ? 0 : 8;
}

// dasp_sample conversions; formats numbered like rb_sample_format.
int ro_convert(const void* in, int in_fmt, void* out, int out_fmt, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        float f;
        switch (in_fmt) {
            case 0: f = ((const float*)in)[i]; break;
            case 1: f = i16_to_f32(((const int16_t*)in)[i]); break;
            case 2: f = u16_to_f32(((const uint16_t*)in)[i]); break;
            case 3: f = i8_to_f32(((const int8_t*)in)[i]); break;
            case 4: f = u8_to_f32(((const uint8_t*)in)[i]); break;
            case 5: f = i32_to_f32(((const int32_t*)in)[i]); break;
            case 6: f = i24_to_f32(((const int32_t*)in)[i]); break;
            default: return 1;
        }
        switch (out_fmt) {
            case 0: ((float*)out)[i] = f; break;
            case 1: ((int16_t*)out)[i] = f32_to_i16(f); break;
            case 2: ((uint16_t*)out)[i] = f32_to_u16(f); break;
            case 3: ((int8_t*)out)[i] = f32_to_i8(f); break;
            case 4: ((uint8_t*)out)[i] = f32_to_u8(f); break;
            case 5: ((int32_t*)out)[i] = f32_to_i32(f); break;
            case 6: ((int32_t*)out)[i] = f32_to_i24(f); break;
            default: return 1;
        }
    }
    return 0;
}

}  // extern "C"
