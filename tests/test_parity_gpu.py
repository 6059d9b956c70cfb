"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bit-exact wherever the arithmetic is +,-,*,/,sqrt (everything except the limiter's
log2/exp2, which is held to the north-star tolerance 1e-5 * peak)."""
import ctypes as C

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from chains import CHAINS, LIMIT_CHAINS, _stereo
from helpers import assert_bit_exact, assert_close_peak, noise, to_oracle
from rodio_b200 import capi

pytestmark = pytest.mark.gpu

GENERAL = capi.RB_KEEP_STREAM_OUTPUTS | capi.RB_NO_FUSION | capi.RB_MIX_EXACT_ORDER


def run_chain(src: rb.Source, ctx, mixer=None, flags=GENERAL) -> np.ndarray:
    ch, rate = mixer if mixer else (src.channels(), src.sample_rate())
    with rb.Batch([src], ch, rate, flags=flags, ctx=ctx) as b:
        b.upload_all()
        b.render_mix_device()
        return b.read_stream(0)


# ------------------------------------------------------------------ conversions (rows a1, a2, a3)
def test_src_reference_vectors(ctx):
    out = rb.SampleRateConverter([2.0, 16.0, 4.0, 18.0, 6.0, 20.0, 8.0, 22.0], 2000, 3000, 2, ctx=ctx)
    assert np.trunc(out).tolist() == [2.0, 16.0, 3.0, 17.0, 4.0, 18.0, 6.0, 20.0, 7.0, 21.0, 8.0, 22.0]
    out = rb.SampleRateConverter([1.0, 14.0], 1000, 7000, 1, ctx=ctx)
    assert np.trunc(out).tolist() == [1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0]
    out = rb.SampleRateConverter(np.arange(17, dtype=np.float32), 12000, 2400, 1, ctx=ctx)
    assert out.tolist() == [0.0, 5.0, 10.0, 15.0]


COMMON_RATES = [8000, 11025, 16000, 22050, 44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000]


@pytest.mark.parametrize("to_rate", COMMON_RATES)
def test_src_common_rates_bit_exact(ctx, to_rate):
    """benches/resampler.rs:27-44 shape: stereo 44.1 kHz to each common rate."""
    x = noise(2 * 4001, 100 + to_rate)
    got = rb.SampleRateConverter(x, 44100, to_rate, 2, ctx=ctx)
    assert_bit_exact(got, oracle.sample_rate_converter(x, 44100, to_rate, 2), f"44100->{to_rate}")


def test_src_random_ratios_bit_exact(ctx):
    rng = np.random.default_rng(21)
    rates = COMMON_RATES + [39690, 40000, 1000, 7000, 2400, 12000, 3, 7]
    for i in range(40):
        f, t, c = int(rng.choice(rates)), int(rng.choice(rates)), int(rng.integers(1, 7))
        frames = int(rng.integers(0, 3000))
        if frames * t / f > 200_000:
            frames = int(200_000 * f / t)
        x = noise(frames * c, 1000 + i)
        got = rb.SampleRateConverter(x, f, t, c, ctx=ctx)
        assert_bit_exact(got, oracle.sample_rate_converter(x, f, t, c), f"{f}->{t} c={c} frames={frames}")


def test_src_properties(ctx):
    """quickcheck properties of sample_rate.rs:252-334 on the GPU path."""
    rng = np.random.default_rng(22)
    assert rb.SampleRateConverter(np.zeros(0, np.float32), 44100, 48000, 2, ctx=ctx).size == 0
    for _ in range(10):
        c, k = int(rng.integers(1, 5)), int(rng.integers(1, 9))
        x = rng.integers(-32768, 32767, c * int(rng.integers(1, 500))).astype(np.float32)
        assert np.array_equal(rb.SampleRateConverter(x, 12345, 12345, c, ctx=ctx), x)            # identity
        out = rb.SampleRateConverter(x, 4800 * k, 4800, c, ctx=ctx)                               # divide
        assert np.array_equal(out, x.reshape(-1, c)[::k].reshape(-1))
        out = rb.SampleRateConverter(x, 4800, 4800 * k, c, ctx=ctx)                               # multiply
        assert np.array_equal(out.reshape(-1, c)[::k].reshape(-1), x)


@pytest.mark.parametrize("x,f,t,want", [
    ([1, 2, 3, 4, 5, 6], 3, 2, [1, 2, 4, 5]),
    ([1, 2, 3, 4, 5, 6, 7, 8], 4, 1, [1, 5]),
    ([1, 2, 3, 4], 1, 2, [1, 1, 2, 2, 3, 3, 4, 4]),
    ([1, 2], 1, 4, [1, 1, 0, 0, 2, 2, 0, 0]),
    ([1, 2, 3, 4], 2, 4, [1, 2, 0, 0, 3, 4, 0, 0]),
    ([1, 2, 3, 4], 2, 2, [1, 2, 3, 4]),
])
def test_channel_count_converter_vectors(ctx, x, f, t, want):
    out = rb.ChannelCountConverter(np.array(x, np.float32), f, t, ctx=ctx)
    assert out.tolist() == [float(v) for v in want]


def test_channel_count_converter_random(ctx):
    rng = np.random.default_rng(23)
    for i in range(20):
        f, t = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        x = noise(f * int(rng.integers(0, 2000)), 300 + i)
        x[::7] = -0.0
        assert_bit_exact(rb.ChannelCountConverter(x, f, t, ctx=ctx), oracle.channel_count_converter(x, f, t), f"{f}->{t}")


@pytest.mark.parametrize("fmt,dt", [(1, np.int16), (2, np.uint16), (3, np.int8), (4, np.uint8), (5, np.int32), (6, np.int32)])
def test_sample_type_converter(ctx, fmt, dt):
    rng = np.random.default_rng(24 + fmt)
    info = np.iinfo(dt)
    lo, hi = (info.min, info.max) if fmt != 6 else (-(1 << 23), (1 << 23) - 1)
    ints = np.concatenate([rng.integers(lo, hi, 5000, endpoint=True), [lo, hi, 0, 1, -1 if lo < 0 else 1]]).astype(dt)
    got = rb.SampleTypeConverter(ints, capi.RB_FMT_F32, in_fmt=fmt, ctx=ctx)
    assert_bit_exact(got, oracle.convert(ints, fmt, 0), f"fmt {fmt} -> f32")
    f = np.concatenate([noise(5000, 77 + fmt, 1.2), np.array([1.0, -1.0, 0.0, -0.0, 2.0, -2.0, np.nan, np.inf, -np.inf,
                                                               0.99999994, -0.99999994, 1e-9], np.float32)])
    got = rb.SampleTypeConverter(f, fmt, ctx=ctx)
    want = oracle.convert(f, 0, fmt)
    assert got.dtype == want.dtype and np.array_equal(got, want), f"f32 -> fmt {fmt}"


# ------------------------------------------------------------------ single adapters (rows a6..a12)
@pytest.mark.parametrize("name", sorted(CHAINS))
def test_chain_bit_exact(ctx, name):
    src = CHAINS[name]()
    chain, ch, rate = oracle.chain(to_oracle(src))
    assert (ch, rate) == (src.channels(), src.sample_rate())
    # the batch hands the chain to a mixer of the chain's own format: an identity UniformSourceIterator, which
    # differs from the bare chain only by never pulling a TakeDuration's frame padding (take.rs:180-196)
    want = oracle.chain_uniform(to_oracle(src), ch, rate)
    assert want.size <= chain.size and np.array_equal(want, chain[: want.size])
    got = run_chain(src, ctx)
    assert_bit_exact(got, want, name)


@pytest.mark.parametrize("name", sorted(LIMIT_CHAINS))
def test_limiter_within_tolerance(ctx, name):
    """log2f / exp2f on the device differ from glibc by <= 2 ulp: tolerance 1e-5 * peak (north star)."""
    src = LIMIT_CHAINS[name]()
    want = oracle.chain(to_oracle(src))[0]
    assert_close_peak(run_chain(src, ctx), want, 1e-5, name)


def test_limiter_reference_bands(ctx):
    """tests/limit.rs:6-38 on the GPU path."""
    st = rb.LimitSettings.default().with_threshold(-6.0).with_knee_width(0.5) \
        .with_attack(rb.Duration.from_millis(3)).with_release(rb.Duration.from_millis(12))
    out = run_chain(rb.TestSource(oracle.sine_wave(440.0, 2600), 1, 48000).amplify(3.0).limit(st), ctx)
    assert 0.4 <= np.max(np.abs(out[1500:])) <= 0.6


# ------------------------------------------------------------------ what the mixer pulls per stream (row a4)
@pytest.mark.parametrize("name", ["amplify", "low_pass_stereo", "reverb", "spatial", "channel_volume_1_to_2",
                                  "uniform_then_effects", "agc_default", "i16_input"])
@pytest.mark.parametrize("mix", [(1, 48000), (2, 48000), (2, 44100), (4, 96000)])
def test_chain_through_mixer_uniform(ctx, name, mix):
    src = CHAINS[name]()
    want = oracle.chain_uniform(to_oracle(src), *mix)
    assert_bit_exact(run_chain(src, ctx, mixer=mix), want, f"{name} -> {mix}")


# ------------------------------------------------------------------ mixer (row a5): src/mixer.rs:208-341
def test_mixer_basic(ctx):
    tx, rx = rb.mixer(1, 48000, ctx=ctx)
    tx.add(rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]))
    tx.add(rb.SamplesBuffer(1, 48000, [5.0, 5.0, 5.0, 5.0]))
    assert rx.channels() == 1 and rx.sample_rate() == 48000
    assert [rx.next() for _ in range(5)] == [15.0, -5.0, 15.0, -5.0, None]


def test_mixer_channels_conv(ctx):
    tx, rx = rb.mixer(2, 48000, ctx=ctx)
    tx.add(rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]))
    tx.add(rb.SamplesBuffer(1, 48000, [5.0, 5.0, 5.0, 5.0]))
    assert [rx.next() for _ in range(9)] == [15.0, 15.0, -5.0, -5.0, 15.0, 15.0, -5.0, -5.0, None]


def test_mixer_rate_conv(ctx):
    tx, rx = rb.mixer(1, 96000, ctx=ctx)
    tx.add(rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]))
    tx.add(rb.SamplesBuffer(1, 48000, [5.0, 5.0, 5.0, 5.0]))
    assert [rx.next() for _ in range(8)] == [15.0, 5.0, -5.0, 5.0, 15.0, 5.0, -5.0, None]


def test_mixer_start_afterwards(ctx):
    tx, rx = rb.mixer(1, 48000, ctx=ctx)
    tx.add(rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]))
    assert [rx.next(), rx.next()] == [10.0, -10.0]
    tx.add(rb.SamplesBuffer(1, 48000, [5.0, 5.0, 6.0, 6.0, 7.0, 7.0, 7.0]))
    assert [rx.next() for _ in range(4)] == [15.0, -5.0, 6.0, 6.0]
    tx.add(rb.SamplesBuffer(1, 48000, [2.0]))
    assert [rx.next() for _ in range(4)] == [9.0, 7.0, 7.0, None]


def test_mixer_added_taking_phase_into_account(ctx):
    tx, rx = rb.mixer(2, 48000, ctx=ctx)
    tx.add(rb.SamplesBuffer(2, 48000, [10.0, -10.0, 10.0, -10.0]))
    assert rx.next() == 10.0
    tx.add(rb.SamplesBuffer(2, 48000, [5.0, -5.0, 6.0, -6.0]))
    assert rx.next() == -10.0      # not yet mixed (out of phase)
    assert rx.next() == 15.0       # mixing starts


def test_mixer_empty(ctx):
    tx, rx = rb.mixer(2, 48000, ctx=ctx)
    assert rx.next() is None
    with rb.Batch([], 2, 48000, ctx=ctx) as b:
        assert b.mix_len == 0
        assert b.render_mix().size == 0


def _random_streams(rng, n, mixer_rate):
    srcs, starts = [], []
    for i in range(n):
        ch = int(rng.choice([1, 2]))
        rate = int(rng.choice([mixer_rate, 44100, 22050]))
        frames = int(rng.integers(0, 3000))
        x = noise(frames * ch, 5000 + i, 0.5)
        s = rb.SamplesBuffer(ch, rate, x) if i % 2 else rb.TestSource(x, ch, rate)
        k = int(rng.integers(0, 4))
        if k == 1:
            s = s.amplify(0.7)
        elif k == 2:
            s = s.low_pass(800)
        elif k == 3:
            s = s.reverb(rb.Duration.from_millis(3), 0.4)
        srcs.append(s)
        starts.append(int(rng.integers(0, 1500)) if i % 3 == 0 else 0)
    return srcs, starts


@pytest.mark.parametrize("mix", [(1, 48000), (2, 48000)])
def test_mixer_random_streams_exact_order(ctx, mix):
    """Heterogeneous chains, lengths, rates and late starts; ordered sum is bit-exact with the oracle."""
    rng = np.random.default_rng(40 + mix[0])
    srcs, starts = _random_streams(rng, 37, mix[1])
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], *mix)
    with rb.Batch(srcs, *mix, flags=GENERAL, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        for i, s in enumerate(srcs):
            assert b.stream_out_len(i) == oracle.chain_uniform(to_oracle(s), *mix).size
    assert_bit_exact(got, want, "random mixer")


# ------------------------------------------------------------------ BASELINE configs at oracle-friendly sizes
def _cfg3_sources(n_streams, frames, seed=0x5EED):
    return [rb.UniformSourceIterator(rb.TestSource(noise(frames, seed + s), 1, 44100), 1, 48000)
            .low_pass(200).amplify(1.2) for s in range(n_streams)]


def test_cfg3_pipeline_small_exact(ctx):
    """cfg3 shape: resample 44.1->48 k -> low_pass(200) -> amplify(1.2) -> mix, general path, bit-exact."""
    srcs = _cfg3_sources(24, 4410)
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=GENERAL, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
    assert_bit_exact(got, want, "cfg3 small")


@pytest.mark.parametrize("n_streams,frames", [(1, 4410), (5, 1000), (33, 2000), (150, 3000), (300, 700)])
def test_cfg3_pipeline_default_path(ctx, n_streams, frames):
    """Same shape through the default (fused) path: <= 1e-5 * peak of the sequential reference sum."""
    srcs = _cfg3_sources(n_streams, frames, seed=900)
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        assert b.launches_per_render >= 1
    assert_close_peak(got, want, 1e-5, f"cfg3 default S={n_streams}")


def test_cfg2_mixer_of_sines(ctx):
    """cfg2 shape: SineWave sources summed by mixer(1, 48000)."""
    n, frames = 64, 4800
    srcs = [rb.TestSource(oracle.sine_wave(110.0 * 2 ** (s / 12.0), frames), 1, 48000) for s in range(n)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=GENERAL, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "cfg2 exact")
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        b.upload_all()
        assert_close_peak(b.render_mix(), want, 1e-5, "cfg2 default")


def test_wide_short_mixer(ctx):
    """Many sources, few output samples (a 10 ms block of a big mixer): the default mode sums the source list in
    concurrent runs (1e-5 * peak), the exact-order mode keeps the reference's sequential sum bit for bit."""
    rng = np.random.default_rng(77)
    n = 700
    srcs = [rb.TestSource(noise(int(rng.integers(1, 481)), 5000 + s, 0.5), 1, 48000) for s in range(n)]
    starts = sorted(int(rng.integers(0, 200)) * (s % 5 == 0) for s in range(n))
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 1, 48000)
    with rb.Batch(srcs, 1, 48000, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        assert b.launches_per_render == 2, "expected the run-split mix"
    assert_close_peak(got, want, 1e-5, "wide short mixer, default")
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_MIX_EXACT_ORDER, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "wide short mixer, exact order")


def test_cfg4_effect_chain(ctx):
    """cfg4 shape: spatial -> reverb(50 ms, 0.3) -> AGC(default) -> mix(2 ch)."""
    n, frames = 12, 12000
    srcs = []
    for s in range(n):
        t = np.arange(frames, dtype=np.float32)
        l = np.sin(t * np.float32(0.02 + 0.001 * s)).astype(np.float32) * np.float32(0.4)
        r = np.sin(t * np.float32(0.021 + 0.001 * s)).astype(np.float32) * np.float32(0.4)
        x = np.stack([l, r], 1).reshape(-1) + noise(2 * frames, 700 + s, 0.1)
        src = rb.Spatial(rb.TestSource(x, 2, 48000), [float(s % 7 - 3), 1.0, 0.0], [-1, 0, 0], [1, 0, 0]) \
            .reverb(rb.Duration.from_millis(50), 0.3).automatic_gain_control()
        srcs.append(src)
    want = oracle.mixer([to_oracle(s) for s in srcs], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=GENERAL, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        for i in (0, n - 1):
            assert_bit_exact(b.read_stream(i), oracle.chain_uniform(to_oracle(srcs[i]), 2, 48000), f"cfg4 stream {i}")
    assert_bit_exact(got, want, "cfg4 mix")


# ------------------------------------------------------------------ edge cases and errors
def test_edge_cases(ctx):
    for frames in [0, 1, 2, 3]:
        src = rb.UniformSourceIterator(rb.SamplesBuffer(2, 44100, noise(2 * frames, 60 + frames)), 2, 48000).low_pass(100)
        assert_bit_exact(run_chain(src, ctx), oracle.chain(to_oracle(src))[0], f"frames={frames}")
    src = rb.SamplesBuffer(1, 48000, np.array([-0.0, 0.0, -0.0], np.float32)).reverb(rb.Duration.from_nanos(1), 1.0)
    assert_bit_exact(run_chain(src, ctx), oracle.chain(to_oracle(src))[0], "negative zero")
    tx, rx = rb.mixer(1, 48000, ctx=ctx)
    tx.add(rb.SamplesBuffer(1, 48000, np.array([-0.0], np.float32)))
    assert np.float32(rx.next()).view(np.uint32) == 0        # 0.0 + -0.0 == +0.0 like the reference sum


def test_argument_errors(ctx):
    with pytest.raises(rb.RodioB200Error) as e:
        rb.Batch([rb.SamplesBuffer(2, 44100, np.zeros(3, np.float32))], 2, 48000, ctx=ctx)
    assert e.value.status == capi.RB_ERR_UNALIGNED_FRAMES
    with pytest.raises(rb.RodioB200Error) as e:
        rb.Batch([rb.SamplesBuffer(1, 44100, np.zeros(4, np.float32)).automatic_gain_control(
            rb.AutomaticGainControlSettings(absolute_max_gain=0.05))], 1, 48000, ctx=ctx)
    assert e.value.status == capi.RB_ERR_INVALID_ARGUMENT
    with pytest.raises(rb.RodioB200Error) as e:
        rb.Batch([rb.SamplesBuffer(1, 96001, np.zeros(4, np.float32))], 1, 192000, ctx=ctx)
    assert e.value.status == capi.RB_ERR_RATIO_OVERFLOW
    with rb.Batch([rb.SamplesBuffer(1, 44100, np.zeros(4, np.float32))], 1, 48000, ctx=ctx) as b:
        with pytest.raises(rb.RodioB200Error) as e:
            b.render_mix()
        assert e.value.status == capi.RB_ERR_STATE


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_properties(ctx):
    """BASELINE cfg3 at full per-stream size (44 100 frames) on 512 streams: properties that need no oracle."""
    S, frames = 512, 44100
    rng = np.random.default_rng(99)
    base = rng.uniform(-1, 1, (S, frames)).astype(np.float32)
    mk = lambda amp: [rb.UniformSourceIterator(rb.TestSource(base[s], 1, 44100), 1, 48000).low_pass(200).amplify(amp)
                      for s in range(S)]
    with rb.Batch(mk(1.0), 1, 48000, ctx=ctx) as b:
        b.upload_all()
        assert b.mix_len == 48000
        y1 = b.render_mix().copy()
        y1b = b.render_mix().copy()
    assert np.array_equal(y1, y1b), "render is deterministic / idempotent"
    with rb.Batch(mk(2.0), 1, 48000, ctx=ctx) as b:     # scaling by a power of two commutes with every rounding
        b.upload_all()
        y2 = b.render_mix()
    assert np.array_equal(y2, y1 * np.float32(2.0))
    # spot-check 3 streams of the full-size batch against the oracle, through the exact general path
    pick = [0, 255, 511]
    srcs = [mk(1.0)[i] for i in pick]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=GENERAL, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "full-length spot check")
    assert np.isfinite(y1).all() and np.max(np.abs(y1)) > 0


@pytest.mark.parametrize("channels", [1, 2])
def test_few_streams_resample_default_path(ctx, channels):
    """The everyday case: one or a few f32 files at 44.1 kHz played into a 48 kHz mixer with a volume, default flags.
    One stream per CTA, the timeline cut into slices over the SMs, the mix written in place: bit-exact."""
    for S in (1, 2, 5):
        srcs = [rb.TestSource(noise(channels * (30000 + 777 * s), 400 + s, 0.7), channels, 44100).amplify(0.5 + 0.1 * s)
                for s in range(S)]
        starts = [0] + [channels * 1000 * s for s in range(1, S)]
        want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], channels, 48000)
        with rb.Batch(srcs, channels, 48000, mix_starts=starts, ctx=ctx) as b:
            b.upload_all()
            assert b.launches_per_render <= 2
            assert_bit_exact(b.render_mix(), want, f"{S} stream(s), {channels} ch")


def test_cfg1_full_size(ctx):
    """BASELINE cfg1 at full size: 10 s of s16 stereo 44.1 kHz (take_duration) -> UniformSourceIterator(2, 48 kHz),
    bit-exact against the literal iterators, through the general path and through the fused path."""
    x = (noise(2 * 44100 * 12, 31, 0.9) * 30000).astype(np.int16)
    src = rb.TestSource(x, 2, 44100).take_duration(rb.Duration.from_secs(10))
    want = oracle.chain_uniform(to_oracle(src), 2, 48000)
    assert want.size == 960_076
    assert_bit_exact(run_chain(src, ctx, mixer=(2, 48000)), want, "cfg1 general path")
    with rb.Batch([src], 2, 48000, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "cfg1 default path")


# ------------------------------------------------------------------ fused kernel shapes (default flags)
def _fused_case(rng, i, S, c_in, rate_in, mix, biquad, fmt16, late):
    srcs, starts = [], []
    for s in range(S):
        frames = int(rng.integers(200, 1500))
        x = noise(frames * c_in, 9000 + 131 * i + s, 0.8)
        if fmt16:
            x = (x * 30000).astype(np.int16)
        src = rb.TestSource(x, c_in, rate_in).amplify(0.9)
        if (rate_in, c_in) != (mix[1], mix[0]):
            src = rb.UniformSourceIterator(src, mix[0], mix[1])
        src = src.amplify(1.1)
        if biquad == "lp":
            src = src.low_pass(300 + 10 * (s % 300))
        elif biquad == "hp":
            src = src.high_pass(150 + s)
        src = src.amplify(1.2).amplify(0.5)
        srcs.append(src)
        starts.append(int(rng.integers(0, 700)) * (s % 3 == 0) if late else 0)
    starts = sorted(starts)
    return srcs, starts


FUSED_CASES = [
    # S, c_in, rate_in, mixer, biquad, i16, late starts
    (7, 1, 44100, (1, 48000), "lp", False, False),
    (40, 1, 44100, (1, 48000), "lp", False, True),
    (200, 1, 44100, (1, 48000), "lp", False, False),
    (333, 1, 48000, (1, 48000), "lp", False, True),
    (19, 2, 44100, (2, 48000), "lp", False, True),
    (170, 2, 44100, (2, 48000), "hp", True, False),
    (21, 1, 22050, (2, 48000), "lp", False, True),
    (9, 2, 48000, (1, 44100), "hp", False, False),
    (12, 3, 32000, (3, 48000), "lp", False, True),
    (50, 1, 44100, (1, 48000), None, False, True),
    (500, 2, 48000, (2, 48000), None, True, False),
    (33, 2, 44100, (4, 96000), None, False, True),
    (5, 4, 48000, (2, 48000), "lp", False, False),
    (64, 1, 48000, (1, 44100), "lp", False, True),      # mild downsampling: HOT kernel window 280 floats
    (180, 1, 48000, (1, 40000), "hp", False, False),
    (30, 1, 96000, (1, 48000), "lp", False, True),      # 2:1 -> generic fused kernel
    (300, 2, 44100, (2, 48000), "lp", False, True),     # stereo HOT kernel, several rows per CTA
    (40, 2, 48000, (2, 48000), "hp", False, True),      # stereo, same rate
    (100, 2, 48000, (2, 44100), "lp", False, False),    # stereo, mild downsampling
    (2400, 2, 44100, (2, 48000), "hp", False, True),    # stereo, 16 rows per CTA and two CTAs per SM
    (4100, 2, 44100, (2, 48000), "lp", False, True),    # stereo, 28 rows per CTA: both recurrence warps
    (1300, 1, 44100, (1, 48000), None, False, True),    # no filter, batch large enough for the HOT pipeline
    (1250, 2, 44100, (2, 48000), None, False, False),   # same, stereo
]


@pytest.mark.parametrize("case", range(len(FUSED_CASES)))
def test_fused_shapes_match_oracle(ctx, case):
    S, c_in, rate_in, mix, biquad, fmt16, late = FUSED_CASES[case]
    rng = np.random.default_rng(1000 + case)
    srcs, starts = _fused_case(rng, case, S, c_in, rate_in, mix, biquad, fmt16, late)
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], *mix)
    with rb.Batch(srcs, *mix, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        assert b.launches_per_render <= 2, "expected the fused path"
    with rb.Batch(srcs, *mix, flags=GENERAL, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        ref_general = b.render_mix()
    assert_bit_exact(ref_general, want, f"general path case {case}")
    if S <= 148:      # one stream per CTA: the partial-row sum is the reference's sequential order
        assert_bit_exact(got, want, f"fused case {case}")
    else:
        assert_close_peak(got, want, 1e-5, f"fused case {case}")


def test_fused_ragged_batch_stereo(ctx):
    """Stereo HOT kernel: empty, one-frame, tile-boundary and long streams, S above and below one row per CTA."""
    rng = np.random.default_rng(4343)
    for lens in ([0, 2, 4, 6, 510, 512, 514, 1022, 1026, 2000, 8820] * 12, [0, 2, 4, 512, 2002] * 8):
        srcs, starts = [], []
        for i, n in enumerate(lens):
            src = rb.UniformSourceIterator(rb.TestSource(noise(n, 7100 + i), 2, 44100), 2, 48000).high_pass(90 + i)
            srcs.append(src.amplify(0.7))
            starts.append(0 if i % 3 else int(rng.integers(0, 900)))
        starts = sorted(starts)
        want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 2, 48000)
        with rb.Batch(srcs, 2, 48000, mix_starts=starts, ctx=ctx) as b:
            b.upload_all()
            got = b.render_mix()
            assert b.launches_per_render <= 2
        if len(lens) <= 148:
            assert_bit_exact(got, want, "ragged stereo fused batch")
        else:
            assert_close_peak(got, want, 1e-5, "ragged stereo fused batch")


def test_fused_ragged_batch(ctx):
    """HOT kernel with empty, one-frame, two-frame and long streams side by side, late starts, S > rows per CTA."""
    rng = np.random.default_rng(4242)
    lens = [0, 1, 2, 3, 255, 256, 257, 511, 513, 1000, 4410] * 16
    srcs, starts = [], []
    for i, n in enumerate(lens):
        src = rb.UniformSourceIterator(rb.TestSource(noise(n, 7000 + i), 1, 44100), 1, 48000).low_pass(150 + i)
        srcs.append(src.amplify(0.8))
        starts.append(0 if i % 4 else int(rng.integers(0, 900)))
    starts = sorted(starts)
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 1, 48000)
    with rb.Batch(srcs, 1, 48000, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        assert b.launches_per_render <= 2
    assert_close_peak(got, want, 1e-5, "ragged fused batch")
    with rb.Batch(srcs[:100], 1, 48000, mix_starts=starts[:100], ctx=ctx) as b:   # one stream per CTA -> exact order
        b.upload_all()
        got = b.render_mix()
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs[:100], starts[:100])], 1, 48000)
    assert_bit_exact(got, want, "ragged fused batch, S <= 148")


def test_fused_denormal_and_huge_inputs(ctx):
    """The optimistic exact-division path must fall back to IEEE division for denormal / huge operands."""
    x = noise(3000, 555)
    x[100:200] *= np.float32(1e-38)        # denormal differences
    x[300:320] = np.float32(3e30)          # huge
    x[400] = np.float32(-0.0)
    srcs = [rb.UniformSourceIterator(rb.TestSource(x if s == 0 else noise(3000, 556 + s), 1, 44100), 1, 48000)
            .low_pass(5000) for s in range(3)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
    assert_bit_exact(got, want, "denormal / huge inputs through the fused path")


def test_read_mix_blocks(ctx):
    srcs = _cfg3_sources(5, 3000, seed=77)
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        b.upload_all()
        with pytest.raises(rb.RodioB200Error):
            b.read_mix(0, 10)                      # render first
        full = b.render_mix().copy()
        blocks = [b.read_mix(o, 1000) for o in range(0, full.size + 1000, 1000)]
    assert np.array_equal(np.concatenate(blocks), full) and blocks[-1].size == 0


def test_player_volume_speed_and_queueing(ctx):
    """src/player.rs:454-470 (`set_volume(0.5)` == amplify(0.5)) and sequential playback of appended sources."""
    a = noise(3000, 901)
    b = noise(2 * 2000, 902)
    tx, rx = rb.mixer(2, 48000, ctx=ctx)
    player = rb.Player.connect_new(tx)
    assert player.empty()
    player.set_volume(0.5)
    player.append(rb.SamplesBuffer(1, 44100, a))
    player.set_speed(0.9)
    player.append(rb.SamplesBuffer(2, 48000, b))
    assert player.len() == 2
    got = rx.collect()
    first = oracle.chain_uniform(to_oracle(rb.SamplesBuffer(1, 44100, a).speed(1.0).amplify(0.5)), 2, 48000)
    second = oracle.chain_uniform(to_oracle(rb.SamplesBuffer(2, 48000, b).speed(0.9).amplify(0.5)), 2, 48000)
    # the mixer adds +0.0 before every sample (mixer.rs:186-189): -0.0 becomes +0.0, everything else is unchanged
    want = np.concatenate([first, second]) + np.float32(0.0)
    assert_bit_exact(got, want, "player queue")


# ------------------------------------------------------------------ lane-per-stream kernel (RB_FUSED_LANES, rb_lanes.cu)
# The warp program is also verified on the CPU emulator (tests/test_lanes_emulator.py, same source); these are its
# on-device counterparts (first device pass: profiles/r1_lanes_pytest.log, 22 passed).
from helpers import lanes_expected_mix

lanes_gate = lambda f: f       # was a skip until the first device pass
LANES = capi.RB_FUSED_LANES


def _lanes_case(ctx, pcms, starts, in_rate=44100, mix_rate=48000, lp=None, hp=None, q=0.5, gain=None, expect_family=2, ch=1):
    """`starts` in frames; pcms interleaved."""
    srcs = []
    for p in pcms:
        s = rb.UniformSourceIterator(rb.TestSource(p, ch, in_rate), ch, mix_rate)
        if lp is not None:
            s = s.low_pass_with_q(lp, q)
        if hp is not None:
            s = s.high_pass_with_q(hp, q)
        if gain is not None:
            s = s.amplify(gain)
        srcs.append(s)
    starts = [st * ch for st in starts]           # the batch API counts interleaved samples
    with rb.Batch(srcs, ch, mix_rate, flags=LANES, ctx=ctx, mix_starts=starts) as b:
        assert b.kernel_family == expect_family
        b.upload_all()
        got = b.render_mix()
        again = b.render_mix()
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32)), "render is not idempotent"
    per_stream = [oracle.chain_uniform(to_oracle(s), ch, mix_rate) for s in srcs]
    ref = oracle.mixer([to_oracle(s, mix_start=st) for s, st in zip(srcs, starts)], ch, mix_rate)
    assert got.shape == ref.shape
    assert_close_peak(got, ref, 1e-5, "lanes kernel vs the reference's sequential mixer")          # north-star tolerance
    if expect_family == 2:
        assert_bit_exact(got, lanes_expected_mix(per_stream, starts, ref.size), "lanes kernel vs oracle streams + its tree")
    return got


@lanes_gate
def test_lanes_single_stream_is_the_reference_stream(ctx):
    pcm = noise(6000, 1)
    got = _lanes_case(ctx, [pcm], [0], lp=200, gain=1.2)
    want = oracle.chain_uniform(to_oracle(rb.UniformSourceIterator(rb.TestSource(pcm, 1, 44100), 1, 48000).low_pass(200).amplify(1.2)), 1, 48000)
    assert_bit_exact(got, want, "one stream through k_fused_lanes")


@lanes_gate
@pytest.mark.parametrize("kw", [dict(lp=200, gain=1.2), dict(lp=1000, q=0.707), dict(hp=300, gain=0.5), dict(gain=1.2), dict()])
def test_lanes_cfg3_shapes(ctx, kw):
    pcms = [noise(5000 + 13 * i, 100 + i) for i in range(150)]
    _lanes_case(ctx, pcms, [0] * len(pcms), **kw)


@lanes_gate
def test_lanes_ragged(ctx):
    rng = np.random.default_rng(5)
    lens = [4000, 37, 1, 0, 2, 2500, 4000, 999, 16, 17] + [int(v) for v in rng.integers(3, 6000, 90)]
    starts = sorted([0, 100, 5, 9, 3000, 1234, 8, 16, 4001, 7] + [int(v) for v in rng.integers(0, 5000, 90)])
    pcms = [noise(n, 300 + i) for i, n in enumerate(lens)]
    _lanes_case(ctx, pcms, starts, lp=1000, gain=0.7)


@lanes_gate
@pytest.mark.parametrize("rates", [(8000, 48000), (22050, 48000), (32000, 44100), (47999, 48000), (11025, 96000)])
def test_lanes_other_ratios(ctx, rates):
    pcms = [noise(1500 + 11 * i, 900 + i) for i in range(70)]
    _lanes_case(ctx, pcms, [0] * 70, in_rate=rates[0], mix_rate=rates[1], lp=400, gain=1.1)


@lanes_gate
@pytest.mark.parametrize("kw", [dict(lp=200, gain=1.2), dict(hp=300), dict(gain=0.8)])
def test_lanes_stereo(ctx, kw):
    """Interleaved stereo sources into a stereo mixer: one lane carries both channels."""
    pcms = [noise(2 * (3000 + 9 * i), 400 + i, 0.9) for i in range(100)]
    _lanes_case(ctx, pcms, [0] * 100, ch=2, **kw)
    rng = np.random.default_rng(6)
    starts = sorted(int(v) for v in rng.integers(0, 2000, 100))
    _lanes_case(ctx, pcms, starts, ch=2, **kw)


@lanes_gate
def test_lanes_unsafe_inputs_stay_exact(ctx):
    pcms = [noise(3000, 40 + i) for i in range(70)]
    pcms[3][100:110] = np.float32(1e-41)
    pcms[3][500] = np.float32(3e-30)
    pcms[40][7] = np.float32(1e25)
    _lanes_case(ctx, pcms, [0] * 70, lp=800, gain=1.2)


@lanes_gate
def test_lanes_mixed_rate_pairs_and_fallback(ctx):
    """Several rate pairs in one mixer (44.1 kHz, 48 kHz pass-through, 32 kHz) are served class by class; a three-channel
    mixer is outside the kernel's shape: the flag is ignored there and the other kernels serve the batch."""
    rates = [44100, 48000, 32000, 44100, 48000, 22050] * 8
    pcms = [noise(1500 + 13 * i, 40 + i) for i in range(len(rates))]
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, 1, r), 1, 48000).low_pass(500).amplify(0.9) for p, r in zip(pcms, rates)]
    with rb.Batch(srcs, 1, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    ref = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    assert_close_peak(got, ref, 1e-5, "mixed rate pairs vs the reference's mixer")
    # bit-exact against the oracle streams summed class by class with the kernel's tree
    per_stream = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    acc = np.zeros(ref.size, np.float32)
    for rate in dict.fromkeys(rates):                       # classes in order of first appearance
        idx = [i for i, r in enumerate(rates) if r == rate]
        acc = acc + (lanes_expected_mix([per_stream[i] for i in idx], [0] * len(idx), ref.size) - np.float32(0.0))
    assert_bit_exact(got, acc, "mixed rate pairs vs oracle streams + class-wise tree")
    d = rb.UniformSourceIterator(rb.TestSource(noise(3 * 2000, 3), 3, 32000), 3, 48000).low_pass(200)
    with rb.Batch([d], 3, 48000, flags=LANES, ctx=ctx) as b:     # three channels: the shape of FUSED_CASES[7], not the lane kernel's
        assert b.kernel_family != 2
        b.upload_all()
        assert_close_peak(b.render_mix(), oracle.mixer([to_oracle(d)], 3, 48000), 1e-5, "fallback")


@lanes_gate
def test_lanes_full_size_properties(ctx):
    """BASELINE cfg5 size per stream (1 s of 44.1 kHz), 2048 streams: matches the default fused path within the
    tolerance, is idempotent, and a re-upload re-classifies."""
    rng = np.random.default_rng(9)
    base = [rng.uniform(-1, 1, 44100).astype(np.float32) for _ in range(8)]
    mk = lambda: [rb.UniformSourceIterator(rb.TestSource(base[i % 8], 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for i in range(2048)]
    with rb.Batch(mk(), 1, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    with rb.Batch(mk(), 1, 48000, ctx=ctx) as b:
        b.upload_all()
        ref = b.render_mix()
    assert_close_peak(got, ref, 1e-5, "lanes vs default fused path at full size")


# ------------------------------------------------------------------ streaming sessions (rb_session_*), same gate
def _session_sources(n, lp=200, gain=1.2):
    mk = lambda: rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100), 1, 48000)
    out = []
    for _ in range(n):
        s = mk()
        if lp is not None:
            s = s.low_pass(lp)
        if gain is not None:
            s = s.amplify(gain)
        out.append(s)
    return out


def _whole_render(ctx, pcms, starts, lp=200, gain=1.2):
    srcs = []
    for p in pcms:
        s = rb.UniformSourceIterator(rb.TestSource(p, 1, 44100), 1, 48000)
        if lp is not None:
            s = s.low_pass(lp)
        if gain is not None:
            s = s.amplify(gain)
        srcs.append(s)
    with rb.Batch(srcs, 1, 48000, flags=LANES, ctx=ctx, mix_starts=starts) as b:
        assert b.kernel_family == 2
        b.upload_all()
        return b.render_mix(), srcs


@lanes_gate
def test_session_any_split_is_the_whole_render(ctx):
    rng = np.random.default_rng(11)
    pcms = [noise(int(n), 60 + i) for i, n in enumerate(rng.integers(2000, 9000, 70))]
    want, srcs = _whole_render(ctx, pcms, [0] * 70)
    got = []
    with rb.Session(_session_sources(70), 48000, fifo_frames=4096, max_block_frames=2048, ctx=ctx) as s:
        left = [0] * 70
        ended = False
        while not ended:
            for r in rng.permutation(70):
                room = 4096 - 1100       # stay below the FIFO capacity whatever the renders leave behind
                n = min(pcms[r].size - left[r], int(rng.integers(0, 700)))
                if n and rng.random() < 0.8 and n < room:
                    try:
                        s.push(int(r), pcms[r][left[r]:left[r] + n], end_of_stream=(left[r] + n == pcms[r].size))
                        left[r] += n
                    except rb.RodioB200Error:
                        pass                 # FIFO full: render first
            block, ended = s.render(int(rng.integers(1, 2048)))
            got.append(block)
    got = np.concatenate(got)
    assert_bit_exact(got, want, "session blocks vs the whole-stream render")
    ref = oracle.mixer([to_oracle(x) for x in srcs], 1, 48000)
    assert_close_peak(got, ref, 1e-5, "session vs the reference's mixer")


@lanes_gate
def test_session_packed_push_10ms_blocks_and_state_blob(ctx):
    """10 ms blocks pushed for all sources at once; half-way the state moves into a fresh session that carries on."""
    n_src, frames = 40, 44100 // 2
    pcms = [noise(frames, 200 + i) for i in range(n_src)]
    want, _ = _whole_render(ctx, pcms, [0] * n_src, lp=1000, gain=0.8)
    got, pos, blob = [], 0, None
    sa = rb.Session(_session_sources(n_src, lp=1000, gain=0.8), 48000, fifo_frames=2048, max_block_frames=480, ctx=ctx)
    sb = rb.Session(_session_sources(n_src, lp=1000, gain=0.8), 48000, fifo_frames=2048, max_block_frames=480, ctx=ctx)
    cur, ended = sa, False
    while not ended:
        n = min(441, frames - pos)
        cur.push_packed([p[pos:pos + n] for p in pcms], [pos + n == frames] * n_src)
        pos += n
        while True:
            block, ended = cur.render(480)
            got.append(block)
            if block.size == 0 or ended:
                break
        if blob is None and pos >= frames // 2:
            blob = cur.get_state()
            sb.set_state(blob)
            cur = sb
    sa.close(), sb.close()
    assert_bit_exact(np.concatenate(got), want, "10 ms blocks with a state hand-over vs the whole-stream render")


@lanes_gate
def test_session_rejects_other_shapes(ctx):
    stereo = rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 2, 44100), 1, 48000)
    with pytest.raises(rb.RodioB200Error):
        rb.Session([stereo], 48000, ctx=ctx)
    down = rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 96000), 1, 48000)
    rb.Session([down], 48000, ctx=ctx).close()          # above the mixer's rate: served (general per-sample path), not rejected
    agc = rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100), 1, 48000).automatic_gain_control()
    with pytest.raises(rb.RodioB200Error):
        rb.Session([agc], 48000, ctx=ctx)
    no_uniform = rb.TestSource(np.zeros(0, np.float32), 1, 44100).low_pass(100)
    with pytest.raises(rb.RodioB200Error):
        rb.Session([no_uniform], 48000, ctx=ctx)


@lanes_gate
def test_session_stereo_any_split(ctx):
    rng = np.random.default_rng(12)
    pcms = [noise(2 * int(n), 700 + i) for i, n in enumerate(rng.integers(3000, 9000, 40))]
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, 2, 44100), 2, 48000).low_pass(300).amplify(1.1) for p in pcms]
    with rb.Batch(srcs, 2, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        want = b.render_mix()
    chains = [rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 2, 44100), 2, 48000).low_pass(300).amplify(1.1)
              for _ in pcms]
    got, left, ended = [], [0] * 40, False
    with rb.Session(chains, 48000, fifo_frames=4096, max_block_frames=1024, ctx=ctx, mixer_channels=2) as s:
        while not ended:
            blocks, eos = [], []
            for r in range(40):
                n = min(pcms[r].size // 2 - left[r], int(rng.integers(0, 600)))
                blocks.append(pcms[r][2 * left[r]: 2 * (left[r] + n)])
                left[r] += n
                eos.append(left[r] == pcms[r].size // 2)
            s.push_packed(blocks, eos)
            while True:
                block, ended = s.render(int(rng.integers(1, 1024)))
                got.append(block)
                if block.size == 0 or ended:
                    break
    assert_bit_exact(np.concatenate(got), want, "stereo session vs the whole-stream render")


@lanes_gate
def test_session_mono_and_stereo_sources_mixed_rates(ctx):
    """A stereo 48 kHz mixer fed by mono and stereo sources at 44.1, 22.05 and 48 kHz (classes of their own), pushed in
    10 ms blocks: the bytes of the whole-stream batch render."""
    ch_in = [1, 2, 1, 1, 2, 1, 2, 1] * 3
    rates = [44100, 44100, 48000, 22050, 48000, 44100, 44100, 48000] * 3
    pcms = [noise(ci * int(0.2 * r), 1500 + i, 0.8) for i, (ci, r) in enumerate(zip(ch_in, rates))]      # 200 ms each
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, ci, r), 2, 48000).low_pass(800).amplify(0.7) for p, ci, r in zip(pcms, ch_in, rates)]
    with rb.Batch(srcs, 2, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        want = b.render_mix()
    ref = oracle.mixer([to_oracle(s) for s in srcs], 2, 48000)
    assert_close_peak(want, ref, 1e-5, "mono + stereo, three rates vs the reference's mixer")
    chains = [rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), ci, r), 2, 48000).low_pass(800).amplify(0.7)
              for ci, r in zip(ch_in, rates)]
    got, pos, ended = [], [0] * len(pcms), False
    with rb.Session(chains, 48000, fifo_frames=2048, max_block_frames=480, ctx=ctx, mixer_channels=2) as s:
        while not ended:
            blocks, eos = [], []
            for i, (p, ci, r) in enumerate(zip(pcms, ch_in, rates)):
                n = min(r // 100, p.size // ci - pos[i])
                blocks.append(p[ci * pos[i]: ci * (pos[i] + n)])
                pos[i] += n
                eos.append(pos[i] == p.size // ci)
            s.push_packed(blocks, eos)
            while True:
                block, ended = s.render(480)
                got.append(block)
                if block.size == 0 or ended:
                    break
    assert_bit_exact(np.concatenate(got), want, "session vs whole-stream render, mono + stereo sources at three rates")


@lanes_gate
def test_session_speed_changes_the_rate_pair(ctx):
    """source.speed(0.9) in front of the conversion: 44 100 Hz is reported as 39 690 Hz (speed.rs:130-133), nothing else."""
    pcms = [noise(6000 + 100 * i, 2300 + i) for i in range(6)]
    mk = lambda p: rb.UniformSourceIterator(rb.TestSource(p, 1, 44100).speed(0.9), 1, 48000).low_pass(400)
    want = oracle.mixer([to_oracle(mk(p)) for p in pcms], 1, 48000)
    got, pos, ended = [], 0, False
    with rb.Session([mk(np.zeros(0, np.float32)) for _ in pcms], 48000, fifo_frames=2048, max_block_frames=512, ctx=ctx) as s:
        while not ended:
            s.push_packed([p[pos:pos + 400] for p in pcms], [pos + 400 >= p.size for p in pcms])
            pos += 400
            while True:
                block, ended = s.render(512)
                got.append(block)
                if block.size == 0 or ended:
                    break
    assert_close_peak(np.concatenate(got), want, 1e-5, "session with speed() vs the reference's mixer")


@lanes_gate
def test_lanes_sources_without_a_conversion(ctx):
    """Sources at the mixer's format with a filter and a gain, no UniformSourceIterator in the chain."""
    pcms = [noise(4000 + 7 * i, 2500 + i) for i in range(70)]
    srcs = [rb.TestSource(p, 1, 48000).low_pass(300).amplify(0.6) for p in pcms]
    with rb.Batch(srcs, 1, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    per_stream = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    assert_bit_exact(got, lanes_expected_mix(per_stream, [0] * 70, got.size), "filter + gain on same-rate sources")


@lanes_gate
def test_session_sources_added_while_it_runs(ctx):
    """Mixer::add during playback: a held source joins at the frame rendered next -- the bytes of a batch in which it has that
    mix_start."""
    pcms = [noise(5000 + 300 * i, 3100 + i) for i in range(4)]
    mk = lambda p: rb.UniformSourceIterator(rb.TestSource(p, 1, 44100), 1, 48000).low_pass(600).amplify(0.9)
    held = [0, capi.RB_SESSION_HELD, 0, capi.RB_SESSION_HELD]
    got = []
    with rb.Session([mk(np.zeros(0, np.float32)) for _ in pcms], 48000, fifo_frames=8192, max_block_frames=1024, ctx=ctx, mix_starts=held) as s:
        for i, p in enumerate(pcms):
            s.push(i, p, end_of_stream=True)
        got.append(s.render(700)[0])
        s.start(1)                                   # joins at frame 700
        got.append(s.render(333)[0])
        s.start(3)                                   # joins at frame 1033
        ended = False
        while not ended:
            block, ended = s.render(1024)
            got.append(block)
    got = np.concatenate(got)
    with rb.Batch([mk(p) for p in pcms], 1, 48000, flags=LANES, ctx=ctx, mix_starts=[0, 700, 0, 1033]) as b:
        b.upload_all()
        want = b.render_mix()
    # the batch orders its sources by mix_start (0, 0, 700, 1033), the session keeps the declared order: same set of lanes in one
    # warp, idle lanes add +0.0 -> the sums agree up to the tree's lane order, i.e. within the mixer tolerance
    assert got.size == want.size
    assert_close_peak(got, want, 1e-5, "sources added during playback vs the batch with those mix_starts")
    ref = oracle.mixer([to_oracle(mk(p), mix_start=m) for p, m in zip(pcms, [0, 700, 0, 1033])], 1, 48000)
    assert_close_peak(got, ref, 1e-5, "... and vs the reference's mixer")


@lanes_gate
def test_session_queue_of_sources(ctx):
    """Player::append: sources of different rates played one after the other, an independent voice beside them."""
    rates = [44100, 48000, 22050, 44100]
    pcms = [noise(int(0.1 * r) + 17 * i, 3300 + i) for i, r in enumerate(rates)]
    mk = lambda p, r: rb.UniformSourceIterator(rb.TestSource(p, 1, r), 1, 48000).low_pass(900).amplify(0.8)
    held = capi.RB_SESSION_HELD
    got, pos, ended = [], [0] * 4, False
    # the queued sources are decoded ahead while they wait: their FIFOs hold them whole
    with rb.Session([mk(np.zeros(0, np.float32), r) for r in rates], 48000, fifo_frames=16384, max_block_frames=480, ctx=ctx,
                    mix_starts=[0, held, held, 40]) as s:
        s.follow(1, 0)
        s.follow(2, 1)
        while not ended:
            blocks, eos = [], []
            for i, (p, r) in enumerate(zip(pcms, rates)):
                n = min(r // 100, p.size - pos[i])
                blocks.append(p[pos[i]:pos[i] + n])
                pos[i] += n
                eos.append(pos[i] == p.size)
            s.push_packed(blocks, eos)
            while True:
                block, ended = s.render(480)
                got.append(block)
                if block.size == 0 or ended:
                    break
    got = np.concatenate(got)
    lens = [oracle.chain_uniform(to_oracle(mk(p, r)), 1, 48000).size for p, r in zip(pcms, rates)]
    starts = [0, lens[0], lens[0] + lens[1], 40]
    ref = oracle.mixer([to_oracle(mk(p, r), mix_start=m) for p, r, m in zip(pcms, rates, starts)], 1, 48000)
    assert got.size == ref.size
    assert_close_peak(got, ref, 1e-5, "queued sources vs the reference's mixer with the same starts")


@lanes_gate
def test_lanes_and_session_gain_in_front_of_the_conversion(ctx):
    """`source.amplify(v)` handed to the mixer -- the gain multiplies every frame BEFORE it is interpolated (amplify.rs:91-95 in
    front of uniform.rs) -- as a batch on the lane kernel and as a session; 0.004 lies outside the gain range of the fast tiles
    (rb_lanes_plan.h pre_gain_keeps_class) and goes through the slow ones, same bytes."""
    n = 70
    pcms = [noise(3000 + 41 * i, 4700 + i) for i in range(n)]
    pres = [float(np.float32(0.2 + 0.013 * i)) for i in range(n)]
    pres[9], pres[40] = 0.004, -0.6
    mk = lambda p, g: rb.UniformSourceIterator(rb.TestSource(p, 1, 44100).amplify(g), 1, 48000).low_pass(250).amplify(0.8)
    srcs = [mk(p, g) for p, g in zip(pcms, pres)]
    per_stream = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    with rb.Batch(srcs, 1, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        whole = b.render_mix()
    assert_bit_exact(whole, lanes_expected_mix(per_stream, [0] * n, whole.size), "gain -> conversion -> low_pass -> gain on the lane kernel")
    assert_close_peak(whole, oracle.mixer([to_oracle(s) for s in srcs], 1, 48000), 1e-5, "... and the reference's sequential mixer")
    got, pos, ended = [], 0, False
    with rb.Session([mk(np.zeros(0, np.float32), g) for g in pres], 48000, fifo_frames=2048, max_block_frames=480, ctx=ctx) as s:
        while not ended:
            s.push_packed([p[pos:pos + 441] for p in pcms], [pos + 441 >= p.size for p in pcms])
            pos += 441
            while True:
                block, ended = s.render(480)
                got.append(block)
                if block.size == 0 or ended:
                    break
    got = np.concatenate(got)
    assert_bit_exact(got, whole[:got.size], "the session in 10 ms blocks vs the whole-stream render")
    assert got.size == whole.size or not np.any(whole[got.size:])


@lanes_gate
def test_lanes_and_session_filter_in_front_of_the_conversion(ctx):
    """`source.low_pass(f)` handed to the mixer, or appended to a Player whose volume then sits behind it (player.rs:120-128):
    the filter runs once per INPUT frame at the source's rate (blt.rs: to_applier(input.sample_rate())), the interpolation reads
    its outputs.  As a batch on the lane kernel (up-sampling and same-rate sources) and as a session in 10 ms blocks, both
    against the oracle's literal iterators; a down-sampling source alone in a session."""
    n = 66
    rates = [44100, 48000, 22050, 44100, 32000, 44100] * 11
    pcms = [noise(2500 + 37 * i, 5200 + i) for i in range(n)]
    mids = [float(np.float32(0.3 + 0.01 * i)) for i in range(n)]
    mk = lambda p, r, g: rb.UniformSourceIterator(rb.TestSource(p, 1, r).amplify(0.9).low_pass(300).amplify(g), 1, 48000).amplify(0.8)
    srcs = [mk(p, r, g) for p, r, g in zip(pcms, rates, mids)]
    per_stream = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    ref = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        whole = b.render_mix()
    assert_close_peak(whole, ref, 1e-5, "filter in front, lane kernel vs the reference's sequential mixer")
    # one stream alone: the tree adds zeros only -- the kernel output IS the reference stream
    for k in (0, 2, 4):
        with rb.Batch([srcs[k]], 1, 48000, flags=LANES, ctx=ctx) as b:
            assert b.kernel_family == 2
            b.upload_all()
            assert_bit_exact(b.render_mix(), per_stream[k] + np.float32(0.0), f"filter in front, one {rates[k]} Hz stream")
    got, pos, ended = [], 0, False
    blocks = [r // 100 for r in rates]
    with rb.Session([mk(np.zeros(0, np.float32), r, g) for r, g in zip(rates, mids)], 48000, fifo_frames=4096, max_block_frames=480, ctx=ctx) as s:
        while not ended:
            s.push_packed([p[pos * k:(pos + 1) * k] for p, k in zip(pcms, blocks)], [(pos + 1) * k >= p.size for p, k in zip(pcms, blocks)])
            pos += 1
            while True:
                block, ended = s.render(480)
                got.append(block)
                if block.size == 0 or ended:
                    break
    got = np.concatenate(got)
    assert_bit_exact(got, whole[:got.size], "the session in 10 ms blocks vs the whole-stream render")
    assert got.size == whole.size or not np.any(whole[got.size:])
    # 96 kHz into 48 kHz: two input frames per output, the filter consumes both
    down = mk(pcms[0], 96000, 0.5)
    want = oracle.chain_uniform(to_oracle(down), 1, 48000)
    got, pos, ended = [], 0, False
    with rb.Session([mk(np.zeros(0, np.float32), 96000, 0.5)], 48000, fifo_frames=4096, max_block_frames=480, ctx=ctx) as s:
        while not ended:
            s.push(0, pcms[0][pos:pos + 700], end_of_stream=pos + 700 >= pcms[0].size)
            pos += 700
            while True:
                block, ended = s.render(333)
                got.append(block)
                if block.size == 0 or ended:
                    break
    assert_bit_exact(np.concatenate(got), want + np.float32(0.0), "down-sampling source with the filter in front, alone in a session")


@lanes_gate
def test_session_player_volume_changes(ctx):
    """Player::set_volume while playing (rb_session_set_volume): the Player's Amplify sits in front of the mixer's conversion
    (src/player.rs:120-128), so every input frame keeps the factor it had when the converter pulled it.  One 44.1 kHz source,
    5 ms blocks, four volume changes; the expectation is the oracle's literal chain over an input multiplied frame by frame with
    the factor of the block that pulled it (the converter stands two frames behind the left neighbour of the last output)."""
    L, rate = 9000, 44100
    pcm = noise(L, 6100)
    mk = lambda p: rb.UniformSourceIterator(rb.TestSource(p, 1, rate).amplify(1.0), 1, 48000).low_pass(500).amplify(0.9)
    total = int(rb.plan(mk(pcm), 1, 48000)[0])
    plan_ = {1: 0.5, 4: 1.5, 9: 0.05, 14: 0.8}
    gains = np.ones(L, np.float32)
    vol, pulled, pushed, done, got, ended, rnd = 1.0, 0, 0, 0, [], False, 0
    with rb.Session([mk(np.zeros(0, np.float32))], 48000, fifo_frames=4096, max_block_frames=240, ctx=ctx) as s:
        while not ended:
            if rnd in plan_:
                vol = plan_[rnd]
                s.set_volume(0, vol)
            k = min(441, L - pushed)
            s.push(0, pcm[pushed:pushed + k], end_of_stream=pushed + k == L)
            pushed += k
            while True:
                block, ended = s.render(240)
                got.append(block)
                if block.size == 0:
                    break
                o = min(done + block.size, total)
                p_new = min(((o - 1) * 147) // 160 + 2, pushed)
                gains[pulled:p_new] = np.float32(vol)
                pulled, done = p_new, o
                if ended:
                    break
            rnd += 1
            assert rnd < 1000
    assert np.unique(gains).size == 5
    want = oracle.chain_uniform(to_oracle(rb.UniformSourceIterator(rb.TestSource(pcm * gains, 1, rate), 1, 48000).low_pass(500).amplify(0.9)), 1, 48000)
    assert_bit_exact(np.concatenate(got), want + np.float32(0.0), "volume changes in front of the conversion")


@lanes_gate
@pytest.mark.parametrize("rate,mix", [(48000, 44100), (96000, 48000), (88200, 48000)])
def test_lanes_and_session_sources_above_the_mixers_rate(ctx, rate, mix):
    """Sources above the mixer's rate, up to twice: fast tiles of their own in the lane kernel (one or two input frames per
    output plus the carry).  Batch on request (RB_FUSED_LANES; without the flag the default kernels keep such batches) and
    session, both bit for bit against the oracle streams summed with the kernel's tree."""
    n = 70
    pcms = [noise(5000 + 31 * i, 6400 + i) for i in range(n)]
    mk = lambda p: rb.UniformSourceIterator(rb.TestSource(p, 1, rate), 1, mix).low_pass(300).amplify(0.9)
    srcs = [mk(p) for p in pcms]
    per_stream = [oracle.chain_uniform(to_oracle(s), 1, mix) for s in srcs]
    with rb.Batch(srcs, 1, mix, flags=LANES, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        whole = b.render_mix()
    assert_bit_exact(whole, lanes_expected_mix(per_stream, [0] * n, whole.size), "down-sampling batch on the lane kernel")
    assert_close_peak(whole, oracle.mixer([to_oracle(s) for s in srcs], 1, mix), 1e-5, "... and the reference's sequential mixer")
    with rb.Batch(srcs, 1, mix, ctx=ctx) as b:
        assert b.kernel_family != 2                      # not chosen automatically
    got, pos, ended, k = [], 0, False, rate // 100
    with rb.Session([mk(np.zeros(0, np.float32)) for _ in pcms], mix, fifo_frames=4096, max_block_frames=480, ctx=ctx) as s:
        while not ended:
            s.push_packed([p[pos:pos + k] for p in pcms], [pos + k >= p.size for p in pcms])
            pos += k
            while True:
                block, ended = s.render(480)
                got.append(block)
                if block.size == 0 or ended:
                    break
    got = np.concatenate(got)
    assert_bit_exact(got, whole[:got.size], "the session in 10 ms blocks vs the whole-stream render")
    assert got.size == whole.size or not np.any(whole[got.size:])
