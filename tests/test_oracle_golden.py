"""Pins the CPU oracle against every golden vector / known-answer test the reference holds for the
path.  CPU only.  Each test names the reference test it restates."""
import math

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import to_oracle


# ---- src/conversions/sample_rate.rs:356-387 -------------------------------------------------
def test_src_upsample():
    x = np.array([2.0, 16.0, 4.0, 18.0, 6.0, 20.0, 8.0, 22.0], dtype=np.float32)
    out = oracle.sample_rate_converter(x, 2000, 3000, 2)
    assert out.size == 12
    assert np.trunc(out).tolist() == [2.0, 16.0, 3.0, 17.0, 4.0, 18.0, 6.0, 20.0, 7.0, 21.0, 8.0, 22.0]


def test_src_upsample2():
    out = oracle.sample_rate_converter(np.array([1.0, 14.0], dtype=np.float32), 1000, 7000, 1)
    assert np.trunc(out).tolist() == [1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0]


def test_src_downsample():
    out = oracle.sample_rate_converter(np.arange(17, dtype=np.float32), 12000, 2400, 1)
    assert out.tolist() == [0.0, 5.0, 10.0, 15.0]


# ---- quickcheck properties, src/conversions/sample_rate.rs:252-334 --------------------------
def test_src_empty():
    rng = np.random.default_rng(1)
    for _ in range(50):
        f, t, c = int(rng.integers(1, 768000)), int(rng.integers(1, 768000)), int(rng.integers(1, 12))
        assert oracle.sample_rate_converter(np.zeros(0, np.float32), f, t, c).size == 0


def test_src_identity():
    rng = np.random.default_rng(2)
    for _ in range(30):
        f, c = int(rng.integers(1, 400000)), int(rng.integers(1, 9))
        x = rng.integers(-32768, 32767, int(rng.integers(0, 200))).astype(np.float32)
        assert np.array_equal(oracle.sample_rate_converter(x, f, f, c), x)


def test_src_divide_sample_rate():
    rng = np.random.default_rng(3)
    for _ in range(40):
        to, k, c = int(rng.integers(1, 48000)), int(rng.integers(1, 12)), int(rng.integers(1, 6))
        x = rng.integers(-32768, 32767, int(rng.integers(0, 300))).astype(np.float32)
        x = x[: c * (x.size // c)]
        out = oracle.sample_rate_converter(x, to * k, to, c)
        want = x.reshape(-1, c)[::k].reshape(-1)
        assert np.array_equal(out, want)


def test_src_multiply_sample_rate():
    rng = np.random.default_rng(4)
    for _ in range(40):
        f, k, c = int(rng.integers(1, 65535)), int(rng.integers(1, 12)), int(rng.integers(1, 6))
        x = rng.integers(-32768, 32767, int(rng.integers(0, 120))).astype(np.float32)
        x = x[: c * (x.size // c)]
        out = oracle.sample_rate_converter(x, f, f * k, c)
        got = out[: c * (out.size // c)].reshape(-1, c)[::k].reshape(-1)
        assert np.array_equal(got, x)


# ---- src/conversions/channels.rs:114-177 ----------------------------------------------------
@pytest.mark.parametrize("x,f,t,want", [
    ([1, 2, 3, 4, 5, 6], 3, 2, [1, 2, 4, 5]),
    ([1, 2, 3, 4, 5, 6, 7, 8], 4, 1, [1, 5]),
    ([1, 2, 3, 4], 1, 2, [1, 1, 2, 2, 3, 3, 4, 4]),
    ([1, 2], 1, 4, [1, 1, 0, 0, 2, 2, 0, 0]),
    ([1, 2, 3, 4], 2, 4, [1, 2, 0, 0, 3, 4, 0, 0]),
])
def test_channel_count_converter(x, f, t, want):
    out = oracle.channel_count_converter(np.array(x, dtype=np.float32), f, t)
    assert out.tolist() == [float(v) for v in want]


def test_channel_len():
    assert oracle.channel_count_converter(np.array([1, 2, 3, 4], np.float32), 2, 3).size == 6
    assert oracle.channel_count_converter(np.array([1, 2, 3, 4], np.float32), 2, 1).size == 2


# ---- src/math.rs:187-339 --------------------------------------------------------------------
def test_lerp_random():
    rng = np.random.default_rng(5)
    n = 0
    while n < 2000:
        a, b = np.float32(rng.uniform(-1, 1)), np.float32(rng.uniform(-1, 1))
        num, den = int(rng.integers(0, 2000)), int(rng.integers(1, 4000))
        g = math.gcd(num, den)
        num, den = num // g, den // g
        c = num / den
        if num > 1000 or not (0.0 <= c <= 1.0):
            continue
        ref = float(a) * (1.0 - c) + float(b) * c
        assert abs(float(oracle.lerp(a, b, num, den)) - ref) < 1e-6
        n += 1


DB_TABLE = [(100., 100000.), (90., 31623.), (80., 10000.), (70., 3162.), (60., 1000.), (50., 316.2), (40., 100.),
            (30., 31.62), (20., 10.), (10., 3.162), (5.998, 1.995), (3.003, 1.413), (1.002, 1.122), (0., 1.),
            (-1.002, 0.891), (-3.003, 0.708), (-5.998, 0.501), (-10., 0.3162), (-20., 0.1), (-30., 0.03162),
            (-40., 0.01), (-50., 0.003162), (-60., 0.001), (-70., 0.0003162), (-80., 0.0001), (-90., 0.00003162),
            (-100., 0.00001)]


def test_db_tables():
    for db, lin in DB_TABLE:
        assert 0.99 < oracle.db_to_linear(db) / lin < 1.01
        if abs(db) > 1e-5:
            assert 0.99 < oracle.linear_to_db(lin) / db < 1.01


def test_db_round_trip():
    eps = float(np.finfo(np.float32).eps)
    for db in [-60.0, -20.0, -6.0, 0.0, 6.0, 20.0, 40.0]:
        assert abs(oracle.linear_to_db(oracle.db_to_linear(db)) - db) < 16 * eps
    for lin in [0.001, 0.1, 1.0, 10.0, 100.0]:
        assert abs((oracle.db_to_linear(oracle.linear_to_db(lin)) - lin) / lin) < 16 * eps


# ---- src/mixer.rs:208-341 -------------------------------------------------------------------
def _sb(ch, rate, data, start=0):
    return oracle.Stream(np.array(data, np.float32), ch, rate, [], span_len=len(data), mix_start=start)


def test_mixer_basic():
    out = oracle.mixer([_sb(1, 48000, [10, -10, 10, -10]), _sb(1, 48000, [5, 5, 5, 5])], 1, 48000)
    assert out.tolist() == [15.0, -5.0, 15.0, -5.0]


def test_mixer_channels_conv():
    out = oracle.mixer([_sb(1, 48000, [10, -10, 10, -10]), _sb(1, 48000, [5, 5, 5, 5])], 2, 48000)
    assert out.tolist() == [15.0, 15.0, -5.0, -5.0, 15.0, 15.0, -5.0, -5.0]


def test_mixer_rate_conv():
    out = oracle.mixer([_sb(1, 48000, [10, -10, 10, -10]), _sb(1, 48000, [5, 5, 5, 5])], 1, 96000)
    assert out.tolist() == [15.0, 5.0, -5.0, 5.0, 15.0, 5.0, -5.0]


def test_mixer_start_afterwards():
    out = oracle.mixer([_sb(1, 48000, [10, -10, 10, -10]), _sb(1, 48000, [5, 5, 6, 6, 7, 7, 7], start=2),
                        _sb(1, 48000, [2], start=6)], 1, 48000)
    assert out.tolist() == [10.0, -10.0, 15.0, -5.0, 6.0, 6.0, 9.0, 7.0, 7.0]


def test_mixer_added_taking_phase_into_account():
    out = oracle.mixer([_sb(2, 48000, [10, -10, 10, -10]), _sb(2, 48000, [5, -5, 6, -6], start=1)], 2, 48000)
    assert out[:3].tolist() == [10.0, -10.0, 15.0]


# ---- src/source/channel_volume.rs:135-166 ---------------------------------------------------
def test_channel_volume_vectors():
    f32 = np.float32
    s = rb.ChannelVolume(rb.SamplesBuffer(1, 44100, [1.0, 2.0, 3.0]), [0.5, 0.8])
    out, ch, _ = oracle.chain(to_oracle(s))
    assert ch == 2
    assert out.tolist() == [float(f32(1) * f32(0.5)), float(f32(1) * f32(0.8)), float(f32(2) * f32(0.5)),
                            float(f32(2) * f32(0.8)), float(f32(3) * f32(0.5)), float(f32(3) * f32(0.8))]
    s = rb.ChannelVolume(rb.SamplesBuffer(2, 44100, [1.0, 2.0, 3.0, 4.0]), [1.0])
    assert oracle.chain(to_oracle(s))[0].tolist() == [1.5, 3.5]
    s = rb.ChannelVolume(rb.SamplesBuffer(2, 44100, [1.0, 3.0, 2.0, 4.0]), [0.5, 2.0])
    assert oracle.chain(to_oracle(s))[0].tolist() == [1.0, 4.0, 1.5, 6.0]


# ---- src/source/signal_generator.rs:181-238 -------------------------------------------------
def test_signal_generators():
    assert oracle.signal(oracle.SQUARE, 2000, 500.0, 8).tolist() == [1.0, 1.0, -1.0, -1.0, 1.0, 1.0, -1.0, -1.0]
    assert oracle.signal(oracle.TRIANGLE, 8000, 1000.0, 16).tolist() == [
        -1.0, -0.5, 0.0, 0.5, 1.0, 0.5, 0.0, -0.5, -1.0, -0.5, 0.0, 0.5, 1.0, 0.5, 0.0, -0.5]
    assert oracle.signal(oracle.SAWTOOTH, 200, 50.0, 7).tolist() == [0.0, 0.5, -1.0, -0.5, 0.0, 0.5, -1.0]
    want = [0.0, 0.58778525, 0.95105652, 0.95105652, 0.58778525, 0.0, -0.58778554]
    got = oracle.signal(oracle.SINE, 1000, 100.0, 7)
    assert np.max(np.abs(got - np.array(want, np.float32))) < 1e-4


# ---- tests/limit.rs:6-155 (behavioural bands) -----------------------------------------------
def _limited(freq, amp, n, settings):
    src = rb.TestSource(oracle.sine_wave(freq, n), 1, 48000).amplify(amp).limit(settings)
    return oracle.chain(to_oracle(src))[0]


def test_limiting_works():
    st = rb.LimitSettings.default().with_threshold(-6.0).with_knee_width(0.5) \
        .with_attack(rb.Duration.from_millis(3)).with_release(rb.Duration.from_millis(12))
    out = _limited(440.0, 3.0, 2600, st)
    peak = np.max(np.abs(out[1500:]))
    assert 0.4 <= peak <= 0.6


def test_passthrough_below_threshold():
    x = oracle.sine_wave(1000.0, 880)
    out = _limited(1000.0, 0.2, 880, rb.LimitSettings.default().with_threshold(-6.0))
    assert np.max(np.abs(out - x * np.float32(0.2))) < 0.01


@pytest.mark.parametrize("thr,expect", [(-1.0, 0.89), (-3.0, 0.71), (-6.0, 0.50)])
def test_limiter_with_different_settings(thr, expect):
    st = rb.LimitSettings.default().with_threshold(thr).with_knee_width(1.0) \
        .with_attack(rb.Duration.from_millis(2)).with_release(rb.Duration.from_millis(10))
    out = _limited(440.0, 2.0, 2000, st)
    peak = np.max(np.abs(out[1000:]))
    assert expect - 0.1 <= peak <= expect + 0.1


def test_limiter_stereo_processing():
    i = np.arange(1000, dtype=np.float32)
    left = np.sin(i * np.float32(0.01)).astype(np.float32) * np.float32(1.5)
    right = np.sin(i * np.float32(0.01)).astype(np.float32) * np.float32(0.8)
    st = np.stack([left, right], 1).reshape(-1)
    out = oracle.chain(to_oracle(rb.SamplesBuffer(2, 44100, st).limit(rb.LimitSettings.default().with_threshold(-3.0))))[0]
    assert np.max(np.abs(out[0::2])) <= 1.5 and np.max(np.abs(out[1::2])) <= 1.5


# ---- src/player.rs:454-470: set_volume(0.5) == amplify(0.5) ---------------------------------
def test_volume_is_amplify():
    x = np.array([0.1, -0.4, 0.7, 1.0], np.float32)
    out = oracle.chain(to_oracle(rb.SamplesBuffer(1, 44100, x).amplify(0.5)))[0]
    assert np.array_equal(out, x * np.float32(0.5))


# ---- coefficient vector quoted in SURVEY.md Appendix A.4 (48 kHz, 200 Hz, q = 0.5) ----------
def test_blt_coefficients_known_values():
    k = oracle.blt_coeffs(False, 200, 0.5, 48000)
    assert np.allclose(k, [1.6696297e-4, 3.3392594e-4, 1.6696297e-4, -1.9483138, 0.94898164], rtol=2e-6)


# ---- the statically dispatched baseline chain is the same arithmetic as the virtual classes ----
def test_static_dispatch_baseline_is_bit_identical():
    rng = np.random.default_rng(77)
    srcs = [rb.UniformSourceIterator(rb.TestSource(rng.uniform(-1, 1, 3000).astype(np.float32), 1, 44100), 1, 48000)
            .low_pass(200).amplify(1.2) for _ in range(9)]
    streams = [to_oracle(s) for s in srcs]
    want = oracle.mixer(streams, 1, 48000)
    got_dyn, _ = oracle.mixer_mt(streams, 1, 48000, 1, want.size + 8)
    got_static, _ = oracle.mixer_mt(streams, 1, 48000, 1, want.size + 8, static_dispatch=True)
    assert np.array_equal(got_dyn.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got_static.view(np.uint32), want.view(np.uint32))
    # sharded over threads: association across shards changes, values stay within the parity tolerance
    got_mt, _ = oracle.mixer_mt(streams, 1, 48000, 3, want.size + 8, static_dispatch=True)
    assert got_mt.shape == want.shape and np.max(np.abs(got_mt - want)) <= 1e-5 * np.max(np.abs(want))
