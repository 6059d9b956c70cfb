"""source::from_iter (src/source/from_iter.rs:16-127) -- ONE source whose sample rate and channel count change from one buffer to
the next -- and what rodio's adapters make of such a source: SpanTracker (src/source/span.rs:66-101) inside the filters
(blt.rs:122-137: state kept, coefficients recomputed behind the first sample of the new span), the mixer's UniformSourceIterator
re-bootstrapping with whatever FromIter reports at that moment (uniform.rs:50-68,:83-96).  Rows a14 and (f-4) of SURVEY.md section 8.
CPU: the oracle on the reference's own test, the host planner against the oracle's literal iterators.  GPU: bit for bit."""
import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, noise, to_oracle
from rodio_b200 import capi


def _buf(n, ch, rate, seed, spans=True):
    x = noise(n, seed)
    return rb.SamplesBuffer(ch, rate, x) if spans else rb.TestSource(x, ch, rate)


def _random_sequence(rng, seed):
    k = int(rng.integers(2, 5))
    parts = []
    for i in range(k):
        ch = int(rng.integers(1, 4))
        rate = int(rng.choice([8000, 22050, 32000, 44100, 48000, 96000]))
        n = int(rng.choice([0, 1, 2, 3, ch, 2 * ch, 7 * ch + 1, 500 * ch, 3000 * ch, 40000, 70001]))
        parts.append(_buf(n, ch, rate, seed * 10 + i, spans=bool(rng.integers(0, 4))))
    src = rb.from_iter(parts)
    kind = int(rng.integers(0, 7))
    if kind == 1:
        src = src.amplify(0.7)
    elif kind == 2:
        src = src.low_pass(int(rng.choice([200, 1000, 5000])))
    elif kind == 3:
        src = src.amplify(1.3).high_pass(300).amplify(0.5)
    elif kind == 4:
        src = src.speed(float(rng.choice([0.5, 0.9, 1.25]))).low_pass(800)
    elif kind == 5:
        src = src.automatic_gain_control()
    elif kind == 6:
        src = src.amplify(2.0).limit()
    return src


# ------------------------------------------------------------------ CPU
def test_oracle_from_iter_reference_vector():
    """from_iter.rs:129-157 `basic` (the same vector as queue.rs:280-303): four mono 48 kHz samples, then four stereo 96 kHz ones."""
    src = rb.from_iter([rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]), rb.SamplesBuffer(2, 96000, [5.0, 5.0, 5.0, 5.0])])
    out, ch, rate = oracle.chain(to_oracle(src))
    assert out.tolist() == [10.0, -10.0, 10.0, -10.0, 5.0, 5.0, 5.0, 5.0]


def test_oracle_filter_follows_the_rate_change():
    """blt.rs:122-137: the first sample of the new span still runs through the old coefficients, everything behind it through the
    new ones, the state carries over -- against a plain numpy restatement."""
    a, b = noise(300, 1), noise(400, 2)
    src = rb.from_iter([rb.SamplesBuffer(1, 48000, a), rb.SamplesBuffer(1, 8000, b)]).low_pass(1000)
    got, _, _ = oracle.chain(to_oracle(src))
    k48 = oracle.blt_coeffs(False, 1000, 0.5, 48000)
    k8 = oracle.blt_coeffs(False, 1000, 0.5, 8000)
    x = np.concatenate([a, b])
    x1 = x2 = y1 = y2 = np.float32(0)
    want = np.empty_like(x)
    for i, xv in enumerate(x):
        k = k48 if i <= 300 else k8
        r = np.float32(k[0] * xv)
        r = np.float32(r + np.float32(k[1] * x1))
        r = np.float32(r + np.float32(k[2] * x2))
        r = np.float32(r - np.float32(k[3] * y1))
        r = np.float32(r - np.float32(k[4] * y2))
        y2, x2, y1, x1 = y1, x1, r, xv
        want[i] = r
    assert_bit_exact(got, want, "filter over a rate change")


def test_oracle_agc_and_limiter_start_over_at_a_format_change():
    """agc.rs:524-548 (coefficients for the new rate, RMS window / peak / gain from scratch) and limit.rs:651-697 (per-channel state
    rebuilt when the channel count changes, coefficients kept): from the SECOND sample of the new span on -- so the literal iterators
    equal independent adapters over the stretches between those points."""
    a, b, c = noise(3000, 1), noise(2 * 2500, 2), noise(4000, 3)
    seq = rb.from_iter([rb.SamplesBuffer(1, 48000, a), rb.SamplesBuffer(2, 44100, b), rb.SamplesBuffer(1, 22050, c)])
    x = np.concatenate([a, b, c])
    cuts = [0, 3000 + 1, 3000 + 5000 + 1, x.size]
    got = oracle.chain(to_oracle(seq.automatic_gain_control()))[0]
    want = np.concatenate([oracle.chain(to_oracle(rb.TestSource(x[cuts[i]:cuts[i + 1]], 1, r).automatic_gain_control()))[0]
                           for i, r in enumerate([48000, 44100, 22050])])
    assert_bit_exact(got, want, "AGC over a format change")
    got = oracle.chain(to_oracle(seq.limit()))[0]
    want = np.concatenate([oracle.chain(to_oracle(rb.TestSource(x[cuts[i]:cuts[i + 1]], ch, 48000).limit()))[0]
                           for i, ch in enumerate([1, 2, 1])])
    assert_bit_exact(got, want, "limiter over a channel-count change")


def test_planner_against_the_literal_iterators():
    """Closed forms of the host planner (converter runs per bootstrap, their lengths) against the oracle's pull iterators."""
    rng = np.random.default_rng(2024)
    for t in range(250):
        src = _random_sequence(rng, 100 + t)
        for mixer in ((1, 48000), (2, 44100)):
            want = oracle.chain_uniform(to_oracle(src), *mixer)
            out_len, ch, rate, chain_len = rb.plan(src, *mixer)
            assert out_len == want.size, (t, mixer, out_len, want.size)


def test_from_iter_argument_errors():
    a, b = _buf(10, 1, 48000, 1), _buf(10, 2, 44100, 2)
    with pytest.raises(rb.RodioB200Error):      # an adapter the block path does not follow across a format change
        rb.plan(rb.from_iter([a, b]).reverb(rb.Duration.from_millis(10), 0.5), 1, 48000)
    with pytest.raises(ValueError):
        rb.from_iter([a, b.amplify(2.0)])


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_from_iter_bit_exact(ctx):
    rng = np.random.default_rng(7)
    for t in range(90):
        src = _random_sequence(rng, 500 + t)
        mixer = (int(rng.integers(1, 3)), int(rng.choice([44100, 48000])))
        want = oracle.mixer([to_oracle(src)], *mixer)
        with rb.Batch([src], *mixer, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
            b.upload_all()
            got = b.render_mix()
        # a filter above the Nyquist rate of an 8 kHz buffer diverges to inf - inf = NaN on both sides: x86 hands on the payload of
        # an operand, the device the canonical NaN -- not a parity question, so NaNs compare equal here
        assert got.shape == want.shape
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), f"from_iter case {t}: NaNs in other places"
        if any(e.kind == capi.RB_FX_LIMIT for e in src.effects):      # device log2 / exp2: the north-star tolerance
            g, w = np.where(nan, np.float32(0), got), np.where(nan, np.float32(0), want)
            assert np.max(np.abs(g - w), initial=0.0) <= 1e-5 * max(float(np.max(np.abs(w), initial=0.0)), 1e-30), f"from_iter case {t}"
        else:
            assert_bit_exact(np.where(nan, np.float32(0), got), np.where(nan, np.float32(0), want), f"from_iter case {t}")


@pytest.mark.gpu
def test_from_iter_reference_vector_and_mixer(ctx):
    seq = rb.from_iter([rb.SamplesBuffer(1, 48000, [10.0, -10.0, 10.0, -10.0]), rb.SamplesBuffer(2, 96000, [5.0, 5.0, 5.0, 5.0])])
    srcs = [seq, rb.TestSource(noise(5000, 3), 1, 44100).low_pass(500),
            rb.from_iter([_buf(2 * 9000, 2, 44100, 4), _buf(12000, 1, 22050, 5), _buf(2 * 4000, 2, 48000, 6)]).low_pass(2000).amplify(0.6)]
    starts = [0, 0, 960]
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=capi.RB_MIX_EXACT_ORDER, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "from_iter sources in a mixer")
