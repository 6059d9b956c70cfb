"""Source::mix / take_crossfade_with (src/source/mod.rs:253-261,:444-454, mix.rs:10-53, crossfade.rs:10-23): the two-input adapter
(RB_FX_MIX: the second input is a descriptor of its own, consumed by the MIX).  CPU: the oracle against the two tests the
reference holds (crossfade.rs:45-80) and the planner's argument checks.  GPU: the general path bit for bit against the oracle."""
import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, noise, to_oracle
from rodio_b200 import capi

S5 = rb.Duration.from_secs(5) + 1          # Duration::from_secs(5) + Duration::from_nanos(1), crossfade.rs:52,:72


def _dummy(length):                        # crossfade.rs:40-43: SamplesBuffer::new(1 ch, 1 Hz, 1..=length)
    return rb.SamplesBuffer(1, 1, np.arange(1, length + 1, dtype=np.float32))


def _chain(src):
    out, ch, rate = oracle.chain(to_oracle(src))
    return out


# ------------------------------------------------------------------ CPU: the oracle on the reference's own vectors
def test_oracle_crossfade_with_self():
    """crossfade.rs:45-63: 1 2 3 4 5 within 1e-6, then None."""
    got = _chain(_dummy(10).take_crossfade_with(_dummy(10), S5))
    assert got.size == 5 and np.all(np.abs(got - np.array([1, 2, 3, 4, 5], np.float32)) < 1e-6), got


def test_oracle_crossfade_against_silence():
    """crossfade.rs:65-80: source2 = Zero (endless silence): [1.0, 2*0.8, 3*0.6, 4*0.4, 5*0.2] within 1e-6."""
    zero = rb.TestSource(np.zeros(64, np.float32), 1, 1)
    got = _chain(_dummy(10).take_crossfade_with(zero, S5))
    assert got.size == 5 and np.all(np.abs(got - np.array([1.0, 1.6, 1.8, 1.6, 1.0], np.float32)) < 1e-6), got


def test_mix_planner_and_argument_errors():
    import ctypes as C
    from rodio_b200.source import pack_descs
    a = rb.TestSource(noise(100, 1), 1, 48000)
    b = rb.TestSource(noise(2 * 300, 2), 2, 24000)
    out_len, ch, rate, chain_len = rb.plan(a.mix(b), 1, 48000)          # rb_streams_plan: the descriptor array
    assert (ch, rate) == (1, 48000) and chain_len == _chain(a.mix(b)).size == out_len
    descs, keep = pack_descs([a.mix(b)])
    n = C.c_uint64()
    L = rb.lib()
    # one descriptor alone cannot name its second input
    assert L.rb_stream_plan(C.byref(descs[0]), 1, 48000, C.byref(n), None, None, None) == capi.RB_ERR_UNSUPPORTED
    # the second input must be marked consumed, and only one MIX may take it
    descs[1].mix_start = 0
    assert L.rb_streams_plan(descs, 2, 0, 1, 48000, C.byref(n), None, None, None) == capi.RB_ERR_INVALID_ARGUMENT
    descs[1].mix_start = capi.RB_MIX_START_CONSUMED
    assert L.rb_streams_plan(descs, 2, 0, 1, 48000, C.byref(n), None, None, None) == capi.RB_OK
    keep[1][0].u32[0] = 0                                                # a source mixed with itself
    assert L.rb_streams_plan(descs, 2, 0, 1, 48000, C.byref(n), None, None, None) == capi.RB_ERR_INVALID_ARGUMENT
    keep[1][0].u32[0] = 5
    assert L.rb_streams_plan(descs, 2, 0, 1, 48000, C.byref(n), None, None, None) == capi.RB_ERR_INVALID_ARGUMENT
    # a descriptor marked consumed that nobody consumes
    lone, _ = pack_descs([a])
    lone[0].mix_start = capi.RB_MIX_START_CONSUMED
    assert L.rb_streams_plan(lone, 1, 0, 1, 48000, C.byref(n), None, None, None) == capi.RB_OK     # planned, but a batch refuses it


# ------------------------------------------------------------------ GPU
GENERAL = capi.RB_MIX_EXACT_ORDER


@pytest.mark.gpu
def test_crossfade_reference_vectors_on_the_device(ctx):
    zero = rb.TestSource(np.zeros(64, np.float32), 1, 1)
    for src, want in ((_dummy(10).take_crossfade_with(_dummy(10), S5), [1, 2, 3, 4, 5]),
                      (_dummy(10).take_crossfade_with(zero, S5), [1.0, 1.6, 1.8, 1.6, 1.0])):
        got = src.collect(ctx)
        assert_bit_exact(got, _chain(src), "crossfade vs oracle")
        assert got.size == 5 and np.all(np.abs(got - np.array(want, np.float32)) < 1e-6)


@pytest.mark.gpu
def test_mix_of_two_sources_bit_exact(ctx):
    """Second input at another rate and channel count, shorter and longer than the first, a generator, adapters on both sides
    and behind the mix, a mix inside a mix."""
    a = rb.TestSource(noise(2 * 3000, 11), 2, 44100)
    cases = {
        "other rate, mono into stereo, shorter": a.mix(rb.TestSource(noise(1000, 12), 1, 32000)),
        "longer second input": rb.TestSource(noise(500, 13), 1, 48000).mix(rb.TestSource(noise(2 * 4000, 14), 2, 44100)),
        "generator as second input": a.amplify(0.5).mix(rb.SineWave(440.0).take(9000).amplify(0.25)).low_pass(2000),
        "spans on both sides": rb.SamplesBuffer(2, 22050, noise(2 * 40000, 15)).mix(rb.SamplesBuffer(1, 48000, noise(70000, 16))),
        "mix inside a mix": a.mix(rb.TestSource(noise(900, 17), 1, 8000).mix(rb.TestSource(noise(2 * 700, 18), 2, 96000)).amplify(0.3)),
        "empty second input": a.mix(rb.TestSource(np.zeros(0, np.float32), 1, 48000)),
        "empty first input": rb.TestSource(np.zeros(0, np.float32), 2, 48000).mix(rb.TestSource(noise(300, 19), 1, 48000)),
    }
    for what, src in cases.items():
        assert_bit_exact(src.collect(ctx), _chain(src), what)


@pytest.mark.gpu
def test_crossfade_in_a_mixer_with_other_sources(ctx):
    """benches/pipeline.rs-sized sources: a crossfade of two stereo 44.1 kHz sounds (0.25 s), a plain source and a generator in
    one 48 kHz stereo mixer; exact order, bit for bit."""
    d = rb.Duration.from_millis(250)
    x1, x2 = noise(2 * 30000, 21), noise(2 * 20000, 22)
    srcs = [rb.TestSource(x1, 2, 44100).take_crossfade_with(rb.TestSource(x2, 2, 44100), d),
            rb.TestSource(noise(15000, 23), 1, 48000).amplify(0.5),
            rb.SamplesBuffer(2, 44100, x1).take_crossfade_with(rb.SamplesBuffer(1, 22050, noise(9000, 24)), d).amplify(0.7),
            rb.TriangleWave(220.0).take(10000)]
    starts = [0, 480, 960, 2000]
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=GENERAL, mix_starts=starts, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "crossfades in a mixer")
    with rb.Batch(srcs, 2, 48000, mix_starts=starts, ctx=ctx) as b:      # default flags: the batch still takes the general path
        b.upload_all()
        assert b.kernel_family == -1
        got = b.render_mix()
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))


def test_planner_of_random_mixes_against_the_literal_iterators():
    """Closed forms for Source::mix / take_crossfade_with (the second input converted to the first one's format, the longer one
    wins; take_duration's frame padding is never pulled by Mix's UniformSourceIterators) against the oracle's pull iterators."""
    rng = np.random.default_rng(4242)

    def rand_src(seed):
        ch = int(rng.integers(1, 4))
        rate = int(rng.choice([8000, 22050, 32000, 44100, 48000]))
        n = int(rng.choice([0, ch, 3 * ch, 200 * ch, 1500 * ch]))      # whole frames (the C ABI refuses anything else)
        x = noise(n, seed)
        s = rb.SamplesBuffer(ch, rate, x) if rng.integers(0, 2) else rb.TestSource(x, ch, rate)
        k = int(rng.integers(0, 4))
        if k == 1:
            s = s.amplify(0.5)
        elif k == 2:
            s = s.low_pass(500)
        elif k == 3:
            s = s.take_duration(rb.Duration.from_millis(int(rng.integers(1, 40))))
        return s

    for t in range(200):
        a, b = rand_src(9000 + 2 * t), rand_src(9001 + 2 * t)
        src = a.mix(b) if t % 3 else a.take_crossfade_with(b, rb.Duration.from_millis(int(rng.integers(1, 30))))
        if t % 5 == 0:
            src = src.amplify(0.9)
        want = _chain(src)
        for mixer in ((1, 48000), (2, 44100)):
            out_len, ch, rate, chain_len = rb.plan(src, *mixer)
            assert chain_len == want.size, (t, chain_len, want.size)
            assert out_len == oracle.chain_uniform(to_oracle(src), *mixer).size, (t, mixer)
