"""Row a13: SignalGenerator / SineWave (src/source/signal_generator.rs:40-135, src/source/sine.rs:23-27) generated on the device
(RB_FX_SIGNAL, k_siggen) -- bit for bit against the oracle's literal generator, whose sinf is this machine's glibc (what rustc's
`f32::sin` calls on linux-gnu).  The CPU part holds the sinf restatement of the kernel against libm and checks the planner."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, assert_close_peak, to_oracle
from rodio_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNCS = [rb.Function.Sine, rb.Function.Triangle, rb.Function.Square, rb.Function.Sawtooth]


# ------------------------------------------------------------------ CPU: planner and the sinf restatement
def test_planner_of_a_generator_stream():
    src = rb.SignalGenerator(44100, 440.0, rb.Function.Triangle).take(1000).amplify(0.5)
    out_len, ch, rate, chain_len = rb.plan(src, 2, 48000)
    assert (ch, rate, chain_len) == (1, 44100, 1000)
    ref = oracle.sample_rate_converter(np.zeros(1000, np.float32), 44100, 48000, 1)
    assert out_len == 2 * ref.size


@pytest.mark.parametrize("bad", ["zero_freq", "negative", "nan", "inf", "not_first", "with_pcm", "stereo", "function"])
def test_generator_argument_errors(bad):
    e = rb.Effect.make(capi.RB_FX_SIGNAL, u32=[0], f32=[440.0], ns=[100])
    src = rb.Source(np.zeros(0, np.float32), 1, 48000, 0, [e])
    if bad in ("zero_freq", "negative", "nan", "inf"):
        f = {"zero_freq": 0.0, "negative": -3.0, "nan": float("nan"), "inf": float("inf")}[bad]
        src = rb.Source(np.zeros(0, np.float32), 1, 48000, 0, [rb.Effect.make(capi.RB_FX_SIGNAL, u32=[0], f32=[f], ns=[100])])
    elif bad == "not_first":
        src = rb.Source(np.zeros(0, np.float32), 1, 48000, 0, [rb.Effect.make(capi.RB_FX_AMPLIFY, f32=[1.0]), e])
    elif bad == "with_pcm":
        src = rb.Source(np.zeros(8, np.float32), 1, 48000, 0, [e])
    elif bad == "stereo":
        src = rb.Source(np.zeros(0, np.float32), 2, 48000, 0, [e])
    elif bad == "function":
        src = rb.Source(np.zeros(0, np.float32), 1, 48000, 0, [rb.Effect.make(capi.RB_FX_SIGNAL, u32=[7], f32=[440.0], ns=[100])])
    with pytest.raises(rb.RodioB200Error) as ei:
        rb.plan(src, 1, 48000)
    assert ei.value.status in (capi.RB_ERR_INVALID_ARGUMENT, capi.RB_ERR_UNSUPPORTED)


def test_sinf_restatement_against_libm(tmp_path):
    """The double-precision polynomial k_siggen evaluates, as plain C++, against this machine's sinf on every 3rd float of
    [0, 2*pi] (tools/microbench/sinf_exhaustive.cpp without an argument checks all 1 086 918 620 of them)."""
    exe = tmp_path / "sinf_x"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-pthread", os.path.join(ROOT, "tools/microbench/sinf_exhaustive.cpp"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "3"], check=True, capture_output=True, text=True).stdout
    assert "mismatches no-fma 0 " in out and ", fma 0 " in out, out


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("fn", FUNCS)
def test_generators_bit_exact(ctx, fn):
    cases = [(48000, 440.0, 5000), (44100, 1000.0, 4411), (8000, 3999.5, 777), (48000, 0.25, 3000), (200, 50.0, 7),
             (1000, 100.0, 0), (48000, 19999.0, 1), (48000, 110.0, 127), (48000, 110.0, 128), (48000, 110.0, 129),
             (1000, 2500.0, 300), (48000, 48000.0, 100), (22050, 1e-3, 500), (96000, 12345.678, 2048)]
    for r, f, n in cases:
        src = rb.SignalGenerator(r, f, fn).take(n)
        want, ch, rate = oracle.chain(to_oracle(src))
        assert (want.size, ch, rate) == (n, 1, r)
        assert_bit_exact(src.collect(ctx), want, f"fn {fn}: {f} Hz at {r} Hz x {n}")      # nothing is uploaded


@pytest.mark.gpu
def test_sine_phase_sweep_bit_exact(ctx):
    """Long sines over many frequencies: the f32 phase drift and every quadrant / polynomial branch of sinf."""
    rng = np.random.default_rng(5)
    freqs = [float(np.float32(v)) for v in np.exp(rng.uniform(np.log(0.01), np.log(23000.0), 40))]
    for f in freqs[:6]:
        got = rb.SineWave(f).take(60000).collect(ctx)
        assert_bit_exact(got, oracle.signal(oracle.SINE, 48000, f, 60000), f"SineWave({f})")
    srcs = [rb.SineWave(f).take(9000) for f in freqs]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert_bit_exact(b.render_mix(), want, "mixer of 40 SineWaves, exact order")


@pytest.mark.gpu
def test_cfg2_literally_1024_sinewaves(ctx):
    """BASELINE cfg2 as the reference states it: 1024 SineWave sources handed to mixer(1, 48000) -- generated AND summed on the
    device, nothing uploaded; exact order bit for bit, the default grouping within 1e-5 * peak."""
    srcs = [rb.SineWave(min(110.0 * 2.0 ** (s / 128.0), 19999.0)).take(4800) for s in range(1024)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert_bit_exact(b.render_mix(), want, "cfg2 from generators, exact order")
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        assert_close_peak(b.render_mix(), want, 1e-5, "cfg2 from generators, default")


@pytest.mark.gpu
def test_generator_through_a_chain_and_the_mixer(ctx):
    """A generator is a Source like any other: effects behind it, another rate and channel count at the mixer, a late start."""
    srcs = [rb.SignalGenerator(44100, 523.25, rb.Function.Sawtooth).take(6000).low_pass(1000).amplify(0.4),
            rb.SineWave(330.0).take(5000).reverb(rb.Duration.from_millis(20), 0.5),
            rb.SquareWave(80.0).take(4000).fade_in(rb.Duration.from_millis(30)),
            rb.TestSource(np.linspace(-1, 1, 3000, dtype=np.float32), 1, 32000)]
    starts = [0, 0, 960, 2000]
    want = oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=capi.RB_MIX_EXACT_ORDER, mix_starts=starts, ctx=ctx) as b:
        b.upload(3)
        assert_bit_exact(b.render_mix(), want, "generators with effects into a stereo 48 kHz mixer")
