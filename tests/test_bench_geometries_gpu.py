"""Parity at the launch geometries the bench lines are quoted on ("parity gaps" 1, 2, 4):
cfg3 with 4096 mono streams (k_fused_hot, 28 rows per CTA incl. the second-row warps), its stereo and filter-free
twins, the automatically selected large-batch kernel (>= 128 streams per SM), cfg2 at 1024 sources, cfg4 at 512 streams.
Streams are short (the oracle finishes in seconds); the launch configuration is the one of the full-length run
because it depends on the number of streams only.

Two bars per fused case: <= 1e-5 * peak against the reference's sequential mixer (src/mixer.rs:185-198) -- the
north-star tolerance -- and BIT-EXACT against the oracle's per-stream outputs added in the kernel's documented
order (rb_batch_mix_group)."""
import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, assert_close_peak, fused_expected_mix, noise, to_oracle
from rodio_b200 import capi

pytestmark = pytest.mark.gpu

GENERAL = capi.RB_KEEP_STREAM_OUTPUTS | capi.RB_NO_FUSION | capi.RB_MIX_EXACT_ORDER


def grouped_expected_mix(per_stream, starts, mix_len, group):
    """Per-CTA sequential sums from +0.0 over `group` consecutive streams (only where a stream is active), the partial
    rows added in CTA order starting from the first (k_fused_hot stage C + k_sum_partials)."""
    total = None
    for g in range(0, len(per_stream), group):
        acc = np.zeros(mix_len, dtype=np.float32)
        for y, s in zip(per_stream[g:g + group], starts[g:g + group]):
            acc[s:s + y.size] = acc[s:s + y.size] + y
        total = acc if total is None else total + acc
    return total


def _cfg3(n, frames, ch=1, lp=200, gain=1.2, seed=31000):
    srcs = []
    for s in range(n):
        src = rb.UniformSourceIterator(rb.TestSource(noise(frames * ch, seed + s), ch, 44100), ch, 48000)
        if lp:
            src = src.low_pass(lp)
        srcs.append(src.amplify(gain))
    return srcs


def _check_fused(ctx, srcs, ch, family, group=None, flags=0):
    with rb.Batch(srcs, ch, 48000, flags=flags, ctx=ctx) as b:
        if isinstance(family, tuple):
            assert b.kernel_family in family, f"kernel family {b.kernel_family}, expected one of {family}"
            family = b.kernel_family
        assert b.kernel_family == family, f"kernel family {b.kernel_family}, expected {family}"
        if group is not None:
            assert b.mix_group == group, f"rows per partial sum {b.mix_group}, expected {group}"
        group = b.mix_group
        b.upload_all()
        got = b.render_mix()
        again = b.render_mix()
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32)), "render is not idempotent"
    streams = [to_oracle(s) for s in srcs]
    ref = oracle.mixer(streams, ch, 48000)
    assert got.shape == ref.shape
    assert_close_peak(got, ref, 1e-5, "fused kernel vs the reference's sequential mixer")
    per_stream = [oracle.chain_uniform(s, ch, 48000) for s in streams]
    starts = [0] * len(srcs)
    want = grouped_expected_mix(per_stream, starts, ref.size, group or len(srcs)) if family in (0, 1, 5, 6) else fused_expected_mix(3 if family == 4 else family, per_stream, starts, ref.size)
    assert_bit_exact(got, want, "fused kernel vs oracle streams added in the kernel's documented order")


def test_cfg3_bench_geometry_mono_4096(ctx):
    """The headline launch: 4096 mono streams, default flags -> k_fused_hot<1, true>, 147 CTAs x 28 rows."""
    _check_fused(ctx, _cfg3(4096, 4410), 1, family=1, group=28)


def test_cfg3_bench_geometry_mono_4096_low_pass_1000(ctx):
    _check_fused(ctx, _cfg3(4096, 2000, lp=1000, seed=32000), 1, family=1, group=28)


def test_cfg3_bench_geometry_stereo_2048(ctx):
    """profiles' cfg3_stereo line: 2048 stereo streams -> k_fused_hot<2, true>."""
    _check_fused(ctx, _cfg3(2048, 2205, ch=2, seed=33000), 2, family=1)


def test_nofilter_bench_geometry_mono_4096(ctx):
    """resample -> amplify -> mix at 4096 streams: a chain without a filter carries nothing from sample to sample, so the planner
    cuts the timeline into segments for the lane-pair kernel (family 4) -- still every stream's serial samples bit for bit; with
    RB_FUSED_LANES | ... pinned off (RB_SEGMENTS_FROM) it is k_fused_hot<1, false> with full CTAs."""
    _check_fused(ctx, _cfg3(4096, 4410, lp=None, seed=34000), 1, family=(1, 4, 6))


def test_cfg5_auto_selected_large_batch(ctx):
    """No flag, 41 000 streams (>= 128 per SM): the planner hands the batch to the large-batch kernels on its own (mono
    sources that start together: the lane-pair kernel)."""
    _check_fused(ctx, _cfg3(41000, 600, seed=35000), 1, family=3)


def test_cfg5_large_batch_16384(ctx):
    """16 384 streams, whatever kernel the planner picks at that size."""
    srcs = _cfg3(16384, 900, seed=36000)
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        family = b.kernel_family
    assert family in (1, 2, 3)
    _check_fused(ctx, srcs, 1, family=family)


def test_cfg2_1024_sources(ctx):
    """BASELINE cfg2 stream count: mixer(1, 48000) of 1024 SineWave sources, exact order and default."""
    n, frames = 1024, 4800
    srcs = [rb.TestSource(oracle.sine_wave(min(55.0 + 19.3 * s, 19900.0), frames), 1, 48000) for s in range(n)]
    want = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "cfg2 1024 sources, exact order")
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:     # the bench's launch (default flags)
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, want, 1e-5, "cfg2 1024 sources, default")


def cfg4_sources(n, frames, seed=700):
    srcs = []
    t = np.arange(frames, dtype=np.float32)
    for s in range(n):
        l = np.sin(t * np.float32(0.02 + 0.00001 * s)).astype(np.float32) * np.float32(0.4)
        r = np.sin(t * np.float32(0.021 + 0.00001 * s)).astype(np.float32) * np.float32(0.4)
        x = np.stack([l, r], 1).reshape(-1) + noise(2 * frames, seed + s, 0.1)
        srcs.append(rb.Spatial(rb.TestSource(x, 2, 48000), [float(s % 7 - 3), 1.0, 0.0], [-1, 0, 0], [1, 0, 0])
                    .reverb(rb.Duration.from_millis(50), 0.3).automatic_gain_control())
    return srcs


def test_cfg4_512_streams(ctx):
    """BASELINE cfg4 stream count: 512 stereo sources, spatial -> reverb(50 ms, 0.3) -> AGC -> mix(2 ch): the default
    launch within the tolerance, the exact-order launch and its per-stream outputs bit for bit."""
    n, frames = 512, 2400 + 2400     # the echo starts after 2400 frames
    srcs = cfg4_sources(n, frames)
    streams = [to_oracle(s) for s in srcs]
    want = oracle.mixer(streams, 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=GENERAL, ctx=ctx) as b:
        b.upload_all()
        got = b.render_mix()
        for i in (0, 255, n - 1):
            assert_bit_exact(b.read_stream(i), oracle.chain_uniform(streams[i], 2, 48000), f"cfg4 stream {i}")
    assert_bit_exact(got, want, "cfg4 512 streams, exact order")
    with rb.Batch(srcs, 2, 48000, ctx=ctx) as b:     # the default launch: k_fused_fx, 4 streams per CTA
        assert b.kernel_family == 5 and b.mix_group == 4 and b.launches_per_render == 2
        b.upload_all()
        got = b.render_mix()
        again = b.render_mix()
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))
    assert_close_peak(got, want, 1e-5, "cfg4 512 streams, default launch")
    per = [oracle.chain_uniform(s, 2, 48000) for s in streams]
    assert_bit_exact(got, grouped_expected_mix(per, [0] * n, want.size, 4), "k_fused_fx vs oracle streams in its documented order")


@pytest.mark.parametrize("shape", ["mono_agc_only", "stereo_reverb_agc", "stereo_spatial_agc", "ragged"])
def test_fx_kernel_shapes(ctx, shape):
    """The other shapes k_fused_fx serves: without the channel volumes, without the echo, mono, streams of different lengths,
    odd echo delays, custom AGC settings, a number of streams that is not a multiple of four."""
    rng = np.random.default_rng(17)
    ch = 1 if shape == "mono_agc_only" else 2
    n = 37
    srcs = []
    for s in range(n):
        frames = 9000 + (int(rng.integers(0, 4000)) if shape == "ragged" else 0)
        src = rb.TestSource(noise(frames * ch, 51000 + s, 0.3), ch, 48000)
        if shape in ("stereo_spatial_agc", "ragged"):
            src = rb.Spatial(src, [float(s % 5 - 2), 1.0, 0.5], [-1, 0, 0], [1, 0, 0])
        if shape in ("stereo_reverb_agc", "ragged"):
            src = src.reverb(rb.Duration.from_micros(20000 + 137 * s), 0.4)
        settings = rb.AutomaticGainControlSettings(target_level=0.8, attack_time=rb.Duration.from_millis(200 + s),
                                                   release_time=rb.Duration.from_millis(50), absolute_max_gain=5.0) if s % 2 else rb.AutomaticGainControlSettings()
        srcs.append(src.automatic_gain_control(settings))
    streams = [to_oracle(s) for s in srcs]
    want = oracle.mixer(streams, ch, 48000)
    with rb.Batch(srcs, ch, 48000, ctx=ctx) as b:
        assert b.kernel_family == 5
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, want, 1e-5, f"k_fused_fx {shape}")
    per = [oracle.chain_uniform(s, ch, 48000) for s in streams]
    assert_bit_exact(got, grouped_expected_mix(per, [0] * n, want.size, 4), f"k_fused_fx {shape}: documented order")


# ------------------------------------------------------------------ the lane-pair kernel and the time-parallel plan
def test_duo_kernel_by_flag(ctx):
    """RB_FUSED_DUO on a batch far below the automatic threshold: k_fused_duo, several CTAs, ragged lengths, late joiners in
    phase (multiples of 4 * to = 640 frames)."""
    rng = np.random.default_rng(5)
    n = 1500
    lens = [int(v) for v in rng.integers(1, 3000, n)]
    lens[:6] = [0, 1, 2, 9, 640, 641]
    starts = sorted(640 * int(v) for v in rng.integers(0, 3, n))
    srcs = [rb.UniformSourceIterator(rb.TestSource(noise(L, 41000 + i), 1, 44100), 1, 48000).low_pass(200 + (i % 50) * 40).amplify(1.2)
            for i, L in enumerate(lens)]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_DUO, mix_starts=starts, ctx=ctx) as b:
        assert b.kernel_family == 3 and b.mix_group == 64
        b.upload_all()
        got = b.render_mix()
    streams = [to_oracle(s, st) for s, st in zip(srcs, starts)]
    ref = oracle.mixer(streams, 1, 48000)
    assert_close_peak(got, ref, 1e-5, "k_fused_duo vs the sequential mixer")
    per = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    assert_bit_exact(got, fused_expected_mix(3, per, starts, ref.size), "k_fused_duo vs oracle streams in its documented order")


@pytest.mark.parametrize("chain", ["no_filter", "high_pass", "other_ratio"])
def test_duo_kernel_shapes(ctx, chain):
    n = 700
    mk = {"no_filter": lambda x: rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).amplify(0.8),
          "high_pass": lambda x: rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).high_pass(300),
          "other_ratio": lambda x: rb.UniformSourceIterator(rb.TestSource(x, 1, 22050), 1, 48000).low_pass(1500).amplify(0.5)}[chain]
    srcs = [mk(noise(900 + 3 * (i % 200), 42000 + i)) for i in range(n)]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_DUO, ctx=ctx) as b:
        assert b.kernel_family == 3
        b.upload_all()
        got = b.render_mix()
    per = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    assert_bit_exact(got, fused_expected_mix(3, per, [0] * n, got.size), f"k_fused_duo, {chain}")


def test_duo_falls_back_when_neighbours_are_out_of_phase(ctx):
    rng = np.random.default_rng(6)
    n = 300
    starts = sorted(int(v) for v in rng.integers(0, 500, n))
    srcs = _cfg3(n, 1500, seed=43000)
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_FUSED_DUO, mix_starts=starts, ctx=ctx) as b:
        assert b.kernel_family == 2
        b.upload_all()
        got = b.render_mix()
    per = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    assert_bit_exact(got, fused_expected_mix(2, per, starts, got.size), "out of phase: k_fused_lanes")


def _f64_truth(srcs, lp, q, gain):
    """The same chain with the biquad's recurrence carried in f64 (coefficients and input as the f32 reference has them)."""
    total = None
    for s in srcs:
        x = oracle.chain_uniform(oracle.Stream(s.pcm, 1, 44100, [e for e in to_oracle(s).effects if e.kind == capi.RB_FX_UNIFORM], 0), 1, 48000).astype(np.float64)
        b0, b1, b2, a1, a2 = (float(v) for v in oracle.blt_coeffs(False, lp, q, 48000))
        import scipy.signal
        y = scipy.signal.lfilter([b0, b1, b2], [1.0, a1, a2], x) * float(np.float32(gain))
        total = y if total is None else total + y
    return total


@pytest.mark.parametrize("lp,q,family", [(1000, 0.5, 4), (2500, 0.707, 4), (200, 0.5, 1)])
def test_time_parallel_plan_4096_streams(ctx, lp, q, family):
    """RB_BIQUAD_TIME_PARALLEL at the bench's stream count: taken for filters inside the accuracy gate (<= 1e-5 * peak against
    the f32 reference AND no further from an f64 run of the same filter than the reference itself, times 1.5), refused for
    low_pass(200), which is then served by the exact default kernel."""
    n, frames = 4096, 11025
    srcs = [rb.UniformSourceIterator(rb.TestSource(noise(frames, 44000 + i), 1, 44100), 1, 48000).low_pass_with_q(lp, q).amplify(1.2)
            for i in range(n)]
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_BIQUAD_TIME_PARALLEL, ctx=ctx) as b:
        assert b.kernel_family == family, b.kernel_family
        b.upload_all()
        got = b.render_mix()
    ref = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    assert_close_peak(got, ref, 1e-5, f"time-parallel low_pass({lp}) vs the f32 reference")
    if family == 4:
        truth = _f64_truth(srcs[:256], lp, q, 1.2)
        with rb.Batch(srcs[:256], 1, 48000, flags=capi.RB_BIQUAD_TIME_PARALLEL, ctx=ctx) as b:
            assert b.kernel_family == 4
            b.upload_all()
            got256 = b.render_mix()
        ref256 = oracle.mixer([to_oracle(s) for s in srcs[:256]], 1, 48000)
        e_ref = float(np.max(np.abs(ref256 - truth)))
        e_tp = float(np.max(np.abs(got256 - truth)))
        assert e_tp <= 1.5 * e_ref + 1e-7 * float(np.max(np.abs(truth))), (e_tp, e_ref)


# ------------------------------------------------------------------ integer PCM in front of the fused kernels
def test_integer_pcm_keeps_a_resident_f32_copy(ctx):
    """s16 sources (what decoders yield, src/decoder/wav.rs:119-151) through the cfg3 chain: the batch converts once per upload
    (SampleTypeConverter, src/conversions/sample.rs:42-44) and the f32 kernels serve it -- same family, same bits as the oracle's
    conversion + chain; a second upload is converted again."""
    n, frames = 600, 3000
    rng = np.random.default_rng(3)

    def make(seed):
        return [(rng.integers(-30000, 30000, frames)).astype(np.int16) for _ in range(n)]
    pcm_a, pcm_b = make(1), make(2)
    mk = lambda pcm: [rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).low_pass(300).amplify(0.9) for x in pcm]
    srcs = mk(pcm_a)
    with rb.Batch(srcs, 1, 48000, ctx=ctx) as b:
        assert b.kernel_family == 1, b.kernel_family            # k_fused_hot on the f32 copy, not the generic integer kernel
        group = b.mix_group
        b.upload_all()
        got_a = b.render_mix()
        for i, x in enumerate(pcm_b):                            # upload other PCM into the same batch
            b.upload(i, x)
        got_b = b.render_mix()
    for pcm, got in ((pcm_a, got_a), (pcm_b, got_b)):
        per = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in mk(pcm)]
        assert_bit_exact(got, grouped_expected_mix(per, [0] * n, got.size, group), "s16 input through the f32 kernels")


@pytest.mark.parametrize("shape", ["limit_only_mono", "reverb_limit_stereo", "spatial_reverb_agc_limit"])
def test_fx_kernel_with_limiter(ctx, shape):
    """`limit` behind (or instead of) the AGC inside k_fused_fx: a fourth recurrence warp (per-channel integrator and peak
    envelopes, the channels coupled through max(peaks), limit.rs:903-988) between the gain computer and the final factor.  The
    limiter's log2 / exp2 are the device's (within 2 ulp of glibc's): the 1e-5 * peak class of the general-path limiter."""
    ch = 1 if shape == "limit_only_mono" else 2
    n, frames = 21, 7000
    presets = [rb.LimitSettings.default(), rb.LimitSettings.dynamic_content(), rb.LimitSettings.broadcast(), rb.LimitSettings.gaming()]
    srcs = []
    for s in range(n):
        src = rb.TestSource(noise(frames * ch, 61000 + s, 1.5), ch, 48000)     # well above the threshold: the limiter works
        if shape == "spatial_reverb_agc_limit":
            src = rb.Spatial(src, [float(s % 5 - 2), 1.0, 0.0], [-1, 0, 0], [1, 0, 0])
        if shape != "limit_only_mono":
            src = src.reverb(rb.Duration.from_millis(10 + s), 0.5)
        if shape == "spatial_reverb_agc_limit":
            src = src.automatic_gain_control()
        srcs.append(src.limit(presets[s % 4]))
    streams = [to_oracle(s) for s in srcs]
    want = oracle.mixer(streams, ch, 48000)
    with rb.Batch(srcs, ch, 48000, ctx=ctx) as b:
        assert b.kernel_family == 5, b.kernel_family
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, want, 1e-5, f"k_fused_fx with a limiter, {shape}")
    with rb.Batch(srcs, ch, 48000, flags=GENERAL, ctx=ctx) as b:       # the general-path limiter uses the same device functions
        b.upload_all()
        ref = b.render_mix()
    per = None
    assert_close_peak(got, ref, 2e-6, f"k_fused_fx against the general path, {shape}")



# ------------------------------------------------------------------ k_lerp_mix: filter-free chains, parallel over the timeline
def test_lerp_mix_kernel_ragged_and_exact_order(ctx):
    """resample -> [gain] -> mix without a filter: k_lerp_mix (family 6).  Ragged lengths (empty, one frame, shorter than a tile),
    late joiners in phase (multiples of `to` = 160 frames), inputs outside the exact-reciprocal class, signed zeros; with
    RB_MIX_EXACT_ORDER the sum is the reference's sequential one, bit for bit."""
    rng = np.random.default_rng(8)
    n = 900
    lens = [int(v) for v in rng.integers(1, 6000, n)]
    lens[:6] = [0, 1, 2, 147, 148, 5999]
    starts = sorted(160 * int(v) for v in rng.integers(0, 12, n))
    pcms = [noise(L, 71000 + i) for i, L in enumerate(lens)]
    pcms[7][3:9] = [1e-42, -0.0, 0.0, 1e25, -1e-41, 5.0]
    pcms[8][:] = 0.0
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, 1, 44100), 1, 48000).amplify(0.7 + 0.001 * i) for i, p in enumerate(pcms)]
    streams = [to_oracle(s, st) for s, st in zip(srcs, starts)]
    ref = oracle.mixer(streams, 1, 48000)
    per = [oracle.chain_uniform(to_oracle(s), 1, 48000) for s in srcs]
    with rb.Batch(srcs, 1, 48000, mix_starts=starts, ctx=ctx) as b:
        assert b.kernel_family == 6, b.kernel_family
        group = b.mix_group
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, ref, 1e-5, "k_lerp_mix vs the sequential mixer")
    assert_bit_exact(got, grouped_expected_mix(per, starts, ref.size, group or n), "k_lerp_mix vs oracle streams in its documented order")
    with rb.Batch(srcs, 1, 48000, mix_starts=starts, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert b.kernel_family == 6 and b.mix_group == 0 and b.launches_per_render == 1
        b.upload_all()
        assert_bit_exact(b.render_mix(), ref, "k_lerp_mix, RB_MIX_EXACT_ORDER: the reference's order")


def test_lerp_mix_other_ratio_no_gain_and_out_of_phase_fallback(ctx):
    srcs = [rb.UniformSourceIterator(rb.TestSource(noise(2000 + i, 72000 + i), 1, 22050), 1, 48000) for i in range(200)]
    ref = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000)
    with rb.Batch(srcs, 1, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert b.kernel_family == 6
        b.upload_all()
        assert_bit_exact(b.render_mix(), ref, "22.05 -> 48 kHz, no gain, exact order")
    starts = sorted(int(v) for v in np.random.default_rng(9).integers(0, 999, 200))     # not in phase: another kernel serves it
    with rb.Batch(srcs, 1, 48000, mix_starts=starts, ctx=ctx) as b:
        assert b.kernel_family != 6
        b.upload_all()
        got = b.render_mix()
    assert_close_peak(got, oracle.mixer([to_oracle(s, st) for s, st in zip(srcs, starts)], 1, 48000), 1e-5, "out of phase")


# ------------------------------------------------------------------ RB_MIX_EXACT_ORDER on k_fused_hot: the running sum handed from CTA to CTA
def _check_exact_chain(ctx, srcs, ch, starts=None, family=1):
    streams = [to_oracle(s, st) for s, st in zip(srcs, starts or [0] * len(srcs))]
    ref = oracle.mixer(streams, ch, 48000)
    with rb.Batch(srcs, ch, 48000, flags=capi.RB_MIX_EXACT_ORDER, mix_starts=starts, ctx=ctx) as b:
        assert b.kernel_family == family, b.kernel_family
        if family == 1:
            assert b.mix_group == 0 and b.launches_per_render == 1
        b.upload_all()
        got = b.render_mix()
        again = b.render_mix()
    assert_bit_exact(got, ref, "exact order on the fused kernel vs the reference's sequential mixer")
    assert_bit_exact(again, ref, "second render (flags and ticket reset)")


def test_exact_order_headline_geometry_mono_4096(ctx):
    """The benchmarked batch with RB_MIX_EXACT_ORDER: k_fused_hot, 147 CTAs x 28 rows, every tile's sum started from the running sum of
    the CTA in front -- the WHOLE mixer output bit-identical to the reference's sequential sum over 4096 streams."""
    _check_exact_chain(ctx, _cfg3(4096, 4410), 1)


def test_exact_order_stereo_2048_and_two_waves(ctx):
    _check_exact_chain(ctx, _cfg3(2048, 2205, ch=2, seed=33000), 2)
    _check_exact_chain(ctx, _cfg3(5000, 700, seed=37000), 1)          # more CTAs than SMs: the ticket orders the chain


def test_exact_order_ragged_and_late_starts(ctx):
    """Streams of different lengths joining at different frames: partial tiles, CTAs whose rows are silent on a tile still hand the
    running sum on."""
    rng = np.random.default_rng(12)
    n = 700
    lens = [int(v) for v in rng.integers(1, 5000, n)]
    lens[:5] = [0, 1, 2, 300, 4999]
    starts = sorted(int(v) for v in rng.integers(0, 3000, n))
    srcs = [rb.UniformSourceIterator(rb.TestSource(noise(L, 91000 + i), 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for i, L in enumerate(lens)]
    _check_exact_chain(ctx, srcs, 1, starts=starts)


def test_exact_order_filter_free_4096(ctx):
    """resample -> amplify -> mix with RB_MIX_EXACT_ORDER at 4096 streams: the chain on k_fused_hot<1, false> (below 1024 streams:
    k_lerp_mix in one group, tests above)."""
    _check_exact_chain(ctx, _cfg3(4096, 3000, lp=None, seed=38000), 1)


def test_exact_order_effect_chain_cfg4(ctx):
    """BASELINE cfg4 with RB_MIX_EXACT_ORDER alone: k_fused_fx in its chain form (the running sum handed from CTA to CTA, one launch) --
    the whole stereo mix bit-identical to the reference's sequential mixer; also with a limiter behind the AGC within the limiter's
    tolerance, ragged lengths, and a render repeated (tickets and tags come round)."""
    n, frames = 512, 2400 + 2400
    srcs = cfg4_sources(n, frames)
    want = oracle.mixer([to_oracle(s) for s in srcs], 2, 48000)
    with rb.Batch(srcs, 2, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert b.kernel_family == 5 and b.mix_group == 0 and b.launches_per_render == 1, (b.kernel_family, b.mix_group, b.launches_per_render)
        b.upload_all()
        got = b.render_mix()
        again = b.render_mix()
    assert_bit_exact(got, want, "cfg4, exact order on k_fused_fx")
    assert_bit_exact(again, want, "cfg4, exact order, second render")
    rng = np.random.default_rng(23)
    ragged = [rb.TestSource(noise(2 * int(rng.integers(1, 3000)), 95000 + i), 2, 48000).reverb(rb.Duration.from_millis(10), 0.4)
              .automatic_gain_control() for i in range(203)]
    want = oracle.mixer([to_oracle(s) for s in ragged], 2, 48000)
    with rb.Batch(ragged, 2, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert b.kernel_family == 5 and b.mix_group == 0
        b.upload_all()
        assert_bit_exact(b.render_mix(), want, "ragged effect chains, exact order")
    lim = [s.limit() for s in srcs[:100]]
    want = oracle.mixer([to_oracle(s) for s in lim], 2, 48000)
    with rb.Batch(lim, 2, 48000, flags=capi.RB_MIX_EXACT_ORDER, ctx=ctx) as b:
        assert b.kernel_family == 5
        b.upload_all()
        assert_close_peak(b.render_mix(), want, 1e-5, "limiter behind the AGC, exact order")
