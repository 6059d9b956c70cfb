"""Executable specification of block-to-block state for the fused family:
    [amplify]* -> UniformSourceIterator (same channels) -> biquad -> [amplify]*
rendered block by block with explicit carried state, against the whole-stream oracle, bit for bit.
This is the contract of the CUDA state carry: rb_session_* (include/rodio_b200.h) implements it for the
uniform -> biquad -> amplify shape and is held to it by tests/test_lanes_emulator.py::test_session_* (CPU emulator) and
tests/test_parity_gpu.py::test_session_* (device); the pre-gain, AGC and reverb forms below are still specification only.
CPU-only, numpy float32 scalars (one rounding per operation, like the kernels)."""
from dataclasses import dataclass, field
from math import gcd

import numpy as np

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, noise, to_oracle

F = np.float32


@dataclass
class StreamState:
    """What has to survive between two blocks of one stream."""
    n_out: int = 0                       # output frames emitted so far (the resampler phase is n_out * from mod to)
    frames_in: int = 0                   # input frames received so far
    carry: np.ndarray = None             # input frames >= the left tap of the next output: [k, channels]
    carry_first: int = 0                 # absolute index of carry[0]
    biquad: list = field(default_factory=list)   # per channel [x1, x2, y1, y2]


def render_block(st: StreamState, block: np.ndarray, channels: int, from_rate: int, to_rate: int, co, pre: float,
                 post: float, final: bool) -> np.ndarray:
    """One block of interleaved input -> the output frames that are computable now."""
    g = gcd(from_rate, to_rate)
    fr, to = from_rate // g, to_rate // g
    x_new = (block.astype(F) * F(pre)).astype(F).reshape(-1, channels)      # amplify before the resampler
    x = x_new if st.carry is None else np.concatenate([st.carry, x_new])
    first = st.carry_first if st.carry is not None else st.frames_in
    st.frames_in += x_new.shape[0]
    L = st.frames_in
    if not st.biquad:
        st.biquad = [[F(0)] * 4 for _ in range(channels)]
    out = []
    last_done = False
    while not last_done:
        n = st.n_out
        i, num = (n * fr) // to, (n * fr) % to
        if fr == to:                       # pass-through (sample_rate.rs:131-134)
            if i >= L:
                break
            frame = x[i - first]
        elif i + 1 < L:                    # both taps have arrived: lerp with its three roundings (math.rs:24-26)
            a, b = x[i - first], x[i + 1 - first]
            frame = (a + ((b - a).astype(F) * F(num)).astype(F) / F(to)).astype(F)
        elif final and i == L - 1:         # the last input frame is emitted raw, once (sample_rate.rs:174-201)
            frame, last_done = x[i - first], True
        else:
            break
        b0, b1, b2, a1, a2 = (F(v) for v in co)
        y = np.empty(channels, F)
        for c in range(channels):          # DF1, strictly left to right (blt.rs:558-560)
            s = st.biquad[c]
            r = F(F(F(F(F(b0 * frame[c]) + F(b1 * s[0])) + F(b2 * s[1])) - F(a1 * s[2])) - F(a2 * s[3]))
            s[3], s[1], s[2], s[0] = s[2], s[0], r, frame[c]
            y[c] = F(r * F(post))
        out.append(y)
        st.n_out += 1
    # keep the frames the next output still needs: its left tap onwards
    keep_from = (st.n_out * fr) // to
    keep_from = min(max(keep_from, first), L)
    st.carry, st.carry_first = x[keep_from - first:].copy(), keep_from
    return np.concatenate(out) if out else np.zeros(0, F)


def test_block_rendering_equals_whole_stream():
    rng = np.random.default_rng(2024)
    for case, (channels, fr, to, hp) in enumerate([(1, 44100, 48000, False), (2, 48000, 44100, True), (1, 48000, 48000, False),
                                                   (2, 22050, 48000, False), (1, 96000, 44100, True)]):
        frames = int(rng.integers(900, 1600))
        x = noise(channels * frames, 50 + case, 0.8)
        src = rb.UniformSourceIterator(rb.TestSource(x, channels, fr).amplify(0.9), channels, to)
        src = (src.high_pass(700) if hp else src.low_pass(1200)).amplify(1.1)
        want = oracle.chain(to_oracle(src))[0]
        co = oracle.blt_coeffs(hp, 700 if hp else 1200, 0.5, to)
        for trial in range(4):
            cuts = sorted(set(int(c) for c in rng.integers(0, frames + 1, int(rng.integers(1, 9))))) + [frames]
            st, got, prev = StreamState(), [], 0
            for k, cut in enumerate(cuts):
                block = x[prev * channels: cut * channels]
                got.append(render_block(st, block, channels, fr, to, co, 0.9, 1.1, final=(k == len(cuts) - 1)))
                prev = cut
            assert st.carry.shape[0] <= 2 + fr // to, "the carried window stays a couple of frames"
            assert_bit_exact(np.concatenate(got), want, f"case {case} trial {trial} cuts {cuts}")


# ---------------------------------------------------------------------------------------------------------------
# AGC and reverb: what their state is (agc.rs:133-171,:433-504; mod.rs:628-634, delay.rs:8-16, mix.rs:43-53)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class AgcState:
    peak: np.float32 = F(0)
    sum: np.float32 = F(0)
    gain: np.float32 = F(1)
    ring: np.ndarray = None      # the last 8192 squares (equivalently: the last 8192 inputs), oldest first at `idx`
    idx: int = 0


def agc_block(st: AgcState, x: np.ndarray, attack, release, target, max_gain, floor) -> np.ndarray:
    if st.ring is None:
        st.ring = np.zeros(8192, F)
    y = np.empty_like(x)
    for n, s in enumerate(x):
        v = F(abs(s))
        c = F(0) if v > st.peak else release
        st.peak = F(F(st.peak * c) + F(v * F(F(1) - c)))
        sq = F(v * v)
        st.sum = F(F(st.sum - st.ring[st.idx]) + sq)
        st.ring[st.idx] = sq
        st.idx = (st.idx + 1) & 8191
        rms = F(np.sqrt(F(st.sum / F(8192))))
        rms_gain = F(target / rms) if rms > 0 else max_gain
        peak_gain = min(F(target / st.peak), max_gain) if st.peak > 0 else max_gain
        desired = max(min(rms_gain, peak_gain), floor)
        k = attack if desired > st.gain else release
        st.gain = max(F(0.1), min(F(F(st.gain * k) + F(desired * F(F(1) - k))), max_gain))
        y[n] = F(s * st.gain)
    return y


def test_agc_state_is_three_scalars_and_the_ring():
    x = np.concatenate([noise(9000, 61, 0.05), noise(5000, 62, 0.9), noise(4000, 63, 0.2)]).astype(F)
    settings = rb.AutomaticGainControlSettings()
    want = oracle.chain(to_oracle(rb.TestSource(x, 2, 44100).automatic_gain_control(settings)))[0]
    coef = oracle.lib().ro_duration_to_coefficient
    attack, release = F(coef(settings.attack_time, 44100)), F(coef(settings.release_time, 44100))
    rng = np.random.default_rng(5)
    cuts = sorted(set(int(c) for c in rng.integers(0, x.size, 6))) + [x.size]
    st, got, prev = AgcState(), [], 0
    for cut in cuts:
        got.append(agc_block(st, x[prev:cut], attack, release, F(settings.target_level), F(settings.absolute_max_gain), F(settings.floor)))
        prev = cut
    assert_bit_exact(np.concatenate(got), want, f"agc in blocks {cuts}")


def test_reverb_state_is_the_last_d_inputs():
    channels, rate, ms, amp = 2, 44100, 7, 0.6
    x = noise(channels * 4000, 71, 0.8)
    want = oracle.chain(to_oracle(rb.TestSource(x, channels, rate).reverb(rb.Duration.from_millis(ms), amp)))[0]
    d = ms * 1_000_000 * channels * rate // 1_000_000_000
    hist = np.zeros(d, F)                      # scaled inputs still on their way through the delay line
    rng = np.random.default_rng(6)
    cuts = sorted(set(int(c) for c in rng.integers(0, x.size, 5))) + [x.size]
    got, prev = [], 0
    for cut in cuts:
        blk = x[prev:cut]
        line = np.concatenate([hist, (blk * F(amp)).astype(F)])
        got.append((blk + line[: blk.size]).astype(F))
        hist, prev = line[blk.size:], cut
    got.append(hist)                           # the tail: the survivor of the Mix (mix.rs:47-50) is the echo alone
    assert_bit_exact(np.concatenate(got), want, f"reverb in blocks {cuts}")
