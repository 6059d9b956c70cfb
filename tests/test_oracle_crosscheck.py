"""Second, independent restatement of the adapters the reference has NO result-pinning tests for (SURVEY 8c:
biquad, AGC, reverb) -- numpy float32 scalars, one rounding per operation, written from the reference files -- held
bit for bit against the C++ oracle.  Two restatements in different languages agreeing on every bit is what stands
in for the missing golden vectors.  (Transcendental inputs -- filter coefficients, exp() of the time constants --
are taken from the oracle's own helpers, which test_oracle_golden.py pins.)"""
import ctypes as C

import numpy as np

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, noise, to_oracle

F = np.float32


def biquad_df1(x, co, channels):
    """blt.rs:397-410 (mono), :431-451 / :472-492 (per-channel state); apply() :558-560, strictly left to right."""
    b0, b1, b2, a1, a2 = (F(v) for v in co)
    st = [[F(0), F(0), F(0), F(0)] for _ in range(channels)]   # x_n1, x_n2, y_n1, y_n2
    y = np.empty_like(x)
    for n, xn in enumerate(x):
        s = st[n % channels]
        r = F(F(F(F(F(b0 * xn) + F(b1 * s[0])) + F(b2 * s[1])) - F(a1 * s[2])) - F(a2 * s[3]))
        s[3], s[1], s[2], s[0] = s[2], s[0], r, xn
        y[n] = r
    return y


def agc(x, rate, target, attack_ns, release_ns, max_gain, floor):
    """agc.rs:433-504 with update_peak_level :397-408, CircularBuffer :139-171 (8192 squares, running sum),
    calculate_peak_gain :424-431; time limits source/mod.rs:432-433; state shared by the interleaved channels."""
    coef = oracle.lib().ro_duration_to_coefficient
    attack = F(coef(min(attack_ns, 10_000_000_000), rate))
    release = F(coef(min(release_ns, 10_000_000_000), rate))
    target, max_gain, floor = F(target), F(max_gain), F(floor)
    ring = np.zeros(8192, F)
    ssum, idx, peak, gain = F(0), 0, F(0), F(1)
    y = np.empty_like(x)
    for n, s in enumerate(x):
        v = F(abs(s))
        c = F(0) if v > peak else release
        peak = F(F(peak * c) + F(v * F(F(1) - c)))
        sq = F(v * v)
        ssum = F(F(ssum - ring[idx]) + sq)
        ring[idx] = sq
        idx = (idx + 1) & 8191
        rms = F(np.sqrt(F(ssum / F(8192))))
        rms_gain = F(target / rms) if rms > 0 else max_gain
        peak_gain = min(F(target / peak), max_gain) if peak > 0 else max_gain
        desired = max(min(rms_gain, peak_gain), floor)
        k = attack if desired > gain else release
        gain = F(F(gain * k) + F(desired * F(F(1) - k)))
        gain = max(F(0.1), min(gain, max_gain))      # clamp(0.1, max): no NaN can reach it
        y[n] = F(s * gain)
    return y


def reverb(x, channels, rate, delay_ns, amplitude):
    """source/mod.rs:628-634: Mix(x, Delay(Amplify(x))) -- delay.rs:8-16 counts interleaved samples in u128, the
    silence is literal 0.0, mix.rs:43-53 adds while both run and passes the survivor through."""
    d = delay_ns * channels * rate // 1_000_000_000
    echo = np.concatenate([np.zeros(d, F), (x * F(amplitude)).astype(F)])
    y = echo.copy()
    y[: x.size] = x + echo[: x.size]
    return y


def test_biquad_recurrence_mono_and_stereo():
    for channels, high, freq, q, fs in [(1, False, 200, 0.5, 48000), (2, True, 3000, 0.7, 44100), (3, False, 9000, 2.0, 96000)]:
        x = noise(channels * 2500, 900 + channels, 0.9)
        co = oracle.blt_coeffs(high, freq, q, fs)
        src = rb.TestSource(x, channels, fs)
        src = src.high_pass_with_q(freq, q) if high else src.low_pass_with_q(freq, q)
        assert_bit_exact(oracle.chain(to_oracle(src))[0], biquad_df1(x, co, channels), f"biquad {channels} ch")


def test_agc_default_and_custom_settings():
    x = np.concatenate([noise(9000, 31, 0.05), noise(6000, 32, 0.9), np.zeros(500, F), noise(4000, 33, 0.3)]).astype(F)
    st = rb.AutomaticGainControlSettings()
    got = oracle.chain(to_oracle(rb.TestSource(x, 2, 44100).automatic_gain_control(st)))[0]
    assert_bit_exact(got, agc(x, 44100, st.target_level, st.attack_time, st.release_time, st.absolute_max_gain, st.floor), "agc default")
    st = rb.AutomaticGainControlSettings(target_level=0.6, attack_time=rb.Duration.from_millis(250),
                                         release_time=rb.Duration.from_millis(40), absolute_max_gain=3.0, floor=0.5)
    got = oracle.chain(to_oracle(rb.TestSource(x, 1, 16000).automatic_gain_control(st)))[0]
    assert_bit_exact(got, agc(x, 16000, st.target_level, st.attack_time, st.release_time, st.absolute_max_gain, st.floor), "agc custom")


def test_reverb_is_source_plus_delayed_scaled_copy():
    for channels, rate, ms, amp in [(1, 48000, 50, 0.3), (2, 44100, 7, 0.7), (3, 32000, 0, 0.5)]:
        x = noise(channels * 3000, 77 + channels, 0.8)
        src = rb.TestSource(x, channels, rate).reverb(rb.Duration.from_millis(ms), amp)
        assert_bit_exact(oracle.chain(to_oracle(src))[0], reverb(x, channels, rate, ms * 1_000_000, amp), f"reverb {channels} ch")
