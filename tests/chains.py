"""Adapter chains shared by the CPU planner tests and the GPU parity tests."""
import numpy as np

import oracle
import rodio_b200 as rb
from helpers import noise


def _stereo(n, seed):
    return noise(2 * n, seed, 0.9)


CHAINS = {
    "amplify": lambda: rb.SamplesBuffer(2, 44100, _stereo(3000, 1)).amplify(1.2),
    "amplify_decibel": lambda: rb.SamplesBuffer(1, 48000, noise(1000, 2)).amplify_decibel(-6.0),
    "low_pass_mono": lambda: rb.TestSource(noise(20000, 3), 1, 48000).low_pass(200),
    "low_pass_stereo": lambda: rb.TestSource(_stereo(12000, 4), 2, 44100).low_pass(200),
    "high_pass_3ch": lambda: rb.TestSource(noise(3 * 5000, 5), 3, 44100).high_pass(300),
    "low_pass_q": lambda: rb.TestSource(noise(8000, 6), 1, 48000).low_pass_with_q(1000, 0.707),
    "reverb": lambda: rb.SamplesBuffer(2, 48000, _stereo(9000, 7)).reverb(rb.Duration.from_millis(50), 0.3),
    "reverb_longer_than_input": lambda: rb.SamplesBuffer(1, 48000, noise(100, 8)).reverb(rb.Duration.from_millis(10), 0.5),
    "delay": lambda: rb.SamplesBuffer(2, 44100, _stereo(500, 9)).delay(rb.Duration.from_millis(7)),
    "agc_default": lambda: rb.TestSource(noise(30000, 10, 0.3), 1, 48000).automatic_gain_control(),
    "agc_stereo_custom": lambda: rb.TestSource(_stereo(15000, 11), 2, 44100).automatic_gain_control(
        rb.AutomaticGainControlSettings(0.8, rb.Duration.from_millis(300), rb.Duration.from_millis(20), 5.0, 0.2)),
    "spatial": lambda: rb.Spatial(rb.SamplesBuffer(2, 48000, _stereo(4000, 12)), [2.0, 1.0, 0.0], [-1, 0, 0], [1, 0, 0]),
    "channel_volume_1_to_2": lambda: rb.ChannelVolume(rb.SamplesBuffer(1, 44100, noise(999, 13)), [0.5, 0.8]),
    "channel_volume_2_to_6": lambda: rb.ChannelVolume(rb.SamplesBuffer(2, 44100, _stereo(999, 14)), [1, .9, .8, .7, .6, .5]),
    "speed_changes_ratio": lambda: rb.UniformSourceIterator(rb.TestSource(_stereo(5000, 15), 2, 44100).speed(0.9), 2, 48000),
    "uniform_spanless": lambda: rb.UniformSourceIterator(rb.TestSource(_stereo(50000, 16), 2, 44100), 2, 48000),
    "uniform_spans_32768": lambda: rb.UniformSourceIterator(rb.SamplesBuffer(2, 44100, _stereo(50000, 17)), 1, 48000),
    "uniform_small_span": lambda: rb.UniformSourceIterator(rb.Source(noise(10000, 18), 1, 44100, span_len=1024), 2, 22050),
    "uniform_then_effects": lambda: rb.UniformSourceIterator(rb.TestSource(noise(20000, 19), 1, 44100), 1, 48000)
        .low_pass(200).amplify(1.2),
    "pipeline_long_like": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(30000, 20), 2, 44100).high_pass(300).amplify(1.2).speed(0.9).automatic_gain_control()
        .delay(rb.Duration.from_millis(20)).reverb(rb.Duration.from_secs_f32(0.05), 0.3), 2, 40000),
    "misaligned_delay_then_uniform": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(3000, 40), 2, 44100).delay(rb.Duration.from_nanos(12000)), 2, 48000),
    "misaligned_delay_then_downmix": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(3000, 41), 2, 44100).delay(rb.Duration.from_nanos(12000)), 1, 48000),
    "misaligned_3ch_partial2": lambda: rb.UniformSourceIterator(
        rb.TestSource(noise(3 * 700, 42), 3, 48000).delay(rb.Duration.from_nanos(14000)), 3, 32000),
    "misaligned_passthrough_upmix": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(500, 43), 2, 48000).delay(rb.Duration.from_nanos(12000)), 4, 48000),
    "misaligned_biquad": lambda: rb.TestSource(_stereo(2000, 44), 2, 44100).delay(rb.Duration.from_nanos(12000)).low_pass(300),
    "bench_long_shape": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(22050, 45), 2, 44100).high_pass(300).amplify(1.2).speed(0.9).automatic_gain_control()
        .delay(rb.Duration.from_secs_f32(0.5)).reverb(rb.Duration.from_secs_f32(0.05), 0.3), 2, 40000),
    "distortion": lambda: rb.TestSource(_stereo(3000, 50), 2, 44100).distortion(4.0, 0.3),
    "fade_in_stereo": lambda: rb.TestSource(_stereo(6000, 51), 2, 44100).fade_in(rb.Duration.from_millis(50)),
    "fade_out_mono": lambda: rb.SamplesBuffer(1, 48000, noise(9000, 52)).fade_out(rb.Duration.from_millis(100)),
    "linear_ramp_3ch": lambda: rb.TestSource(noise(3 * 2500, 53), 3, 32000).linear_gain_ramp(
        rb.Duration.from_secs_f32(0.03), 0.25, 2.0, True),
    "take_duration": lambda: rb.TestSource(_stereo(8000, 54), 2, 44100).take_duration(rb.Duration.from_millis(123)),
    "take_duration_fadeout": lambda: rb.TestSource(noise(20000, 55), 1, 48000).take_duration(
        rb.Duration.from_millis(300), filter_fadeout=True),
    "take_longer_than_input": lambda: rb.SamplesBuffer(2, 44100, _stereo(1000, 56)).take_duration(rb.Duration.from_secs(3)),
    "take_pads_frame": lambda: rb.TestSource(noise(3 * 4000, 57), 3, 44100).take_duration(rb.Duration.from_nanos(7_566_000)),
    "bench_long_full": lambda: rb.UniformSourceIterator(
        rb.TestSource(_stereo(44100, 58), 2, 44100).high_pass(300).amplify(1.2).speed(0.9).automatic_gain_control()
        .delay(rb.Duration.from_secs_f32(0.1)).fade_in(rb.Duration.from_secs_f32(0.4))
        .take_duration(rb.Duration.from_secs(1), filter_fadeout=True).reverb(rb.Duration.from_secs_f32(0.05), 0.3), 2, 40000),
    "i16_input": lambda: rb.SamplesBuffer(2, 44100, (noise(4000, 21) * 30000).astype(np.int16)).amplify(0.5).low_pass(500),
    # round 2, second half: sources generated on the device, two-input adapters
    "signal_sine_chain": lambda: rb.SineWave(440.0).take(5000).amplify(0.5).low_pass(2000),
    "signal_triangle": lambda: rb.SignalGenerator(44100, 523.25, rb.Function.Triangle).take(3000),
    "signal_square_fade": lambda: rb.SquareWave(80.0).take(4000).fade_in(rb.Duration.from_millis(30)),
    "mix_other_rate_and_channels": lambda: rb.TestSource(_stereo(3000, 60), 2, 44100).mix(rb.TestSource(noise(1000, 61), 1, 32000)),
    "mix_generator_longer": lambda: rb.SamplesBuffer(1, 48000, noise(700, 62)).amplify(0.5)
        .mix(rb.SawtoothWave(220.0).take(2500).amplify(0.25)).high_pass(100),
    "crossfade_stereo": lambda: rb.TestSource(_stereo(30000, 63), 2, 44100).take_crossfade_with(
        rb.TestSource(_stereo(20000, 64), 2, 44100), rb.Duration.from_millis(250)),
    # Pausable with Player::pause / play scripted (pausable.rs:85-97): the filter in front is not pulled while paused
    "pause_behind_filter": lambda: rb.TestSource(_stereo(4000, 67), 2, 44100).low_pass(300).pause_at(1001, 50).amplify(0.5),
    "pause_at_start_and_end": lambda: rb.SamplesBuffer(1, 48000, noise(700, 68)).pause_at(0, 10).pause_at(710, 7),
    "pause_then_mixer_conversion": lambda: rb.UniformSourceIterator(rb.TestSource(noise(3 * 900, 69), 3, 32000).pause_at(300, 40), 2, 48000),
    # a take inside a take that outlasts it: the inner padding is handed on and still never pulled by a UniformSourceIterator
    "take_inside_take_then_uniform": lambda: rb.UniformSourceIterator(
        rb.SamplesBuffer(2, 44100, _stereo(1500, 70)).take_duration(rb.Duration.from_millis(26))
        .take_duration(rb.Duration.from_millis(28)).fade_in(rb.Duration.from_millis(28)), 1, 48000),
    "crossfade_of_a_taken_source": lambda: rb.TestSource(noise(1, 71), 1, 48000).amplify(0.5).take_crossfade_with(
        rb.SamplesBuffer(2, 44100, _stereo(1500, 72)).take_duration(rb.Duration.from_millis(26)), rb.Duration.from_millis(28)),
    "crossfade_other_format": lambda: rb.SamplesBuffer(2, 44100, _stereo(30000, 65)).take_crossfade_with(
        rb.SamplesBuffer(1, 22050, noise(9000, 66)), rb.Duration.from_millis(150)).amplify(0.7),
}



LIMIT_CHAINS = {
    "limit_default_mono": lambda: rb.TestSource(oracle.sine_wave(440.0, 6000), 1, 48000).amplify(3.0).limit(),
    "limit_stereo": lambda: rb.SamplesBuffer(2, 44100, _stereo(8000, 30) * np.float32(1.6)).limit(
        rb.LimitSettings.default().with_threshold(-3.0)),
    "limit_multi": lambda: rb.TestSource(noise(4 * 3000, 31, 1.5), 4, 48000).limit(rb.LimitSettings.broadcast()),
    "limit_hard_knee_ish": lambda: rb.TestSource(noise(6000, 32, 2.0), 1, 48000).limit(
        rb.LimitSettings.mastering()),
}


