//! `impl Source` on top of the C ABI in include/rodio_b200.h.
//!
//! The per-sample pull model of `rodio::Source` (reference src/source/mod.rs:179-218) survives only
//! here: `GpuMixerSource::next()` hands out samples from a block that one `rb_batch_render_mix` call
//! produced on the GPU.  NOT COMPILED in the build image (no Rust toolchain) — reviewed by hand against
//! the header; every `extern "C"` item below is declared 1:1 in include/rodio_b200.h.
#![allow(non_camel_case_types)]
use std::ffi::c_void;
use std::num::NonZero;
use std::time::Duration;

use rodio::source::SeekError;
use rodio::{ChannelCount, Sample, SampleRate, Source};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rb_effect {
    pub kind: u32,
    pub u32_: [u32; 3],
    pub f32_: [f32; 12],
    pub ns: [u64; 2],
}

#[repr(C)]
pub struct rb_stream_desc {
    pub sample_rate: u32,
    pub channels: u16,
    pub format: u16,
    pub n_samples: u64,
    pub span_len: u32,
    pub n_effects: u32,
    pub effects: *const rb_effect,
    pub mix_start: u64,
}

pub enum rb_context {}
pub enum rb_batch {}

extern "C" {
    fn rb_context_create(device: i32, out: *mut *mut rb_context) -> i32;
    fn rb_context_destroy(ctx: *mut rb_context) -> i32;
    fn rb_batch_create(ctx: *mut rb_context, mixer_channels: u16, mixer_rate: u32, descs: *const rb_stream_desc,
                       n: usize, flags: u32, out: *mut *mut rb_batch) -> i32;
    fn rb_batch_destroy(b: *mut rb_batch) -> i32;
    fn rb_batch_upload(b: *mut rb_batch, stream: usize, pcm: *const c_void, n_samples: u64) -> i32;
    fn rb_batch_mix_len(b: *mut rb_batch, n: *mut u64) -> i32;
    fn rb_batch_render_mix(b: *mut rb_batch, out: *mut f32, max_samples: u64, written: *mut u64) -> i32;
    fn rb_speed_sample_rate(input_rate: u32, factor: f32) -> u32;
}

pub const RB_FX_AMPLIFY: u32 = 1;
pub const RB_FX_SPEED: u32 = 2;
pub const RB_FX_LOW_PASS: u32 = 3;
pub const RB_FX_HIGH_PASS: u32 = 4;
pub const RB_FX_REVERB: u32 = 5;
pub const RB_FX_AGC: u32 = 6;
pub const RB_FX_LIMIT: u32 = 7;
pub const RB_FX_SPATIAL: u32 = 8;
pub const RB_FX_CHANNEL_VOLUME: u32 = 9;
pub const RB_FX_UNIFORM: u32 = 10;
pub const RB_FX_DELAY: u32 = 11;
pub const RB_FX_DISTORTION: u32 = 12;
pub const RB_FX_LINEAR_RAMP: u32 = 13;
pub const RB_FX_TAKE_DURATION: u32 = 14;

/// An in-memory source plus the adapters recorded on it (what `SamplesBuffer::new(..).amplify(..)` builds).
pub struct GpuSource {
    pcm: Vec<Sample>,
    channels: ChannelCount,
    sample_rate: SampleRate,
    span_len: u32,
    effects: Vec<rb_effect>,
    reported_rate: SampleRate,
}

impl GpuSource {
    /// `SamplesBuffer::new(channels, sample_rate, data)` (src/buffer.rs:40-60)
    pub fn from_samples(channels: ChannelCount, sample_rate: SampleRate, data: Vec<Sample>) -> Self {
        let span_len = data.len().min(u32::MAX as usize) as u32;
        Self { pcm: data, channels, sample_rate, span_len, effects: vec![], reported_rate: sample_rate }
    }
    fn push(mut self, kind: u32, u: [u32; 3], f: &[f32], ns: [u64; 2]) -> Self {
        let mut f32_ = [0f32; 12];
        f32_[..f.len()].copy_from_slice(f);
        self.effects.push(rb_effect { kind, u32_: u, f32_, ns });
        self
    }
    /// `Source::amplify` (src/source/mod.rs:307-314)
    pub fn amplify(self, value: f32) -> Self { self.push(RB_FX_AMPLIFY, [0; 3], &[value], [0; 2]) }
    /// `Source::speed` (src/source/speed.rs:103-105,:130-133)
    pub fn speed(mut self, ratio: f32) -> Self {
        let r = unsafe { rb_speed_sample_rate(self.reported_rate.get(), ratio) };
        self.reported_rate = NonZero::new(r).expect("minimum is 1");
        self.push(RB_FX_SPEED, [0; 3], &[ratio], [0; 2])
    }
    /// `Source::low_pass` (src/source/mod.rs:686-692)
    pub fn low_pass(self, freq: u32) -> Self { self.push(RB_FX_LOW_PASS, [freq, 0, 0], &[0.5], [0; 2]) }
    /// `Source::high_pass`
    pub fn high_pass(self, freq: u32) -> Self { self.push(RB_FX_HIGH_PASS, [freq, 0, 0], &[0.5], [0; 2]) }
    /// `Source::reverb` (src/source/mod.rs:628-634)
    pub fn reverb(self, duration: Duration, amplitude: f32) -> Self {
        self.push(RB_FX_REVERB, [0; 3], &[amplitude], [duration.as_nanos() as u64, 0])
    }
    /// `Source::delay` (src/source/delay.rs:19-29)
    pub fn delay(self, duration: Duration) -> Self { self.push(RB_FX_DELAY, [0; 3], &[], [duration.as_nanos() as u64, 0]) }
    /// `Source::distortion` (src/source/mod.rs:726-731)
    pub fn distortion(self, gain: f32, threshold: f32) -> Self { self.push(RB_FX_DISTORTION, [0; 3], &[gain, threshold], [0; 2]) }
    /// `Source::linear_gain_ramp` (src/source/mod.rs:534-546); `fade_in` / `fade_out` are the (0,1,false) / (1,0,true) cases
    pub fn linear_gain_ramp(self, duration: Duration, start: f32, end: f32, clamp_end: bool) -> Self {
        self.push(RB_FX_LINEAR_RAMP, [clamp_end as u32, 0, 0], &[start, end], [duration.as_nanos() as u64, 0])
    }
    /// `Source::take_duration` (src/source/take.rs:9-26); `fadeout` = `TakeDuration::set_filter_fadeout`
    pub fn take_duration(self, duration: Duration, fadeout: bool) -> Self {
        self.push(RB_FX_TAKE_DURATION, [fadeout as u32, 0, 0], &[], [duration.as_nanos() as u64, 0])
    }
    /// `Source::automatic_gain_control` (src/source/mod.rs:415-446): target, max gain, floor; attack / release times
    pub fn automatic_gain_control(self, target: f32, attack: Duration, release: Duration, max_gain: f32) -> Self {
        self.push(RB_FX_AGC, [0; 3], &[target, max_gain, 0.0], [attack.as_nanos() as u64, release.as_nanos() as u64])
    }
    /// `Source::limit` (src/source/limit.rs:94-130): threshold dB, knee dB, attack, release
    pub fn limit(self, threshold: f32, knee_width: f32, attack: Duration, release: Duration) -> Self {
        self.push(RB_FX_LIMIT, [0; 3], &[threshold, knee_width], [attack.as_nanos() as u64, release.as_nanos() as u64])
    }
}

/// `mixer::mixer(channels, sample_rate)` (src/mixer.rs:25-43): `add` is infallible like the reference.
pub struct GpuMixer {
    ctx: *mut rb_context,
    channels: ChannelCount,
    sample_rate: SampleRate,
    sources: Vec<(GpuSource, u64)>,
    position: u64,
}

impl GpuMixer {
    pub fn new(channels: ChannelCount, sample_rate: SampleRate) -> Result<Self, i32> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { rb_context_create(0, &mut ctx) };
        if st != 0 { return Err(st); }
        Ok(Self { ctx, channels, sample_rate, sources: vec![], position: 0 })
    }
    /// `Mixer::add` (src/mixer.rs:58-66)
    pub fn add(&mut self, source: GpuSource) { self.sources.push((source, self.position)); }

    /// Drain everything added so far into a block-backed `Source`.
    pub fn into_source(self) -> Result<GpuMixerSource, i32> {
        let descs: Vec<rb_stream_desc> = self.sources.iter().map(|(s, start)| rb_stream_desc {
            sample_rate: s.sample_rate.get(), channels: s.channels.get(), format: 0, n_samples: s.pcm.len() as u64,
            span_len: s.span_len, n_effects: s.effects.len() as u32, effects: s.effects.as_ptr(), mix_start: *start,
        }).collect();
        let mut b = std::ptr::null_mut();
        let st = unsafe { rb_batch_create(self.ctx, self.channels.get(), self.sample_rate.get(), descs.as_ptr(),
                                          descs.len(), 0, &mut b) };
        if st != 0 { return Err(st); }
        for (i, (s, _)) in self.sources.iter().enumerate() {
            let st = unsafe { rb_batch_upload(b, i, s.pcm.as_ptr() as *const c_void, s.pcm.len() as u64) };
            if st != 0 { return Err(st); }
        }
        let mut n = 0u64;
        unsafe { rb_batch_mix_len(b, &mut n) };
        let mut block = vec![0f32; n as usize];
        let mut written = 0u64;
        let st = unsafe { rb_batch_render_mix(b, block.as_mut_ptr(), n, &mut written) };
        unsafe { rb_batch_destroy(b); rb_context_destroy(self.ctx); }
        if st != 0 { return Err(st); }
        block.truncate(written as usize);
        Ok(GpuMixerSource { block, pos: 0, channels: self.channels, sample_rate: self.sample_rate })
    }
}

/// `mixer::MixerSource` (src/mixer.rs:70-136) backed by one rendered block.
pub struct GpuMixerSource {
    block: Vec<Sample>,
    pos: usize,
    channels: ChannelCount,
    sample_rate: SampleRate,
}

impl Iterator for GpuMixerSource {
    type Item = Sample;
    #[inline]
    fn next(&mut self) -> Option<Sample> {
        let s = self.block.get(self.pos).copied();
        self.pos += 1;
        s
    }
}

impl Source for GpuMixerSource {
    fn current_span_len(&self) -> Option<usize> { None }               // src/mixer.rs:88-90
    fn channels(&self) -> ChannelCount { self.channels }
    fn sample_rate(&self) -> SampleRate { self.sample_rate }
    fn total_duration(&self) -> Option<Duration> { None }
    fn try_seek(&mut self, _: Duration) -> Result<(), SeekError> {       // src/mixer.rs:109-113
        Err(SeekError::NotSupported { underlying_source: std::any::type_name::<Self>() })
    }
}
