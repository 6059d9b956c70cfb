"""rodio_b200 — B200-native block implementation of rodio's resample -> channel-map -> effects -> mix path.

The compute path is the in-tree CUDA library `rodio_b200/librodio_b200.so` (sm_100a) behind the C ABI in
`include/rodio_b200.h`; this package is the host-side mirror of rodio's `Source` / `Mixer` surface.
There is no CPU fallback: importing works without a GPU (so the CPU test-suite can check symbols), but
the first call that needs a device fails loudly.
"""
from . import _capi as capi
from ._capi import RodioB200Error, lib
from .source import (AutomaticGainControlSettings, Batch, ChannelCountConverter, ChannelVolume, Comm, Context, Duration, wav_source,
                     Effect, LimitSettings, Mixer, MixerSource, Player, SampleRateConverter, SamplesBuffer,
                     SampleTypeConverter, Session, Source, Spatial, TestSource, UniformSourceIterator, default_context, mixer, plan,
                     Function, from_iter, SignalGenerator, SineWave, SquareWave, TriangleWave, SawtoothWave)

__all__ = [
    "capi", "lib", "RodioB200Error", "AutomaticGainControlSettings", "Batch", "ChannelCountConverter",
    "ChannelVolume", "Comm", "Context", "Duration", "Effect", "LimitSettings", "Mixer", "MixerSource", "Player",
    "SampleRateConverter", "SamplesBuffer", "SampleTypeConverter", "Session", "Source", "Spatial", "TestSource",
    "UniformSourceIterator", "default_context", "mixer", "plan", "wav_source",
    "Function", "from_iter", "SignalGenerator", "SineWave", "SquareWave", "TriangleWave", "SawtoothWave",
]
