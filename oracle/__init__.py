"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE — see oracle/rodio_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this package.  The product package `rodio_b200` never does.

Effects are passed duck-typed: any object with `.kind`, `.u32` (3 ints), `.f32` (12 floats) and
`.ns` (2 ints) works, so this module needs nothing from the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librodio_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/librodio_oracle.so with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("rodio_oracle_capi.cpp", "rodio_oracle.hpp", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "librodio_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Effect(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("u32", C.c_uint32 * 3), ("f32", C.c_float * 12), ("ns", C.c_uint64 * 2)]


class _Stream(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("channels", C.c_uint16),
        ("format", C.c_uint16),
        ("n_samples", C.c_uint64),
        ("span_len", C.c_uint32),
        ("n_effects", C.c_uint32),
        ("effects", C.POINTER(_Effect)),
        ("mix_start", C.c_uint64),
        ("pcm", C.POINTER(C.c_float)),
    ]


assert C.sizeof(_Effect) == 80

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.ro_lerp.restype = C.c_float
        L.ro_lerp.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32]
        L.ro_db_to_linear.restype = C.c_float
        L.ro_db_to_linear.argtypes = [C.c_float]
        L.ro_linear_to_db.restype = C.c_float
        L.ro_linear_to_db.argtypes = [C.c_float]
        L.ro_duration_to_coefficient.restype = C.c_float
        L.ro_duration_to_coefficient.argtypes = [C.c_uint64, C.c_uint32]
        L.ro_speed_sample_rate.restype = C.c_uint32
        L.ro_speed_sample_rate.argtypes = [C.c_uint32, C.c_float]
        L.ro_delay_samples.restype = C.c_uint64
        L.ro_delay_samples.argtypes = [C.c_uint64, C.c_uint32, C.c_uint16]
        for fn in (L.ro_mixer_mt, L.ro_mixer_mt_static):
            fn.argtypes = [C.POINTER(_Stream), C.c_uint64, C.c_uint16, C.c_uint32, C.c_int,
                           C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib = L
    return _lib


@dataclass
class Stream:
    """Neutral description of one mixer input: PCM + adapter chain (same meaning as rb_stream_desc)."""
    pcm: np.ndarray                      # float32, interleaved
    channels: int
    sample_rate: int
    effects: Sequence = field(default_factory=list)
    span_len: int = 0                    # 0 = None
    mix_start: int = 0


@dataclass
class Fx:
    """One adapter of a chain in the oracle's own terms (same layout as the 80-byte effect record of the C side), so that
    callers which must not touch the product package (bench.py --impl reference) can build chains."""
    kind: int
    u32: Sequence = (0, 0, 0)
    f32: Sequence = (0.0,) * 12
    ns: Sequence = (0, 0)
    other: Optional["Stream"] = None     # FX_MIX: the second input of Source::mix (a Stream of its own)


FX_AMPLIFY, FX_LOW_PASS, FX_HIGH_PASS, FX_UNIFORM, FX_SIGNAL, FX_MIX, FX_APPEND = 1, 3, 4, 10, 15, 16, 17     # rodio_oracle_capi.cpp, enum of adapter kinds
MIX_START_CONSUMED = 0xFFFFFFFFFFFFFFFF


def fx(kind: int, u32=(), f32=(), ns=()) -> Fx:
    return Fx(kind, tuple(int(v) for v in u32) + (0,) * (3 - len(u32)), tuple(float(v) for v in f32) + (0.0,) * (12 - len(f32)),
              tuple(int(v) for v in ns) + (0,) * (2 - len(ns)))


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pack(streams: Sequence[Stream]):
    """The ro_stream array of `streams`, followed by the second inputs of their FX_MIX adapters (entries of their own, marked
    consumed: the mixer skips them; the adapter finds its second input through the entry's address in u32[1] / u32[2])."""
    flat = list(streams)
    i = 0
    while i < len(flat):
        for e in flat[i].effects:
            o = getattr(e, "other", None)
            if int(e.kind) in (FX_MIX, FX_APPEND) and o is not None and not any(o is f for f in flat):
                flat.append(o)
        i += 1
    keep = []
    arr = (_Stream * max(1, len(flat)))()
    base = C.addressof(arr)
    for i, s in enumerate(flat):
        pcm = np.ascontiguousarray(s.pcm, dtype=np.float32)
        fx = (_Effect * max(1, len(s.effects)))()
        for j, e in enumerate(s.effects):
            fx[j].kind = int(e.kind)
            for k in range(3):
                fx[j].u32[k] = int(e.u32[k])
            for k in range(12):
                fx[j].f32[k] = float(e.f32[k])
            for k in range(2):
                fx[j].ns[k] = int(e.ns[k])
            if int(e.kind) in (FX_MIX, FX_APPEND):
                addr = base + C.sizeof(_Stream) * next(k for k, f in enumerate(flat) if f is e.other)
                fx[j].u32[1], fx[j].u32[2] = addr & 0xFFFFFFFF, addr >> 32
        keep += [pcm, fx]
        arr[i].sample_rate = s.sample_rate
        arr[i].channels = s.channels
        arr[i].format = 0
        arr[i].n_samples = pcm.size
        arr[i].span_len = s.span_len
        arr[i].n_effects = len(s.effects)
        arr[i].effects = C.cast(fx, C.POINTER(_Effect))
        arr[i].mix_start = s.mix_start if i < len(streams) else MIX_START_CONSUMED
        arr[i].pcm = _fptr(pcm)
    return arr, keep


def _grow(call, guess: int) -> np.ndarray:
    cap = max(16, int(guess))
    while True:
        out = np.empty(cap, dtype=np.float32)
        n = C.c_uint64(0)
        rc = call(out, cap, n)
        if rc == 0:
            return out[: n.value].copy()
        if rc != 8:
            raise RuntimeError(f"oracle call failed rc={rc}")
        cap = int(n.value) + 16


def sample_rate_converter(x: np.ndarray, from_rate: int, to_rate: int, channels: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    L = lib()
    guess = x.size * (to_rate / from_rate) + 4 * channels + 16
    return _grow(lambda out, cap, n: L.ro_sample_rate_converter(
        _fptr(x), C.c_uint64(x.size), C.c_uint32(from_rate), C.c_uint32(to_rate), C.c_uint16(channels),
        _fptr(out), C.c_uint64(cap), C.byref(n)), guess)


def channel_count_converter(x: np.ndarray, from_ch: int, to_ch: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    L = lib()
    guess = (x.size // from_ch + 2) * to_ch + 16
    return _grow(lambda out, cap, n: L.ro_channel_count_converter(
        _fptr(x), C.c_uint64(x.size), C.c_uint16(from_ch), C.c_uint16(to_ch), _fptr(out), C.c_uint64(cap),
        C.byref(n)), guess)


SINE, TRIANGLE, SQUARE, SAWTOOTH = 0, 1, 2, 3


def signal(fn: int, rate: int, freq: float, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lib().ro_signal(C.c_int(fn), C.c_uint32(rate), C.c_float(freq), C.c_uint64(n), _fptr(out))
    return out


def sine_wave(freq: float, n: int) -> np.ndarray:
    """SineWave::new(freq).take(n) — 48 kHz mono."""
    return signal(SINE, 48000, freq, n)


def chain(s: Stream):
    """Drain the stream's own adapter chain.  Returns (samples, channels, sample_rate)."""
    arr, keep = _pack([s])
    L = lib()
    ch, rate = C.c_uint16(0), C.c_uint32(0)
    out = _grow(lambda out, cap, n: L.ro_chain(C.byref(arr[0]), _fptr(out), C.c_uint64(cap), C.byref(n),
                                               C.byref(ch), C.byref(rate)), s.pcm.size * 2 + 65536)
    return out, ch.value, rate.value


def chain_uniform(s: Stream, mixer_channels: int, mixer_rate: int) -> np.ndarray:
    """What MixerSource pulls from this source (chain wrapped in UniformSourceIterator)."""
    arr, keep = _pack([s])
    L = lib()
    guess = s.pcm.size * 4 * max(1.0, mixer_rate / max(1, s.sample_rate)) * mixer_channels + (1 << 16)
    return _grow(lambda out, cap, n: L.ro_chain_uniform(C.byref(arr[0]), C.c_uint16(mixer_channels),
                                                        C.c_uint32(mixer_rate), _fptr(out), C.c_uint64(cap),
                                                        C.byref(n)), guess)


def mixer(streams: Sequence[Stream], channels: int, rate: int, return_gaps: bool = False):
    arr, keep = _pack(streams)
    L = lib()
    gaps = C.c_uint64(0)
    guess = max([1 << 16] + [int(s.pcm.size * 4 * max(1.0, rate / s.sample_rate) * channels) + s.mix_start
                             for s in streams])
    out = _grow(lambda out, cap, n: L.ro_mixer(arr, C.c_uint64(len(streams)), C.c_uint16(channels),
                                               C.c_uint32(rate), _fptr(out), C.c_uint64(cap), C.byref(n),
                                               C.byref(gaps)), guess)
    return (out, gaps.value) if return_gaps else out


def mixer_mt(streams: Sequence[Stream], channels: int, rate: int, n_threads: int, out_cap: int,
             static_dispatch: bool = False):
    """CPU baseline: pull-model mixer drain sharded over host threads.  Returns (mix, seconds).
    static_dispatch=True builds the bench chain shape from the monomorphised templates (what rustc emits)."""
    arr, keep = _pack(streams)
    L = lib()
    out = np.empty(out_cap, dtype=np.float32)
    n = C.c_uint64(0)
    secs = C.c_double(0.0)
    fn = L.ro_mixer_mt_static if static_dispatch else L.ro_mixer_mt
    rc = fn(arr, C.c_uint64(len(streams)), C.c_uint16(channels), C.c_uint32(rate), C.c_int(n_threads),
                       out.ctypes.data_as(C.c_void_p), C.c_uint64(out_cap), C.byref(n), C.byref(secs))
    if rc not in (0, 8):
        raise RuntimeError(f"ro_mixer_mt rc={rc}")
    return out[: min(n.value, out_cap)], secs.value


_NP_FMT = {0: np.float32, 1: np.int16, 2: np.uint16, 3: np.int8, 4: np.uint8, 5: np.int32, 6: np.int32}


def convert(x: np.ndarray, in_fmt: int, out_fmt: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=_NP_FMT[in_fmt])
    out = np.empty(x.size, dtype=_NP_FMT[out_fmt])
    rc = lib().ro_convert(x.ctypes.data_as(C.c_void_p), C.c_int(in_fmt), out.ctypes.data_as(C.c_void_p),
                          C.c_int(out_fmt), C.c_uint64(x.size))
    if rc:
        raise RuntimeError("ro_convert failed")
    return out


def lerp(a, b, num, den):
    return lib().ro_lerp(a, b, num, den)


def db_to_linear(d):
    return lib().ro_db_to_linear(d)


def linear_to_db(x):
    return lib().ro_linear_to_db(x)


def blt_coeffs(high: bool, freq: int, q: float, fs: int) -> np.ndarray:
    out = np.empty(5, dtype=np.float32)
    lib().ro_blt_coeffs(C.c_int(int(high)), C.c_uint32(freq), C.c_float(q), C.c_uint32(fs), _fptr(out))
    return out


def spatial_volumes(emitter, left, right) -> np.ndarray:
    e = np.asarray(emitter, dtype=np.float32)
    l = np.asarray(left, dtype=np.float32)
    r = np.asarray(right, dtype=np.float32)
    out = np.empty(2, dtype=np.float32)
    lib().ro_spatial_volumes(_fptr(e), _fptr(l), _fptr(r), _fptr(out))
    return out
