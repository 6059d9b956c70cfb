"""Test helpers: turn a rodio_b200 Source into the oracle's neutral Stream, compare results."""
from __future__ import annotations

import numpy as np

import oracle
import rodio_b200 as rb


def to_oracle(src: rb.Source, mix_start: int = 0) -> oracle.Stream:
    pcm = src.pcm
    if pcm.dtype != np.float32:
        fmt = {np.dtype(np.int16): 1, np.dtype(np.uint16): 2, np.dtype(np.int8): 3, np.dtype(np.uint8): 4,
               np.dtype(np.int32): 5}[pcm.dtype]
        if getattr(src, "fmt_override", None) is not None:
            fmt = src.fmt_override
        pcm = oracle.convert(pcm, fmt, 0)
    effects = [oracle.Fx(e.kind, e.u32, e.f32, e.ns, to_oracle(e.other)) if getattr(e, "other", None) is not None else e
               for e in src.effects]
    return oracle.Stream(pcm=pcm, channels=src.base_channels, sample_rate=src.base_rate, effects=effects,
                         span_len=src.span_len, mix_start=mix_start)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_exact(got: np.ndarray, want: np.ndarray, what: str = ""):
    assert got.shape == want.shape, f"{what}: length {got.shape} != {want.shape}"
    if got.size == 0:
        return
    g, w = bits(got), bits(want)
    bad = np.nonzero(g != w)[0]
    assert bad.size == 0, (f"{what}: {bad.size}/{got.size} samples differ, first at {bad[0]}: "
                           f"got {got[bad[0]]!r} want {want[bad[0]]!r}")


def assert_close_peak(got: np.ndarray, want: np.ndarray, tol: float = 1e-5, what: str = ""):
    """Parity metric of the north star: max|got - want| <= tol * max|want| (peak-normalised)."""
    assert got.shape == want.shape, f"{what}: length {got.shape} != {want.shape}"
    if got.size == 0:
        return
    peak = float(np.max(np.abs(want)))
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    assert err <= tol * max(peak, 1e-30), f"{what}: max err {err:.3e} > {tol:g} * peak {peak:.3e}"


def noise(n: int, seed: int, amp: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1.0, 1.0, n) * amp).astype(np.float32)


def lanes_tree_sum(rows) -> np.ndarray:
    """The reduction tree of k_fused_lanes over the 32 lanes of a warp (rb_lanes_core.h reduce_tile), in float32."""
    y = np.asarray(rows, dtype=np.float32)
    a = y[:16] + y[16:]
    b = a[:8] + a[8:]
    c = b[:4] + b[4:]
    return ((c[0] + c[1]) + (c[2] + c[3])) + np.float32(0.0)


def lanes_expected_mix(per_stream, starts, mix_len: int) -> np.ndarray:
    """Mixer output of the lane-per-stream kernel given every stream's exact samples: groups of 32 streams (insertion
    order) summed with the tree, the groups added in order from +0.0."""
    acc = np.zeros(mix_len, dtype=np.float32)
    for g in range(0, len(per_stream), 32):
        lanes_ = np.zeros((32, mix_len), dtype=np.float32)
        for l, (y, s) in enumerate(zip(per_stream[g:g + 32], starts[g:g + 32])):
            lanes_[l, s:s + y.size] = y
        acc = acc + lanes_tree_sum(lanes_)
    return acc


def duo_expected_mix(per_stream, starts, mix_len: int) -> np.ndarray:
    """Mixer output of the lane-pair kernel (rb_duo_core.h) given every stream's exact samples: groups of 64 streams
    (insertion order); lane l = stream 2l + stream 2l+1 (an absent or silent one counts +0.0), the 32 lane values summed with
    the tree of k_fused_lanes, the groups added in order from +0.0."""
    acc = np.zeros(mix_len, dtype=np.float32)
    for g in range(0, len(per_stream), 64):
        halves = np.zeros((64, mix_len), dtype=np.float32)
        for l, (y, s) in enumerate(zip(per_stream[g:g + 64], starts[g:g + 64])):
            halves[l, s:s + y.size] = y
        acc = acc + lanes_tree_sum(halves[0::2] + halves[1::2])
    return acc


def fused_expected_mix(family: int, per_stream, starts, mix_len: int) -> np.ndarray:
    """The documented order of kernel family 2 (k_fused_lanes) or 3 (k_fused_duo)."""
    return (duo_expected_mix if family == 3 else lanes_expected_mix)(per_stream, starts, mix_len)
