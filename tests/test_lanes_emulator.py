"""The lane-per-stream fused kernel (rodio_b200/csrc/rb_lanes_core.h) on the CPU: tests/emu/lanes_emu.cpp compiles
the kernel's own source against the SIMT emulator of rb_simt.h (32 host threads per warp; cp.async copies land only
at their wait_group and poison their destination when issued) and this file holds it bit for bit against the
oracle: per-stream outputs from the literal pull iterators, summed with the kernel's documented reduction tree.
CPU only -- the GPU counterpart is tests/test_parity_gpu.py::test_lanes_*."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, assert_close_peak, lanes_expected_mix as expected_mix, noise, to_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "lanes_emu.cpp")
LIB = os.path.join(HERE, "emu", "liblanes_emu.so")
DEPS = [SRC] + [os.path.join(ROOT, "rodio_b200", "csrc", f) for f in ("rb_lanes_core.h", "rb_lanes_plan.h", "rb_simt.h")]


@pytest.fixture(scope="module")
def emu(built):
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               "-fPIC", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.rb_lanes_emulate.restype = C.c_int
    return lib


def counters(emu, reset=True):
    out = (C.c_uint64 * 4)()
    emu.rb_lanes_emu_counters(out, int(reset))
    return {"fast": out[0], "slow": out[1], "refills": out[2]}


def run_emu(emu, pcms, outs_len, starts, coefs, posts, from_, to, mix_len, hasb, ff2, npost):
    n = len(pcms)
    pcms = [np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
    ptrs = (C.POINTER(C.c_float) * n)(*[p.ctypes.data_as(C.POINTER(C.c_float)) for p in pcms])
    u64 = lambda v: (C.c_uint64 * n)(*[int(x) for x in v])
    co = np.ascontiguousarray(coefs, dtype=np.float32).reshape(-1)
    po = np.ascontiguousarray(posts, dtype=np.float32)
    out = np.full(mix_len, np.nan, dtype=np.float32)
    used, unsafe = C.c_int(0), C.c_uint32(0)
    rc = emu.rb_lanes_emulate(ptrs, u64([p.size for p in pcms]), u64(outs_len), u64(starts),
                              co.ctypes.data_as(C.POINTER(C.c_float)), po.ctypes.data_as(C.POINTER(C.c_float)),
                              C.c_uint32(n), C.c_uint32(from_), C.c_uint32(to), C.c_uint64(mix_len), int(hasb), int(ff2),
                              int(npost), out.ctypes.data_as(C.POINTER(C.c_float)), None, C.byref(used), C.byref(unsafe))
    assert rc == 0
    return out, bool(used.value), unsafe.value


def make_case(pcms, in_rate, mix_rate, starts, lp=None, hp=None, q=0.5, gain=None):
    """Sources as a rodio user writes them + everything the emulator needs, the expectation from the oracle."""
    srcs, per_stream = [], []
    for p in pcms:
        s = rb.UniformSourceIterator(rb.TestSource(p, 1, in_rate), 1, mix_rate)
        if lp is not None:
            s = s.low_pass_with_q(lp, q)
        if hp is not None:
            s = s.high_pass_with_q(hp, q)
        if gain is not None:
            s = s.amplify(gain)
        srcs.append(s)
        per_stream.append(oracle.chain_uniform(to_oracle(s), 1, mix_rate))
    g = math.gcd(in_rate, mix_rate)
    hasb = lp is not None or hp is not None
    co = oracle.blt_coeffs(hp is not None, lp if lp is not None else (hp or 1), q, mix_rate) if hasb else np.zeros(5, np.float32)
    coefs = np.tile(co, (len(pcms), 1))
    mix_len = max([s + y.size for s, y in zip(starts, per_stream)] + [0])
    return dict(per_stream=per_stream, outs_len=[y.size for y in per_stream], coefs=coefs,
                posts=np.full(len(pcms), gain if gain is not None else 1.0, np.float32), from_=in_rate // g, to=mix_rate // g,
                mix_len=mix_len, hasb=hasb, npost=gain is not None, srcs=srcs)


def check(emu, pcms, in_rate, mix_rate, starts, ff2=True, expect_ff2=None, **kw):
    c = make_case(pcms, in_rate, mix_rate, starts, **kw)
    got, used_ff2, unsafe = run_emu(emu, pcms, c["outs_len"], starts, c["coefs"], c["posts"], c["from_"], c["to"], c["mix_len"],
                                    c["hasb"], ff2, c["npost"])
    if expect_ff2 is not None:
        assert used_ff2 == expect_ff2
    want = expected_mix(c["per_stream"], starts, c["mix_len"])
    assert_bit_exact(got, want, "emulated kernel vs oracle streams summed with the kernel's tree")
    # and the north-star tolerance against the reference's sequential mixer
    ref = oracle.mixer([to_oracle(s, mix_start=st) for s, st in zip(c["srcs"], starts)], 1, mix_rate)
    assert_close_peak(got, ref, 1e-5, "emulated kernel vs the reference's sequential mixer sum")
    return unsafe


def test_single_stream_bit_exact_with_reference_order(emu):
    """One stream: the tree adds zeros only, so the kernel output IS the reference stream (uniform -> low_pass -> amplify)."""
    pcm = noise(5000, 1)
    c = make_case([pcm], 44100, 48000, [0], lp=200, gain=1.2)
    got, used_ff2, _ = run_emu(emu, [pcm], c["outs_len"], [0], c["coefs"], c["posts"], 147, 160, c["mix_len"], True, True, True)
    assert used_ff2
    assert_bit_exact(got, c["per_stream"][0], "single stream")


@pytest.mark.parametrize("ff2", [True, False])
def test_cfg3_shape_bit_exact(emu, ff2):
    """BASELINE cfg3 shape in small: 70 streams (three warps, the last one partly filled), 44.1 -> 48 kHz, low_pass(200),
    amplify(1.2); both feed-forward variants."""
    pcms = [noise(3000 + 7 * i, 100 + i) for i in range(70)]
    counters(emu)
    check(emu, pcms, 44100, 48000, [0] * 70, ff2=ff2, expect_ff2=ff2, lp=200, gain=1.2)
    c = counters(emu)
    # ~3265..4020 output samples per stream: nearly all tiles of the three warps are fast ones, the ring turns over
    assert c["fast"] > 1100 and c["slow"] < 0.4 * c["fast"] and c["refills"] > 500, c


def test_ragged_starts_and_lengths(emu):
    """Late joiners, short streams inside long ones, a one-frame and an empty stream: run management."""
    rng = np.random.default_rng(5)
    lens = [4000, 37, 1, 0, 2, 2500, 4000, 999, 16, 17] + [int(v) for v in rng.integers(3, 3000, 30)]
    starts = [0, 100, 5, 9, 3000, 1234, 8, 16, 4001, 7] + [int(v) for v in rng.integers(0, 2500, 30)]
    pcms = [noise(n, 300 + i) for i, n in enumerate(lens)]
    check(emu, pcms, 44100, 48000, starts, lp=1000, gain=0.7)


def test_high_pass_and_no_gain(emu):
    pcms = [noise(2000 + i, 500 + i) for i in range(33)]
    check(emu, pcms, 44100, 48000, [0] * 33, hp=300, expect_ff2=True)


def test_no_filter(emu):
    """resample -> amplify -> mix and resample -> mix."""
    pcms = [noise(1500 + 3 * i, 700 + i) for i in range(40)]
    check(emu, pcms, 44100, 48000, [i % 5 for i in range(40)], gain=1.2)
    check(emu, pcms[:9], 44100, 48000, [0] * 9)


@pytest.mark.parametrize("rates", [(8000, 48000), (22050, 48000), (32000, 44100), (47999, 48000), (11025, 96000), (1, 3)])
def test_other_ratios(emu, rates):
    pcms = [noise(700 + 11 * i, 900 + i) for i in range(35)]
    check(emu, pcms, rates[0], rates[1], [3 * (i % 4) for i in range(35)], lp=400, gain=1.1)


def test_unsafe_streams_take_the_ieee_path(emu):
    """Denormal / huge / non-finite samples: the stream is classified out of the reciprocal fast path; its warp divides
    with IEEE division and stays bit-exact, the other warps keep the fast path."""
    pcms = [noise(1200, 40 + i) for i in range(40)]
    pcms[3][100:110] = np.float32(1e-41)          # denormals
    pcms[3][500] = np.float32(3e-30)              # tiny normal next to normal samples
    pcms[35][7] = np.float32(1e25)
    counters(emu)
    unsafe = check(emu, pcms, 44100, 48000, [0] * 40, lp=800, gain=1.2)
    c = counters(emu)
    assert unsafe == 2
    assert c["fast"] == 0 and c["slow"] > 300, c   # both warps hold an unsafe stream here: slow tiles only


def test_silence_and_negative_zero(emu):
    """Exact zeros (either sign) go through the reciprocal step; the mixer output carries +0 like the reference's."""
    pcms = [np.zeros(600, np.float32), -np.zeros(600, np.float32), noise(600, 1)]
    pcms[2][200:300] = 0.0
    check(emu, pcms, 44100, 48000, [0, 0, 0], gain=-1.0)
    check(emu, pcms[:2], 44100, 48000, [0, 0], lp=300, gain=-1.0)
