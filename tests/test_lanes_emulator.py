"""The lane-per-stream fused kernel (rodio_b200/csrc/rb_lanes_core.h) on the CPU: tests/emu/lanes_emu.cpp compiles
the kernel's own source against the SIMT emulator of rb_simt.h (32 fibers per warp; cp.async copies land only
at their wait_group and poison their destination when issued) and this file holds it bit for bit against the
oracle: per-stream outputs from the literal pull iterators, summed with the kernel's documented reduction tree.
CPU only -- the GPU counterpart is tests/test_parity_gpu.py::test_lanes_*."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import rodio_b200 as rb
from helpers import assert_bit_exact, assert_close_peak, lanes_expected_mix as expected_mix, noise, to_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "emu", "liblanes_emu.so")      # built by tests/emu/build_emu.py


@pytest.fixture(scope="module")
def emu(built):
    sys.path[:0] = [os.path.join(HERE, "emu")]
    import build_emu
    lib = C.CDLL(build_emu.lanes_lib())
    lib.rb_lanes_emulate.restype = C.c_int
    return lib


def counters(emu, reset=True):
    out = (C.c_uint64 * 4)()
    emu.rb_lanes_emu_counters(out, int(reset))
    return {"fast": out[0], "slow": out[1], "refills": out[2], "divided": out[3]}


def _u32(v, n):
    v = list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] * n
    return (C.c_uint32 * n)(*[int(x) for x in v]), v


def expected_mix_classes(per_stream, starts, mix_len, froms, tos):
    """Streams with several rate pairs are served class by class (rb_lanes_plan.h classes_by_ratio): per class, groups of
    32 in stream order summed with the tree; all groups added in class order from +0.0."""
    keys, classes = [], []
    for i, k in enumerate(zip(froms, tos)):
        if k not in keys:
            keys.append(k)
            classes.append([])
        classes[keys.index(k)].append(i)
    acc = np.zeros(mix_len, dtype=np.float32)
    for cls in classes:
        for g in range(0, len(cls), 32):
            rows = np.zeros((32, mix_len), dtype=np.float32)
            for l, i in enumerate(cls[g:g + 32]):
                rows[l, starts[i]:starts[i] + per_stream[i].size] = per_stream[i]
            a = rows[:16] + rows[16:]
            b = a[:8] + a[8:]
            c = b[:4] + b[4:]
            acc = acc + (((c[0] + c[1]) + (c[2] + c[3])) + np.float32(0.0))
    return acc


def run_emu(emu, pcms, outs_len, starts, coefs, posts, from_, to, mix_len, hasb, ff2, npost, channels=1, ch_in=None, pres=None, front=False,
            mids=None):
    """Frames everywhere (outs_len, starts, mix_len); the result holds frames * channels floats, pcms[r] frames * ch_in[r]."""
    ch_in = [channels] * len(pcms) if ch_in is None else list(ch_in)
    n = len(pcms)
    pcms = [np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
    ptrs = (C.POINTER(C.c_float) * n)(*[p.ctypes.data_as(C.POINTER(C.c_float)) for p in pcms])
    u64 = lambda v: (C.c_uint64 * n)(*[int(x) for x in v])
    co = np.ascontiguousarray(coefs, dtype=np.float32).reshape(-1)
    po = np.ascontiguousarray(posts, dtype=np.float32)
    out = np.full(mix_len * channels, np.nan, dtype=np.float32)
    used, unsafe = C.c_int(0), C.c_uint32(0)
    pr = None if pres is None else np.ascontiguousarray(pres, dtype=np.float32)
    mi = None if mids is None else np.ascontiguousarray(mids, dtype=np.float32)
    rc = emu.rb_lanes_emulate(ptrs, u64([p.size // c for p, c in zip(pcms, ch_in)]), u64(outs_len), u64(starts),
                              co.ctypes.data_as(C.POINTER(C.c_float)), po.ctypes.data_as(C.POINTER(C.c_float)),
                              C.c_uint32(n), C.c_uint32(channels), _u32(ch_in, n)[0], _u32(from_, n)[0], _u32(to, n)[0], C.c_uint64(mix_len), int(hasb), int(ff2),
                              int(npost), out.ctypes.data_as(C.POINTER(C.c_float)), None, C.byref(used), C.byref(unsafe),
                              None if pr is None else pr.ctypes.data_as(C.POINTER(C.c_float)), int(front),
                              None if mi is None else mi.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return out, bool(used.value), unsafe.value


def make_case(pcms, in_rate, mix_rate, starts, lp=None, hp=None, q=0.5, gain=None, channels=1, ch_in=None, pre=None, front=False, mid=None):
    """Sources as a rodio user writes them + everything the emulator needs, the expectation from the oracle.
    `starts` and the lengths in the result are frames."""
    srcs, per_stream = [], []
    in_rates = list(in_rate) if isinstance(in_rate, (list, tuple)) else [in_rate] * len(pcms)
    ch_in = [channels] * len(pcms) if ch_in is None else list(ch_in)
    pres = None if pre is None else (list(pre) if isinstance(pre, (list, tuple, np.ndarray)) else [pre] * len(pcms))
    mids = None if mid is None else (list(mid) if isinstance(mid, (list, tuple, np.ndarray)) else [mid] * len(pcms))
    for k, (p, rate, ci) in enumerate(zip(pcms, in_rates, ch_in)):
        s = rb.TestSource(p, ci, rate)
        if pres is not None:
            s = s.amplify(float(pres[k]))       # source.amplify(v) handed to the mixer: the gain sits in front of the conversion
        if front:                               # source.low_pass(f) [.amplify(v)] handed to the mixer: the filter runs at the source's rate
            s = s.low_pass_with_q(lp, q) if lp is not None else s.high_pass_with_q(hp, q)
            if mids is not None:
                s = s.amplify(float(mids[k]))
        s = rb.UniformSourceIterator(s, channels, mix_rate)
        if front:
            pass
        elif lp is not None:
            s = s.low_pass_with_q(lp, q)
        if hp is not None and not front:
            s = s.high_pass_with_q(hp, q)
        if gain is not None:
            s = s.amplify(gain)
        srcs.append(s)
        per_stream.append(oracle.chain_uniform(to_oracle(s), channels, mix_rate))
    froms = [r // math.gcd(r, mix_rate) for r in in_rates]
    tos = [mix_rate // math.gcd(r, mix_rate) for r in in_rates]
    hasb = lp is not None or hp is not None
    # the filter's coefficients follow the rate of the stream it sees (blt.rs: to_applier(input.sample_rate()))
    coef_at = lambda rate: oracle.blt_coeffs(hp is not None, lp if lp is not None else (hp or 1), q, rate) if hasb else np.zeros(5, np.float32)
    coefs = np.stack([coef_at(r if front else mix_rate) for r in in_rates])
    mix_len = max([s + y.size // channels for s, y in zip(starts, per_stream)] + [0])
    return dict(per_stream=per_stream, outs_len=[y.size // channels for y in per_stream], coefs=coefs, channels=channels, ch_in=ch_in,
                posts=np.full(len(pcms), gain if gain is not None else 1.0, np.float32), from_=froms, to=tos,
                mix_len=mix_len, hasb=hasb, npost=gain is not None, srcs=srcs, pres=pres, front=front, mids=mids)


def check(emu, pcms, in_rate, mix_rate, starts, ff2=True, expect_ff2=None, **kw):
    c = make_case(pcms, in_rate, mix_rate, starts, **kw)
    ch = c["channels"]
    got, used_ff2, unsafe = run_emu(emu, pcms, c["outs_len"], starts, c["coefs"], c["posts"], c["from_"], c["to"], c["mix_len"],
                                    c["hasb"], ff2, c["npost"], channels=ch, ch_in=c["ch_in"], pres=c["pres"], front=c["front"], mids=c["mids"])
    if expect_ff2 is not None:
        assert used_ff2 == expect_ff2
    want = expected_mix_classes(c["per_stream"], [st * ch for st in starts], c["mix_len"] * ch, c["from_"], list(zip(c["to"], c["ch_in"])))
    assert_bit_exact(got, want, "emulated kernel vs oracle streams summed with the kernel's tree")
    # and the north-star tolerance against the reference's sequential mixer
    ref = oracle.mixer([to_oracle(s, mix_start=st * ch) for s, st in zip(c["srcs"], starts)], ch, mix_rate)
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


# ------------------------------------------------------------------------------------------------- streaming sessions
def run_session(emu, pcms, starts, coefs, posts, from_, to, hasb, npost, ops, out_cap, channels=1, ch_in=None):
    ch_in = [channels] * len(pcms) if ch_in is None else list(ch_in)
    n = len(pcms)
    pcms = [np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
    ptrs = (C.POINTER(C.c_float) * n)(*[p.ctypes.data_as(C.POINTER(C.c_float)) for p in pcms])
    u64 = lambda v: (C.c_uint64 * len(v))(*[int(x) for x in v])
    co = np.ascontiguousarray(coefs, dtype=np.float32).reshape(-1)
    po = np.ascontiguousarray(posts, dtype=np.float32)
    flat = [int(x) for op in ops for x in op]
    out = np.full(out_cap, np.nan, dtype=np.float32)
    renders = C.c_uint64(0)
    pushed = (C.c_uint64 * n)()
    joined = (C.c_uint64 * n)()
    emu.rb_session_emulate.restype = C.c_longlong
    w = emu.rb_session_emulate(ptrs, u64([p.size // c for p, c in zip(pcms, ch_in)]), u64(starts), co.ctypes.data_as(C.POINTER(C.c_float)),
                               po.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(n), C.c_uint32(channels), _u32(ch_in, n)[0], _u32(from_, n)[0], _u32(to, n)[0],
                               int(hasb), int(npost), u64(flat), C.c_uint64(len(ops)), out.ctypes.data_as(C.POINTER(C.c_float)),
                               C.c_uint64(out_cap), C.byref(renders), pushed, joined)
    assert w >= 0
    run_session.joined_at = [int(v) for v in joined]
    return out[:w * channels], renders.value, [int(v) for v in pushed]


def session_case(emu, pcms, starts, ops, in_rate=44100, mix_rate=48000, lp=None, gain=None, channels=1, ch_in=None):
    """Any split of the streams into pushed blocks and of the mixer output into rendered blocks gives the bytes of the
    whole-stream render (tests/test_block_state_spec.py is the numpy form of the same contract)."""
    c = make_case(pcms, in_rate, mix_rate, starts, lp=lp, gain=gain, channels=channels, ch_in=ch_in)
    got, renders, pushed = run_session(emu, pcms, starts, c["coefs"], c["posts"], c["from_"], c["to"], c["hasb"], c["npost"], ops,
                                       (c["mix_len"] + 64) * channels, channels=channels, ch_in=c["ch_in"])
    assert pushed == [p.size // ci for p, ci in zip(pcms, c["ch_in"])]
    want = expected_mix_classes(c["per_stream"], [st * channels for st in starts], c["mix_len"] * channels, c["from_"],
                                list(zip(c["to"], c["ch_in"])))
    assert_bit_exact(got, want, "session blocks vs whole-stream render")
    return renders


def test_session_any_split_is_the_whole(emu):
    rng = np.random.default_rng(11)
    pcms = [noise(int(n), 60 + i) for i, n in enumerate(rng.integers(400, 2600, 37))]
    starts = [0] * 37
    ops = []
    left = [p.size for p in pcms]
    while any(left):                      # random interleaving of pushes (1..700 frames) and renders (1..900 frames)
        for r in rng.permutation(37):
            n = min(left[r], int(rng.integers(1, 700)))
            if n and rng.random() < 0.8:
                ops.append((0, r, n))
                left[r] -= n
        ops.append((1, 0, int(rng.integers(1, 900))))
    renders = session_case(emu, pcms, starts, ops, lp=200, gain=1.2)
    assert renders > 5


def test_session_tiny_blocks_and_late_joiners(emu):
    """10 ms style blocks (and blocks of 1..7 frames, shorter than the kernel's tile), streams that join later."""
    pcms = [noise(900 + 50 * i, 80 + i) for i in range(5)]
    starts = [0, 0, 480, 481, 1000]
    ops = []
    for step in range(200):
        for r in range(5):
            ops.append((0, r, 97 + r))
        ops.append((1, 0, [441, 1, 7, 3, 480, 5][step % 6]))
    session_case(emu, pcms, starts, ops, lp=1000, gain=0.9)


def test_session_without_filter_and_starved_render(emu):
    """A render with nothing new pushed produces nothing and leaves the state alone."""
    pcms = [noise(1500, 90 + i) for i in range(3)]
    ops = [(0, 0, 500), (0, 1, 500), (0, 2, 10), (1, 0, 10000), (1, 0, 10000), (1, 0, 10000),
           (0, 2, 1490), (1, 0, 100), (0, 0, 1000), (0, 1, 1000)]
    session_case(emu, pcms, [0, 0, 0], ops, gain=1.2)


def test_session_early_end_of_stream(emu):
    """A stream that is cut short (eof before all of its PCM was pushed) renders like the shorter whole stream."""
    pcms = [noise(2000, 95 + i) for i in range(4)]
    ops = [(0, r, 700) for r in range(4)] + [(1, 0, 300), (2, 1, 0), (0, 0, 1300), (0, 2, 1300), (0, 3, 1300), (1, 0, 5000)]
    c_full = make_case(pcms, 44100, 48000, [0] * 4, lp=300, gain=1.1)
    got, _, pushed = run_session(emu, pcms, [0] * 4, c_full["coefs"], c_full["posts"], 147, 160, True, True, ops, 4000)
    assert pushed == [2000, 700, 2000, 2000]
    cut = [pcms[0], pcms[1][:700], pcms[2], pcms[3]]
    c = make_case(cut, 44100, 48000, [0] * 4, lp=300, gain=1.1)
    assert_bit_exact(got, expected_mix(c["per_stream"], [0] * 4, c["mix_len"]), "cut stream")


def test_session_plan_fuzz(emu):
    """rb_session_plan.h alone (no kernel): 3000 random sessions on random rate pairs -- every output rendered exactly
    once and where the timeline says, taps always inside the FIFO, FIFO front aligned, totals equal the closed form."""
    emu.rb_session_plan_fuzz.restype = C.c_int
    for seed in range(6):
        line = emu.rb_session_plan_fuzz(C.c_uint64(seed), C.c_uint32(500))
        assert line == 0, f"invariant at tests/emu/lanes_emu.cpp:{line} violated (seed {seed})"


# ------------------------------------------------------------------------------------------------- stereo (C = 2)
def test_stereo_cfg3_shape(emu):
    """Interleaved stereo sources into a stereo mixer: one lane carries both channels (shared index state, two filters)."""
    pcms = [noise(2 * (2000 + 9 * i), 400 + i, 0.9) for i in range(40)]
    counters(emu)
    check(emu, pcms, 44100, 48000, [0] * 40, lp=200, gain=1.2, channels=2, expect_ff2=True)
    c = counters(emu)
    assert c["fast"] > 800 and c["refills"] > 200, c
    check(emu, pcms[:33], 44100, 48000, [0] * 33, ff2=False, hp=300, channels=2, expect_ff2=False)


def test_stereo_ragged_and_no_filter(emu):
    rng = np.random.default_rng(6)
    lens = [1500, 3, 1, 0, 2, 900] + [int(v) for v in rng.integers(3, 1500, 30)]
    starts = [0, 10, 5, 9, 700, 123] + [int(v) for v in rng.integers(0, 900, 30)]
    pcms = [noise(2 * n, 600 + i) for i, n in enumerate(lens)]
    check(emu, pcms, 32000, 44100, starts, lp=900, gain=0.6, channels=2)
    check(emu, pcms[:12], 22050, 48000, starts[:12], gain=1.3, channels=2)


def test_stereo_session_any_split(emu):
    rng = np.random.default_rng(12)
    pcms = [noise(2 * int(n), 700 + i) for i, n in enumerate(rng.integers(300, 1800, 9))]
    starts = [0, 0, 0, 100, 0, 37, 0, 0, 512]
    ops, left = [], [p.size // 2 for p in pcms]
    while any(left):
        for r in rng.permutation(9):
            n = min(left[r], int(rng.integers(1, 500)))
            if n and rng.random() < 0.8:
                ops.append((0, r, n))
                left[r] -= n
        ops.append((1, 0, int(rng.integers(1, 600))))
    session_case(emu, pcms, starts, ops, lp=300, gain=1.1, channels=2)


def test_session_gain_changes_between_blocks(emu):
    """rb_session_set_amplify: the gain of a live source changes from one rendered block to the next (Player::set_volume
    through its 5 ms periodic access, src/player.rs:138-166); inside a block it is constant."""
    pcms = [noise(1500, 800 + i) for i in range(3)]
    c = make_case(pcms, 44100, 48000, [0] * 3, lp=500, gain=1.0)        # oracle streams with gain 1.0: y * 1.0 = y
    ops = [(0, r, 1500) for r in range(3)]
    gains, block = [], 240                                               # 5 ms at 48 kHz
    n_blocks = -(-c["mix_len"] // block)
    rng = np.random.default_rng(3)
    for k in range(n_blocks):
        g = [np.float32(v) for v in rng.uniform(0.0, 1.5, 3)]
        gains.append(g)
        ops += [(3, r, int(np.float32(g[r]).view(np.uint32))) for r in range(3)] + [(1, 0, block)]
    got, renders, _ = run_session(emu, pcms, [0] * 3, c["coefs"], c["posts"], 147, 160, True, True, ops, c["mix_len"] + 64)
    assert renders == n_blocks
    scaled = []
    for r in range(3):
        y = c["per_stream"][r].copy()
        for k in range(n_blocks):
            y[k * block:(k + 1) * block] = y[k * block:(k + 1) * block] * gains[k][r]     # one f32 rounding, like Amplify::next
        scaled.append(y)
    assert_bit_exact(got, expected_mix(scaled, [0] * 3, c["mix_len"]), "per-block gains")


# ------------------------------------------------------------------------------------------------- same-rate sources
def test_same_rate_sources_pass_through(emu):
    """Sources already at the mixer's rate (SampleRateConverter hands them through, sample_rate.rs:131-136): raw taps,
    every bit kept -- denormals, huge values and signed zeros included, no input class needed."""
    pcms = [noise(900 + 13 * i, 30 + i) for i in range(35)]
    pcms[2][10:20] = np.float32(1e-41)
    pcms[5][7] = np.float32(1e30)
    pcms[6][:50] = -0.0
    counters(emu)
    unsafe = check(emu, pcms, 48000, 48000, [i % 3 for i in range(35)], lp=250, gain=1.2)
    c = counters(emu)
    assert c["fast"] > 150, c            # the streams outside the class still take the fast path: nothing is divided
    check(emu, pcms[:9], 44100, 44100, [0] * 9, gain=0.5)
    stereo = [noise(2 * (400 + i), 70 + i) for i in range(8)]
    check(emu, stereo, 48000, 48000, [0] * 8, hp=300, channels=2)


def test_same_rate_session(emu):
    pcms = [noise(1200 + 31 * i, 130 + i) for i in range(6)]
    ops = []
    for step in range(40):
        ops += [(0, r, 100 + 7 * r) for r in range(6)] + [(1, 0, 97)]
    session_case(emu, pcms, [0, 0, 5, 0, 300, 0], ops, in_rate=48000, mix_rate=48000, lp=400, gain=0.9)


# ------------------------------------------------------------------------------------------------- several rate pairs
def test_mixed_rates_in_one_mixer(emu):
    """44.1 kHz, 48 kHz (pass-through), 22.05 kHz and 32 kHz sources in one 48 kHz mixer: one launch per rate pair over its
    own rows, all partial rows added in order."""
    rates = [44100, 48000, 22050, 44100, 32000, 48000, 44100] * 6
    pcms = [noise(700 + 17 * i, 500 + i) for i in range(len(rates))]
    starts = [(5 * i) % 97 for i in range(len(rates))]
    check(emu, pcms, rates, 48000, starts, lp=600, gain=1.1)
    check(emu, pcms[:10], rates[:10], 48000, [0] * 10)


def test_mixed_rates_session(emu):
    rates = [44100, 48000, 22050, 48000, 44100]
    pcms = [noise(int(1.0 * r / 40) + 11 * i, 900 + i) for i, r in enumerate(rates)]      # about 25 ms each
    ops = []
    for step in range(60):
        ops += [(0, r, int(rates[r] / 1000) + r) for r in range(5)] + [(1, 0, 45 + step % 7)]
    session_case(emu, pcms, [0, 0, 0, 100, 3], ops, in_rate=rates, mix_rate=48000, lp=700, gain=0.8)


# ------------------------------------------------------------------------------------------------- mono sources, stereo mixer
def test_mono_and_stereo_sources_in_a_stereo_mixer(emu):
    """ChannelCountConverter repeats a mono source on both channels (src/conversions/channels.rs:57-85); the lane computes
    the frame once and emits it twice.  Mono and stereo sources, two rates, one stereo mixer."""
    ch_in = [1, 2, 1, 1, 2, 1, 2, 1] * 5
    rates = [44100, 44100, 48000, 22050, 48000, 44100, 44100, 48000] * 5
    pcms = [noise(ci * (600 + 11 * i), 1500 + i) for i, ci in enumerate(ch_in)]
    starts = [(7 * i) % 50 for i in range(len(ch_in))]
    check(emu, pcms, rates, 48000, starts, lp=800, gain=0.7, channels=2, ch_in=ch_in)
    check(emu, pcms[:6], rates[:6], 48000, [0] * 6, channels=2, ch_in=ch_in[:6])


def test_mono_and_stereo_session(emu):
    ch_in = [1, 2, 1, 2]
    rates = [44100, 48000, 48000, 44100]
    pcms = [noise(ci * (1100 + 13 * i), 1700 + i) for i, ci in enumerate(ch_in)]
    ops = []
    for step in range(50):
        ops += [(0, r, 40 + 3 * r) for r in range(4)] + [(1, 0, 41 + step % 5)]
    session_case(emu, pcms, [0, 0, 10, 0], ops, in_rate=rates, mix_rate=48000, lp=900, gain=1.05, channels=2, ch_in=ch_in)


def test_session_with_sources_above_the_mixer_rate(emu):
    """48 kHz and 96 kHz sources in a 44.1 kHz mixer: down-sampling classes run on the kernel's slow tiles (closed form, IEEE
    division) -- exact, block-split invariant like everything else, just not fast."""
    rates = [48000, 96000, 44100, 22050, 48000]
    pcms = [noise(int(0.02 * r) + 7 * i, 2100 + i) for i, r in enumerate(rates)]       # about 20 ms each
    ops = []
    for step in range(40):
        ops += [(0, r, int(rates[r] / 1500) + r) for r in range(5)] + [(1, 0, 30 + step % 9)]
    counters(emu)
    session_case(emu, pcms, [0, 0, 0, 50, 3], ops, in_rate=rates, mix_rate=44100, lp=500, gain=0.9)
    c = counters(emu)
    assert c["slow"] > 50, c


def test_session_sources_added_while_it_runs(emu):
    """Sources declared when the session is created but handed to the mixer later (Mixer::add during playback,
    src/mixer.rs:58-66,:175-183): they join at the frame rendered next -- the bytes of a session in which they had that
    mix_start from the beginning; a source that is never added contributes nothing; while nothing plays the timeline stands
    still."""
    HELD = (1 << 64) - 1
    pcms = [noise(1200 + 50 * i, 3100 + i) for i in range(5)]
    starts = [0, HELD, 0, HELD, HELD]
    ops = [(0, r, pcms[r].size) for r in range(5)]                 # everything is decoded already
    ops += [(1, 0, 300), (4, 1, 0), (1, 0, 211)]                  # source 1 joins at frame 300
    ops += [(1, 0, 5000)] * 5                                      # a render stops where a source ends: play the three out
    ops += [(1, 0, 100), (4, 3, 0), (1, 0, 5000)]                  # nothing plays: no frames; source 3 joins where the timeline stopped
    c = make_case(pcms, 44100, 48000, [0] * 5, lp=600, gain=0.9)
    got, renders, pushed = run_session(emu, pcms, starts, c["coefs"], c["posts"], 147, 160, True, True, ops, 20000)
    joined = run_session.joined_at
    assert joined[0] == 0 and joined[1] == 300 and joined[2] == 0 and joined[4] == HELD
    end_of_first_three = max(joined[i] + c["per_stream"][i].size for i in (0, 1, 2))
    assert joined[3] == end_of_first_three                        # the timeline did not move while nothing was playing
    keep = [0, 1, 2, 3]
    want = expected_mix_classes([c["per_stream"][i] if i in keep else np.zeros(0, np.float32) for i in range(5)],
                                [joined[i] if i in keep else 0 for i in range(5)], got.size, c["from_"], list(zip(c["to"], c["ch_in"])))
    assert_bit_exact(got, want, "sources added during playback")


def test_session_queue_of_sources(emu):
    """Sources queued one behind the other (Player::append -> queue, src/queue.rs:128-192): the next one starts on the frame
    after the current one has played out -- known as soon as the current one has all of its input --, different rates and
    a second, independent voice beside the queue."""
    HELD = (1 << 64) - 1
    rates = [44100, 48000, 22050, 44100]
    pcms = [noise(int(0.03 * r) + 17 * i, 3300 + i) for i, r in enumerate(rates)]
    starts = [0, HELD, HELD, 40]                       # 0 -> 1 -> 2 is the queue, 3 plays beside it
    ops = [(5, 1, 0), (5, 2, 1)]
    left = [p.size for p in pcms]
    step = 0
    while any(left):
        for r in range(4):
            n = min(left[r], 90 + 20 * r)
            if n:
                ops.append((0, r, n))
                left[r] -= n
        ops.append((1, 0, 77 + step % 13))
        step += 1
    c = make_case(pcms, rates, 48000, [0] * 4, lp=900, gain=0.8)
    got, renders, pushed = run_session(emu, pcms, starts, c["coefs"], c["posts"], c["from_"], c["to"], True, True, ops, 20000)
    joined = run_session.joined_at
    assert joined[0] == 0 and joined[3] == 40
    assert joined[1] == c["per_stream"][0].size and joined[2] == joined[1] + c["per_stream"][1].size
    want = expected_mix_classes(c["per_stream"], joined, got.size, c["from_"], list(zip(c["to"], c["ch_in"])))
    assert got.size == max(j + y.size for j, y in zip(joined, c["per_stream"]))
    assert_bit_exact(got, want, "queued sources")


def test_gain_in_front_of_the_conversion(emu):
    """`source.amplify(v)` handed to the mixer: the gain multiplies every input frame before the interpolation
    (src/source/amplify.rs:91-95 in front of src/source/uniform.rs).  Per-stream gains; 0.001 and 100 lie outside the range
    the unguarded tile accepts: their class runs the guarded twin (every quotient checked like with a filter in front), still on
    fast tiles."""
    n = 40
    pcms = [noise(1500 + 13 * i, 900 + i) for i in range(n)]
    pres = [[0.8, -0.5, 1.0, 0.3][i % 4] for i in range(n)]
    counters(emu)
    check(emu, pcms, 44100, 48000, [0] * n, lp=200, gain=0.7, pre=pres, expect_ff2=True)
    c = counters(emu)
    assert c["fast"] > 4 * c["slow"]
    pres[5], pres[33], pres[20] = 0.001, 100.0, 0.0      # ... and a muted source (zero taps are exact: fast tiles)
    check(emu, pcms, 44100, 48000, [7 * (i % 5) for i in range(n)], lp=200, gain=0.7, pre=pres)
    c = counters(emu)
    assert c["fast"] > 4 * c["slow"]   # a gain out of range does not cost the fast tiles: the class runs the guarded twin
    # tiny taps (2^-116) behind an absurdly small gain: quotients below the reciprocal's exact range, divided instead
    quiet = [noise(1500, 990 + i) * np.float32(1e-15) for i in range(4)]
    counters(emu)
    check(emu, quiet, 44100, 48000, [0] * 4, lp=200, pre=[1e-20, 1.0, 0.5, 1e-20])
    c = counters(emu)
    assert c["divided"] > 0 and c["fast"] > 4 * c["slow"]
    # no filter, stereo and mono-in-stereo sources, sources at the mixer's rate (taps used raw, times the gain)
    ch_in = [2 if i % 3 else 1 for i in range(12)]
    pcms = [noise(ci * (700 + 9 * i), 950 + i) for i, ci in enumerate(ch_in)]
    check(emu, pcms, [44100, 48000, 22050] * 4, 48000, [0] * 12, channels=2, ch_in=ch_in, pre=[0.5 + 0.1 * i for i in range(12)])
    check(emu, pcms, 48000, 48000, [3 * i for i in range(12)], hp=300, channels=2, ch_in=ch_in, pre=0.9)


@pytest.mark.parametrize("what,seed,cases", [("batches", 11, 40), ("sessions", 12, 15)])
def test_randomised_soak(emu, what, seed, cases):
    """tests/emu/stress.py: random mixers, rate pairs (up and down), mono / stereo / mono-in-stereo sources, ragged lengths and
    starts, chain variants with and without a gain in front -- every case bit for bit against the oracle.  (Hundreds of cases
    per seed run in a minute from the command line; the suite keeps one short seed of each kind.)"""
    sys.path[:0] = [os.path.join(HERE, "emu")]
    import stress
    assert getattr(stress, what)(seed, cases) == 0


def test_filter_in_front_of_the_conversion(emu):
    """`source.low_pass(f)` handed to the mixer (or appended to a Player, whose volume then sits behind it, player.rs:120-128):
    the filter runs at the SOURCE's rate on every input frame, the interpolation reads its outputs.  Up-sampling on the fast
    tiles, down-sampling and ragged edges on the slow ones, same-rate sources, mono / stereo / mono-in-stereo."""
    n = 40
    pcms = [noise(1800 + 13 * i, 1300 + i) for i in range(n)]
    counters(emu)
    check(emu, pcms, 44100, 48000, [0] * n, lp=300, front=True, pre=0.9, mid=[0.5 + 0.01 * i for i in range(n)], gain=0.8)
    c = counters(emu)
    assert c["fast"] > 4 * c["slow"]
    check(emu, pcms, [44100, 22050, 48000, 32000] * 10, 48000, [5 * (i % 7) for i in range(n)], hp=500, front=True)
    counters(emu)
    check(emu, pcms, 48000, 44100, [0] * n, lp=1000, front=True, mid=0.7)                 # above the mixer's rate: one or two
    check(emu, pcms, 96000, 48000, [2 * (i % 3) for i in range(n)], lp=1000, front=True, gain=0.9)   # filter steps per output
    c = counters(emu)
    assert c["fast"] > 4 * c["slow"]
    counters(emu)
    check(emu, pcms[:8], 96000, 44100, [3 * i for i in range(8)], lp=1000, front=True)    # more than two input frames per output:
    assert counters(emu)["fast"] == 0                                                      # slow tiles
    ch_in = [2 if i % 3 else 1 for i in range(12)]
    st = [noise(ci * (900 + 9 * i), 1400 + i) for i, ci in enumerate(ch_in)]
    check(emu, st, [44100, 48000, 22050] * 4, 48000, [0] * 12, channels=2, ch_in=ch_in, lp=400, front=True, mid=0.6, gain=1.1)
    # a filter that rings down into denormals behind a burst: the guarded division of the fast tiles
    quiet = [np.concatenate([noise(200, 1500 + i) * np.float32(1e-30), np.zeros(4000, np.float32)]) for i in range(4)]
    counters(emu)
    check(emu, quiet, 44100, 48000, [0] * 4, lp=4000, front=True)
    assert counters(emu)["divided"] > 0        # lane 0 met quotients outside the reciprocal's exact range and divided
    # tiny streams
    check(emu, [noise(k, 1600 + k) for k in (0, 1, 2, 3, 5)], 44100, 48000, [0, 1, 2, 3, 4], lp=300, front=True, mid=0.5)


def test_sources_above_the_mixers_rate_on_fast_tiles(emu):
    """48 kHz sources in a 44.1 kHz mixer, 96 kHz in a 48 kHz one (exactly two frames per output), 88.2 kHz in 48 kHz: up to
    twice the mixer's rate the lane kernel has fast tiles of its own (DOWN: one or two frames per output plus the carry, both
    taps reloaded); beyond that (96 kHz into 44.1 kHz) the slow tiles."""
    n = 40
    pcms = [noise(4000 + 29 * i, 1700 + i) for i in range(n)]
    for in_rate, mix_rate, kw in ((48000, 44100, dict(lp=300, gain=0.8)), (96000, 48000, dict(lp=300, gain=0.8)), (88200, 48000, dict(hp=400)),
                                  (48000, 44100, dict(gain=1.1)), (48000, 44100, dict(lp=300, gain=0.8, pre=[0.3 + 0.01 * i for i in range(n)]))):
        counters(emu)
        check(emu, pcms, in_rate, mix_rate, [3 * (i % 5) for i in range(n)], **kw)
        c = counters(emu)
        assert c["fast"] > 4 * c["slow"] and c["refills"] > 20, (in_rate, mix_rate, c)
    counters(emu)
    check(emu, pcms[:8], 96000, 44100, [0] * 8, lp=300, gain=0.8)
    assert counters(emu)["fast"] == 0
    ch_in = [2 if i % 3 else 1 for i in range(12)]
    st = [noise(ci * (2500 + 9 * i), 1800 + i) for i, ci in enumerate(ch_in)]
    counters(emu)
    check(emu, st, [48000, 44100, 88200] * 4, 44100, [0] * 12, channels=2, ch_in=ch_in, lp=400, gain=1.1)
    assert counters(emu)["fast"] > 100


def test_three_slot_ring_variant(built):
    """The A/B knob -DRB_LANES_UP_SLOTS=3 (three ring slots, one chunk of look-ahead: rb_lanes_core.h) stays correct: a subset of
    this file against an emulator library built with it (RB_EMU_VARIANT=slots3 runs everything)."""
    env = dict(os.environ, RB_EMU_VARIANT="slots3")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-k",
                        "cfg3 or stereo_ragged or mixed_rates or gain_in_front or filter_in_front or any_split"],
                       capture_output=True, text=True, env=env, timeout=1200, cwd=os.path.dirname(HERE))
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout + r.stderr)[-3000:]
