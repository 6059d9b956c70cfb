#!/usr/bin/env python
"""Randomised soak of the lane kernel on the CPU emulator (not collected by pytest: minutes per run).
    python tests/emu/stress.py batches [seed] [cases]     random batches: channels, rate pairs, lengths, starts, chain variants
    python tests/emu/stress.py sessions [seed] [cases]    random sessions: random pushes / renders over random sources
Every case is held bit for bit against the oracle streams summed with the kernel's tree (tests/test_lanes_emulator.py).
The emulator library is built on demand (tests/emu/build_emu.py)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import build_emu                         # noqa: E402
import test_lanes_emulator as T          # noqa: E402
from helpers import noise                # noqa: E402


def batches(seed, cases):
    emu = C.CDLL(build_emu.lanes_lib())
    emu.rb_lanes_emulate.restype = C.c_int
    rng = np.random.default_rng(seed)
    mixers = [(48000, [44100, 22050, 48000, 32000, 8000, 11025, 47999]), (44100, [22050, 44100, 32000, 8000, 11025]), (96000, [44100, 48000, 96000, 88200])]
    bad = 0
    for case in range(cases):
        mix_rate, in_rates = mixers[rng.integers(len(mixers))]
        ch = int(rng.choice([1, 2]))
        n = int(rng.integers(1, 80))
        rates = [int(rng.choice(in_rates)) for _ in range(n)] if rng.random() < 0.6 else [int(rng.choice(in_rates))] * n
        ch_in = [ch if (ch == 1 or rng.random() < 0.6) else 1 for _ in range(n)]
        lens = [int(rng.choice([0, 1, 2, 3, 17, int(rng.integers(4, 2500))])) if rng.random() < 0.3 else int(rng.integers(50, 2500)) for _ in range(n)]
        starts = [0] * n if rng.random() < 0.5 else [int(rng.integers(0, 1500)) for _ in range(n)]
        pcms = [noise(ci * L, 10000 * case + i) for i, (ci, L) in enumerate(zip(ch_in, lens))]
        kind = rng.integers(4)
        kw = [dict(lp=int(rng.choice([200, 1000, 3000])), gain=float(np.float32(rng.uniform(0.2, 1.5)))), dict(hp=300), dict(gain=1.2), dict()][kind]
        if kind < 2 and rng.random() < 0.35:   # the filter in front of the conversion, at the source's rate, gains around it
            kw = dict(kw, front=True)
            if rng.random() < 0.6:
                kw["mid"] = [float(np.float32(rng.uniform(0.1, 1.3))) for _ in range(n)]
        if rng.random() < 0.4:   # source.amplify(v) in front of the mixer's conversion; a few outside the fast-path gain range
            kw = dict(kw, pre=[float(np.float32(rng.choice([0.001, -0.5, 100.0, rng.uniform(0.05, 2.0)]))) for _ in range(n)])
        try:
            T.check(emu, pcms, rates, mix_rate, starts, channels=ch, ch_in=ch_in, **kw)
        except AssertionError as e:
            bad += 1
            print("CASE", case, "FAILED", dict(mix=mix_rate, ch=ch, n=n, kw=kw), str(e)[:300], flush=True)
    print("batches:", cases, "cases, failures:", bad, flush=True)
    return bad


def sessions(seed, cases):
    emu = C.CDLL(build_emu.lanes_lib())
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(cases):
        mix_rate = int(rng.choice([48000, 44100]))
        in_rates = [44100, 22050, 48000, 32000, 96000, 8000]
        ch = int(rng.choice([1, 2]))
        n = int(rng.integers(1, 40))
        rates = [int(rng.choice(in_rates)) for _ in range(n)]
        ch_in = [ch if (ch == 1 or rng.random() < 0.6) else 1 for _ in range(n)]
        lens = [int(rng.integers(0, 1500)) for _ in range(n)]
        starts = [0 if rng.random() < 0.6 else int(rng.integers(0, 600)) for _ in range(n)]
        pcms = [noise(ci * L, 777 * case + i) for i, (ci, L) in enumerate(zip(ch_in, lens))]
        ops, left = [], list(lens)
        while any(left):
            for r in rng.permutation(n):
                k = min(left[r], int(rng.integers(1, 400)))
                if k and rng.random() < 0.7:
                    ops.append((0, int(r), k)); left[r] -= k
            ops.append((1, 0, int(rng.choice([1, 3, 7, 64, 240, 480, 1000]))))
        kw = [dict(lp=int(rng.choice([200, 2000])), gain=0.8), dict(gain=1.1), dict()][rng.integers(3)]
        try:
            T.session_case(emu, pcms, starts, ops, in_rate=rates, mix_rate=mix_rate, channels=ch, ch_in=ch_in, **kw)
        except AssertionError as e:
            bad += 1
            print("CASE", case, "FAILED", dict(mix=mix_rate, ch=ch, n=n, kw=kw), str(e)[:300], flush=True)
    print("sessions:", cases, "cases, failures:", bad, flush=True)
    return bad


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "batches"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cases = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    sys.exit(1 if (batches if what == "batches" else sessions)(seed, cases) else 0)
