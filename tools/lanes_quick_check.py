#!/usr/bin/env python
"""Smallest possible device check of k_fused_lanes (no torch, no pytest, no rebuild): a 70-stream cfg3-shaped batch through
RB_FUSED_LANES against the oracle streams summed with the kernel's tree, a stereo batch, one streaming session, then a
wall-clock look at 16 384 streams x 1 s.  Prints one line per step; exits non-zero on the first mismatch."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle                                                   # noqa: E402
import rodio_b200 as rb                                         # noqa: E402
from helpers import lanes_expected_mix, noise, to_oracle        # noqa: E402

LANES = rb.capi.RB_FUSED_LANES


def case(ctx, n, ch, lp, gain, name):
    pcms = [noise(ch * (3000 + 7 * i), 100 + i) for i in range(n)]
    srcs = [rb.UniformSourceIterator(rb.TestSource(p, ch, 44100), ch, 48000) for p in pcms]
    if lp:
        srcs = [s.low_pass(lp) for s in srcs]
    if gain:
        srcs = [s.amplify(gain) for s in srcs]
    with rb.Batch(srcs, ch, 48000, flags=LANES, ctx=ctx) as b:
        fam = b.kernel_family
        b.upload_all()
        got = b.render_mix()
    per = [oracle.chain_uniform(to_oracle(s), ch, 48000) for s in srcs]
    want = lanes_expected_mix(per, [0] * n, got.size)
    bad = int(np.count_nonzero(got.view(np.uint32) != want.view(np.uint32)))
    nan = int(np.count_nonzero(np.isnan(got)))
    err = float(np.max(np.abs(got.astype(np.float64) - want))) if got.size else 0.0
    print(f"{name}: family {fam}, {got.size} samples, {bad} differ bitwise ({nan} NaN), max abs err {err:.3e}", flush=True)
    return fam == 2 and bad == 0, pcms, got


def main():
    ctx = rb.Context(0)
    ok = True
    r, _, _ = case(ctx, 70, 1, 200, 1.2, "mono  cfg3 shape, 70 streams")
    ok &= r
    r, _, _ = case(ctx, 40, 2, 300, 0.9, "stereo cfg3 shape, 40 streams")
    ok &= r
    r, pcms, whole = case(ctx, 33, 1, 1000, 0.8, "mono  low_pass(1000), 33 streams")
    ok &= r
    chains = [rb.UniformSourceIterator(rb.TestSource(np.zeros(0, np.float32), 1, 44100), 1, 48000).low_pass(1000).amplify(0.8)
              for _ in pcms]
    got, pos, ended = [], 0, False
    with rb.Session(chains, 48000, fifo_frames=2048, max_block_frames=480, ctx=ctx) as s:
        longest = max(p.size for p in pcms)
        while not ended:
            s.push_packed([p[pos:pos + 441] for p in pcms], [pos + 441 >= p.size for p in pcms])
            pos += 441
            while True:
                blk, ended = s.render(480)
                got.append(blk)
                if blk.size == 0 or ended:
                    break
            if pos > longest + 10000:
                break
    got = np.concatenate(got)
    same = got.size == whole.size and bool(np.array_equal(got.view(np.uint32), whole.view(np.uint32)))
    print(f"session, 10 ms blocks: {got.size} samples, identical to the whole render: {same}", flush=True)
    ok &= same
    sizes = ([16384] if "--time" in sys.argv else []) + ([65536] if "--time-big" in sys.argv else [])
    for S in (sizes if ok else []):
        one = np.zeros(44100, np.float32)
        for flags, nm in ((LANES, "k_fused_lanes"),) + (() if "--lanes-only" in sys.argv else ((0, "default"),)):
            srcs = [rb.UniformSourceIterator(rb.TestSource(one, 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for _ in range(S)]
            with rb.Batch(srcs, 1, 48000, flags=flags, ctx=ctx) as b:
                for i in range(S):                  # silence: classified safe, same arithmetic cost as any normal input
                    b.upload(i, one)
                b.render_mix_device()
                ctx.sync()
                t0 = time.perf_counter()
                for _ in range(5):
                    b.render_mix_device()
                ctx.sync()
                ms = (time.perf_counter() - t0) / 5 * 1e3
                print(f"{S} x 1 s [{nm}, family {b.kernel_family}]: {ms:.3f} ms per render, "
                      f"{b.algorithmic_bytes / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    print("QUICK CHECK " + ("PASSED" if ok else "FAILED"), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
