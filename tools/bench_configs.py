#!/usr/bin/env python
"""Time the other BASELINE.json configs (cfg2 mixer of sines, cfg4 effect chain, cfg5 sweep) on one GPU.
Inputs are resident in HBM; times are CUDA events on the context's stream.  Prints one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rodio_b200 as rb
from rodio_b200 import dist as rbd


def time_batch(name, srcs, mixer, flags=0, steps=10, fill="uniform"):
    ctx = rb.default_context(0)
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    with rb.Batch(srcs, *mixer, flags=flags, ctx=ctx) as b:
        S = len(srcs)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        for i in range(S if fill is not None else 0):
            p, cap = b.input_device_ptr(i)
            t = torch.as_tensor(rbd.DeviceArray(p, cap), device=dev)
            with torch.cuda.stream(ext):
                t.uniform_(-0.5, 0.5, generator=gen)
        for _ in range(3):
            b.render_mix_device()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            b.render_mix_device()
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        samples = sum(b.stream_out_len(i) for i in range(S))
        out = {"case": name, "streams": S, "ms": round(ms, 4), "Msamples_s": round(samples / ms / 1e3, 1),
               "algo_GBs": round(b.algorithmic_bytes / ms / 1e6, 1), "frac_6570": round(b.algorithmic_bytes / ms / 1e6 / 6570, 4),
               "launches": b.launches_per_render, "kernel_family": b.kernel_family}
        print(json.dumps(out), flush=True)


def main():
    which = sys.argv[1:] or ["cfg2", "cfg4", "cfg3_stereo"]
    z = lambda n: np.zeros(n, np.float32)
    if "cfg2" in which:
        for S, secs in [(1024, 10), (4096, 2), (16384, 1)]:
            srcs = [rb.TestSource(z(48000 * secs), 1, 48000) for _ in range(S)]
            time_batch(f"cfg2 mixer of {S} mono 48k sources x {secs}s (fused)", srcs, (1, 48000))
        for S in (1024, 16384):
            srcs = [rb.TestSource(z(480), 1, 48000) for _ in range(S)]
            time_batch(f"cfg2 mixer of {S} sources x 10 ms block (concurrent source runs)", srcs, (1, 48000), steps=20)
            time_batch(f"cfg2 mixer of {S} sources x 10 ms block, exact order", srcs, (1, 48000), steps=20,
                       flags=rb.capi.RB_MIX_EXACT_ORDER)
        srcs = [rb.TestSource(z(48000 * 10), 1, 48000) for _ in range(1024)]
        time_batch("cfg2 1024 x 10s exact-order general path", srcs, (1, 48000), flags=rb.capi.RB_MIX_EXACT_ORDER)
    if "cfg4" in which:
        S, frames = 512, 48000
        srcs = [rb.Spatial(rb.TestSource(z(2 * frames), 2, 48000), [float(s % 7 - 3), 1.0, 0.0], [-1, 0, 0], [1, 0, 0])
                .reverb(rb.Duration.from_millis(50), 0.3).automatic_gain_control() for s in range(S)]
        time_batch("cfg4 512 stereo: spatial -> reverb -> agc -> mix (general path)", srcs, (2, 48000), steps=3)
    if "limit" in which:
        S, frames = 512, 48000
        srcs = [rb.TestSource(z(2 * frames), 2, 48000).limit() for s in range(S)]
        time_batch("limit: 512 stereo x 1 s, limit(default) -> mix [k_fused_fx]", srcs, (2, 48000), steps=3)
        time_batch("limit: the same through the general path (k_limit_tile)", srcs, (2, 48000), steps=3, flags=rb.capi.RB_NO_FUSION)
        srcs = [rb.Spatial(rb.TestSource(z(2 * frames), 2, 48000), [float(s % 7 - 3), 1.0, 0.0], [-1, 0, 0], [1, 0, 0])
                .reverb(rb.Duration.from_millis(50), 0.3).automatic_gain_control().limit() for s in range(S)]
        time_batch("cfg4 + limit: spatial -> reverb -> agc -> limit -> mix [k_fused_fx]", srcs, (2, 48000), steps=3)
    if "nofilter" in which:
        for ch, S in ((1, 4096), (2, 2048)):
            srcs = [rb.UniformSourceIterator(rb.TestSource(z(44100 * 2 * ch), ch, 44100), ch, 48000).amplify(0.8)
                    for _ in range(S)]
            time_batch(f"resample 44.1k->48k -> amplify -> mix, {S} streams x {ch} ch x 2s (HOT pipeline, no filter)", srcs, (ch, 48000))
            time_batch(f"same, {S // 8} streams (generic fused kernel, timeline split over CTAs)", srcs[: S // 8], (ch, 48000))
    if "cfg3_stereo" in which:
        S, frames = 2048, 44100 * 2
        srcs = [rb.UniformSourceIterator(rb.TestSource(z(2 * frames), 2, 44100), 2, 48000).low_pass(200).amplify(1.2)
                for _ in range(S)]
        time_batch("cfg3 shape, 2048 STEREO streams x 2s (stereo HOT kernel)", srcs, (2, 48000))
        srcs = [rb.UniformSourceIterator(rb.TestSource(z(44100 * 2), 1, 44100), 1, 48000).low_pass(200).amplify(1.2)
                for _ in range(4096)]
        time_batch("cfg3 4096 mono x 2s, general path (RB_NO_FUSION)", srcs, (1, 48000), flags=rb.capi.RB_NO_FUSION, steps=3)
    if "cfg5" in which:
        # HBM sweep: the cfg3 pipeline at 1 s of 44.1 kHz mono per stream, S from 1 to 65536
        for S in [1, 4, 16, 64, 256, 1024, 4096, 16384, 65536]:
            one = z(44100)
            srcs = [rb.UniformSourceIterator(rb.TestSource(one, 1, 44100), 1, 48000).low_pass(200).amplify(1.2)
                    for _ in range(S)]
            time_batch(f"cfg5 sweep S={S}: 44.1k mono x 1s -> uniform(1,48k) -> low_pass(200) -> amplify -> mix", srcs,
                       (1, 48000), steps=5 if S >= 16384 else 10)
    if "lanes" in which:
        # the lane-per-stream kernel (RB_FUSED_LANES) beside the default fused kernel on the large end of the sweep,
        # at the bench's low_pass(200), at low_pass(1000), and without a filter
        for S in [4096, 16384, 65536]:
            one = z(44100)
            for label, mk in (("low_pass(200) -> amplify", lambda s: s.low_pass(200).amplify(1.2)),
                              ("low_pass(1000) -> amplify", lambda s: s.low_pass(1000).amplify(1.2)),
                              ("amplify (no filter)", lambda s: s.amplify(1.2))):
                srcs = [mk(rb.UniformSourceIterator(rb.TestSource(one, 1, 44100), 1, 48000)) for _ in range(S)]
                for fl, nm in ((rb.capi.RB_FUSED_LANES, "k_fused_lanes"), (0, "default")):
                    time_batch(f"lanes sweep S={S}: 44.1k mono x 1s -> uniform(1,48k) -> {label} -> mix [{nm}]", srcs,
                               (1, 48000), flags=fl, steps=5)
    if "duo" in which:
        # the lane-pair kernel and the time-parallel plan beside k_fused_lanes and the default planner, over the batch sizes
        # where the planner's thresholds lie; 1 s of 44.1 kHz per stream except the 4096-stream points (2 s: the bench config)
        one = z(44100)
        for S, secs in [(2048, 2), (4096, 2), (8192, 1), (16384, 1), (32768, 1), (65536, 1)]:
            x = z(44100 * secs)
            for label, mk in (("low_pass(200) -> amplify", lambda s: s.low_pass(200).amplify(1.2)),
                              ("low_pass(1000) -> amplify", lambda s: s.low_pass(1000).amplify(1.2)),
                              ("amplify (no filter)", lambda s: s.amplify(1.2))):
                srcs = [mk(rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000)) for _ in range(S)]
                kinds = [(0, "default"), (rb.capi.RB_FUSED_LANES, "k_fused_lanes"), (rb.capi.RB_FUSED_DUO, "k_fused_duo")]
                if "1000" in label:
                    kinds.append((rb.capi.RB_BIQUAD_TIME_PARALLEL, "time-parallel"))
                for fl, nm in kinds:
                    if nm == "k_fused_lanes" and S < 8192:
                        continue
                    time_batch(f"duo sweep S={S} x {secs}s: uniform(1,48k) -> {label} -> mix [{nm}]", srcs, (1, 48000), flags=fl, steps=5)
    if "cfg5big" in which:
        one = z(44100)
        for S in (16384, 32768, 65536):
            srcs = [rb.UniformSourceIterator(rb.TestSource(one, 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for _ in range(S)]
            time_batch(f"cfg5 S={S} x 1s low_pass(200) [RB_FUSED_DUO]", srcs, (1, 48000), flags=rb.capi.RB_FUSED_DUO, steps=5)
    if "exact" in which:
        # the benchmarked batch with RB_MIX_EXACT_ORDER (the reference's sequential sum over all streams) beside the default grouping
        x = z(44100 * 2)
        srcs = [rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for _ in range(4096)]
        time_batch("cfg3 4096 x 2s, default order (partial sums per CTA)", srcs, (1, 48000), steps=10)
        time_batch("cfg3 4096 x 2s, RB_MIX_EXACT_ORDER (running sum handed from CTA to CTA)", srcs, (1, 48000), flags=rb.capi.RB_MIX_EXACT_ORDER, steps=10)
        srcs = [rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).amplify(1.2) for _ in range(4096)]
        time_batch("no filter 4096 x 2s, RB_MIX_EXACT_ORDER (k_lerp_mix, one group)", srcs, (1, 48000), flags=rb.capi.RB_MIX_EXACT_ORDER, steps=5)
    if "exact4" in which:
        S, frames = 512, 48000
        srcs = [rb.Spatial(rb.TestSource(z(2 * frames), 2, 48000), [float(s % 7 - 3), 1.0, 0.0], [-1, 0, 0], [1, 0, 0])
                .reverb(rb.Duration.from_millis(50), 0.3).automatic_gain_control() for s in range(S)]
        time_batch("cfg4 512 stereo, default order [k_fused_fx]", srcs, (2, 48000), steps=5)
        time_batch("cfg4 512 stereo, RB_MIX_EXACT_ORDER [k_fused_fx, chain]", srcs, (2, 48000), flags=rb.capi.RB_MIX_EXACT_ORDER, steps=5)
    if "gen" in which:
        srcs = [rb.SineWave(min(110.0 * 2.0 ** (s / 128.0), 19999.0)).take(48000 * 10) for s in range(1024)]
        time_batch("cfg2 from generators: 1024 SineWave x 10 s generated and mixed on the device", srcs, (1, 48000), steps=3, fill=None)
    if "tp" in which:
        # the two plans that cut the timeline into segments: time-parallel low_pass(1000) and the filter-free chain, 4096 x 2 s
        x = z(44100 * 2)
        srcs = [rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).low_pass(1000).amplify(1.2) for _ in range(4096)]
        time_batch("tp: 4096 x 2s low_pass(1000) time-parallel", srcs, (1, 48000), flags=rb.capi.RB_BIQUAD_TIME_PARALLEL, steps=10)
        srcs = [rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).amplify(1.2) for _ in range(4096)]
        time_batch("tp: 4096 x 2s no filter (segments)", srcs, (1, 48000), steps=10)
    if "lanes_shapes" in which:
        # the lane kernel on the shapes added after its first device runs: stereo (ring geometry: see RB_LANES_STEREO_CHW in
        # rb_lanes_core.h, A/B via RODIO_B200_LIB), mono sources in a stereo mixer, and the chain the way rodio users write it
        # (gain / filter in front of the mixer's conversion)
        S = 16384
        mono, stereo = z(44100), z(2 * 44100)
        shapes = [
            ("stereo: uniform -> low_pass(200) -> amplify", 2, lambda: rb.UniformSourceIterator(rb.TestSource(stereo, 2, 44100), 2, 48000).low_pass(200).amplify(1.2)),
            ("mono in a stereo mixer: uniform -> low_pass(200) -> amplify", 2, lambda: rb.UniformSourceIterator(rb.TestSource(mono, 1, 44100), 2, 48000).low_pass(200).amplify(1.2)),
            ("mono: amplify -> uniform -> low_pass(200) -> amplify", 1, lambda: rb.UniformSourceIterator(rb.TestSource(mono, 1, 44100).amplify(0.9), 1, 48000).low_pass(200).amplify(1.2)),
            ("mono: low_pass(200) -> amplify -> uniform (filter in front)", 1, lambda: rb.UniformSourceIterator(rb.TestSource(mono, 1, 44100).low_pass(200).amplify(0.9), 1, 48000)),
            ("stereo: low_pass(200) -> amplify -> uniform (filter in front)", 2, lambda: rb.UniformSourceIterator(rb.TestSource(stereo, 2, 44100).low_pass(200).amplify(0.9), 2, 48000)),
            ("mono 96 kHz into 48 kHz (DOWN tiles): uniform -> low_pass(200) -> amplify", 1, lambda: rb.UniformSourceIterator(rb.TestSource(z(96000), 1, 96000), 1, 48000).low_pass(200).amplify(1.2)),
        ]
        for label, ch, mk in shapes:
            srcs = [mk() for _ in range(S)]
            time_batch(f"lanes shapes S={S} x 1s: {label} [k_fused_lanes]", srcs, (ch, 48000), flags=rb.capi.RB_FUSED_LANES, steps=5)
    if "session" in which:
        # streaming sessions: every source receives 10 ms of 44.1 kHz PCM per step (one packed push), the mixer output is
        # pulled as it becomes available; wall clock per step through the public calls (host -> device -> host included)
        import time
        shapes = {"uniform -> low_pass(200) -> amplify": lambda ch: rb.UniformSourceIterator(rb.TestSource(z(0), ch, 44100), ch, 48000).low_pass(200).amplify(1.2),
                  # the chain of a Player with a user filter: both in front of the mixer's conversion (src/player.rs:120-128)
                  "low_pass(200) -> amplify(volume) -> uniform [Player]": lambda ch: rb.UniformSourceIterator(rb.TestSource(z(0), ch, 44100).low_pass(200).amplify(0.8), ch, 48000)}
        names = list(shapes)
        for S, ch, k in [(256, 1, 0), (4096, 1, 0), (1024, 2, 0), (16384, 1, 0), (4096, 1, 1), (1024, 2, 1)]:
            label, mk = names[k], shapes[names[k]]
            chains = [mk(ch) for _ in range(S)]
            rng = np.random.default_rng(1)
            block = [rng.uniform(-0.5, 0.5, 441 * ch).astype(np.float32) for _ in range(min(S, 64))]
            blocks = [block[i % len(block)] for i in range(S)]
            with rb.Session(chains, 48000, fifo_frames=2048, max_block_frames=1024, ctx=rb.default_context(0)) as sess:
                for _ in range(20):                       # warm-up
                    sess.push_packed(blocks)
                    sess.render(1024)
                t0, frames, steps = time.perf_counter(), 0, 200
                for k in range(steps):
                    if "Player" in label and k % 10 == 0:
                        sess.set_volume(k % S, 0.5 + 0.001 * (k % 100))   # Player::set_volume now and then
                    sess.push_packed(blocks)
                    out, _ = sess.render(1024)
                    frames += out.size // ch
                dt = time.perf_counter() - t0
            print(json.dumps({"case": f"session: {S} x {ch} ch sources, 10 ms pushes -> {label} -> mix", "streams": S,
                              "ms_per_10ms_block": round(1e3 * dt / steps, 3), "realtime_factor": round(frames / 48000 / dt, 2),
                              "Msamples_s": round(S * frames * ch / dt / 1e6, 1)}), flush=True)
    if "cpu" in which:
        # C++ restatement of rodio's CPU path on this box's host cores: one audio thread (what rodio itself runs)
        # and all cores with a final partial-mix reduction; dynamic dispatch (Box<dyn Source>) and monomorphised.
        import oracle

        def to_oracle_stream(x):
            src = rb.UniformSourceIterator(rb.TestSource(x, 1, 44100), 1, 48000).low_pass(200).amplify(1.2)
            return oracle.Stream(pcm=src.pcm, channels=src.base_channels, sample_rate=src.base_rate, effects=src.effects,
                                 span_len=src.span_len, mix_start=0)

        rng = np.random.default_rng(5)
        ncpu = os.cpu_count() or 1
        for threads, S in [(1, 64), (ncpu, 64 * min(ncpu, 64))]:
            xs = [rng.uniform(-1, 1, 44100).astype(np.float32) for _ in range(min(S, 256))]
            streams = [to_oracle_stream(xs[i % len(xs)]) for i in range(S)]
            for static in (False, True):
                _, secs = oracle.mixer_mt(streams, 1, 48000, threads, 48000, static_dispatch=static)
                print(json.dumps({"case": f"cpu cfg3 pipeline, {threads} thread(s), {'static' if static else 'dyn'} dispatch",
                                  "streams": S, "seconds": round(secs, 4),
                                  "Msamples_s": round(S * 48000 / secs / 1e6, 1), "cores": threads}), flush=True)


if __name__ == "__main__":
    main()
