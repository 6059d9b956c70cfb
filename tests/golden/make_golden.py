"""Regenerates tests/golden/*.json.

Two kinds of fixtures:

* reference_vectors.json -- the literal input/output vectors of the reference's own unit tests for the path,
  transcribed as DATA with the file:line they come from (rodio @ 1f927962).  They were typed from the reference
  test sources; this script only re-serialises the table below.
* oracle_chains.json -- fingerprints (length, SHA-256 of the f32 bytes, first samples) of the CPU oracle's output
  for every adapter chain in tests/chains.py on the seeded inputs defined there.  The reference is Rust and
  cannot run in this image (no rustc / cargo), so for the rows the reference itself does not pin (biquad, AGC,
  reverb, spatial, uniform ...) the fixture pins the *restatement*: any drift of oracle/rodio_oracle.hpp shows
  up as a CPU test failure, and the CUDA path is checked against the committed fingerprints on the GPU box
  without building the oracle.  Limiter chains (log2/exp2: tolerance, not bit-exact) are stored as full arrays
  in limiter_chains.npz.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

REFERENCE_VECTORS = {
    "sample_rate_converter": [
        # src/conversions/sample_rate.rs:356-366 (upsample, compared after truncation like the reference does)
        {"cite": "src/conversions/sample_rate.rs:356-366", "input": [2, 16, 4, 18, 6, 20, 8, 22], "from": 2000,
         "to": 3000, "channels": 2, "trunc": True, "output": [2, 16, 3, 17, 4, 18, 6, 20, 7, 21, 8, 22]},
        {"cite": "src/conversions/sample_rate.rs:368-376", "input": [1, 14], "from": 1000, "to": 7000, "channels": 1,
         "trunc": True, "output": [1, 2, 4, 6, 8, 10, 12, 14]},
        {"cite": "src/conversions/sample_rate.rs:378-387", "input": list(range(17)), "from": 12000, "to": 2400,
         "channels": 1, "trunc": False, "output": [0, 5, 10, 15]},
    ],
    "channel_count_converter": [
        {"cite": "src/conversions/channels.rs:114-125", "input": [1, 2, 3, 4, 5, 6], "from": 3, "to": 2, "output": [1, 2, 4, 5]},
        {"cite": "src/conversions/channels.rs:114-125", "input": [1, 2, 3, 4, 5, 6, 7, 8], "from": 4, "to": 1, "output": [1, 5]},
        {"cite": "src/conversions/channels.rs:127-143", "input": [1, 2, 3, 4], "from": 1, "to": 2, "output": [1, 1, 2, 2, 3, 3, 4, 4]},
        {"cite": "src/conversions/channels.rs:127-143", "input": [1, 2], "from": 1, "to": 4, "output": [1, 1, 0, 0, 2, 2, 0, 0]},
        {"cite": "src/conversions/channels.rs:127-143", "input": [1, 2, 3, 4], "from": 2, "to": 4, "output": [1, 2, 0, 0, 3, 4, 0, 0]},
    ],
    "mixer": [
        # src/mixer.rs:208-230 basic: two mono sources at the mixer's own format
        {"cite": "src/mixer.rs:208-230", "mixer": [1, 48000],
         "sources": [{"channels": 1, "rate": 48000, "pcm": [10, -10, 10, -10]}, {"channels": 1, "rate": 48000, "pcm": [5, 5, 5, 5]}],
         "output": [15, -5, 15, -5]},
        # src/mixer.rs:232-258 channels: mono sources into a stereo mixer
        {"cite": "src/mixer.rs:232-258", "mixer": [2, 48000],
         "sources": [{"channels": 1, "rate": 48000, "pcm": [10, -10, 10, -10]}, {"channels": 1, "rate": 48000, "pcm": [5, 5, 5, 5]}],
         "output": [15, 15, -5, -5, 15, 15, -5, -5]},
        # src/mixer.rs:260-285 rates: 48 kHz sources into a 96 kHz mixer
        {"cite": "src/mixer.rs:260-285", "mixer": [1, 96000],
         "sources": [{"channels": 1, "rate": 48000, "pcm": [10, -10, 10, -10]}, {"channels": 1, "rate": 48000, "pcm": [5, 5, 5, 5]}],
         "output": [15, 5, -5, 5, 15, 5, -5]},
    ],
    "channel_volume": [
        {"cite": "src/source/channel_volume.rs:135-146", "channels": 1, "rate": 44100, "pcm": [1.0, 2.0, 3.0],
         "volumes": [0.5, 0.8], "output": [0.5, 0.8, 1.0, 1.6, 1.5, 2.4]},
        {"cite": "src/source/channel_volume.rs:148-155", "channels": 2, "rate": 44100, "pcm": [1.0, 2.0, 3.0, 4.0],
         "volumes": [1.0], "output": [1.5, 3.5]},
        {"cite": "src/source/channel_volume.rs:157-166", "channels": 2, "rate": 44100, "pcm": [1.0, 3.0, 2.0, 4.0],
         "volumes": [0.5, 2.0], "output": [1.0, 4.0, 1.5, 6.0]},
    ],
    # src/source/crossfade.rs:45-80: 1 Hz mono sources 1..=10, crossfade(a, b, 5 s + 1 ns); within 1e-6
    "crossfade": [
        {"cite": "src/source/crossfade.rs:45-63", "a": list(range(1, 11)), "b": list(range(1, 11)), "rate": 1,
         "duration_ns": 5_000_000_001, "tolerance": 1e-6, "output": [1, 2, 3, 4, 5]},
        {"cite": "src/source/crossfade.rs:65-80", "a": list(range(1, 11)), "b": [0] * 64, "b_spanless": True, "rate": 1,
         "duration_ns": 5_000_000_001, "tolerance": 1e-6, "output": [1.0, 1.6, 1.8, 1.6, 1.0]},
    ],
    # src/source/from_iter.rs:129-157 (the same vector as src/queue.rs:280-303): the format changes between the buffers
    "from_iter": [
        {"cite": "src/source/from_iter.rs:129-157",
         "buffers": [{"channels": 1, "rate": 48000, "pcm": [10, -10, 10, -10]}, {"channels": 2, "rate": 96000, "pcm": [5, 5, 5, 5]}],
         "output": [10, -10, 10, -10, 5, 5, 5, 5]},
    ],
    # src/source/signal_generator.rs:181-238
    "signal_generator": [
        {"cite": "src/source/signal_generator.rs:182-193", "function": 2, "rate": 2000, "frequency": 500.0, "output": [1, 1, -1, -1, 1, 1, -1, -1]},
        {"cite": "src/source/signal_generator.rs:195-214", "function": 1, "rate": 8000, "frequency": 1000.0,
         "output": [-1.0, -0.5, 0.0, 0.5, 1.0, 0.5, 0.0, -0.5, -1.0, -0.5, 0.0, 0.5, 1.0, 0.5, 0.0, -0.5]},
        {"cite": "src/source/signal_generator.rs:216-226", "function": 3, "rate": 200, "frequency": 50.0, "output": [0.0, 0.5, -1.0, -0.5, 0.0, 0.5, -1.0]},
        {"cite": "src/source/signal_generator.rs:228-238", "function": 0, "rate": 1000, "frequency": 100.0, "tolerance": 1e-4,
         "output": [0.0, 0.58778525, 0.95105652, 0.95105652, 0.58778525, 0.0, -0.58778554]},
    ],
    # src/math.rs:238-266 DECIBELS_LINEAR_TABLE (Wikipedia values; the reference asserts a ratio within 1 %, :268-316)
    "db_table": {"cite": "src/math.rs:238-316", "ratio_tolerance": 0.01, "rows": [
        [100., 100000.], [90., 31623.], [80., 10000.], [70., 3162.], [60., 1000.], [50., 316.2], [40., 100.],
        [30., 31.62], [20., 10.], [10., 3.162], [5.998, 1.995], [3.003, 1.413], [1.002, 1.122], [0., 1.],
        [-1.002, 0.891], [-3.003, 0.708], [-5.998, 0.501], [-10., 0.3162], [-20., 0.1], [-30., 0.03162],
        [-40., 0.01], [-50., 0.003162], [-60., 0.001], [-70., 0.0003162], [-80., 0.0001], [-90., 0.00003162],
        [-100., 0.00001]]},
}


def fingerprint(a: np.ndarray) -> dict:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return {"len": int(a.size), "sha256": hashlib.sha256(a.tobytes()).hexdigest(),
            "head_bits": [int(v) for v in a[:6].view(np.uint32)]}


def main() -> None:
    import oracle
    from chains import CHAINS, LIMIT_CHAINS
    from helpers import to_oracle

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(REFERENCE_VECTORS, f, indent=1)

    chains = {}
    for name in sorted(CHAINS):
        src = CHAINS[name]()
        _, ch, rate = oracle.chain(to_oracle(src))
        want = oracle.chain_uniform(to_oracle(src), ch, rate)     # what a mixer of the chain's own format pulls
        chains[name] = dict(fingerprint(want), channels=int(ch), sample_rate=int(rate))
    with open(os.path.join(HERE, "oracle_chains.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/rodio_oracle.hpp (C++ restatement)",
                   "chains": chains}, f, indent=1)

    lim = {}
    for name in sorted(LIMIT_CHAINS):
        lim[name] = oracle.chain(to_oracle(LIMIT_CHAINS[name]()))[0]
    np.savez_compressed(os.path.join(HERE, "limiter_chains.npz"), **lim)
    print(f"wrote {len(chains)} chain fingerprints, {len(lim)} limiter arrays")


if __name__ == "__main__":
    main()
