"""N>1 plumbing on CPU: two processes, gloo backend.  Each rank computes the partial mix of its
contiguous slice of the streams (with the oracle — this is a test of the sharding + all-reduce logic in
rodio_b200/dist.py, not of the kernels) and the all-reduced result must equal the single-process mixer
within the cross-shard tolerance (association across shards differs from the strictly sequential sum)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _streams():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rodio_b200 as rb
    from helpers import noise
    return [rb.UniformSourceIterator(rb.TestSource(noise(2000 + 10 * s, 300 + s), 1, 44100), 1, 48000)
            .low_pass(200).amplify(1.2) for s in range(11)]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import to_oracle
    from rodio_b200 import dist as rbd
    r, lr, w = rbd.init_process_group("gloo")
    assert (r, w) == (rank, world)
    srcs = _streams()
    lo, hi = rbd.shard_range(len(srcs), rank, world)
    full_len = oracle.mixer([to_oracle(s) for s in srcs], 1, 48000).size
    part = oracle.mixer([to_oracle(s) for s in srcs[lo:hi]], 1, 48000) if hi > lo else np.zeros(0, np.float32)
    mix = torch.zeros(full_len, dtype=torch.float32)
    mix[: part.size] = torch.from_numpy(part)
    rbd.all_reduce_mix(mix)
    t = rbd.max_over_ranks(float(rank + 1))
    assert t == float(world)
    np.save(os.path.join(out_dir, f"mix_{rank}.npy"), mix.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_in_order():
    sys.path.insert(0, ROOT)
    from rodio_b200 import dist as rbd
    for n in [0, 1, 7, 4096, 65537]:
        for w in [1, 2, 3, 8]:
            ranges = [rbd.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    with pytest.raises(ValueError):
        rbd.shard_range(4, 2, 2)


def test_two_rank_all_reduce_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import to_oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = oracle.mixer([to_oracle(s) for s in _streams()], 1, 48000)
    m0 = np.load(tmp_path / "mix_0.npy")
    m1 = np.load(tmp_path / "mix_1.npy")
    assert np.array_equal(m0, m1), "every rank holds the same all-reduced mix"
    assert m0.shape == want.shape
    assert np.max(np.abs(m0 - want)) <= 1e-5 * np.max(np.abs(want))
