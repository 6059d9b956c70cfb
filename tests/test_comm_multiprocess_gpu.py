"""The cross-shard mixer sum between PROCESSES (one per GPU, the form bench.py and a Rust host use): rb_comm_init_rank +
rb_batch_render_mix_allreduce.  On an NVLink box the ranks map each other's mailboxes (cudaIpc) and k_mix_exchange sums the shards
in rank order from +0.0 -- bit for bit what the host computes from the shards' own mixes, identical on every rank; with
RB_COMM_NCCL_ONLY=1 the same call goes through ncclAllReduce (<= 1e-5 * peak).  Needs two GPUs: skipped on a one-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, nccl_only):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    if nccl_only:
        os.environ["RB_COMM_NCCL_ONLY"] = "1"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import rodio_b200 as rb
    from helpers import noise
    from rodio_b200 import dist as rbd
    torch.cuda.set_device(rank)
    rbd.init_process_group("nccl")
    ctx = rb.Context(rank)
    ids = [rb.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = rb.Comm(ctx, world, rank, ids[0])
    n = 600                                              # 300 streams per rank: several CTAs, the partial rows are summed in the exchange kernel
    srcs = [rb.UniformSourceIterator(rb.TestSource(noise(3000, 4000 + s), 1, 44100), 1, 48000).low_pass(200).amplify(1.2) for s in range(n)]
    lo, hi = rbd.shard_range(n, rank, world)
    with rb.Batch(srcs[lo:hi], 1, 48000, ctx=ctx) as b:
        b.upload_all()
        own = b.render_mix()                             # this shard alone
        for _ in range(3):
            comm.render_mix_allreduce(b)                 # three renders: both mailbox buffers and the tags come round
        got = b.read_mix(0, b.mix_len)
        # a second batch with another kernel family (plain mixer: general path) on the same communicator
        plain = [rb.TestSource(noise(b.mix_len, 5000 + s), 1, 48000) for s in range(lo, lo + 5)]
        with rb.Batch(plain, 1, 48000, ctx=ctx) as b2:
            b2.upload_all()
            own2 = b2.render_mix()
            comm.render_mix_allreduce(b2)
            got2 = b2.read_mix(0, b2.mix_len)
    np.save(os.path.join(out_dir, f"own_{rank}.npy"), own)
    np.save(os.path.join(out_dir, f"got_{rank}.npy"), got)
    np.save(os.path.join(out_dir, f"own2_{rank}.npy"), own2)
    np.save(os.path.join(out_dir, f"got2_{rank}.npy"), got2)
    with open(os.path.join(out_dir, f"transport_{rank}.txt"), "w") as f:
        f.write(comm.transport)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("nccl_only", [False, True])
def test_two_process_allreduce(tmp_path, nccl_only):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), nccl_only), nprocs=world, join=True)
    transport = [open(tmp_path / f"transport_{r}.txt").read() for r in range(world)]
    assert transport[0] == transport[1]
    for tag in ("", "2"):
        own = [np.load(tmp_path / f"own{tag}_{r}.npy") for r in range(world)]
        got = [np.load(tmp_path / f"got{tag}_{r}.npy") for r in range(world)]
        ordered = np.zeros_like(own[0])
        for o in own:
            ordered = ordered + o                        # rank order, from +0.0
        peak = float(np.max(np.abs(ordered)))
        for r in range(world):
            assert np.max(np.abs(got[r] - ordered)) <= 1e-5 * peak
        if not nccl_only and transport[0].startswith("p2p"):
            for r in range(world):
                assert np.array_equal(got[r].view(np.uint32), ordered.view(np.uint32)), f"rank {r}: not the rank-ordered sum"
    if nccl_only:
        assert transport[0].startswith("nccl")
    print("transport:", transport[0])
