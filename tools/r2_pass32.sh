#!/usr/bin/env bash
# pass 32 (2 GPUs): the bench line through torchrun with the final code, the two-GPU communicator test through the C ABI
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass32
mkdir -p "$OUT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > "$OUT/bench_2gpu.json" 2> "$OUT/bench_2gpu.err"; echo "bench 2gpu exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python -m pytest tests -q -m gpu -k "comm or two_gpus or dist" > "$OUT/pytest_comm.log" 2>&1; echo "pytest comm exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_comm.log" >> "$OUT/summary.txt"
tail -c 1500 "$OUT/bench_2gpu.json"
cat "$OUT/summary.txt"
