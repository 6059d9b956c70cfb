#!/usr/bin/env bash
# Round 2, two GPUs: the communicator behind the C ABI (C++ test, bench under torchrun), cfg4 after the role reassignment.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass13
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
timeout 300 python tools/bench_configs.py cfg4 > "$OUT/cfg4.jsonl" 2>&1
timeout 600 python -m pytest tests/test_cpp_mirror.py tests/test_bench_geometries_gpu.py -q -m gpu -k "comm or fx or cfg4" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/pytest.log" >> "$OUT/summary.txt"
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_2gpu.json" 2> "$OUT/bench_2gpu.err"; echo "bench 2 gpus exit $?" | tee -a "$OUT/summary.txt"
grep -c "Init COMPLETE\|nranks 2" "$OUT/bench_2gpu.err" >> "$OUT/summary.txt"
cat "$OUT/summary.txt"
