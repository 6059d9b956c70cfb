#!/usr/bin/env bash
# pass 33 (2 GPUs): the cross-shard sum over NVLink peer memory (k_mix_exchange) -- one process / two GPUs through the C ABI, two
# processes (cudaIpc), A/B against NCCL on the bench line
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass33
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -x -k "comm" -s > "$OUT/pytest_comm.log" 2>&1; echo "pytest comm exit $?" | tee -a "$OUT/summary.txt"
grep -E "transport|passed|failed|differ|GPU [0-9]" "$OUT/pytest_comm.log" | tail -20 >> "$OUT/summary.txt"
for v in p2p nccl; do
  if [ $v = nccl ]; then export RB_COMM_NCCL_ONLY=1; else unset RB_COMM_NCCL_ONLY; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/bench_2gpu_$v.json" 2> "$OUT/bench_2gpu_$v.err"; echo "bench $v exit $?" | tee -a "$OUT/summary.txt"
  python -c "
import json
d=json.loads(open('$OUT/bench_2gpu_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']), round(d['ms_per_step'],4), d['allreduce'])" | tee -a "$OUT/summary.txt"
done
cat "$OUT/summary.txt"
