#!/usr/bin/env bash
# pass 35 (N GPUs, N = $1): the bench line through torchrun, peer-memory exchange against NCCL
set -u
cd "$(dirname "$0")/.."
N=${1:-4}
OUT=gpurun_out/r2_pass35_n$N
mkdir -p "$OUT"
for v in p2p nccl; do
  if [ $v = nccl ]; then export RB_COMM_NCCL_ONLY=1; else unset RB_COMM_NCCL_ONLY; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"; echo "bench $v exit $?" | tee -a "$OUT/summary.txt"
  python -c "
import json
d=json.loads(open('$OUT/bench_$v.json').read().strip().splitlines()[-1])
print('$v', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), d['allreduce'], json.dumps(d.get('strong_scaling'))[:600])" | tee -a "$OUT/summary.txt"
done
cat "$OUT/summary.txt"
