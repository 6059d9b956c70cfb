#!/usr/bin/env bash
# pass 31: A/B of the default k_fused_hot launch before / after the exact-order chain went into the kernel (same box, interleaved)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass31
mkdir -p "$OUT"
for i in 1 2 3; do
  for v in prechain now; do
    if [ $v = prechain ]; then export RODIO_B200_LIB=$PWD/rodio_b200/librodio_b200_prechain.so; else unset RODIO_B200_LIB; fi
    timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'variant':'$v','run':$i,'ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac'],'clocks':d['clocks']}))" | tee -a "$OUT/ab.jsonl"
  done
done
unset RODIO_B200_LIB
timeout 600 python -m pytest tests -q -m gpu -x -k "exact_order or cfg3_bench or smoke or nofilter" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -3 "$OUT/pytest_gpu.log"
timeout 300 python tools/bench_configs.py exact > "$OUT/exact.jsonl" 2>/dev/null; cut -c1-200 "$OUT/exact.jsonl"
