#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass12
mkdir -p "$OUT"
for rep in 1 2; do
  for v in "" _scalarb _order1 _both; do
    RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200$v.so" timeout 300 python tools/bench_configs.py cfg5big tp > "$OUT/ab${v}_$rep.jsonl" 2> "$OUT/ab${v}_$rep.err"
  done
done
RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200_scalarb.so" timeout 600 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu -k "duo or time_parallel or cfg5 or nofilter" > "$OUT/pytest_scalarb.log" 2>&1; echo "pytest scalarb exit $?" | tee -a "$OUT/summary.txt"
echo done | tee -a "$OUT/summary.txt"
