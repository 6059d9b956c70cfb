#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass8
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu -k "duo or time_parallel or cfg5 or nofilter" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" >> "$OUT/summary.txt"
for w in 7 10; do RB_TP_WARPS_PER_SM=$w timeout 300 python tools/bench_configs.py tp > "$OUT/tp_w$w.jsonl" 2> "$OUT/tp_w$w.err"; done
timeout 300 python tools/bench_configs.py cfg5big > "$OUT/cfg5big.jsonl" 2> "$OUT/cfg5big.err"
cat "$OUT/summary.txt"
