#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass15
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py cfg4 limit > "$OUT/fx.jsonl" 2>&1
cat "$OUT/summary.txt"
