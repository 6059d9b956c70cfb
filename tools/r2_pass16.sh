#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass16
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py tp nofilter > "$OUT/nofilter.jsonl" 2> "$OUT/nofilter.err"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lerp_mix -s 3 -c 1 -o "$OUT/lerpmix_full" python tools/bench_configs.py tp > "$OUT/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
