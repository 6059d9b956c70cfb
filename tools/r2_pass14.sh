#!/usr/bin/env bash
# Round 2: k_fused_hot stage A without the per-sample range guard (inputs classified per upload) and with the one-product feed-forward.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass14
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py cfg3_stereo > "$OUT/stereo.jsonl" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_hot -s 3 -c 1 -o "$OUT/hot_full" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_hot.log" 2>&1
echo "ncu hot exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
