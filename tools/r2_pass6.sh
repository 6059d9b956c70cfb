#!/usr/bin/env bash
# Round 2, sixth device pass: k_fused_fx with eight workers; segment rows per SM for the time-parallel / filter-free plans;
# integer PCM in front of the fused kernels.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass6
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py cfg4 > "$OUT/cfg4.jsonl" 2>&1
for w in 8 12 16 24; do
  RB_TP_WARPS_PER_SM=$w timeout 300 python tools/bench_configs.py tp > "$OUT/tp_w$w.jsonl" 2> "$OUT/tp_w$w.err"
done
timeout 600 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
