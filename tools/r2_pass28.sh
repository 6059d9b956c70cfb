#!/usr/bin/env bash
# pass 28: RB_MIX_EXACT_ORDER on k_fused_hot (running sum handed from CTA to CTA), k_siggen with 8 generators per CTA
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass28
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -x -k "exact_order or exact or signal or cfg3" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py exact gen > "$OUT/exact.jsonl" 2> "$OUT/exact.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
cut -c1-300 "$OUT/exact.jsonl"; tail -3 "$OUT/exact.err"
cat "$OUT/summary.txt"
