#!/usr/bin/env bash
# pass 40: RB_MIX_EXACT_ORDER on k_fused_fx (cfg4): parity, timing against the default launch, the other fx tests unchanged
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass40
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -x -k "cfg4 or fx_kernel or exact_order or effect_chain" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py exact4 > "$OUT/exact4.jsonl" 2> "$OUT/exact4.err"; cut -c1-230 "$OUT/exact4.jsonl"; tail -2 "$OUT/exact4.err"
cat "$OUT/summary.txt"
