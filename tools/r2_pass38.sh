#!/usr/bin/env bash
# pass 38: k_siggen with 4 generators per CTA and 256-sample tiles (parity + timing)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass38
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -k "signal or generator" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -3 "$OUT/pytest_gpu.log"
timeout 300 python tools/bench_configs.py gen > "$OUT/gen.jsonl" 2>/dev/null; cut -c1-220 "$OUT/gen.jsonl"
