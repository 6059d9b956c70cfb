#!/usr/bin/env bash
# pass 41: whole GPU suite with the final code (pause adapter, C++ mirror tests, fixtures), smoke(), default bench line
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass41
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
