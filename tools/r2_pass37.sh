#!/usr/bin/env bash
# pass 37: AGC and limiter over a source whose format changes (from_iter), memcheck (the AGC movers on unaligned runs), cfg4 general path unchanged
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass37
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -k "from_iter or agc or cfg4 or limit or pause" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "from_iter or pause" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/memcheck.log" >> "$OUT/summary.txt"
cat "$OUT/summary.txt"
