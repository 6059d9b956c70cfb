#!/usr/bin/env bash
# pass 22: whole GPU suite with the new rows (signal generators, mix / crossfade, from_iter), memcheck of the new general-path kernels
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass22
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -25 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "from_iter or crossfade or mix_of_two or generators_bit_exact or generator_through" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/memcheck.log" >> "$OUT/summary.txt"
cat "$OUT/summary.txt"
