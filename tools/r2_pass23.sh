#!/usr/bin/env bash
# pass 23: from_iter on the device (NaN-insensitive comparison), memcheck of the new general-path kernels, default bench line
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass23
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -k "from_iter or crossfade or mix_of_two or signal" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "from_iter or crossfade or mix_of_two or generators_bit_exact or generator_through" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/memcheck.log" >> "$OUT/summary.txt"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
