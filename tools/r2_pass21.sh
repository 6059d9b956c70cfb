#!/usr/bin/env bash
# pass 21: k_lerp_mix with cp.async-staged windows (parity, memcheck, timing, ncu)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass21
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -x -k "lerp or nofilter" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "lerp_mix" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/memcheck.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py tp nofilter > "$OUT/nofilter.jsonl" 2> "$OUT/nofilter.err"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lerp_mix -s 3 -c 1 -o "$OUT/lerpmix_full" python tools/bench_configs.py tp > "$OUT/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
cut -c1-260 "$OUT/nofilter.jsonl"
cat "$OUT/summary.txt"
