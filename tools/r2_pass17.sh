#!/usr/bin/env bash
# pass 17: state after the container restore -- whole GPU suite, default bench line, filter-free chains on k_lerp_mix, ncu of k_lerp_mix
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass17
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py tp nofilter > "$OUT/nofilter.jsonl" 2> "$OUT/nofilter.err"
RB_NO_LERPMIX=1 timeout 300 python tools/bench_configs.py tp > "$OUT/nofilter_segments.jsonl" 2>> "$OUT/nofilter.err"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lerp_mix -s 3 -c 1 -o "$OUT/lerpmix_full" python tools/bench_configs.py tp > "$OUT/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/nofilter.jsonl" "$OUT/nofilter_segments.jsonl" | cut -c1-300
cat "$OUT/summary.txt"
