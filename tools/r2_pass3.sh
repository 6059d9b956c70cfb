#!/usr/bin/env bash
# Round 2, third device pass: k_fused_duo as a three-stage software pipeline; segments for filter-free chains.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass3
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu -x > "$OUT/pytest_geo.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -8 "$OUT/pytest_geo.log" >> "$OUT/summary.txt"
timeout 900 python tools/bench_configs.py duo > "$OUT/duo_sweep.jsonl" 2> "$OUT/duo_sweep.err"; echo "duo sweep exit $?" | tee -a "$OUT/summary.txt"
RB_SEGMENTS_FROM=1000000 timeout 300 python tools/bench_configs.py nofilter > "$OUT/nofilter_hot.jsonl" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_duo -s 3 -c 1 -o "$OUT/duo_full" \
    python bench.py --streams 65536 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_duo.log" 2>&1
echo "ncu duo exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
