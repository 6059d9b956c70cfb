#!/usr/bin/env bash
# Round 2, fourth device pass: 3-slot rings / 1-warp CTAs / 12 warps of segment rows per SM for k_fused_duo; k_fused_fx (cfg4).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass4
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu > "$OUT/pytest_geo.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_geo.log" >> "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
for w in 8 16; do
  RB_TP_WARPS_PER_SM=$w timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 > "$OUT/bench_tpw$w.json" 2> "$OUT/bench_tpw$w.err"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/launches_cfg4.csv" python tools/bench_configs.py cfg4 > /dev/null 2>&1
cat "$OUT/summary.txt"
