#!/usr/bin/env bash
# Round 2, fifth device pass: k_fused_duo_split (producer / consumer warps) against the single-warp form; ncu of k_fused_fx.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass5
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu > "$OUT/pytest_geo.log" 2>&1; echo "pytest (split) exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_geo.log" >> "$OUT/summary.txt"
RB_DUO_SPLIT=0 timeout 900 python -m pytest tests/test_bench_geometries_gpu.py -q -m gpu -k "duo or time_parallel or cfg5 or nofilter" > "$OUT/pytest_geo_unsplit.log" 2>&1; echo "pytest (single warp) exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_geo_unsplit.log" >> "$OUT/summary.txt"
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 > "$OUT/bench_split.json" 2> "$OUT/bench_split.err"; echo "bench split exit $?" | tee -a "$OUT/summary.txt"
RB_DUO_SPLIT=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 > "$OUT/bench_unsplit.json" 2> "$OUT/bench_unsplit.err"; echo "bench unsplit exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_fx -c 1 -o "$OUT/fx_full" python tools/bench_configs.py cfg4 > "$OUT/ncu_fx.log" 2>&1
echo "ncu fx exit $?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_duo_split -s 3 -c 1 -o "$OUT/duo_split_full" \
    python bench.py --streams 65536 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_duo_split.log" 2>&1
echo "ncu duo split exit $?" | tee -a "$OUT/summary.txt"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_bench_geometries_gpu.py -x -q -m gpu -k "duo_kernel_shapes" > "$OUT/racecheck.log" 2>&1; echo "racecheck exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
