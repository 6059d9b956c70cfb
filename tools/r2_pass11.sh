#!/usr/bin/env bash
# Round 2: full GPU suite, the bench line, final ncu captures (k_fused_duo at 65 536 streams, k_fused_fx), launch list of the bench.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass11
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench reference exit $?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_duo -s 3 -c 1 -o "$OUT/duo_full" \
    python bench.py --streams 65536 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_duo.log" 2>&1
echo "ncu duo exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_fx -c 1 -o "$OUT/fx_full" python tools/bench_configs.py cfg4 > "$OUT/ncu_fx.log" 2>&1
echo "ncu fx exit $?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "launch list exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
