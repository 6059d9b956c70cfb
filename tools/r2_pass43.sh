#!/usr/bin/env bash
# pass 43: the ncu launch list of the bench command with the final code (shares only: cold caches, serialised)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass43
mkdir -p "$OUT"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench_default.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1; echo "launch list exit $?"
wc -l "$OUT/launches_bench_default.csv"
