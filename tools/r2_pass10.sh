#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass10
mkdir -p "$OUT"
for rep in 1 2; do
  for v in "" _pass3; do
    RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200$v.so" RB_TP_WARPS_PER_SM=7 timeout 300 python tools/bench_configs.py cfg5big tp > "$OUT/ab${v}_$rep.jsonl" 2> "$OUT/ab${v}_$rep.err"
  done
done
echo done | tee "$OUT/summary.txt"
