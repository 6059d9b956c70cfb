#!/usr/bin/env bash
# A/B of the k_fused_duo launch geometry: warps per CTA (2 / 1) and the ring rotation, every variant twice, interleaved; SM clocks sampled.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass9
mkdir -p "$OUT"
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 250 > "$OUT/smi.csv" 2>/dev/null &
SMI=$!
for rep in 1 2; do
  for v in "" _w1 _rot _w1rot; do
    RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200$v.so" RB_TP_WARPS_PER_SM=7 timeout 300 python tools/bench_configs.py cfg5big tp > "$OUT/ab${v}_$rep.jsonl" 2> "$OUT/ab${v}_$rep.err"
  done
done
kill $SMI
echo done | tee "$OUT/summary.txt"
