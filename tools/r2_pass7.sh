#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass7
mkdir -p "$OUT"
RB_TP_WARPS_PER_SM=12 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_duo -s 3 -c 1 -o "$OUT/tp_full" python tools/bench_configs.py tp > "$OUT/ncu_tp.log" 2>&1
echo "ncu tp exit $?" | tee -a "$OUT/summary.txt"
