#!/usr/bin/env bash
# Round 2, first device pass: FP32 issue microbenchmark, the whole GPU suite (incl. the new bench-geometry parity tests),
# the lane-kernel sweep, full ncu captures of k_fused_lanes (65 536 streams) and k_fused_hot<1,false>, and a default bench line.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass1
mkdir -p "$OUT"
timeout 120 tools/microbench/fp32_throughput > "$OUT/fp32_throughput.txt" 2>&1; echo "microbench exit $?" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 600 python tools/bench_configs.py lanes > "$OUT/lanes_sweep.jsonl" 2> "$OUT/lanes_sweep.err"; echo "lanes sweep exit $?" | tee -a "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py nofilter cfg4 > "$OUT/configs.jsonl" 2> "$OUT/configs.err"; echo "configs exit $?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_lanes -s 3 -c 1 -o "$OUT/lanes_full" \
    python bench.py --streams 65536 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_lanes.log" 2>&1
echo "ncu lanes exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_hot -s 3 -c 1 -o "$OUT/hot_nofilter_full" \
    python tools/bench_configs.py nofilter > "$OUT/ncu_hot_nofilter.log" 2>&1
echo "ncu hot nofilter exit $?" | tee -a "$OUT/summary.txt"
timeout 400 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
