#!/usr/bin/env bash
# Round 2, second device pass: the lane-pair kernel (k_fused_duo) and the time-parallel plan -- parity, sweep, ncu.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass2
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 900 python tools/bench_configs.py duo > "$OUT/duo_sweep.jsonl" 2> "$OUT/duo_sweep.err"; echo "duo sweep exit $?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_duo -s 3 -c 1 -o "$OUT/duo_full" \
    python bench.py --streams 65536 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_duo.log" 2>&1
echo "ncu duo exit $?" | tee -a "$OUT/summary.txt"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bench_geometries_gpu.py -x -q -m gpu -k "duo_kernel_shapes or out_of_phase" \
    > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
