#!/usr/bin/env bash
# pass 29: final state -- whole GPU suite, smoke(), default bench line, reference arm, ncu launch list of the bench command,
# full ncu captures of k_fused_hot in chain mode (RB_MIX_EXACT_ORDER) and of k_siggen
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass29
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_gpu.log" >> "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" | tee -a "$OUT/summary.txt"
tail -2 "$OUT/smoke.log" >> "$OUT/summary.txt"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "reference arm exit $?" | tee -a "$OUT/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_default.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1; echo "launch list exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_hot -s 14 -c 1 -o "$OUT/hot_chain_full" python tools/bench_configs.py exact > "$OUT/ncu_chain.log" 2>&1; echo "ncu chain exit $?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_siggen -s 1 -c 1 -o "$OUT/siggen_full" python tools/bench_configs.py gen > "$OUT/ncu_siggen.log" 2>&1; echo "ncu siggen exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
