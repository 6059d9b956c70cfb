#!/usr/bin/env bash
# The commands behind profiles/ (run from the repo root on a B200 box, e.g. through `gpurun -- 'bash tools/reproduce_profiles.sh'`).
# Everything lands in gpurun_out/; copy what should be tracked into profiles/ and run tools/ncu_summary.py here.
set -u
mkdir -p gpurun_out
# 1. bench lines (un-profiled)
timeout 300 python bench.py > gpurun_out/r1_bench_default.json
timeout 300 python bench.py --impl reference > gpurun_out/r1_bench_reference.json
# 2. the other configurations, the batch sweep and the CPU restatement on this box's cores
timeout 400 python tools/bench_configs.py cfg2 cfg4 cfg3_stereo nofilter cfg5 cpu > gpurun_out/r1_configs.jsonl
# 3. launch lists (shares only: cold caches, serialised)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r1_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_(agc|echo|chanvol|mix|uniform|amplify|convert)" -c 36 --csv \
    --log-file gpurun_out/r1_launches_cfg4.csv python tools/bench_configs.py cfg4 > /dev/null
# 4. one full capture of the dominant kernel (then: python tools/ncu_summary.py gpurun_out/prof_fused_r1_final.ncu-rep)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_hot -s 3 -c 1 \
    -o gpurun_out/prof_fused_r1_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e
# 5. per-role timing of the kernel and the chain-latency microbenchmark
timeout 120 python tools/hot_timing.py 4096 --skips > gpurun_out/r1_hot_timing.txt
(cd tools/microbench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false chain_latency.cu -o chain_latency) \
    && timeout 60 tools/microbench/chain_latency > gpurun_out/r1_chain_latency.txt
# 6. two GPUs (gpurun --gpus 2):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
#       bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r1_bench_2gpu.json

# --- k_fused_lanes / sessions: round 1 had two ten-second runs (profiles/r1_lanes_*.log: python tools/lanes_quick_check.py
#     --time, --time-big --lanes-only, and the pytest line below); the full measurement plan is tools/first_device_pass.sh ---
# python -m pytest tests -m gpu -x -q -k "lanes or session"      (round 1: with RB_TEST_LANES=1, the gate of the time)
# python tools/bench_configs.py lanes > gpurun_out/lanes_sweep.jsonl
# ncu --set full --clock-control none --import-source on -k regex:k_fused_lanes -s 3 -c 1 -o gpurun_out/lanes_full python tools/bench_configs.py lanes

# --- round 2: one script per GPU call, tools/r2_pass1.sh ... r2_pass37.sh (each writes gpurun_out/r2_passNN/; profiles/README.md says
#     which file came from which pass).  The ones behind the second half of the round:
#   bash tools/r2_pass17.sh      whole GPU suite + bench line + k_lerp_mix (first version) with a full ncu capture
#   bash tools/r2_pass21.sh      k_lerp_mix as it is now: parity, memcheck, timing, ncu
#   bash tools/r2_pass28.sh      RB_MIX_EXACT_ORDER on k_fused_hot (tagged-pair chain): parity + python tools/bench_configs.py exact gen
#   bash tools/r2_pass29.sh      final-state evidence: suite, smoke(), bench, reference arm, launch list, ncu of the chain kernel and k_siggen
#   bash tools/r2_pass31.sh      A/B of the default k_fused_hot launch against the kernel before the chain (needs a pre-chain build)
#   gpurun --gpus 2 -- 'bash tools/r2_pass33.sh'; gpurun --gpus 2 -- 'bash tools/r2_pass34.sh'     the peer-memory exchange: tests, A/B against NCCL
#   gpurun --gpus 4 -- 'bash tools/r2_pass35.sh 4'; gpurun --gpus 8 -- 'bash tools/r2_pass35.sh 8'
#   bash tools/r2_pass36.sh      whole GPU suite + smoke() + bench line with the final code
#   bash tools/r2_pass37.sh      AGC / limiter over from_iter sources, memcheck
