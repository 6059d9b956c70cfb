#!/usr/bin/env bash
# Measurement pass for k_fused_lanes / rb_session_* on a B200 (round 1 only had two ten-second runs:
# all parity checks green, two wall-clock timings).
# One gpurun call:   gpurun --timeout 2400 -- 'bash tools/first_device_pass.sh'   (about 25-30 min; the steps are ordered by what they are worth)
# Everything lands in gpurun_out/lanes_first/; every step is bounded by `timeout` so a hang cannot eat the box.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/lanes_first
mkdir -p "$OUT"
# (the tests of the kernel and the sessions are part of the normal -m gpu suite since the first device pass)

# 1. the parity tests of the kernel and the sessions
timeout 600 python -m pytest tests -q -m gpu -k "lanes or session" > "$OUT/pytest_lanes.log" 2>&1
echo "pytest lanes/session exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_lanes.log" >> "$OUT/summary.txt"

# 2. throughput: large end of the cfg5 sweep with and without the flag, then the streaming case
timeout 600 python tools/bench_configs.py lanes > "$OUT/lanes_sweep.jsonl" 2> "$OUT/lanes_sweep.err"
echo "lanes sweep exit $?" | tee -a "$OUT/summary.txt"
timeout 300 python tools/bench_configs.py session > "$OUT/session.jsonl" 2> "$OUT/session.err"
echo "session bench exit $?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --streams 65536 --seconds 1 --flags 16 --steps 5 --warmup 3 --no-cpu-baseline \
    > "$OUT/bench_65536_lanes.json" 2> "$OUT/bench_65536_lanes.err"
echo "bench 65536 lanes exit $?" | tee -a "$OUT/summary.txt"

# 3a. the shapes added after the first device runs, and the stereo ring geometry A/B (64-byte against 128-byte chunks per lane)
timeout 400 python tools/bench_configs.py lanes_shapes > "$OUT/lanes_shapes.jsonl" 2> "$OUT/lanes_shapes.err"
echo "lanes shapes exit $?" | tee -a "$OUT/summary.txt"
timeout 300 python -c "from rodio_b200 import build; build.build(force=True, extra_flags=['-DRB_LANES_STEREO_CHW=32'], out='rodio_b200/librodio_b200_chw32.so')" \
    > "$OUT/build_chw32.log" 2>&1 && \
RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200_chw32.so" timeout 400 python tools/bench_configs.py lanes_shapes \
    > "$OUT/lanes_shapes_chw32.jsonl" 2> "$OUT/lanes_shapes_chw32.err"
echo "lanes shapes (128-byte stereo chunks) exit $?" | tee -a "$OUT/summary.txt"

# 3b. ring with three slots and one chunk of look-ahead (68 instead of 84 words per lane: 24 instead of 20 warps per SM)
timeout 300 python -c "from rodio_b200 import build; build.build(force=True, extra_flags=['-DRB_LANES_UP_SLOTS=3'], out='rodio_b200/librodio_b200_slots3.so')" \
    > "$OUT/build_slots3.log" 2>&1 && \
RODIO_B200_LIB="$PWD/rodio_b200/librodio_b200_slots3.so" timeout 600 python tools/bench_configs.py lanes lanes_shapes \
    > "$OUT/lanes_slots3.jsonl" 2> "$OUT/lanes_slots3.err"
echo "lanes sweep + shapes (3 ring slots) exit $?" | tee -a "$OUT/summary.txt"

# 4. launch list and one full capture of the new kernel (numbers printed under ncu are never bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/launches_lanes.csv" \
    python bench.py --streams 16384 --seconds 1 --flags 16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused_lanes -s 3 -c 1 -o "$OUT/lanes_full" \
    python bench.py --streams 16384 --seconds 1 --flags 16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "ncu done" | tee -a "$OUT/summary.txt"
# 5. memory errors, on the smallest cases (compute-sanitizer is slow: keep it to a handful of tests)
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 \
    python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "lanes_single or lanes_ragged or session_rejects" \
    > "$OUT/memcheck.log" 2>&1
echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/summary.txt"
