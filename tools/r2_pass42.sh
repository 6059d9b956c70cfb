#!/usr/bin/env bash
# pass 42: the Player mirror with pauses (one new GPU test) + the pause / player tests
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass42
mkdir -p "$OUT"
timeout 300 python -m pytest tests -q -m gpu -k "pause or player" > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
