#!/usr/bin/env bash
# pass 34 (2 GPUs): the communicator tests (one process / two GPUs through the C ABI; two processes over cudaIpc; both also on NCCL)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_pass34
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -k "comm" -s > "$OUT/pytest_comm.log" 2>&1; echo "pytest comm exit $?" | tee -a "$OUT/summary.txt"
grep -E "transport|passed|failed|differ|GPU [0-9]|Error" "$OUT/pytest_comm.log" | tail -30 >> "$OUT/summary.txt"
cat "$OUT/summary.txt"
