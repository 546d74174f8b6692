#!/bin/bash
# Runs the stand-alone kernel self-tests on the GPU box, one process per case with its own timeout,
# so a deadlocked kernel cannot hang the whole call. Usage: scripts/gpu_selftest.sh [case ...]
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
BIN=pytorch_models_b200/b200enc_selftest
LOG=gpurun_out/selftest.log
: > "$LOG"
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv >> "$LOG" 2>&1
cases=("$@")
if [ ${#cases[@]} -eq 0 ]; then mapfile -t cases < <($BIN list); fi
fail=0
for c in "${cases[@]}"; do
  echo "=== $c" >> "$LOG"
  timeout ${CASE_TIMEOUT:-120} $BIN "$c" >> "$LOG" 2>&1
  rc=$?
  echo "=== $c rc=$rc" >> "$LOG"
  if [ $rc -ne 0 ]; then fail=$((fail+1)); fi
  if [ $rc -eq 124 ]; then echo "TIMEOUT (hang) in $c" >> "$LOG"; fi
done
echo "failed cases: $fail" >> "$LOG"
grep -E "^\s+\[|time |=== .* rc=|failed cases|TIMEOUT|error" "$LOG" | tail -80
exit 0
