#!/bin/bash
# timing experiments: B200_DEBUG_FLAGS values given as arguments (0 = normal)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
BIN=pytorch_models_b200/b200enc_selftest
for dbg in "$@"; do
  for c in perf_qkv perf_out perf_fc1 perf_fc2; do
    echo "== dbg=$dbg $c"; B200_DEBUG_FLAGS=$dbg timeout 120 $BIN linear:$c 2>&1 | grep -E "time|FAIL" | head -3
  done
done
