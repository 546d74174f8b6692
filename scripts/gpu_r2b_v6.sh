#!/bin/bash
# The shipped streaming attention kernel (attention_v6.cuh) against the previous one (attention.cuh, B200ENC_ATTN_V5=1)
# on one box: self-tests (fault injection included), timings, then the GPU test suite.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_v6_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; env $3 timeout 120 pytorch_models_b200/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
run v6 attn:all
run v6 attn:fault
run v5 attn:all B200ENC_ATTN_V5=1
for rep in 1 2; do
  for c in perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500 perf_vitb_b1024; do run v6 attn:$c; run v5 attn:$c B200ENC_ATTN_V5=1; done
done
grep -E "^=== \[.*\] attn:perf|TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-6s %-26s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}' | tail -20
grep -A8 "fault injection" $OUT | head -12
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu.log
fi
