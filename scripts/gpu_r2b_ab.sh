#!/bin/bash
# Same-box A/B of compile-time variants of the attention kernels: the shipped library (.) against every
# pytorch_models_b200/ab_* directory (make -C pytorch_models_b200/csrc variant NAME=ab_x DEFS=-D...).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; timeout 120 pytorch_models_b200/$1/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
VARS=$(cd pytorch_models_b200 && ls -d ab_* 2>/dev/null)
for d in . $VARS . $VARS; do
  for c in ${CASES:-perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500 perf_vitb_b1024}; do run $d attn:$c; done
done
grep -E "^=== \[.*\] attn:perf|TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-10s %-24s %s %s %s %s\n", $2, $3, $5, $6, $7, $8}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
