#!/bin/bash
# A/B of the streaming attention kernels on one box: shipped kernel (.) against the v6 variants built under
# pytorch_models_b200/ab_v6* (see DESIGN.md section 3.2 for what each one is).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_v6_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; timeout 120 pytorch_models_b200/$1/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
VARS=$(cd pytorch_models_b200 && ls -d ab_v6* 2>/dev/null)
for d in $VARS; do run $d attn:all; done
for d in . $VARS . $VARS; do
  for c in ${CASES:-perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500}; do run $d attn:$c; done
done
grep -E "^=== \[.*\] attn:perf|TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-12s %-26s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}' | tail -n +$((1))
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
