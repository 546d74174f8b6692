#!/bin/bash
# A/B of the streaming attention kernels on one box: shipped (fused issue), v6 with the lean issuer (bounded / plain
# softmax waits), v6 as first written; then the launch-plan timing at several queue depths.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_v6_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; timeout 120 pytorch_models_b200/$1/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
for d in ab_v6l ab_v6lp; do run $d attn:all; done
for d in . ab_v6l ab_v6lp ab_v6o . ab_v6l; do
  [ -x pytorch_models_b200/$d/b200enc_selftest ] || continue
  for c in perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500; do run $d attn:$c; done
done
grep -E "^=== \[.*\] attn:perf|TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-12s %-26s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
timeout 600 python -m pytest tests/test_gpu_plans.py -x -q > gpurun_out/r2b_pytest_plans.log 2>&1
echo "pytest plans rc=$?"; tail -4 gpurun_out/r2b_pytest_plans.log
timeout 300 python scripts/gpu_graph_check.py 128 > gpurun_out/r2b_graph_check_b128.txt 2>&1; grep -v Warn gpurun_out/r2b_graph_check_b128.txt
