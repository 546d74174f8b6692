#!/bin/bash
# Round 2, second session: launch plans (tests + CPU enqueue time), row-maximum A/B of the streaming attention kernel,
# event trace of the streaming kernel at L = 1500. Outputs under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plans.py -x -q > gpurun_out/r2b_pytest_plans.log 2>&1
echo "pytest plans rc=$?"; tail -12 gpurun_out/r2b_pytest_plans.log
for b in 128 1024; do timeout 300 python scripts/gpu_graph_check.py $b > gpurun_out/r2b_graph_check_b$b.txt 2>&1; cat gpurun_out/r2b_graph_check_b$b.txt | grep -v Warn; done
OUT=gpurun_out/r2b_attn_ab.txt
: > $OUT
for d in . ab_max4 ab_max8; do
  bin=pytorch_models_b200/$d/b200enc_selftest
  [ -x "$bin" ] || continue
  for c in attn:l576_tmem attn:l1370_tmem attn:causal_l448 attn:l1500_wide attn:perf_siglip_b256 attn:perf_dinov2_b128 attn:perf_whisper_b64 attn:perf_causal_1500 attn:perf_siglip_b256 attn:perf_whisper_b64; do
    echo "=== [$d] $c" >> $OUT
    timeout 120 $bin $c >> $OUT 2>&1
    echo "=== [$d] $c rc=$?" >> $OUT
  done
done
grep -E "^=== \[.*\] attn:perf|time " $OUT | grep -v "rc=" | paste - - | awk '{printf "%-10s %-26s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
timeout 120 pytorch_models_b200/b200enc_trace attn:trace 4 20 1500 > gpurun_out/r2b_trace_l1500.txt 2>&1
echo "trace rc=$? lines=$(wc -l < gpurun_out/r2b_trace_l1500.txt)"
if [ "${FULL:-0}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
fi
