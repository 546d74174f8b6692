#!/bin/bash
# Round 2, second session: launch plans (tests + CPU enqueue time), fused MMA issue in the attention kernels (A/B against
# the previous issue order, event trace of the streaming kernel at L = 1500), optionally the whole GPU suite + bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_attn_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; timeout 120 pytorch_models_b200/$1/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
run . attn:all
run . attn:fault
for d in . ab_nofuse . ab_nofuse; do
  [ -x pytorch_models_b200/$d/b200enc_selftest ] || continue
  for c in perf_vitb_b1024 perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500; do run $d attn:$c; done
done
grep -E "^=== \[.*\] attn:perf|time |TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-12s %-26s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
timeout 120 pytorch_models_b200/b200enc_trace attn:trace 4 20 1500 > gpurun_out/r2b_trace_l1500.txt 2>&1
echo "trace rc=$? lines=$(wc -l < gpurun_out/r2b_trace_l1500.txt)"
timeout 120 pytorch_models_b200/b200enc_trace attn:trace 128 12 197 > gpurun_out/r2b_trace_l197.txt 2>&1
if [ "${SKIP_PLANS:-0}" != "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_plans.py -x -q > gpurun_out/r2b_pytest_plans.log 2>&1
  echo "pytest plans rc=$?"; tail -12 gpurun_out/r2b_pytest_plans.log
  for b in 128; do timeout 300 python scripts/gpu_graph_check.py $b > gpurun_out/r2b_graph_check_b$b.txt 2>&1; cat gpurun_out/r2b_graph_check_b$b.txt | grep -v Warn; done
fi
if [ "${FULL:-0}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
fi
if [ "${BENCH:-0}" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
  echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
fi
