#!/bin/bash
# Attention kernel check: every attn:* self-test case (correctness + perf), then the event trace at two shapes.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
mapfile -t cases < <(pytorch_models_b200/b200enc_selftest list | grep "^attn:")
bash scripts/gpu_selftest.sh "${cases[@]}"
if [ -x pytorch_models_b200/b200enc_trace ]; then
  timeout 60 pytorch_models_b200/b200enc_trace attn:trace 128 12 197 > gpurun_out/trace_l197.txt 2>&1
  timeout 60 pytorch_models_b200/b200enc_trace attn:trace 8 20 1500 > gpurun_out/trace_l1500.txt 2>&1
fi
