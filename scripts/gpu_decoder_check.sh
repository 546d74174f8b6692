#!/bin/bash
# Decoder widening check: causal / tanh-GELU self-tests, then the whole GPU suite without -x so every failure shows.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
bash scripts/gpu_selftest.sh attn:causal_l16 attn:causal_l197 attn:causal_l448 attn:causal_l700 attn:causal_many \
  attn:perf_causal_1500 attn:perf_whisper linear:fold_gelu_tanh linear:gelu_tanh linear:fold_gelu
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -40
