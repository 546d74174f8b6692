#!/bin/bash
# ncu evidence for one bench configuration: launch list (gpu__time_duration) + one --set full capture of a full layer.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_kernel|layernorm_kernel" -s 98 -c 7 -o gpurun_out/prof_layer $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
