#!/bin/bash
# ncu evidence for the bench configuration (one warmed-up forward = 65 launches):
#   launches.csv          every launch with its device time (cold-cache, serialised: compare SHARES)
#   prof_layer.ncu-rep    --set full capture of one encoder layer (QKV gemm, attention, out_proj, FC1, FC2)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
B200_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
B200_PROFILE_STEP=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_kernel" -s 21 -c 5 -o gpurun_out/prof_layer $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu_full.log
