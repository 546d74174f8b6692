#!/bin/bash
# Re-capture of the C2 head only (patch rows, patch-embed GEMM, class-token rows, the five launches of layer 0) and
# regeneration of the tracked summaries; see scripts/gpu_r2_ncu.sh for the full evidence run.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$BENCH > gpurun_out/ncu_plain_c2b.log 2>&1 &&
B200_PROFILE_STEP=1 ncu --clock-control none --profile-from-start off --set full -c 9 -f -o gpurun_out/r2_c2_head $BENCH > gpurun_out/ncu_f_c2.log 2>&1
echo "c2 head rc=$?"
python scripts/ncu_summary_r2.py gpurun_out/r2_ncu_summaries > gpurun_out/r2_ncu_summary.log 2>&1
echo "summaries rc=$?"
rm -f gpurun_out/r2_c2_head.ncu-rep
