#!/bin/bash
# Refresh of the round-2 ncu evidence after the im2col-free patch embedding, programmatic dependent launch and the
# LayerNorm rewrite: C2 launch list, C2 head (--set full), row kernels (--set full). Summaries -> gpurun_out/r2_ncu_summaries.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
NCU="ncu --clock-control none"
$BENCH > gpurun_out/ncu_plain_c2.log 2>&1 &&
B200_PROFILE_STEP=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_c2.csv $BENCH > gpurun_out/ncu_l_c2.log 2>&1
echo "c2 launch list rc=$?"
$BENCH > gpurun_out/ncu_plain_c2b.log 2>&1 &&
B200_PROFILE_STEP=1 $NCU --profile-from-start off --set full -c 8 -f -o gpurun_out/r2_c2_head $BENCH > gpurun_out/ncu_f_c2.log 2>&1
echo "c2 head rc=$?"
ROW_ITERS=1 python tests/tools/gpu_row_kernels.py > gpurun_out/ncu_plain_rows.log 2>&1 &&
ROW_ITERS=1 $NCU --set full -k regex:"layernorm|mean_tokens|patchify|cls_rows|time_rows|embed_rows|logmel" -f -o gpurun_out/r2_rows python tests/tools/gpu_row_kernels.py > gpurun_out/ncu_f_rows.log 2>&1
echo "rows rc=$?"
python scripts/ncu_summary_r2.py gpurun_out/r2_ncu_summaries > gpurun_out/r2_ncu_summary.log 2>&1
echo "summaries rc=$?"
rm -f gpurun_out/r2_c2_head.ncu-rep gpurun_out/r2_rows.ncu-rep
