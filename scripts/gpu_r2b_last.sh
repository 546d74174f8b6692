#!/bin/bash
# Last evidence refresh of the round on one GPU: GPU test log, ncu --set full of one C3 and one C4 encoder layer (the
# shipped streaming attention kernel in situ). Every profiled command first exits 0 without ncu.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest_gpu.log
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --graph off --prewarm 0"
NCU="ncu --clock-control none"
for c in c3 c4; do
  $BENCH --config $c > gpurun_out/ncu_plain_$c.log 2>&1 &&
  B200_PROFILE_STEP=1 $NCU --profile-from-start off --set full -k regex:"gemm_bf16|attention" -s 8 -c 5 -f -o gpurun_out/r2_${c}_layer $BENCH --config $c > gpurun_out/ncu_f_$c.log 2>&1
  echo "$c layer rc=$?"
  sleep 2
done
python scripts/ncu_summary_r2.py gpurun_out/r2_ncu_summaries > gpurun_out/r2_ncu_summary.log 2>&1
echo "summaries rc=$?"
rm -f gpurun_out/r2_c3_layer.ncu-rep gpurun_out/r2_c4_layer.ncu-rep
