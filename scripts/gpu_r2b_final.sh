#!/bin/bash
# Final evidence of the second round-2 session (one GPU): bench lines (C2 + C3-C5), ncu launch lists at 1024 and 128
# images, ncu --set full of the streaming attention kernel at L = 1500, library baselines, parity report, CPU enqueue
# time per forward, GPU test suite. Every profiled command first exits 0 without ncu. Outputs under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err
  echo "bench $c rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2b_bench_$c.json'));print(d['value'],d['ms_per_step'],d['roofline']['step_breakdown_ms'],d['roofline_attention']['tflops'])"
done
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --graph off"
NCU="ncu --clock-control none"
$BENCH > gpurun_out/ncu_plain_c2.log 2>&1 &&
B200_PROFILE_STEP=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_c2.csv $BENCH > gpurun_out/ncu_l_c2.log 2>&1
echo "c2 launch list rc=$?"
$BENCH --batch 128 > gpurun_out/ncu_plain_c2_b128.log 2>&1 &&
B200_PROFILE_STEP=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2b_launches_c2_b128.csv $BENCH --batch 128 > gpurun_out/ncu_l_c2_b128.log 2>&1
echo "c2 b128 launch list rc=$?"
ST=pytorch_models_b200/b200enc_selftest
$ST attn:perf_whisper_b64 > gpurun_out/ncu_plain_a1500.log 2>&1 &&
$NCU --set full --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/r2_attn_l1500 $ST attn:perf_whisper_b64 > gpurun_out/ncu_f_a1500.log 2>&1
echo "attn l1500 rc=$?"
python scripts/ncu_summary_r2.py gpurun_out/r2_ncu_summaries > gpurun_out/r2_ncu_summary.log 2>&1
echo "summaries rc=$?"
ncu -i gpurun_out/r2_attn_l1500.ncu-rep --page source --csv > gpurun_out/r2_attn_l1500.source.csv 2>/dev/null
rm -f gpurun_out/r2_attn_l1500.ncu-rep
timeout 900 python tests/tools/gpu_library_baseline.py > gpurun_out/r2b_library_baseline.log 2>&1
echo "library baseline rc=$?"; grep -E "^attention|^ours|^torch" gpurun_out/r2b_library_baseline.log | cut -c1-400
timeout 900 python tests/tools/gpu_parity_report.py > gpurun_out/r2b_parity_report.log 2>&1
echo "parity report rc=$?"; tail -6 gpurun_out/r2b_parity_report.log
for b in 128 1024; do timeout 300 python scripts/gpu_graph_check.py $b > gpurun_out/r2b_graph_check_b$b.txt 2>&1; grep -v Warn gpurun_out/r2b_graph_check_b$b.txt; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest_gpu.log
