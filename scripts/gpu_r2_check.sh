#!/bin/bash
# Round 2 validation on one box: attention A/B (shipped kernel vs round-1 kernel vs the experimental v6), the GPU test
# suite, one bench line. Outputs under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2_attn_ab.txt
: > $OUT
run() {  # dir case
  local bin=pytorch_models_b200/$1/b200enc_selftest
  [ "$1" = "new" ] && bin=pytorch_models_b200/b200enc_selftest
  [ -x "$bin" ] || return
  echo "=== [$1] $2" >> $OUT
  timeout 120 $bin $2 >> $OUT 2>&1
  echo "=== [$1] $2 rc=$?" >> $OUT
}
if [ "${SKIP_AB:-0}" != "1" ]; then
for c in l197_tmem l576_tmem l1370_tmem causal_l448 causal_many many_items l1500_wide fault l197_tmem; do run new attn:$c; done
grep -A6 "fault injection" $OUT
for d in new ab_v5 $(cd pytorch_models_b200 && ls -d ab_* 2>/dev/null | grep -v '^ab_v5$'); do
  for c in perf_vitb perf_vitb_b1024 perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500; do run $d attn:$c; done
  for c in perf_out_big perf_fc2_big perf_qkv_big perf_fc1_big; do run $d linear:$c; done
done
grep -E "^=== \[|time " $OUT | grep -B1 "time" | grep -v "^--" | paste - - | awk '{printf "%-12s %-28s %s %s %s %s\n", $2, $3, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
  echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json
fi
