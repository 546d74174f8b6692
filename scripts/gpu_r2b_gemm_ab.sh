#!/bin/bash
# Same-box A/B of GEMM variants: self-test correctness + stand-alone timings, then the bench step (sustained, power-capped
# regime) with each library (B200ENC_LIB), alternating twice.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r2b_gemm_ab.txt
: > $OUT
run() { echo "=== [$1] $2" >> $OUT; timeout 300 pytorch_models_b200/$1/b200enc_selftest $2 >> $OUT 2>&1; echo "=== [$1] $2 rc=$?" >> $OUT; }
VARS=$(cd pytorch_models_b200 && ls -d ab_* 2>/dev/null)
run . linear:all
run . patch:all
for d in . $VARS . $VARS; do
  for c in perf_qkv_big perf_out_big perf_fc1_big perf_fc2_big; do run $d linear:$c; done
done
grep -E "^=== \[.*\] linear:perf|TFLOP" $OUT | grep -v "rc=" | paste - - | awk '{printf "%-8s %-22s %s %s %s %s %s\n", $2, $3, $5, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
for rep in 1 2; do
  for d in . $VARS; do
    lib=$PWD/pytorch_models_b200/$d/libb200enc.so
    B200ENC_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2b_gemm_ab_bench.json 2> gpurun_out/r2b_gemm_ab_bench.err
    python -c "import json;d=json.load(open('gpurun_out/r2b_gemm_ab_bench.json'));r=d['roofline'];print('[$d] rep $rep', round(d['value']), round(d['ms_per_step'],3), {k:v['tflops'] for k,v in r['by_shape'].items()}, d['clocks']['sm_mhz'])"
  done
done
