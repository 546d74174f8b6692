#!/bin/bash
# Round 2: attention kernel A/B on one box. New kernel = pytorch_models_b200/b200enc_selftest, round-1 kernel =
# pytorch_models_b200/ab_v5/b200enc_selftest (built from the round-1 commit), extra variants = pytorch_models_b200/ab_*/.
# Every case runs under its own timeout so that a protocol slip cannot hold the box.
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
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $OUT
# correctness of the new kernel first (all small cases + tails + causal)
for c in l64_tmem l128_tmem l197_tmem l576_tmem l1370_tmem cross_q1 l16 causal_l16 causal_l197 causal_l448 causal_l700 \
         causal_many l130 l256 cross_q300_kv200 cross_q1_kv200 many_short many_items l1500_wide causal_l1100; do
  run new attn:$c
done
PERF="perf_vitb perf_siglip perf_whisper perf_causal_1500 perf_vitb_b1024 perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64"
for rep in 1; do
  for d in new ab_v5 $(cd pytorch_models_b200 && ls -d ab_* 2>/dev/null | grep -v '^ab_v5$'); do
    for c in $PERF; do
      run $d attn:$c
    done
  done
done
# GEMM epilogue with the coalesced residual loads: correctness, then old vs new on the residual shapes
for c in embed_like residual multi_tile tails_tma many_tiles fold_gelu; do run new linear:$c; done
for d in new ab_v5; do
  for c in perf_out perf_fc2 perf_out_big perf_fc2_big perf_qkv_big perf_fc1_big; do run $d linear:$c; done
done
if [ -x pytorch_models_b200/b200enc_trace ]; then
  timeout 120 pytorch_models_b200/b200enc_trace attn:trace 8 20 1500 > gpurun_out/r2_trace_l1500.txt 2>&1
  timeout 120 pytorch_models_b200/b200enc_trace attn:trace 128 12 197 > gpurun_out/r2_trace_l197.txt 2>&1
fi
grep -E "^=== |time |FAIL|MISMATCH|error" $OUT | grep -v "rc=0" | head -150
echo "--- failures:"; grep -c "rc=[1-9]" $OUT
