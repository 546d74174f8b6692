#!/bin/bash
# Same-box A/B of attention variants: `new` (the library build) and every pytorch_models_b200/ab_*/ directory
# (built with `make variant NAME=ab_x DEFS=...`). Correctness cases on `new` first, then the perf cases everywhere.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab.txt
: > $OUT
run() {
  local bin=pytorch_models_b200/$1/b200enc_selftest
  [ "$1" = "new" ] && bin=pytorch_models_b200/b200enc_selftest
  [ -x "$bin" ] || return
  echo "=== [$1] $2" >> $OUT
  timeout 120 $bin $2 >> $OUT 2>&1
  echo "=== [$1] $2 rc=$?" >> $OUT
}
CASES=${CASES:-"perf_vitb_b1024 perf_siglip_b256 perf_dinov2_b128 perf_whisper_b64 perf_causal_1500"}
for c in ${CHECKS:-l197_tmem l576_tmem l1370_tmem causal_l448 causal_many many_items l1500_wide}; do run new attn:$c; done
for rep in 1 2; do
for d in new $(cd pytorch_models_b200 && ls -d ab_* 2>/dev/null); do
  for c in $CASES; do run $d attn:$c; done
done
done
grep -E "^=== \[|time " $OUT | grep -B1 "time" | grep -v "^--" | paste - - | awk '{printf "%-12s %-28s %s %s %s %s\n", $2, $3, $6, $7, $8, $9}'
grep -E "FAIL|MISMATCH|rc=[1-9]" $OUT | head
