#!/bin/bash
# sustained-run clocks/power for linear cases in normal and no-store mode (time-ordered samples)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
BIN=pytorch_models_b200/b200enc_selftest
for c in "$@"; do
for dbg in 0 1; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > gpurun_out/smi_${c}_$dbg.csv &
  SMI=$!
  sleep 0.5
  B200_ITERS=${ITERS:-30000} B200_DEBUG_FLAGS=$dbg timeout 120 $BIN linear:$c 2>&1 | grep -E "time"
  kill $SMI
  echo "-- $c dbg=$dbg (clock MHz, W, power_cap) every 100 ms:"; tr '\n' '|' < gpurun_out/smi_${c}_$dbg.csv | cut -c1-1500; echo
done
done
