#!/bin/bash
# bench lines of the second round-2 session: the 128-image shard on one GPU (graph replay vs launch plan), optionally
# the 2-GPU strong-scaled run, the full GPU test suite.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
N=${NGPU:-1}
if [ "$N" = "1" ]; then
  for g in on off; do
    timeout 600 python bench.py --batch 128 --graph $g --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_b128_graph_$g.json 2> gpurun_out/r2b_bench_b128_graph_$g.err
    echo "bench b128 graph=$g rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2b_bench_b128_graph_$g.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['config']['launch'][:30])"
  done
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
  echo "bench n$N rc=$?"; tail -3 gpurun_out/r2b_bench_n$N.err; cat gpurun_out/r2b_bench_n$N.json
fi
