#!/bin/bash
# Full GPU check on the B200 box: pytest -m gpu, smoke(), bench.py. Logs land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.log
