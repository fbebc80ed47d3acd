#!/bin/bash
# Regular GPU check: full GPU test suite + smoke + bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit=$?"
tail -12 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
