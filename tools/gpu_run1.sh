#!/bin/bash
# First GPU bring-up: every stage under its own timeout, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
for s in umma fwd bwd sim; do
  echo "=== diag $s" | tee -a gpurun_out/diag.log
  timeout 180 python tools/diag.py $s >> gpurun_out/diag.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/diag.log
done
echo "=== pytest gpu" 
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_x.log 2>&1
echo "pytest -x exit=$?"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest all exit=$?"
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"
echo "=== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"
tail -5 gpurun_out/diag.log
tail -15 gpurun_out/pytest_gpu_all.log
tail -3 gpurun_out/smoke.log
tail -2 gpurun_out/bench.log
