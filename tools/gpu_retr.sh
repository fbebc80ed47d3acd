#!/bin/bash
# Retrieval check: tensor-core retrieval tests + the fp32 ones + C2 timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval_tc.py tests/test_gpu_retrieval.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_retr.log 2>&1
echo "pytest exit=$?"; tail -30 gpurun_out/pytest_retr.log
timeout 300 python tools/exp_retrieval.py > gpurun_out/exp_retrieval.log 2>&1
echo "exp exit=$?"; tail -5 gpurun_out/exp_retrieval.log
