#!/bin/bash
# retrieval timings only (tools/exp_retrieval.py), no tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/exp_retrieval.py > gpurun_out/exp_retrieval.log 2>&1
echo "exp exit=$?"; tail -3 gpurun_out/exp_retrieval.log | cut -c1-420
