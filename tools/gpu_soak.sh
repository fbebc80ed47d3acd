#!/bin/bash
# flakiness soak: the whole -m gpu suite N times on one box (tolerances vs atomic-order nondeterminism)
cd "$(dirname "$0")/.."
N=${1:-6}
mkdir -p gpurun_out
for i in $(seq 1 $N); do
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -p no:randomly 2>&1 | tail -2 | tr '\n' ' '
  echo " [run $i]"
done
