#!/bin/bash
# ncu evidence for the bank-CE kernels (one GPU): launch list + full capture of forward and backward kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 2 > gpurun_out/launches_run.log 2>&1
echo "launch list exit=$?"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:bank_ce_kernel -s 1 -c 1 -f -o gpurun_out/prof_bank_ce \
    python bench.py --profile --steps 1 > gpurun_out/prof_run.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/prof_run.log
