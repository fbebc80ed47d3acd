#!/bin/bash
# one GPU: full GPU test suite, smoke, the bench line (C5 headline + secondary); with "ncu" as $1 also the launch list +
# a full capture of the fused kernel at C5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit=$?"
tail -15 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -c 1500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -k regex:"bank_ce|prep_query|finalize|shard_|cast_bf16|query_assemble|fused_init" -c 40 --csv --log-file gpurun_out/launches_c5.csv \
    python bench.py --profile --steps 3 > gpurun_out/launches_run.log 2>&1
echo "launch list exit=$?"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:bank_ce_kernel -s 1 -c 1 -f -o gpurun_out/prof_bank_ce_c5 \
    python bench.py --profile --steps 1 > gpurun_out/prof_run.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out/*.ncu-rep
fi
