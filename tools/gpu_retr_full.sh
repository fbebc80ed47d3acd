#!/bin/bash
# ncu --set full of the tensor-core retrieval kernels (row kernels + sweep) at C2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:"row_|sim_tc_gemm|target_score_warp|rank_prep|rank_resolve|split_rows" -s 14 -c 8 -f -o gpurun_out/prof_retr \
    python tools/prof_retr.py > gpurun_out/prof_retr_run.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out/prof_retr.ncu-rep
