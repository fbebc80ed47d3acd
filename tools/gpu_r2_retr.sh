#!/bin/bash
# round-2 retrieval evidence: launch list + ncu --set full of the retrieval kernels at C2
cd "$(dirname "$0")/.."
bash tools/gpu_retr_ncu.sh
bash tools/gpu_retr_full.sh
