#!/bin/bash
# N-GPU check: distributed GPU tests + bench (default: queries data-parallel; --bank-sharded: NCCL exchange).
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
for extra in "" "--bank-sharded"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 $extra 2> gpurun_out/bench_n${N}.err | tee gpurun_out/bench_n${N}${extra}.json | tail -1
  tail -3 gpurun_out/bench_n${N}.err
done
