#!/bin/bash
# 8-GPU record: C3 data-parallel (the driver's scaling run), C5 bank-sharded (BASELINE config 5), distributed tests.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-retrieval $2 2> gpurun_out/bench_n${N}_$1.err | tee gpurun_out/bench_n${N}_$1.json | tail -c 1800
  tail -2 gpurun_out/bench_n${N}_$1.err
}
run dp ""
run c5_sharded "--config C5 --bank-sharded"
run c3_sharded "--bank-sharded"
