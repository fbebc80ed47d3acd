#!/bin/bash
# N-GPU check (default 2): NCCL parity test + the sharded bench line (C5 headline) + C3 sharded (the shape that hung in round 1)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 500 -p no:cacheprovider -x 2>&1 | tail -15
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 $2 2> gpurun_out/bench_n${N}_$1.err > gpurun_out/bench_n${N}_$1.json
  echo "bench $1 exit=$?"; tail -c 2500 gpurun_out/bench_n${N}_$1.json; tail -3 gpurun_out/bench_n${N}_$1.err
}
run c5 ""
run c3 "--config C3"
