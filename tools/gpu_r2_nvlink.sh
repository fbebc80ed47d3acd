#!/bin/bash
# NVLink byte counters around sharded steps (N GPUs, default 2): C5 and C3, symmetric-memory vs NCCL exchange
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 2>&1 | head -12
for c in C5 C3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tools/nvlink_counters.py --config $c --steps 200 --out gpurun_out/nvlink_n${N}_$c.json 2> gpurun_out/nvlink_n${N}_$c.err | tail -c 1800
  echo "nvlink $c exit=$?"; tail -3 gpurun_out/nvlink_n${N}_$c.err
done
