#!/bin/bash
# compute-sanitizer memcheck over small cases of the kernels added in round 2 (rank window, shard kernels, fused rank
# epilogue + pair sweep, single-pass TMA epilogue, query-assemble backward)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=$(which compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import spn4cir_b200 as spn
from spn4cir_b200 import synth, compat
from spn4cir_b200.dist import _CudaShardCompute
dev = "cuda"
# rank window
f = torch.nn.functional.normalize(torch.randn(600, 64), dim=-1).to(dev)
w = compat.i2i_rank_window(f, 100, 300, block_rows=256)
# fused rank + pair sweep + topk
q, g, _, ref, tgt, grp = synth.make_retrieval_case(300, 700, 128, 1, 3, noise=6.0)
r = spn.sim_rank(q.to(dev), g.to(dev), tgt, exclude=ref, group=grp, mode="tc")
v = spn.sim_topk(q.to(dev), g.to(dev), 20, exclude=ref, mode="tc")
# bank step: two-pass, single-pass, T = 32; fused bank_large_step
for (B, N, D, T) in [(256, 1500, 640, 1), (256, 1500, 256, 1), (128, 200, 256, 32)]:
    qq, bank, lab = synth.make_bank_case(B, N, D, T, "S", 5)
    db = spn.DeviceBank.from_tensor(bank, dev)
    x = qq.to(dev).requires_grad_(True)
    spn.bank_infonce(x, db, lab.to(dev), 0.03).backward()
# shard kernels (fake ranks on one device)
qq, bank, lab = synth.make_bank_case(256, 2000, 640, 1, "S", 6)
shards = [spn.DeviceBank.from_tensor(bank, dev, shard_rank=i, shard_world=2) for i in range(2)]
c = _CudaShardCompute()
q16 = c.cast(qq.to(dev)); labd = lab.to(dev)
mx = torch.stack([s.stats()[0] for s in shards]).max().reshape(1)
ref = torch.stack([c.prep(q16, labd, s, 50.0, mx) for s in shards]).min(0).values.contiguous()
outs = [c.fused(q16, labd, ref, s, 50.0, True) for s in shards]
vec = torch.stack([c.pack(o[1], o[2], o[3], o[4]) for o in outs]).sum(0).contiguous()
for o, s in zip(outs, shards):
    c.fold(o[0], vec, ref, o[4], s, 0.1)
torch.cuda.synchronize()
print("SANITIZE_CASES_DONE")
PY
timeout 900 $CS --tool memcheck --error-exitcode 7 python /tmp/san_case.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit=$?"
tail -8 gpurun_out/sanitize_memcheck.log
