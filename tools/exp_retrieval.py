"""C2 retrieval timing: fused sim + top-50, rank-of-target, and the dense score matrix (GEMM share of the kernel)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import synth, compat
c = synth.CONFIGS["C2"]
q, gal, _, ref, tgt, grp = synth.make_retrieval_case(c["Q"], c["G"], c["D"], 1, 0, noise=8.0)
dev = torch.device("cuda:0")
qd, gd = q.to(dev), gal.to(dev)
refd, tgtd, grpd = ref.to(dev, torch.int32), tgt.to(dev, torch.int32), grp.to(dev, torch.int32)
from spn4cir_b200._lib import call
dense_out = torch.empty((c["Q"], c["G"]), dtype=torch.float32, device=dev)
def dense(a, b):
    call("spn_sim_matrix", a.data_ptr(), b.data_ptr(), c["Q"], c["G"], c["D"], 1, dense_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return round(s.elapsed_time(e) / iters * 1e3, 1)
out = {}
for mode in ("fp32", "tc"):
    out.update({f"topk50_{mode}_us": t(lambda: spn.sim_topk(qd, gd, 50, exclude=refd, mode=mode)),
                f"topk10_{mode}_us": t(lambda: spn.sim_topk(qd, gd, 10, exclude=refd, mode=mode)),
                f"rank_{mode}_us": t(lambda: spn.sim_rank(qd, gd, tgtd, exclude=refd, group=grpd, mode=mode))})
out["topk50_tc_index_exact_us"] = t(lambda: spn.sim_topk(qd, gd, 50, exclude=refd, mode="tc", exact_values=False))
out["topk10_tc_index_exact_us"] = t(lambda: spn.sim_topk(qd, gd, 10, exclude=refd, mode="tc", exact_values=False))
n2 = torch.zeros(1, dtype=torch.int32, device=dev)
spn.sim_topk(qd, gd, 50, exclude=refd, mode="tc", n_second_out=n2)
out["topk50_tc_rows_second_scan"] = int(n2)
# the three kernels of the tensor-core path, timed one by one through the profiler-free event pairs of a CUDA graph-less loop
out.update({
       "matrix_us": t(lambda: dense(qd, gd)),
       "torch_matmul_fp32_us": t(lambda: qd @ gd.T)})
print(json.dumps(out))
