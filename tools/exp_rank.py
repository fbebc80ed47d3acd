"""C2 rank-of-target: fused epilogue path vs S~ + row kernel path (CUDA events, warm L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spn4cir_b200 as spn
from spn4cir_b200 import synth
dev = torch.device("cuda")
c = synth.CONFIGS["C2"]
q, gal, _, ref, tgt, grp = synth.make_retrieval_case(c["Q"], c["G"], c["D"], 1, 0, noise=8.0)
qd, gd = q.to(dev), gal.to(dev)
refd, tgtd, grpd = ref.to(dev, torch.int32), tgt.to(dev, torch.int32), grp.to(dev, torch.int32)
def timed(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
for fused in ("1", "0"):
    os.environ["SPN_TC_RANK_FUSED"] = fused
    t = timed(lambda: spn.sim_rank(qd, gd, tgtd, exclude=refd, group=grpd, mode="tc"))
    t2 = timed(lambda: spn.sim_rank(qd, gd, tgtd, exclude=refd, mode="tc"))
    print(f"fused={fused}: rank+subset {t:.1f} us, rank only {t2:.1f} us")
