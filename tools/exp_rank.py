"""C2 retrieval: pair sweep vs 1-CTA sweep (SPN_TC_SWEEP_CG=1), fused rank epilogue vs S~ + row kernel (CUDA events, warm L2)."""
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
for cg in ("2", "1"):
    os.environ["SPN_TC_SWEEP_CG"] = cg
    for tn in ([None, "256", "128"] if cg == "2" else [None]):
        if tn: os.environ["SPN_TC_TILE_N"] = tn
        else: os.environ.pop("SPN_TC_TILE_N", None)
        for fused in ("1", "0"):
            os.environ["SPN_TC_RANK_FUSED"] = fused
            t = timed(lambda: spn.sim_rank(qd, gd, tgtd, exclude=refd, group=grpd, mode="tc"))
            print(f"sweep cg={cg} tile_n={tn} fused={fused}: rank+subset {t:.1f} us")
        tk = timed(lambda: spn.sim_topk(qd, gd, c["K"], exclude=refd, mode="tc", exact_values=False))
        print(f"sweep cg={cg} tile_n={tn}: top-50 {tk:.1f} us")
