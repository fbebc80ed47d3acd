"""A few tensor-core retrieval calls at C2 for an ncu launch list (tools/gpu_retr_ncu.sh)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import synth
c = synth.CONFIGS["C2"]
q, gal, _, ref, tgt, grp = synth.make_retrieval_case(c["Q"], c["G"], c["D"], 1, 0, noise=8.0)
dev = torch.device("cuda:0")
qd, gd = q.to(dev), gal.to(dev)
refd, tgtd, grpd = ref.to(dev, torch.int32), tgt.to(dev, torch.int32), grp.to(dev, torch.int32)
for _ in range(3):
    spn.sim_topk(qd, gd, 50, exclude=refd, mode="tc")
    spn.sim_topk(qd, gd, 50, exclude=refd, mode="tc", exact_values=False)
    spn.sim_rank(qd, gd, tgtd, exclude=refd, group=grpd, mode="tc")
torch.cuda.synchronize()
