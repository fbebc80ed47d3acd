"""Per 64-column block difference between the two-pass fused kernel and the D-split recompute kernel."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.ops import _fused_sweep
B, N, D = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (256, 5000, 512)
dev = torch.device("cuda:0")
q, bank, labels = synth.make_bank_case(B, N, D, 1, "S", 25)
db = spn.DeviceBank.from_tensor(bank, dev)
res = {}
for tp in ("1", "0"):
    os.environ["SPN_DEBUG_TWO_PASS"] = tp
    sw = _fused_sweep(q.to(dev), db, labels.to(dev), 50.0, False)
    torch.cuda.synchronize()
    res[tp] = sw["o_acc"].cpu()
a, b = res["1"], res["0"]
for c in range(D // 64):
    x, y = a[:, c * 64:(c + 1) * 64], b[:, c * 64:(c + 1) * 64]
    print(f"cols {c*64:4d}..: rel err {float((x - y).norm() / y.norm()):.3e}   |two-pass| {float(x.norm()):.3e} |split| {float(y.norm()):.3e}")
for r0 in range(0, B, 32):
    x, y = a[r0:r0 + 32], b[r0:r0 + 32]
    print(f"rows {r0:4d}..: rel err {float((x - y).norm() / y.norm()):.3e}")
