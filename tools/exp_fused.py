"""Time the fused bank-CE kernel alone (CUDA events, L2 flushed) for a few shapes; SPN_DEBUG_CG=1 forces the single-CTA kernel."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import _lib
from spn4cir_b200.ops import _fused_sweep

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def time_fused(B, N, D, T, tau, iters=10):
    g = torch.Generator(device="cuda").manual_seed(0)
    bank = torch.nn.functional.normalize(torch.randn(N * T, D, device="cuda", generator=g), dim=-1)
    db = spn.DeviceBank(bank.to(torch.bfloat16).contiguous(), T)
    lab = torch.randint(0, N, (B,), device="cuda", generator=g)
    q = torch.nn.functional.normalize(bank[lab * T].float() + 3.2 * torch.randn(B, D, device="cuda", generator=g) / D ** 0.5, dim=-1)
    sw = _fused_sweep(q, db, lab, 1.0 / tau, False)
    from spn4cir_b200.ops import fused_scratch
    ws, ws_bytes = fused_scratch(B, db, q.device)
    def run():
        _lib.call("spn_bank_ce_fused", sw["q16"].data_ptr(), db.data.data_ptr(), lab.data_ptr(), sw["ref"].data_ptr(),
                  1.0 / tau, B, db.N, D, db.T, db.row_base, 0, sw["o_acc"].data_ptr(), sw["l_acc"].data_ptr(),
                  sw["s_t"].data_ptr(), sw["lbl_row"].data_ptr(), None, None, ws.data_ptr() if ws is not None else None, ws_bytes,
                  torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

cases = [("B384", 384, 100000, 640, 1, 0.02), ("B128", 128, 100000, 640, 1, 0.02), ("N25k", 512, 25000, 640, 1, 0.02), ("N50k", 512, 50000, 640, 1, 0.02), ("N200k", 512, 200000, 640, 1, 0.02),
         ("N400k", 512, 400000, 640, 1, 0.02), ("C3", 512, 100000, 640, 1, 0.02), ("C5/8", 1024, 125000, 768, 1, 0.02), ("D512", 512, 100000, 512, 1, 0.02),
         ("D256", 512, 100000, 256, 1, 0.03), ("C4/8", 256, 25000, 256, 32, 0.05), ("1Mx640", 1024, 1000000, 640, 1, 0.02), ("C5", 1024, 1000000, 768, 1, 0.02)]
which = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] != "all" else None
for name, B, N, D, T, tau in cases:
    if which and name not in which:
        continue
    variants = os.environ.get("EXP_VARIANTS", "tp,split,cg1").split(",")
    for var in variants:
        for k in ("SPN_DEBUG_CG", "SPN_DEBUG_TWO_PASS", "SPN_DEBUG_HG"):
            os.environ.pop(k, None)
        if var == "cg1":
            os.environ["SPN_DEBUG_CG"] = "1"
        elif var == "split":
            os.environ["SPN_DEBUG_TWO_PASS"] = "0"
        elif var.startswith("hg"):
            os.environ["SPN_DEBUG_HG"] = var[2:]
        t = time_fused(B, N, D, T, tau)
        fl = (2 * 32 + 2 if T == 32 else 4) * B * N * D
        print(json.dumps(dict(case=name, B=B, N=N, D=D, T=T, variant=var, us=round(t * 1e6, 1), alg_tflops=round(fl / t / 1e12, 1))), flush=True)
