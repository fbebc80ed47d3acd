"""Per-role cycle accounting of the fused bank-CE kernel (needs tools/build_timeline.sh).  SPN_DEBUG_HG selects
variants (2: no second pass, 4: no P~ stores)."""
import ctypes, os, sys, json
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
os.environ["SPN_LIB_PATH"] = str(REPO / "spn4cir_b200" / "libspn_b200_timeline.so")
import numpy as np
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import _lib
from spn4cir_b200.ops import _fused_sweep

def read(ncta):
    buf = (ctypes.c_ulonglong * (ncta * 16))()
    _lib.call("spn_debug_read_timeline", ctypes.addressof(buf), ncta * 16)
    return np.frombuffer(buf, dtype=np.uint64).reshape(ncta, 16).astype(np.float64)

B, N, D = 512, int(os.environ.get("EXP_N", "100000")), 640
bank = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=-1)
lab = torch.randint(0, N, (B,), device="cuda")
q = torch.nn.functional.normalize(bank[lab] + 3.2 * torch.randn(B, D, device="cuda") / D ** 0.5, dim=-1)
db = spn.DeviceBank(bank.to(torch.bfloat16).contiguous())
names = ["prod_wait_empty", "prod_issue_g1", "prod_issue_g2", "prod_total", "mma_wait_full", "mma_wait_sfree", "mma_wait_pfull", "mma_total",
         "sm_wait_sfull", "sm_tmem_ld", "sm_wait_pfree", "sm_math", "sm_tmem_st", "sm_arrive_sfree", "sm_arrive_pfull", "sm_total"]
for hg in sys.argv[1:] or ["6", "2", "0"]:
    os.environ["SPN_DEBUG_HG"] = hg
    for it in range(3):
        _fused_sweep(q, db, lab, 50.0, False)
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 16))()
    _lib.call("spn_debug_read_timeline", ctypes.addressof(buf), -148 * 16)
    ev = np.frombuffer(buf, dtype=np.uint64).reshape(148, 16).astype(np.float64)
    evn = {0: "setup", 1: "q_resident", 2: "p1_issued", 3: "tmem_free", 4: "p2_first_P", 5: "p2_issued", 8: "p1_retired",
           9: "epi1_done", 10: "p2_retired", 11: "end", 12: "p2_wait_P", 13: "p2_wait_bank"}
    print(f"hg={hg} events (rank-0 CTAs, mean cycles since kernel entry):", json.dumps({n: round(float(ev[0::2, i].mean())) for i, n in evn.items()}), flush=True)
    t = read(148)
    for rk in (0, 1):
        print(f"hg={hg} rank {rk}:", json.dumps({n: round(float(t[rk::2, i].mean())) for i, n in enumerate(names)}), flush=True)
