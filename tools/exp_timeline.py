"""Per-role cycle accounting of the bank-CE kernels (needs the -DSPN_TIMELINE build: tools/build_timeline.sh)."""
import ctypes, os, sys, json
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
os.environ["SPN_LIB_PATH"] = str(REPO / "spn4cir_b200" / "libspn_b200_timeline.so")
import numpy as np
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import _lib
from spn4cir_b200.ops import _bwd, _fwd_stats, _loss_from_stats, cast_to_bf16

def read(ncta):
    buf = (ctypes.c_ulonglong * (ncta * 16))()
    _lib.call("spn_debug_read_timeline", ctypes.addressof(buf), ncta * 16)
    return np.frombuffer(buf, dtype=np.uint64).reshape(ncta, 16).astype(np.float64)

B, N = 512, 100000
for D in (256, 640):
    bank = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=-1)
    lab = torch.randint(0, N, (B,), device="cuda")
    db = spn.DeviceBank(bank.to(torch.bfloat16).contiguous())
    q16 = cast_to_bf16(q)
    for it in range(2):
        stats = _fwd_stats(q16, db, lab, 50.0, False)
        torch.cuda.synchronize()
    t = read(148)[:144]
    names = ["prod_wait_empty", "prod_issue_g1", "prod_issue_g2", "prod_total", "mma_wait_full", "mma_wait_sfree", "mma_wait_pfull", "mma_total",
             "sm_wait_sfull", "sm_tmem_ld", "sm_wait_pfree", "sm_math", "sm_tmem_st", "sm_arrive_sfree", "sm_arrive_pfull", "sm_total"]
    print(f"FWD D={D}:", json.dumps({n: round(float(t[:, i].mean())) for i, n in enumerate(names) if n != "-"}), flush=True)
    lse, _, _ = _loss_from_stats(stats, N, 0.0)
    dq = torch.zeros(B, D, device="cuda")
    for it in range(2):
        _bwd(q16, db, lab, 50.0, lse, 0.1, None, 0.0, dq, None)
        torch.cuda.synchronize()
    t = read(148)[:144]
    print(f"BWD D={D}:", json.dumps({n: round(float(t[:, i].mean())) for i, n in enumerate(names) if n != "-"}), flush=True)
    for rk in (0, 1):
        print(f"BWD D={D} even/odd blocks {rk}:", json.dumps({n: round(float(t[rk::2, i].mean())) for i, n in enumerate(names) if n != "-"}), flush=True)
