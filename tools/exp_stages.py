"""Experiment: forward/backward kernel time vs feature dim and ring depth (latency sensitivity)."""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.ops import _bwd, _fwd_stats, _loss_from_stats, cast_to_bf16

def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e-3

B, N = 512, 100000
for D in (256, 512, 640, 768):
    bank = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=-1)
    lab = torch.randint(0, N, (B,), device="cuda")
    db = spn.DeviceBank(bank.to(torch.bfloat16).contiguous())
    q16 = cast_to_bf16(q)
    for st in (None,):
        if st is None: os.environ.pop("SPN_DEBUG_STAGES", None)
        else: os.environ["SPN_DEBUG_STAGES"] = str(st)
        stats = _fwd_stats(q16, db, lab, 50.0, False)
        lse, _, _ = _loss_from_stats(stats, N, 0.0)
        dq = torch.zeros(B, D, device="cuda")
        tf = timeit(lambda: _fwd_stats(q16, db, lab, 50.0, False))
        tb = timeit(lambda: _bwd(q16, db, lab, 50.0, lse, 0.1, None, 0.0, dq, None))
        fl = 2.0 * B * N * D
        print(json.dumps(dict(D=D, stages=st, fwd_us=round(tf * 1e6, 1), fwd_tflops=round(fl / tf / 1e12, 1),
                              bwd_us=round(tb * 1e6, 1), bwd_alg_tflops=round(fl / tb / 1e12, 1))), flush=True)
