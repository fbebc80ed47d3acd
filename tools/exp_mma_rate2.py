"""Compile-time tcgen05.mma issue patterns (umma_bench2_kernel): cycles per iteration of 4 k-steps, per SM."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spn4cir_b200 import _lib
names = {0: ("SS128", 64), 1: ("TS64mn", 32), 2: ("SS128+TS64k sameB", 96), 3: ("SS128+TS64k otherB", 96),
         4: ("SS128+TS64mn otherB", 96), 5: ("4xSS128 then 4xTS64mn", 96), 6: ("TS256mn", 128), 7: ("TS128mn", 64),
         8: ("SS128+2xTS64mn", 128), 9: ("SS256", 128), 10: ("SS256+TS64mn", 160), 11: ("SS256+2xTS64mn", 192),
         12: ("SS128+TS128mn", 128), 13: ("TS128k+TS64mn", 96), 14: ("SS64+TS64mn", 64)}
grid, iters = 148, 2000
out = torch.zeros(grid, dtype=torch.int64, device="cuda")
for pat, (name, ideal) in names.items():
    for rep in range(2):
        _lib.call("spn_debug_umma_bench", None, 0, 0, iters, pat, 0, out.data_ptr(), grid, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    cyc = out.double().mean().item() / (iters * 4)
    print(json.dumps(dict(pat=pat, pattern=name, cyc_per_kstep=round(cyc, 1), ideal=ideal, eff=round(ideal / cyc, 3))), flush=True)
