"""Raw tcgen05.mma throughput per SM for PATTERNS of MMAs (SS / TS, N = 64/128/256, K-major / MN-major B, shared or
distinct B tiles, switch granularity).  See umma_bench_kernel for the op encoding."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spn4cir_b200 import _lib

def op(ts=0, n=128, mn=0, region=0, acc2=0):
    return 1 | (ts << 1) | ({64: 0, 128: 1, 256: 2}[n] << 2) | (mn << 4) | (region << 5) | (acc2 << 7)

def ideal(o):
    return 64 * (1 << ((o >> 2) & 3)) / 2  # M128: N/2 cycles per K=16 MMA

grid, iters = 148, 1000
out = torch.zeros(grid, dtype=torch.int64, device="cuda")
pats = {
    "SS128": [op()], "SS256": [op(n=256)], "TS64 k": [op(ts=1, n=64)], "TS128 k": [op(ts=1, n=128)], "TS256 k": [op(ts=1, n=256)],
    "TS64 mn": [op(ts=1, n=64, mn=1, region=2)], "TS128 mn": [op(ts=1, n=128, mn=1, region=2)], "TS256 mn": [op(ts=1, n=256, mn=1, region=0)],
    "SS64 mn": [op(n=64, mn=1, region=2)],
    "SS128 + TS64 k sameB": [op(), op(ts=1, n=64, acc2=1)],
    "SS128 + TS64 k otherB": [op(), op(ts=1, n=64, region=2, acc2=1)],
    "SS128 + TS64 mn otherB": [op(), op(ts=1, n=64, mn=1, region=2, acc2=1)],
    "SS128 + TS128 mn otherB": [op(), op(ts=1, n=128, mn=1, region=2, acc2=1)],
    "SS128 + SS64 k otherB": [op(), op(n=64, region=2, acc2=1)],
    "SS128 + SS64 mn otherB": [op(), op(n=64, mn=1, region=2, acc2=1)],
    "SS128 + TS64 mn sameacc": [op(), op(ts=1, n=64, mn=1, region=2)],
    "TS128 k + TS64 mn": [op(ts=1), op(ts=1, n=64, mn=1, region=2, acc2=1)],
    "TS128 k + TS64 k otherB": [op(ts=1), op(ts=1, n=64, region=2, acc2=1)],
}
for name, ops in pats.items():
    for group in (1, 4, 8) if len(ops) > 1 else (4,):
        pat = 0
        for i, o in enumerate(ops):
            pat |= o << (8 * i)
        for rep in range(2):
            _lib.call("spn_debug_umma_bench", None, 0, group, iters, pat & 0xffffffff, pat >> 32, out.data_ptr(), grid,
                      torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        cyc = out.double().mean().item() / (iters * group)
        idl = sum(ideal(o) for o in ops)
        print(json.dumps(dict(pattern=name, group=group, cyc_per_pattern=round(cyc, 1), ideal=idl, eff=round(idl / cyc, 3))), flush=True)
