"""Raw tcgen05.mma throughput per SM: SS vs TS, N = 64/128/256, with and without concurrent TMA writes."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spn4cir_b200 import _lib
grid = 148
src = torch.zeros((grid * 512, 512), dtype=torch.bfloat16, device="cuda")
out = torch.zeros(grid, dtype=torch.int64, device="cuda")
iters = 2000
for tma_on in (0,):
    for mode in (0, 1, 2, 3, 4):
        for N in ((64, 128) if mode >= 2 else (64, 128, 256)):
            for rep in range(2):
                _lib.call("spn_debug_umma_bench", src.data_ptr(), src.shape[0], N, iters, mode, tma_on, out.data_ptr(), grid,
                          torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            cyc = out.double().mean().item()
            per_mma = cyc / (iters * 4)
            ideal = 128 * N * 16 / 4096
            print(json.dumps(dict(tma=tma_on, mode=mode, A=("fused SS N + TS 64" if mode & 4 else ("tmem" if mode & 1 else "smem")) + (" alt-acc" if mode & 2 else ""), N=N, cyc_per_mma=round(per_mma, 1), ideal=ideal,
                                  eff=round(ideal / per_mma, 3))), flush=True)
