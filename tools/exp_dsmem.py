"""DSMEM store bandwidth (st.shared::cluster.v4) per writer CTA, clusters of 4."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spn4cir_b200 import _lib
for grid in (4, 148):
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    for ndst in (1, 2):
        iters = 200
        for rep in range(2):
            _lib.call("spn_debug_dsmem_bench", iters, ndst, out.data_ptr(), grid, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        w = out[out > 0].double()
        cyc = w.mean().item()
        print(json.dumps(dict(grid=grid, ndst=ndst, writers=int((out > 0).sum()), cycles_per_32KiB=round(cyc / iters, 1),
                              bytes_per_cycle_out=round(32768 * ndst * iters / cyc, 1))), flush=True)
