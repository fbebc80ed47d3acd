"""single-pass pair kernel: TMA reduce-add epilogue (default) vs per-thread red.global.add (SPN_DEBUG_TMA_SINGLE=0)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spn4cir_b200 as spn
import bench
dev = torch.device("cuda")
peaks = bench.measured_peaks()
timer = bench.Timer(dev)
for rep in range(2):
    for flag in ("1", "0"):
        os.environ["SPN_DEBUG_TMA_SINGLE"] = flag
        for (B, N, D, T, tau) in [(512, 100000, 256, 1, 0.02), (512, 100000, 320, 1, 0.02), (1024, 400000, 256, 1, 0.02), (256, 50000, 256, 32, 0.05)]:
            r = bench.bench_shape(spn, timer, dev, peaks, f"tma_single={flag}", B, N, D, T, tau, 20)
            print(json.dumps({k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in ("workload", "B", "N", "D", "T", "ms_per_step", "kernel_ms", "kernel_frac_of_bf16_peak")}))
