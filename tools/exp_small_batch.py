"""B = 128 / 32 shapes: pair kernel with an empty half (default) vs the single-CTA plan (SPN_DEBUG_PAD1=0)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spn4cir_b200 as spn
from spn4cir_b200 import synth
import bench
dev = torch.device("cuda")
peaks = bench.measured_peaks()
timer = bench.Timer(dev)
for pad in ("1", "0"):
    os.environ["SPN_DEBUG_PAD1"] = pad
    spn.ops._STEP_WS.clear()
    for (B, N, D, T, tau) in [(128, 100000, 640, 1, 0.02), (64, 100000, 640, 1, 0.02), (32, 50000, 256, 32, 0.05), (128, 200000, 256, 32, 0.05), (128, 100000, 768, 1, 0.02)]:
        r = bench.bench_shape(spn, timer, dev, peaks, f"pad1={pad}", B, N, D, T, tau, 20)
        print(json.dumps({k: r[k] for k in ("workload", "B", "N", "D", "T", "ms_per_step", "kernel_ms", "kernel_frac_of_bf16_peak", "kernel_frac_of_hbm_peak")}))
