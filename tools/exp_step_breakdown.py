"""Where a C3 training step spends its time: CPU issue time per step (no sync) vs GPU time (events), per component."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.ops import _fused_step

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
B, N, D, tau = cfg["B"], cfg["N"], cfg["D"], cfg["tau"]
dev = torch.device("cuda:0")
bank_cpu = synth.make_bank(N, D, 1, seed=1234)
labels = synth.make_labels(B, N, seed=1235)
q = synth.make_queries(bank_cpu, labels, "S", seed=1236).to(dev)
labels = labels.to(dev)
bank = spn.DeviceBank.from_tensor(bank_cpu, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def full():
    qx = q.detach().requires_grad_(True)
    loss = spn.bank_infonce(qx, bank, labels, tau)
    loss.backward()
    return qx.grad

def lib_only():
    return _fused_step(q, bank, labels, 1.0 / tau, 0.0, False)

def measure(fn, iters=50, do_flush=True):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs, cpu = [], 0.0
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        t0 = time.perf_counter()
        fn()
        cpu += time.perf_counter() - t0
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return dict(gpu_us=round(sum(s.elapsed_time(e) for s, e in evs) / iters * 1e3, 1), cpu_issue_us=round(cpu / iters * 1e6, 1))

print(json.dumps({"autograd step (flush)": measure(full), "autograd step (no flush)": measure(full, do_flush=False),
                  "spn_bank_ce_step only (flush)": measure(lib_only), "spn_bank_ce_step only (no flush)": measure(lib_only, do_flush=False)}))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    lib_only()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = lib_only()
torch.cuda.synchronize()
print(json.dumps({"graph replay of spn_bank_ce_step (flush)": measure(g.replay)}))
# the same after 1.5 s of sustained load (what bench.py does before its timed region)
t_end = time.perf_counter() + 1.5
while time.perf_counter() < t_end:
    for _ in range(20):
        full()
    torch.cuda.synchronize()
print(json.dumps({"after 1.5 s load: autograd step (flush)": measure(full, iters=20), "lib only": measure(lib_only, iters=20)}))
ev = []
for name, fn in (("prep+main+finalize", lib_only),):
    pass
import spn4cir_b200.ops as ops
from spn4cir_b200._lib import call as _call
sw = ops._fused_sweep(q, bank, labels, 1.0 / tau, False)
ws, wsb = ops.fused_scratch(B, bank, dev)
def main_only():
    _call("spn_bank_ce_fused", sw["q16"].data_ptr(), bank.data.data_ptr(), labels.data_ptr(), sw["ref"].data_ptr(), 1.0 / tau, B, bank.N, D,
          bank.T, 0, 0, sw["o_acc"].data_ptr(), sw["l_acc"].data_ptr(), sw["s_t"].data_ptr(), sw["lbl_row"].data_ptr(), None, None,
          ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
print(json.dumps({"after load: main kernel only": measure(main_only, iters=20)}))
