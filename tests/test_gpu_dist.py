"""Multi-GPU parity (needs >= 2 GPUs, NCCL): the row-sharded fused path must reproduce the single-GPU loss and
gradient for every rank's queries.  Launched as a subprocess under torch.distributed.run.  The comparison is against this
repo's own single-GPU path on the full bank, which tests/test_gpu_bank_ce.py pins to the oracle and the executed-reference
fixtures -- so parity with the reference holds transitively; tests/test_gpu_shard.py checks the same per-rank kernels
against the oracle directly (fake ranks on one GPU)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SPN_REPO"])
import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.dist import bank_infonce_sharded, sim_topk_sharded, sim_rank_sharded, shard_bounds
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, N, D, tau = 96, 5000, 256, 0.02
bank_cpu = synth.make_bank(N, D, 1, seed=77)
qs, labs = [], []
for r in range(world):
    lab = synth.make_labels(B, N, seed=100 + r)
    qs.append(synth.make_queries(bank_cpu, lab, "S", seed=200 + r)); labs.append(lab)
shard = spn.DeviceBank.from_tensor(bank_cpu, dev, shard_rank=rank, shard_world=world)
tt = torch.tensor(tau, device=dev, requires_grad=True)
q = qs[rank].to(dev).requires_grad_(True)
loss = bank_infonce_sharded(q, shard, labs[rank].to(dev), tt)
(loss * 3.0).backward()
full = spn.DeviceBank.from_tensor(bank_cpu, dev)
t2 = torch.tensor(tau, device=dev, requires_grad=True)
q2 = qs[rank].to(dev).requires_grad_(True)
ref = spn.bank_infonce(q2, full, labs[rank].to(dev), t2, single_pass=True)
(ref * 3.0).backward()
assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item())), (rank, loss.item(), ref.item())
rel = float((q.grad - q2.grad).norm() / q2.grad.norm())
assert rel < 1e-4, (rank, rel)
assert abs(tt.grad.item() - t2.grad.item()) <= 1e-3 * max(abs(t2.grad.item()), 1e-3), (tt.grad, t2.grad)
# label smoothing over the sharded bank (zscir/models_bank.py:35): all-reduced column sums
q3 = qs[rank].to(dev).requires_grad_(True)
ls = bank_infonce_sharded(q3, shard, labs[rank].to(dev), tau, label_smoothing=0.1)
ls.backward()
q4 = qs[rank].to(dev).requires_grad_(True)
lr = spn.bank_infonce(q4, full, labs[rank].to(dev), tau, label_smoothing=0.1, single_pass=True)
lr.backward()
assert abs(ls.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item())), (rank, ls.item(), lr.item())
assert float((q3.grad - q4.grad).norm() / q4.grad.norm()) < 1e-4
# the step primitive (what bench.py times at N > 1) and a bigger two-pass shape with B a multiple of 128
from spn4cir_b200.dist import bank_ce_step_sharded
bank2 = synth.make_bank(6000, 640, 1, seed=78)
lab2 = synth.make_labels(256, 6000, seed=300 + rank)
qq2 = synth.make_queries(bank2, lab2, "S", seed=400 + rank).to(dev)
sh2 = spn.DeviceBank.from_tensor(bank2, dev, shard_rank=rank, shard_world=world)
l_s, dq_s = bank_ce_step_sharded(qq2, sh2, lab2.to(dev), tau)
l_f, dq_f = spn.bank_ce_step(qq2, spn.DeviceBank.from_tensor(bank2, dev), lab2.to(dev), tau)
assert abs(l_s.item() - l_f.item()) <= 1e-5 * max(1.0, abs(l_f.item())), (rank, l_s.item(), l_f.item())
assert float((dq_s - dq_f).norm() / dq_f.norm()) < 1e-4
# retrieval
qq, gal, _, refi, tgt, grp = synth.make_retrieval_case(70, 1000, 128, 1, 10, noise=6.0)
lo, hi = shard_bounds(1000, rank, world)
v, i = sim_topk_sharded(qq.to(dev), gal[lo:hi].to(dev), 50, lo, exclude=refi)
v1, i1 = spn.sim_topk(qq.to(dev), gal.to(dev), 50, exclude=refi)
assert torch.equal(i, i1) and torch.equal(v, v1)
r, g = sim_rank_sharded(qq.to(dev), gal[lo:hi].to(dev), tgt, lo, exclude=refi, group_members=grp)
r1, g1 = spn.sim_rank(qq.to(dev), gal.to(dev), tgt, exclude=refi, group=grp)
assert torch.equal(r, r1) and torch.equal(g, g1)
# the NCCL-free exchange (peer-mapped symmetric memory, reduce-scatter inside the sweep's epilogue) against the NCCL exchange
# and the single-GPU step: B % 128 == 0 selects it automatically; two-pass (D = 640) and single-pass (T = 32, D = 256) kernels
from spn4cir_b200 import dist as sdist
assert sdist.symm_path_ok(sh2, 256), "symmetric-memory path unavailable on this box"
assert len(sdist._exchanges) >= 1, "bank_ce_step_sharded did not take the symmetric-memory path"
os.environ["SPN_EXCHANGE"] = "nccl"
l_n, dq_n = bank_ce_step_sharded(qq2, sh2, lab2.to(dev), tau)
os.environ.pop("SPN_EXCHANGE")
assert abs(l_s.item() - l_n.item()) <= 1e-5 * max(1.0, abs(l_n.item())), (rank, l_s.item(), l_n.item())
assert float((dq_s - dq_n).norm() / dq_n.norm()) < 1e-4
# both forms of the NVLink exchange: push (TMA reduce-add from the sweep's epilogue into the home window) and pull (local
# accumulation, the home rank loads and sums its rows from every rank: a deterministic cross-rank sum)
assert sdist.symm_exchange_mode(256, world, sh2.N, 640) == "pull" and sdist.symm_exchange_mode(1024, 8, 125000, 768) == "push"
assert sdist.symm_exchange_mode(96, world, 5000, 256) == "pull"          # B % 128 != 0: only the pull form applies
for mode in ("push", "pull", "push", "pull"):  # alternating: the two forms share the windows and the barrier epoch
    os.environ["SPN_EXCHANGE"] = mode
    l_m, dq_m = bank_ce_step_sharded(qq2, sh2, lab2.to(dev), tau)
    os.environ.pop("SPN_EXCHANGE")
    assert abs(l_m.item() - l_n.item()) <= 1e-5 * max(1.0, abs(l_n.item())), (rank, mode, l_m.item(), l_n.item())
    assert float((dq_m - dq_n).norm() / dq_n.norm()) < 1e-4, (rank, mode)
    if mode == "pull":  # no atomics across ranks: bit-identical from step to step up to the local sweep's own atomics
        l_m2, dq_m2 = bank_ce_step_sharded(qq2, sh2, lab2.to(dev), tau)
        assert float((dq_m2 - dq_m).norm() / dq_m.norm()) < 1e-5
for it in range(3):  # repeated steps reuse the windows (zeroing + barriers between steps)
    l_r, dq_r = bank_ce_step_sharded(qq2, sh2, lab2.to(dev), tau)
    assert abs(l_r.item() - l_s.item()) <= 1e-6 * max(1.0, abs(l_s.item())) and float((dq_r - dq_s).norm() / dq_s.norm()) < 1e-5
tq = torch.tensor(tau, device=dev, requires_grad=True)
q5 = qq2.clone().requires_grad_(True)
l5 = bank_infonce_sharded(q5, sh2, lab2.to(dev), tq, label_smoothing=0.1)
l5.backward()
tf = torch.tensor(tau, device=dev, requires_grad=True)
q6 = qq2.clone().requires_grad_(True)
l6 = spn.bank_infonce(q6, spn.DeviceBank.from_tensor(bank2, dev), lab2.to(dev), tf, label_smoothing=0.1, single_pass=True)
l6.backward()
assert abs(l5.item() - l6.item()) <= 1e-5 * max(1.0, abs(l6.item())), (rank, l5.item(), l6.item())
assert float((q5.grad - q6.grad).norm() / q6.grad.norm()) < 1e-4
assert abs(tq.grad.item() - tf.grad.item()) <= 1e-3 * max(abs(tf.grad.item()), 1e-3), (tq.grad, tf.grad)
bank3 = synth.make_bank(1600, 256, 32, seed=79)                      # BLIP-2 shaped: whole entries per rank
lab3 = synth.make_labels(128, 1600, seed=500 + rank)
qq3 = synth.make_queries(bank3, lab3, "S", seed=600 + rank).to(dev)
sh3 = spn.DeviceBank.from_tensor(bank3, dev, shard_rank=rank, shard_world=world)
l_s3, dq_s3 = bank_ce_step_sharded(qq3, sh3, lab3.to(dev), 0.05)
l_f3, dq_f3 = spn.bank_ce_step(qq3, spn.DeviceBank.from_tensor(bank3, dev), lab3.to(dev), 0.05)
assert abs(l_s3.item() - l_f3.item()) <= 1e-5 * max(1.0, abs(l_f3.item())), (rank, l_s3.item(), l_f3.item())
assert float((dq_s3 - dq_f3).norm() / dq_f3.norm()) < 1e-4
for mode in ("push", "pull"):
    os.environ["SPN_EXCHANGE"] = mode
    l_m3, dq_m3 = bank_ce_step_sharded(qq3, sh3, lab3.to(dev), 0.05)
    os.environ.pop("SPN_EXCHANGE")
    assert abs(l_m3.item() - l_f3.item()) <= 1e-5 * max(1.0, abs(l_f3.item())), (rank, mode, l_m3.item(), l_f3.item())
    assert float((dq_m3 - dq_f3).norm() / dq_f3.norm()) < 1e-4, (rank, mode)
# cold temperature (the fused sweep's static reference is not provably safe below tau ~ 0.0092) and label smoothing over a
# multi-token bank: both are routed to the two-pass sharded path and still agree with the single-GPU result
for (qq_, sh_, full_, lab_, tau_, eps_) in ((qq2, sh2, spn.DeviceBank.from_tensor(bank2, dev), lab2, 0.008, 0.0),
                                            (qq3, sh3, spn.DeviceBank.from_tensor(bank3, dev), lab3, 0.05, 0.1)):
    qa = qq_.clone().requires_grad_(True)
    la = bank_infonce_sharded(qa, sh_, lab_.to(dev), tau_, label_smoothing=eps_)
    la.backward()
    qb = qq_.clone().requires_grad_(True)
    lb = spn.bank_infonce(qb, full_, lab_.to(dev), tau_, label_smoothing=eps_)
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-5 * max(1.0, abs(lb.item())), (rank, tau_, eps_, la.item(), lb.item())
    # (at tau = 0.008 the soft-max is so peaked that 1 - p_t ~ 1e-4: the fp32 rounding of the merged lse alone moves dq by
    # ~1e-3 relative -- same kernels, differently merged statistics)
    assert float((qa.grad - qb.grad).norm() / qb.grad.norm()) < (5e-3 if tau_ < 0.01 else 1e-4), (rank, tau_, eps_)
bad = lab2.clone(); bad[0] = 10 ** 7                                  # a label nobody owns: +inf, as on one GPU
l_bad, _ = bank_ce_step_sharded(qq2, sh2, bad.to(dev), tau)
assert torch.isinf(l_bad).item() and l_bad.item() > 0
dist.barrier()
if rank == 0:
    print("DIST_OK world=%d" % world)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_matches_single_gpu(tmp_path):
    n = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SPN_REPO=str(REPO))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                         capture_output=True, text=True, env=env, timeout=420)
    assert res.returncode == 0 and "DIST_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
