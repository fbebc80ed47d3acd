"""Row-sharded bank on ONE GPU ("fake ranks"): the per-rank kernels of the sharded step (spn_bank_ce_prep / _fused,
spn_shard_pack, spn_shard_fold, spn_bank_ce_finalize with the folded target rows) with the collectives replaced by
explicit sums / minima, against the un-sharded step, the oracle, and -- at BASELINE sizes -- size-independent properties.

The NCCL version of the same exchange is tests/test_gpu_dist.py (>= 2 GPUs) and, over gloo with a CPU emulation of the
kernels, tests/test_dist_gloo.py."""
import numpy as np
import pytest
import torch

import spn4cir_b200 as spn
from conftest import rel_l2
from oracle import bank_ce as obank
from spn4cir_b200 import synth
from spn4cir_b200.dist import _CudaShardCompute

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def fake_rank_step(q_all, labels, shards, tau, eps, tau_grad=True):
    """All queries against every shard on one device; min / sum stand in for the MIN all-reduce, the SUM all-reduce and
    the reduce-scatter.  Returns (loss, dq_unit, dlogit_dot)."""
    c = _CudaShardCompute()
    inv_tau = 1.0 / tau
    B, D = q_all.shape
    q16 = c.cast(q_all)
    maxnorm = torch.stack([s.stats()[0] for s in shards]).max().reshape(1)
    colsum = torch.stack([s.stats(True)[1] for s in shards]).sum(0) if eps != 0.0 else None
    refs = [c.prep(q16, labels, s, inv_tau, maxnorm) for s in shards]
    ref = torch.stack(refs).min(0).values.contiguous()
    outs = [c.fused(q16, labels, ref, s, inv_tau, tau_grad) for s in shards]
    vec = torch.stack([c.pack(o[1], o[2], o[3], o[4]) for o in outs]).sum(0).contiguous()
    st = None
    for o, s in zip(outs, shards):
        st = c.fold(o[0], vec, ref, o[4], s, eps)
    o_sum = torch.stack([o[0] for o in outs]).sum(0).contiguous()
    return c.finalize(o_sum, vec[0], st, ref, vec[2] if tau_grad else None, q16, colsum, shards[0], inv_tau, B, eps)


@pytest.mark.parametrize("B,N,D,R,eps,seed", [
    (96, 4000, 128, 8, 0.0, 14),
    (96, 4000, 128, 8, 0.1, 14),      # label smoothing over shards (zscir/models_bank.py:35)
    (256, 5000, 640, 4, 0.0, 15),     # two-pass pair kernel per shard
    (300, 3001, 768, 3, 0.05, 16),    # ragged shards, padded query tiles
])
def test_fake_ranks_match_unsharded_step_and_oracle(B, N, D, R, eps, seed):
    q, bank, labels = synth.make_bank_case(B, N, D, 1, "S", seed)
    tau = 0.02
    qd, lab = q.to(DEV), labels.to(DEV)
    shards = [spn.DeviceBank.from_tensor(bank, DEV, shard_rank=r, shard_world=R) for r in range(R)]
    loss, dq, ddot = fake_rank_step(qd, lab, shards, tau, eps)
    full = spn.DeviceBank.from_tensor(bank, DEV)
    loss1, dq1, dtau1 = spn.bank_ce_step(qd, full, lab, tau, label_smoothing=eps, want_dtau=True)
    assert abs(loss.item() - loss1.item()) <= 2e-6 * max(1.0, abs(loss1.item())), (loss.item(), loss1.item())
    assert rel_l2(dq, dq1) < 2e-5, rel_l2(dq, dq1)
    dtau = -(1.0 / tau / B) * ddot.sum().item()
    assert abs(dtau - dtau1.item()) <= 1e-3 * max(abs(dtau1.item()), 1e-2), (dtau, dtau1.item())
    # oracle on the bf16-rounded operands (the kernel arithmetic), fp64
    qo = bf16_round(q).double().requires_grad_(True)
    dense = torch.nn.functional.cross_entropy(qo @ bf16_round(bank).double().T / tau, labels, label_smoothing=eps)
    dense.backward()
    assert abs(loss.item() - dense.item()) <= 2e-4 * max(1.0, abs(dense.item())), (loss.item(), dense.item())
    assert rel_l2(dq, qo.grad) < 3e-3, rel_l2(dq, qo.grad)
    if eps == 0.0:
        o_loss, o_dq, _ = obank.infonce_loss_and_grads(q, bank, labels, tau, dtype=torch.float64)
        assert abs(loss.item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
        assert rel_l2(dq, o_dq) < 1e-2


def test_label_nobody_owns_gives_inf():
    q, bank, labels = synth.make_bank_case(64, 1000, 128, 1, "S", 17)
    labels[3] = 5000
    shards = [spn.DeviceBank.from_tensor(bank, DEV, shard_rank=r, shard_world=2) for r in range(2)]
    loss, _, _ = fake_rank_step(q.to(DEV), labels.to(DEV), shards, 0.02, 0.0)
    assert torch.isinf(loss).item() and loss.item() > 0


def dense_fp32(q, rows_bf16, labels, tau, T=1, chunk=1 << 16):
    """dense torch evaluation on the GPU in fp32 from the SAME bf16 rows, chunked over the bank (online log-sum-exp):
    (loss, dq) -- the size-independent 'plain fp32' check of the floating-point kernel.  T > 1: the logit of an entry is
    the max over its T token rows and the gradient flows through the first arg-max token
    (blip2_qformer_cir_align_prompt.py:257-260)."""
    B, D = q.shape
    dev = q.device
    qf = q.to(torch.bfloat16).float()
    m = torch.full((B,), -float("inf"), device=dev)
    l = torch.zeros(B, device=dev)
    st = torch.zeros(B, device=dev)
    o = torch.zeros((B, D), device=dev)
    tgt = torch.zeros((B, D), device=dev)
    for r0 in range(0, rows_bf16.shape[0], chunk * T):
        blk = rows_bf16[r0:r0 + chunk * T].float()
        s = (qf @ blk.T) / tau
        e0 = r0 // T
        if T > 1:
            s, arg = s.view(B, -1, T).max(-1)                      # [B, n] logits, arg-max token per (query, entry)
            g = blk.view(-1, T, D)
        m_new = torch.maximum(m, s.max(-1).values)
        p = torch.exp(s - m_new[:, None])
        a = torch.exp(m - m_new)
        l = l * a + p.sum(-1)
        if T > 1:
            acc = torch.zeros((B, D), device=dev)
            for t in range(T):                                     # sum_n p[b, n] [arg == t] bank[n, t, :]
                acc += (p * (arg == t)) @ g[:, t, :]
            o = o * a[:, None] + acc
        else:
            o = o * a[:, None] + p @ blk
        m = m_new
        inb = (labels >= e0) & (labels < e0 + s.shape[1])
        if bool(inb.any()):
            jj = labels[inb] - e0
            rows_i = torch.nonzero(inb).squeeze(1)
            st[inb] = s[rows_i, jj]
            tgt[inb] = g[jj, arg[rows_i, jj]] if T > 1 else blk[jj]
    lse = m + torch.log(l)
    loss = (lse - st).mean()
    dq = (o / l[:, None] - tgt) / (tau * B)
    return loss, dq


@pytest.mark.parametrize("Bt,n_shard,D", [(4096, 12500, 640), (2048, 100, 640), (1024, 900, 768), (8192, 125000, 768)])
def test_sharded_sweep_shapes_that_hung_in_round_1(Bt, n_shard, D):
    """Regression for the C3 x 8 bank-sharded run that never finished in round 1 (VERDICT weak #2): its per-rank kernel
    shape (4096 queries x 12 500 rows = 98 bank tiles for 16 query-tile pairs x 4 chunks) and shapes with fewer bank tiles
    than CTA pairs, as one rank of a sharded job sees them (row_base != 0, most labels elsewhere)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    rows = torch.nn.functional.normalize(torch.randn((n_shard, D), generator=g, device=DEV), dim=-1).to(torch.bfloat16)
    row_base, n_total = 3 * n_shard, 8 * n_shard
    labels = torch.randint(0, n_total, (Bt,), generator=g, device=DEV)
    q = torch.nn.functional.normalize(torch.randn((Bt, D), generator=g, device=DEV), dim=-1)
    shard = spn.DeviceBank(rows, 1, row_base=row_base, n_total=n_total)
    c = _CudaShardCompute()
    q16 = c.cast(q)
    ref = c.prep(q16, labels, shard, 50.0, shard.stats()[0])
    o_acc, l_acc, s_t, _, lbl_row = c.fused(q16, labels, ref, shard, 50.0, False)
    torch.cuda.synchronize()
    own = (labels >= row_base) & (labels < row_base + n_shard)
    assert torch.equal(lbl_row >= 0, own)
    s = 50.0 * (q16.float() @ rows.float().T)
    p = torch.exp(s - ref[:, None])
    idx = torch.nonzero(own).squeeze(1)
    p[idx, labels[own] - row_base] = 0.0
    assert rel_l2(l_acc, p.sum(-1)) < 1e-4
    assert rel_l2(o_acc, p.to(torch.bfloat16).float() @ rows.float()) < 2e-3
    np.testing.assert_allclose(s_t[own].cpu().numpy(), s[idx, labels[own] - row_base].cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_c5_shard_full_size_properties():
    """BASELINE C5 (1024 queries x 1 M x 768) at full size: the bank in eighths (the rows 8 ranks would own), eight shard
    sweeps merged exactly like the exchange does (min of the references, sums of the partials, owner fold, finalize)
    against (a) a dense fp32 torch evaluation of the whole 1 M x 768 bank on the GPU (loss and dq) and (b) the un-sharded
    single-GPU step on the concatenated bank.  (One rank's kernel shape at C5 x 8 -- 8192 queries x 125 000 rows -- is in
    test_sharded_sweep_shapes_that_hung_in_round_1.)"""
    cfg = synth.CONFIGS["C5"]
    B, N, D, tau = cfg["B"], cfg["N"], cfg["D"], cfg["tau"]
    R = 8
    shards, q, labels = [], None, None
    for r in range(R):
        rows, rb, q_r, lab_r = synth.make_device_case(B, N, D, 1, torch.device(DEV), rank=r, world=R, sharded=True)
        shards.append(spn.DeviceBank(rows, 1, row_base=rb, n_total=N))
        if r == 0:
            q, labels = q_r, lab_r
    loss, dq, _ = fake_rank_step(q, labels, shards, tau, 0.0, tau_grad=False)
    all_rows = torch.cat([s.data for s in shards])
    d_loss, d_dq = dense_fp32(q, all_rows, labels, tau)
    assert abs(loss.item() - d_loss.item()) <= 2e-4 * max(1.0, abs(d_loss.item())), (loss.item(), d_loss.item())
    assert rel_l2(dq, d_dq) < 5e-3, rel_l2(dq, d_dq)
    full = spn.DeviceBank(all_rows, 1)
    loss1, dq1 = spn.bank_ce_step(q, full, labels, tau)
    assert abs(loss1.item() - d_loss.item()) <= 2e-4 * max(1.0, abs(d_loss.item())), (loss1.item(), d_loss.item())
    assert rel_l2(dq1, d_dq) < 5e-3, rel_l2(dq1, d_dq)
    assert rel_l2(dq, dq1) < 1e-4, rel_l2(dq, dq1)


def test_c4_full_size_properties():
    """BASELINE C4 (BLIP-2 stage 2: 256 queries x 200 000 entries x 32 tokens x 256, 3.28 GB of bf16 rows): the fused step
    and the two-pass path against each other and against a dense fp32 torch evaluation (max over the 32 token rows, the
    gradient through the first arg-max token) of the same bf16 rows on the GPU."""
    cfg = synth.CONFIGS["C4"]
    B, N, D, T, tau = cfg["B"], cfg["N"], cfg["D"], cfg["T"], cfg["tau"]
    rows, _, q, labels = synth.make_device_case(B, N, D, T, torch.device(DEV))
    bank = spn.DeviceBank(rows, T)
    loss, dq, dtau = spn.bank_ce_step(q, bank, labels, tau, want_dtau=True)
    qd = q.clone().requires_grad_(True)
    tt = torch.tensor(tau, device=DEV, requires_grad=True)
    loss2 = spn.bank_infonce(qd, bank, labels, tt, single_pass=False)
    loss2.backward()
    assert abs(loss.item() - loss2.item()) <= 1e-4 * max(1.0, abs(loss2.item())), (loss.item(), loss2.item())
    assert rel_l2(dq, qd.grad) < 5e-3, rel_l2(dq, qd.grad)
    assert abs(dtau.item() - tt.grad.item()) <= 2e-2 * max(abs(tt.grad.item()), 1e-2), (dtau.item(), tt.grad.item())
    d_loss, d_dq = dense_fp32(q, rows, labels, tau, T=T, chunk=1 << 13)
    assert abs(loss.item() - d_loss.item()) <= 2e-4 * max(1.0, abs(d_loss.item())), (loss.item(), d_loss.item())
    assert rel_l2(dq, d_dq) < 5e-3, rel_l2(dq, d_dq)
