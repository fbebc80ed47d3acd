"""Edge shapes of the hot path on the GPU: banks smaller than one tile (down to a single entry), a single query, strided
query views, a non-default stream, galleries smaller than K -- against a float64 evaluation of the reference formula
(`F.cross_entropy((q @ bank.T) / tau, labels)`, clip4cir/models_negplus.py:150-154; max over 32 tokens for BLIP-2,
blip2_qformer_cir_align_prompt.py:254-268) on the same bf16-rounded operands, and against `argsort` for retrieval
(clip4cir/validate.py:126-127, ties -> lower index)."""
import pytest
import torch

import spn4cir_b200 as spn
from spn4cir_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def reference_step(q, bank_rows, lab, tau, T):
    q = q.detach().double().requires_grad_(True)
    s = q @ bank_rows.double().T
    if T > 1:
        s = s.view(q.shape[0], -1, T).max(-1).values
    loss = torch.nn.functional.cross_entropy(s / tau, lab)
    loss.backward()
    return loss.item(), q.grad.float()


@pytest.mark.parametrize("single_pass", [True, False], ids=["fused", "two_pass"])
@pytest.mark.parametrize("B,N,D,T,tau", [
    (3, 1, 64, 1, 0.05),        # one bank entry: loss 0, gradient 0
    (3, 2, 64, 1, 0.05),
    (5, 7, 128, 1, 0.05),
    (130, 127, 640, 1, 0.02),   # one partial bank tile, two query tiles
    (2, 1, 256, 32, 0.05),      # one BLIP-2 entry (32 token rows)
    (4, 3, 256, 32, 0.05),
    (1, 5000, 768, 1, 0.02),    # a single query
])
def test_tiny_banks_and_batches(B, N, D, T, tau, single_pass):
    bank = synth.make_bank(N, D, T, seed=N + D)
    lab = synth.make_labels(B, N, seed=B)
    q = synth.make_queries(bank, lab, "S", seed=B + 1)
    rows16 = bank.to(torch.bfloat16).float().view(N * T, D)
    ref_loss, ref_dq = reference_step(q.to(torch.bfloat16).float(), rows16, lab, tau, T)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    qd = q.to(DEV).requires_grad_(True)
    loss = spn.bank_infonce(qd, db, lab.to(DEV), tau, single_pass=single_pass)
    loss.backward()
    dq = qd.grad.cpu()
    assert abs(loss.item() - ref_loss) <= 2e-3 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    scale = (1.0 / tau) / B   # size of one query's gradient when its soft-max is not peaked
    assert float((dq - ref_dq).norm()) <= 5e-3 * float(ref_dq.norm()) + 1e-5 * scale, (float((dq - ref_dq).norm()), float(ref_dq.norm()))


def test_strided_queries_and_side_stream():
    bank = synth.make_bank(1000, 128, 1, seed=3)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    lab = synth.make_labels(64, 1000, seed=4)
    q = synth.make_queries(bank, lab, "S", seed=5).to(DEV)
    lab = lab.to(DEV)
    base, dq_base = spn.bank_ce_step(q, db, lab, 0.02)
    col_major = q.t().contiguous().t()
    wide = torch.zeros(64, 256, device=DEV)
    wide[:, ::2] = q
    for view in (col_major, wide[:, ::2]):
        assert not view.is_contiguous()
        l, d = spn.bank_ce_step(view, db, lab, 0.02)
        assert abs(l.item() - base.item()) < 1e-6 and float((d - dq_base).norm() / dq_base.norm()) < 1e-5
        qv = view.detach().requires_grad_(True)
        spn.bank_infonce(qv, db, lab, 0.02).backward()
        assert float((qv.grad - dq_base).norm() / dq_base.norm()) < 1e-5
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        l, d = spn.bank_ce_step(q, db, lab, 0.02)
    side.synchronize()
    assert abs(l.item() - base.item()) < 1e-6 and float((d - dq_base).norm() / dq_base.norm()) < 1e-5


@pytest.mark.parametrize("Q,G,D,K", [(1, 1, 64, 1), (3, 2, 64, 5), (5, 70, 640, 50), (130, 3, 128, 3)])
def test_tiny_galleries(Q, G, D, K):
    g = torch.Generator().manual_seed(Q * 1000 + G)
    qq = torch.nn.functional.normalize(torch.randn(Q, D, generator=g))
    gg = torch.nn.functional.normalize(torch.randn(G, D, generator=g))
    order = torch.argsort(-(qq @ gg.T), dim=-1, stable=True)
    _, idx = spn.sim_topk(qq.to(DEV), gg.to(DEV), K)
    kk = min(K, G)
    assert torch.equal(idx[:, :kk].cpu().long(), order[:, :kk])
    if K > G:
        assert (idx[:, G:] == -1).all()   # fewer candidates than K: padded
    tgt = torch.randint(0, G, (Q,), generator=g, dtype=torch.int32)
    rank, _ = spn.sim_rank(qq.to(DEV), gg.to(DEV), tgt.to(DEV))
    assert torch.equal(rank.cpu().long(), (order == tgt.long()[:, None]).float().argmax(-1))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_non_current_device():
    """Everything runs on the device of its inputs, not on the current device (one process driving two GPUs)."""
    assert torch.cuda.current_device() == 0
    bank = synth.make_bank(3000, 640, 1, seed=11)
    lab = synth.make_labels(200, 3000, seed=12)
    q = synth.make_queries(bank, lab, "S", seed=13)
    out = {}
    for d in ("cuda:0", "cuda:1"):
        db = spn.DeviceBank.from_tensor(bank, d)
        qd = q.to(d).requires_grad_(True)
        loss = spn.bank_infonce(qd, db, lab.to(d), 0.02)
        loss.backward()
        l2, dq2 = spn.bank_ce_step(q.to(d), db, lab.to(d), 0.02)
        gal = torch.nn.functional.normalize(bank[:500].float(), dim=-1).to(d)
        qq = torch.nn.functional.normalize(q[:50], dim=-1).to(d)
        _, idx = spn.sim_topk(qq, gal, 10)
        _, idx_tc = spn.sim_topk(torch.nn.functional.normalize(q, dim=-1).to(d), gal, 10, mode="tc")
        torch.cuda.synchronize(d)
        assert qd.grad.device == torch.device(d) and dq2.device == torch.device(d) and idx.device == torch.device(d)
        out[d] = (loss.item(), qd.grad.cpu(), l2.item(), idx.cpu(), idx_tc.cpu())
    a, b = out["cuda:0"], out["cuda:1"]
    assert abs(a[0] - b[0]) < 1e-6 and abs(a[2] - b[2]) < 1e-6
    assert float((a[1] - b[1]).norm() / a[1].norm()) < 1e-5
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    assert torch.cuda.current_device() == 0
