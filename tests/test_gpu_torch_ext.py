"""The torch.library operators (C++ autograd nodes over the C ABI) on the GPU: same numbers as the ctypes + Python
autograd.Function route (SPN_TORCH_EXT=0), as the oracle, and as the executed-reference fixtures; the reference call shape
`scaler.scale(model.infonce_loss(...)).backward()` (clip4cir/train_negplus.py:110-121) with fp16 / bf16 inputs."""
import numpy as np
import pytest
import torch

import spn4cir_b200 as spn
from conftest import rel_l2
from oracle import bank_ce as obank
from spn4cir_b200 import compat, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(q, bank, labels, tau, eps, dtype, ext, monkeypatch, scale=65536.0):
    monkeypatch.setenv("SPN_TORCH_EXT", "1" if ext else "0")
    qd = q.to(DEV, dtype).requires_grad_(True)
    tt = torch.tensor(tau, device=DEV, requires_grad=True)
    loss = spn.bank_infonce(qd, bank, labels.to(DEV), tt, label_smoothing=eps, single_pass=True)
    (loss * scale).backward()  # GradScaler-style upstream gradient
    return loss.detach().float().cpu(), (qd.grad.float() / scale).cpu(), (tt.grad / scale).cpu()


@pytest.mark.parametrize("B,N,D,T,eps,dtype", [
    (96, 3000, 128, 1, 0.0, torch.float32),
    (256, 5000, 640, 1, 0.0, torch.float16),
    (256, 5000, 640, 1, 0.1, torch.float32),
    (130, 2000, 768, 1, 0.0, torch.bfloat16),
    (64, 300, 256, 32, 0.0, torch.float32),
])
def test_operator_matches_python_route_and_oracle(B, N, D, T, eps, dtype, monkeypatch):
    q, bank, labels = synth.make_bank_case(B, N, D, T, "S", 51)
    tau = 0.03
    db = spn.DeviceBank.from_tensor(bank, DEV)
    l1, dq1, dt1 = run(q, db, labels, tau, eps, dtype, True, monkeypatch)
    l0, dq0, dt0 = run(q, db, labels, tau, eps, dtype, False, monkeypatch)
    assert abs(l1.item() - l0.item()) <= 1e-6 * max(1.0, abs(l0.item())), (l1, l0)
    assert rel_l2(dq1, dq0) < 1e-6 + (2e-3 if dtype != torch.float32 else 0.0), rel_l2(dq1, dq0)
    assert abs(dt1.item() - dt0.item()) <= 1e-5 * max(abs(dt0.item()), 1e-3)
    if T == 1 and eps == 0.0:
        o_loss, o_dq, o_dtau = obank.infonce_loss_and_grads(q, bank, labels, tau, dtype=torch.float64)
        assert abs(l1.item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
        assert rel_l2(dq1, o_dq) < 1e-2
        assert abs(dt1.item() - o_dtau.item()) <= 2e-2 * max(abs(o_dtau.item()), 1e-2)


def test_step_primitive_through_the_operator(monkeypatch):
    q, bank, labels = synth.make_bank_case(300, 7000, 512, 1, "S", 52)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    out = {}
    for ext in ("1", "0"):
        monkeypatch.setenv("SPN_TORCH_EXT", ext)
        out[ext] = spn.bank_ce_step(q.to(DEV), db, labels.to(DEV), 0.02, want_dtau=True, grad_scale=3.0)
    for a, b in zip(out["1"], out["0"]):
        assert rel_l2(a.reshape(-1), b.reshape(-1)) < 1e-6


@pytest.mark.parametrize("name", ["bank_step_plus", "bank_step_noplus"])
def test_fused_bank_large_step_against_executed_reference_and_eager(golden, name, monkeypatch):
    """bank_large_step as ONE operator (query assembly prologue + fused sweep + normalize-backward kernel)
    against the fixtures written by executing the reference method (clip4cir/models_negplus.py:130-142) and against the
    eager composition (gather + add + F.normalize in torch, autograd through the normalisation)."""
    from spn4cir_b200.compat import BankLossMixin
    g = golden(name)
    rb, tb, text, ix, rix, tix = synth.make_bank_step_case(int(g["seed"]), int(g["B"]), int(g["R"]), int(g["N"]), int(g["D"]))

    class Model(BankLossMixin):
        device, tau, plus = torch.device(DEV), float(g["tau"]), bool(int(g["plus"]))

    m = Model()
    m.set_banks(rb, tb)
    res = {}
    for ext in ("1", "0"):
        monkeypatch.setenv("SPN_TORCH_EXT", ext)
        assert m._fused_large_step_ok() == (ext == "1")
        tt = text.to(DEV).requires_grad_(True)
        loss = {}
        m.bank_large_step(loss, tt, ix.to(DEV), tix.to(DEV), rix.to(DEV))
        (loss["bank_loss"] * 128.0).backward()
        res[ext] = (loss["bank_loss"].item(), tt.grad.cpu() / 128.0)
    assert abs(res["1"][0] - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    assert rel_l2(res["1"][1], torch.from_numpy(g["dtext"])) < 2e-2
    assert abs(res["1"][0] - res["0"][0]) <= 1e-5 * max(1.0, abs(res["0"][0]))
    assert rel_l2(res["1"][1], res["0"][1]) < 2e-3, rel_l2(res["1"][1], res["0"][1])


def test_bank_file_loader_on_device(tmp_path):
    """SURVEY 8a-a10 on the GPU: torch.save([refer, target]) + the unlabeled concat -> resident bf16 bank; a step on the
    loaded bank equals a step on the same rows handed over as tensors; labels only point into the first M rows."""
    g = torch.Generator().manual_seed(9)
    refer = torch.randn(40, 128, generator=g)
    target = torch.nn.functional.normalize(torch.randn(900, 128, generator=g))
    unl = torch.nn.functional.normalize(torch.randn(600, 128, generator=g))
    torch.save([refer, target], tmp_path / "bank.pth")
    torch.save([unl], tmp_path / "unl.pth")

    class Model(compat.BankLossMixin):
        device, tau, plus = torch.device(DEV), 0.02, True

    m = Model()
    m.load_banks(str(tmp_path / "bank.pth"), str(tmp_path / "unl.pth"))
    assert m.M == 900 and m.target_bank.N == 1500 and m.refer_bank.shape == (40, 128)
    db = spn.DeviceBank.from_file(str(tmp_path / "bank.pth"), DEV, unlabeled_path=str(tmp_path / "unl.pth"))
    assert db.M == 900 and torch.equal(db.data, m.target_bank.data)
    assert torch.equal(db.data.cpu(), torch.cat([target, unl]).to(torch.bfloat16))
    text = 0.5 * torch.randn(32, 128, generator=g)
    idx = torch.randint(0, 40, (32,), generator=g)
    tix = torch.randint(0, m.M, (32,), generator=g)
    tt = text.to(DEV).requires_grad_(True)
    loss = {}
    m.bank_large_step(loss, tt, idx.to(DEV), tix.to(DEV), idx.to(DEV))
    loss["bank_loss"].backward()
    want = obank.bank_large_step_query(refer, idx, text)
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(want, torch.cat([target, unl]), tix, 0.02, dtype=torch.float64)
    assert abs(loss["bank_loss"].item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
    assert torch.isfinite(tt.grad).all()


def test_label_smoothing_multi_token_takes_the_exact_path():
    """with T = 32 the fused finalize cannot form sum_n s_n from column sums (that would sum every token
    row, not the per-entry max): such calls are routed to the two-pass path, and both entry points agree with the oracle."""
    q, bank, labels = synth.make_bank_case(48, 200, 256, 32, "S", 53)
    tau, eps = 0.05, 0.1
    db = spn.DeviceBank.from_tensor(bank, DEV)
    qd = q.to(DEV).requires_grad_(True)
    loss = spn.bank_infonce(qd, db, labels.to(DEV), tau, label_smoothing=eps)
    loss.backward()
    loss_s, dq_s = spn.bank_ce_step(q.to(DEV), db, labels.to(DEV), tau, label_smoothing=eps)
    qo = q.to(torch.bfloat16).double().requires_grad_(True)
    s = torch.einsum("bd,ntd->bnt", qo, bank.to(torch.bfloat16).double()).amax(-1) / tau
    ref = torch.nn.functional.cross_entropy(s, labels, label_smoothing=eps)
    ref.backward()
    for l, d in ((loss, qd.grad), (loss_s, dq_s)):
        assert abs(l.item() - ref.item()) <= 2e-4 * max(1.0, abs(ref.item())), (l.item(), ref.item())
        assert rel_l2(d, qo.grad) < 3e-3, rel_l2(d, qo.grad)


def test_query_assemble_rejects_bad_index_loudly():
    refer = torch.randn(10, 64, device=DEV)
    text = torch.randn(4, 64, device=DEV)
    idx = torch.tensor([0, 3, 10, -1], device=DEV)
    q16, q32, inv = spn.query_assemble(refer, idx, text)
    assert torch.isfinite(q32[:2]).all() and torch.isnan(q32[2:]).all() and torch.isnan(inv[2:]).all()


def test_device_bank_is_tensor_like_for_unchanged_reference_code():
    """Reference methods kept verbatim around the swapped loss (blip4cir/models.py:95-105, tgcir/models.py:272-282: their own
    fusion network, then `self.target_bank.detach().to(self.device)` and `infonce_loss`) must run with a resident bank."""
    q, bank, labels = synth.make_bank_case(64, 900, 128, 1, "S", 54)

    class Model(compat.BankLossMixin):
        device, tau, plus = torch.device(DEV), 0.02, False

        def bank_large_step(self, loss, query_feats, indexs, target_indexs, refer_indexs=None):  # tgcir-style body
            target_feats_all = self.target_bank.detach().to(self.device)
            target_indexs = target_indexs.to(self.device)
            target_feats = target_feats_all[target_indexs]  # the reference's (unused) gather
            assert target_feats.shape == (64, 128)
            loss["bank_loss"] = self.infonce_loss(query_feats, target_feats_all, target_indexs, tau=self.tau)

    m = Model()
    m.set_banks(torch.zeros(4, 128), bank)
    assert m.target_bank.shape == (900, 128) and len(m.target_bank) == 900
    qd = q.to(DEV).requires_grad_(True)
    loss = {}
    m.bank_large_step(loss, qd, None, labels)
    loss["bank_loss"].backward()
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(q, bank, labels, 0.02, dtype=torch.float64)
    assert abs(loss["bank_loss"].item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
    assert rel_l2(qd.grad, o_dq) < 1e-2
    np.testing.assert_allclose(m.target_bank[labels[:5]].cpu().numpy(), bank[labels[:5]].to(torch.bfloat16).float().numpy())
