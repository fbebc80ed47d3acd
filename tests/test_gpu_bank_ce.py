"""Parity of the CUDA bank-CE path (through the C ABI) against the oracle and the reference-generated fixtures.

Tolerances (north_star): loss and dL/dq within 1e-2 relative of the fp32 reference (bf16 operands, fp32
accumulation).  Against the oracle fed with the SAME bf16-rounded operands the kernel arithmetic itself is
checked much tighter (2e-3 on dq: P is rounded to bf16 before the second contraction)."""
import math

import numpy as np
import pytest
import torch

import spn4cir_b200 as spn
from conftest import rel_l2
from oracle import bank_ce as obank
from spn4cir_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16_round(x):
    return x.to(torch.bfloat16).float()


PATHS = pytest.mark.parametrize("single_pass", [False, True], ids=["two_pass", "fused"])


def run_ours(q, bank, labels, tau, eps=0.0, grad_out=1.0, tau_tensor=False, single_pass=None):
    qd = q.to(DEV).requires_grad_(True)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    t = torch.tensor(tau, device=DEV, requires_grad=True) if tau_tensor else tau
    loss = spn.bank_infonce(qd, db, labels.to(DEV), t, label_smoothing=eps, single_pass=single_pass)
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), qd.grad.detach().cpu(), (t.grad.detach().cpu() if tau_tensor else None)


CASES = [
    # B, N, D, mode, tau, seed
    (16, 333, 128, "S", 0.02, 1),       # ragged N, ragged B, label in partial tile possible
    (8, 1000, 64, "U", 0.05, 2),
    (128, 4096, 256, "S", 0.03, 3),     # blip4cir dim
    (130, 5000, 512, "S", 0.01, 4),     # tgcir dim, B not a tile multiple, tau = 0.01
    (64, 3000, 640, "U", 0.02, 5),      # clip RN50x4 dim (2 d-splits in backward)
    (64, 2500, 768, "S", 0.02, 6),      # ViT-L/14 dim (3 d-splits)
    (1, 129, 64, "S", 0.02, 7),         # single query
    (257, 128, 64, "U", 0.02, 8),       # exactly one bank tile, 3 query tiles
]


@PATHS
@pytest.mark.parametrize("B,N,D,mode,tau,seed", CASES)
def test_loss_and_grad_match_oracle(B, N, D, mode, tau, seed, single_pass):
    q, bank, labels = synth.make_bank_case(B, N, D, 1, mode, seed)
    loss, dq, dtau = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=single_pass)
    # (i) end-to-end vs the fp32 oracle on the original operands
    o_loss, o_dq, o_dtau = obank.infonce_loss_and_grads(q, bank, labels, tau, dtype=torch.float64)
    assert abs(loss.item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3), (loss, o_loss)
    assert rel_l2(dq, o_dq) < 1e-2, rel_l2(dq, o_dq)
    assert abs(dtau.item() - o_dtau.item()) <= 2e-2 * max(abs(o_dtau.item()), 1e-2), (dtau, o_dtau)
    # (ii) kernel arithmetic vs the oracle on bf16-rounded operands
    r_loss, r_dq, _ = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, tau, dtype=torch.float64)
    assert abs(loss.item() - r_loss.item()) <= 2e-4 * max(abs(r_loss.item()), 1.0), (loss, r_loss)
    assert rel_l2(dq, r_dq) < 3e-3, rel_l2(dq, r_dq)


PAIR_CASES = [
    # B (even number of 128-query tiles -> cta_group::2 pairs), N, D, T, tau, seed
    (256, 1000, 64, 1, 0.05, 21),     # one k-block: half-filled GEMM1 stage, one narrow (2 x 32 column) dq group
    (256, 777, 128, 1, 0.03, 22),     # one wide dq group, ragged last tile
    (200, 3000, 256, 1, 0.03, 23),    # second tile of the pair is partly empty (B < 256)
    (512, 2100, 320, 1, 0.02, 24),    # odd number of k-blocks and of dq blocks
    (256, 5000, 512, 1, 0.02, 25),    # 2 d-splits of 256
    (256, 6000, 640, 1, 0.02, 26),    # 2 d-splits of 320 = 128 + 128 + 64
    (256, 4000, 768, 1, 0.02, 27),    # 3 d-splits of 256
    (256, 300, 256, 32, 0.05, 28),    # BLIP-2 max over tokens
    (256, 3000, 384, 1, 0.02, 29),    # two-pass fused: X = 128, one 256-column group in pass 2
    (512, 9000, 448, 1, 0.03, 30),    # two-pass fused: X = 192 (odd block count in pass 1)
    (256, 2000, 576, 1, 0.02, 31),    # two-pass fused: X = 320, one group
    (512, 20000, 704, 1, 0.02, 32),   # two-pass fused: X = 192, two groups, several tiles per CTA
    (300, 5000, 640, 1, 0.02, 33),    # 3 query tiles: padded to 2 pairs (the 4th tile is empty)
    (600, 3000, 256, 1, 0.03, 34),    # 5 query tiles -> 3 pairs, single-pass shape
]


@PATHS
@pytest.mark.parametrize("B,N,D,T,tau,seed", PAIR_CASES)
def test_cta_pair_kernels_match_oracle_and_single_cta(B, N, D, T, tau, seed, single_pass, monkeypatch):
    """The cta_group::2 kernels (rank 0 issues M = 256 MMAs for a pair of query tiles) against the oracle and against
    the single-CTA kernels on the same inputs (SPN_DEBUG_CG=1 forces those)."""
    q, bank, labels = synth.make_bank_case(B, N, D, T, "S", seed)
    loss, dq, dtau = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=single_pass)
    monkeypatch.setenv("SPN_DEBUG_CG", "1")
    loss1, dq1, dtau1 = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=single_pass)
    monkeypatch.delenv("SPN_DEBUG_CG")
    assert abs(loss.item() - loss1.item()) <= 1e-5 * max(abs(loss1.item()), 1e-3), (loss, loss1)
    assert rel_l2(dq, dq1) < 1e-4, rel_l2(dq, dq1)
    assert abs(dtau.item() - dtau1.item()) <= 1e-3 * max(abs(dtau1.item()), 1e-2), (dtau, dtau1)
    if T == 1:
        r_loss, r_dq, _ = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, tau, dtype=torch.float64)
    else:
        r_loss, r_dq, _ = obank.blip2_loss_and_grads(bf16_round(q), bf16_round(bank), labels, tau, dtype=torch.float64)
    assert abs(loss.item() - r_loss.item()) <= 2e-4 * max(abs(r_loss.item()), 1.0), (loss, r_loss)
    assert rel_l2(dq, r_dq) < 3e-3, rel_l2(dq, r_dq)


@pytest.mark.parametrize("B,N,D,tau,seed", [(512, 40000, 640, 0.02, 41), (256, 30000, 768, 0.02, 42), (300, 7000, 512, 0.05, 43)])
def test_fused_two_pass_matches_split_recompute(B, N, D, tau, seed, monkeypatch):
    """The two-pass fused kernel (P~ through the global scratch, second pass = plain GEMM) against the D-split kernel
    that recomputes S per split (SPN_DEBUG_TWO_PASS=0): same bf16 P~, same fp32 accumulation up to summation order."""
    q, bank, labels = synth.make_bank_case(B, N, D, 1, "S", seed)
    loss, dq, dtau = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=True)
    monkeypatch.setenv("SPN_DEBUG_TWO_PASS", "0")
    loss1, dq1, dtau1 = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=True)
    monkeypatch.delenv("SPN_DEBUG_TWO_PASS")
    assert abs(loss.item() - loss1.item()) <= 1e-5 * max(abs(loss1.item()), 1e-3), (loss, loss1)
    assert rel_l2(dq, dq1) < 1e-5, rel_l2(dq, dq1)
    assert abs(dtau.item() - dtau1.item()) <= 1e-4 * max(abs(dtau1.item()), 1e-2), (dtau, dtau1)
    r_loss, r_dq, _ = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, tau, dtype=torch.float64)
    assert abs(loss.item() - r_loss.item()) <= 2e-4 * max(abs(r_loss.item()), 1.0), (loss, r_loss)
    assert rel_l2(dq, r_dq) < 3e-3, rel_l2(dq, r_dq)


@PATHS
@pytest.mark.parametrize("name", ["small_ragged", "small_uniform", "c1", "c1_uniform"])
def test_against_reference_fixtures(golden, name, single_pass):
    g = golden(f"bank_ce_{name}")
    q, bank, labels = synth.make_bank_case(int(g["B"]), int(g["N"]), int(g["D"]), 1, str(g["mode"]), int(g["seed"]))
    loss, dq, _ = run_ours(q, bank, labels, float(g["tau"]), single_pass=single_pass)
    ref_loss, ref_dq = float(g["loss"]), torch.from_numpy(g["dq"])
    assert abs(loss.item() - ref_loss) <= 1e-2 * max(abs(ref_loss), 1e-3), (loss.item(), ref_loss)
    assert rel_l2(dq, ref_dq) < 1e-2, rel_l2(dq, ref_dq)


def test_default_labels_and_upstream_gradient(golden):
    g = golden("bank_ce_nolabels")
    q, bank, _ = synth.make_bank_case(int(g["B"]), int(g["N"]), int(g["D"]), 1, "S", int(g["seed"]))
    qd = q.to(DEV).requires_grad_(True)
    loss = spn.compat.infonce_loss(qd, bank.to(DEV), None, tau=float(g["tau"]))
    (loss * 65536.0).backward()  # GradScaler-style scalar upstream gradient (train_negplus.py:121)
    assert abs(loss.item() - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    _, o_dq, _ = obank.infonce_loss_and_grads(q, bank, torch.arange(q.shape[0]), float(g["tau"]), grad_out=65536.0)
    assert rel_l2(qd.grad.cpu(), o_dq) < 1e-2


def test_input_dtypes_fp16_bf16():
    q, bank, labels = synth.make_bank_case(32, 700, 128, 1, "S", 9)
    ref, _, _ = obank.infonce_loss_and_grads(q, bank, labels, 0.02)
    for dt in (torch.float16, torch.bfloat16):
        qd = q.to(DEV, dt).requires_grad_(True)
        loss = spn.bank_infonce(qd, bank.to(DEV, dt), labels.to(DEV), 0.02)
        loss.backward()
        assert qd.grad.dtype == dt and loss.dtype == torch.float32
        assert abs(loss.item() - ref.item()) <= 2e-2 * abs(ref.item())


@PATHS
def test_label_smoothing(single_pass):
    q, bank, labels = synth.make_bank_case(40, 900, 128, 1, "U", 10)
    tau, eps = 0.05, 0.1
    loss, dq, _ = run_ours(q, bank, labels, tau, eps=eps, single_pass=single_pass)
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(q, bank, labels, tau, label_smoothing=eps, dtype=torch.float64)
    assert abs(loss.item() - o_loss.item()) <= 1e-2 * abs(o_loss.item())
    assert rel_l2(dq, o_dq) < 1e-2


@PATHS
@pytest.mark.parametrize("B,N,seed", [(6, 50, 11), (130, 1000, 12), (32, 37, 13)])
def test_blip2_max_over_tokens(B, N, seed, single_pass):
    T, D, temp = 32, 256, 0.05
    f, bank3, labels = synth.make_bank_case(B, N, D, T, "S", seed)
    fd = f.to(DEV).requires_grad_(True)
    tp = torch.nn.Parameter(torch.tensor(temp, device=DEV))
    out = {"loss_qtc": spn.bank_infonce(fd, bank3.to(DEV), labels.to(DEV), tp, single_pass=single_pass)}
    out["loss_qtc"].backward()
    o_loss, o_df, o_dt = obank.blip2_loss_and_grads(f, bank3, labels, temp, dtype=torch.float64)
    assert abs(out["loss_qtc"].item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
    assert rel_l2(fd.grad.cpu(), o_df) < 1.5e-2, rel_l2(fd.grad.cpu(), o_df)
    assert abs(tp.grad.item() - o_dt.item()) <= 3e-2 * max(abs(o_dt.item()), 1e-2)
    r_loss, r_df, _ = obank.blip2_loss_and_grads(bf16_round(f), bf16_round(bank3), labels, temp, dtype=torch.float64)
    assert abs(out["loss_qtc"].item() - r_loss.item()) <= 2e-4 * max(abs(r_loss.item()), 1.0)
    assert rel_l2(fd.grad.cpu(), r_df) < 3e-3


def test_blip2_fixture(golden):
    g = golden("blip2_small")
    f, bank3, labels = synth.make_bank_case(int(g["B"]), int(g["N"]), int(g["D"]), int(g["T"]), "S", int(g["seed"]))
    fd = f.to(DEV).requires_grad_(True)
    loss = spn.bank_infonce(fd, bank3.to(DEV), labels.to(DEV), float(g["temp"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    assert rel_l2(fd.grad.cpu(), torch.from_numpy(g["df"])) < 1.5e-2


@pytest.mark.parametrize("name", ["blip2_ref_small", "blip2_ref_ragged"])
def test_blip2_against_executed_reference_lines(golden, name):
    """The CUDA path (T = 32: max over token rows, learnable temp) against the fixtures produced by executing the
    reference's own loss loop (blip2_qformer_cir_align_prompt.py:254-265) and inference tail (:314-318)."""
    from spn4cir_b200 import compat
    g = golden(name)
    f, bank3, labels = synth.make_bank_case(int(g["B"]), int(g["N"]), int(g["D"]), int(g["T"]), "S", int(g["seed"]))
    fd = f.to(DEV).requires_grad_(True)
    tp = torch.tensor(float(g["temp"]), device=DEV, requires_grad=True)
    out = compat.forward_stage2_loss(fd, bank3.to(DEV), labels.to(DEV), tp)
    out["loss_qtc"].backward()
    assert abs(out["loss_qtc"].item() - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))          # bf16 vs the reference's fp32
    assert rel_l2(fd.grad.cpu(), torch.from_numpy(g["df"])) < 1.5e-2
    # (bf16 rounding of the inputs alone moves d loss / d temp by 1.5 % on the ragged case: CPU emulation)
    assert abs(tp.grad.item() - float(g["dtemp"])) <= 6e-2 * max(abs(float(g["dtemp"])), 1e-2)
    sim = compat.inference_sim(f.to(DEV), bank3.to(DEV)).cpu().numpy()                               # exact fp32 chain
    np.testing.assert_allclose(sim, g["sim_i2t"], rtol=1e-5, atol=2e-6)


def test_shards_merge_to_the_unsharded_result():
    """8 'fake ranks' on one GPU: per-shard statistics merged by spn_merge_stats, dq accumulated shard by shard."""
    q, bank, labels = synth.make_bank_case(96, 4000, 128, 1, "S", 14)
    tau, R = 0.02, 8
    full = spn.DeviceBank.from_tensor(bank, DEV)
    st_full = spn.bank_ce_stats(q.to(DEV), full, labels.to(DEV), tau)
    shards = [spn.DeviceBank.from_tensor(bank, DEV, shard_rank=r, shard_world=R) for r in range(R)]
    assert sum(s.N for s in shards) == 4000 and shards[3].row_base == 1500
    parts = torch.stack([spn.bank_ce_stats(q.to(DEV), s, labels.to(DEV), tau) for s in shards])
    merged = spn.merge_stats(parts)
    lse_m = merged[0] + torch.log(merged[1])
    lse_f = st_full[0] + torch.log(st_full[1])
    np.testing.assert_allclose(lse_m.cpu().numpy(), lse_f.cpu().numpy(), rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(merged[2].cpu().numpy(), st_full[2].cpu().numpy(), rtol=1e-6, atol=1e-6)
    m, l, st, _ = obank.shard_stats(bf16_round(q), bf16_round(bank), labels, tau)
    np.testing.assert_allclose(lse_m.cpu().numpy(), (m + torch.log(l)).numpy(), rtol=1e-4, atol=1e-3)
    # backward: shard-by-shard accumulation equals the single call
    from spn4cir_b200.ops import _bwd, cast_to_bf16
    q16 = cast_to_bf16(q.to(DEV))
    lab = labels.to(DEV)
    dq_full = torch.zeros((96, 128), device=DEV)
    _bwd(q16, full, lab, 1 / tau, lse_f.contiguous(), (1 / tau) / 96, None, 0.0, dq_full, None)
    dq_sh = torch.zeros((96, 128), device=DEV)
    for s in shards:
        _bwd(q16, s, lab, 1 / tau, lse_f.contiguous(), (1 / tau) / 96, None, 0.0, dq_sh, None)
    assert rel_l2(dq_sh, dq_full) < 1e-5
    _, o_dq, _ = obank.infonce_loss_and_grads(q, bank, labels, tau, dtype=torch.float64)
    assert rel_l2(dq_sh, o_dq) < 1e-2


@pytest.mark.parametrize("name", ["bank_step_plus", "bank_step_noplus"])
def test_bank_large_step_mixin_against_executed_reference(golden, name):
    """compat.BankLossMixin.bank_large_step (the drop-in for clip4cir/models_negplus.py:130-142: resident bf16 bank, no per-step
    upload) against the fixtures written by executing the reference method: loss and the gradient w.r.t. text_feats."""
    from spn4cir_b200.compat import BankLossMixin
    g = golden(name)
    rb, tb, text, ix, rix, tix = synth.make_bank_step_case(int(g["seed"]), int(g["B"]), int(g["R"]), int(g["N"]), int(g["D"]))

    class Model(BankLossMixin):
        device, tau, plus = torch.device(DEV), float(g["tau"]), bool(int(g["plus"]))

    m = Model()
    m.set_banks(rb, tb)
    tt = text.to(DEV).requires_grad_(True)
    loss = {}
    m.bank_large_step(loss, tt, ix.to(DEV), tix.to(DEV), rix.to(DEV))
    loss["bank_loss"].backward()
    assert abs(loss["bank_loss"].item() - float(g["loss"])) <= 1e-2 * abs(float(g["loss"]))
    assert rel_l2(tt.grad.cpu(), torch.from_numpy(g["dtext"])) < 2e-2, rel_l2(tt.grad.cpu(), torch.from_numpy(g["dtext"]))


def test_query_assemble_prologue():
    g = torch.Generator().manual_seed(5)
    refer = torch.randn((50, 640), generator=g)
    text = torch.randn((20, 640), generator=g)
    idx = torch.randint(0, 50, (20,), generator=g)
    q16, q32, inv = spn.query_assemble(refer.to(DEV), idx.to(DEV), text.to(DEV))
    want = obank.bank_large_step_query(refer, idx, text)
    np.testing.assert_allclose(q32.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(q16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-3)
    np.testing.assert_allclose(inv.cpu().numpy(), (1 / (refer[idx] + text).norm(dim=1)).numpy(), rtol=1e-5)


def test_c3_full_size_properties():
    """BASELINE C3 (512 x 100k x 640): size-independent checks instead of a dense CPU oracle --
    (a) agreement with a dense fp32 torch evaluation on the GPU, (b) an 8-way row-sharded evaluation (statistics
    merged by spn_merge_stats, dq accumulated shard by shard) reproduces the unsharded loss and gradient,
    (c) every dlogit row sums to zero, so dq is invariant to adding a constant row to the bank:
    sum_n (p_n - onehot_n) * (bank_n + c) = sum_n (p_n - onehot_n) * bank_n."""
    cfg = synth.CONFIGS["C3"]
    q, bank, labels = synth.make_bank_case(cfg["B"], cfg["N"], cfg["D"], 1, "S", 0)
    tau = cfg["tau"]
    lab = labels.to(DEV)
    qd = q.to(DEV).requires_grad_(True)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    loss = spn.bank_infonce(qd, db, lab, tau, single_pass=False)
    loss.backward()
    qs = q.to(DEV).requires_grad_(True)
    loss_fused = spn.bank_infonce(qs, db, lab, tau, single_pass=True)
    loss_fused.backward()
    assert abs(loss_fused.item() - loss.item()) <= 1e-4 * max(abs(loss.item()), 1e-3), (loss_fused.item(), loss.item())
    assert rel_l2(qs.grad, qd.grad) < 5e-3, rel_l2(qs.grad, qd.grad)
    # (a) dense fp32 reference on the device
    qf = q.to(DEV).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy((qf @ bank.to(DEV).T) / tau, lab)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-2 * abs(ref.item()), (loss.item(), ref.item())
    assert rel_l2(qd.grad, qf.grad) < 1e-2, rel_l2(qd.grad, qf.grad)
    assert rel_l2(qs.grad, qf.grad) < 1e-2, rel_l2(qs.grad, qf.grad)
    # (b) 8 shards
    from spn4cir_b200.ops import _bwd, _loss_from_stats, cast_to_bf16
    R = 8
    shards = [spn.DeviceBank.from_tensor(bank, DEV, shard_rank=r, shard_world=R) for r in range(R)]
    merged = spn.merge_stats(torch.stack([spn.bank_ce_stats(q.to(DEV), s, lab, tau) for s in shards]))
    lse, _, loss_sh = _loss_from_stats(merged, cfg["N"], 0.0)
    assert abs(loss_sh.item() - loss.item()) <= 1e-5 * max(abs(loss.item()), 1e-3), (loss_sh.item(), loss.item())
    q16 = cast_to_bf16(q.to(DEV))
    dq_sh = torch.zeros((cfg["B"], cfg["D"]), device=DEV)
    for s in shards:
        _bwd(q16, s, lab, 1 / tau, lse, (1 / tau) / cfg["B"], None, 0.0, dq_sh, None)
    assert rel_l2(dq_sh, qd.grad) < 1e-4, rel_l2(dq_sh, qd.grad)
    # (c) shift invariance (the shift is exactly representable in bf16 so only the arithmetic is probed)
    shifted = spn.DeviceBank((db.data.float() + 0.0078125).to(torch.bfloat16).contiguous())
    st = spn.bank_ce_stats(q.to(DEV), shifted, lab, tau)
    lse_s, _, _ = _loss_from_stats(st, cfg["N"], 0.0)
    dq_a = torch.zeros((cfg["B"], cfg["D"]), device=DEV)
    _bwd(q16, shifted, lab, 1 / tau, lse_s, (1 / tau) / cfg["B"], None, 0.0, dq_a, None)
    dense = torch.softmax((q16.float() @ shifted.data.float().T) / tau, dim=1)
    dense[torch.arange(cfg["B"], device=DEV), lab] -= 1.0
    want = (dense @ shifted.data.float()) * ((1 / tau) / cfg["B"])
    assert rel_l2(dq_a, want) < 3e-3, rel_l2(dq_a, want)


def test_fused_static_reference_extremes():
    """Inside its envelope (inv_tau |q| max|bank| <= 109 nats) the static reference must survive the worst rows:
    a query anti-aligned with its target and aligned with another entry (s_max - s_t ~ 1.4 / tau)."""
    g = torch.Generator().manual_seed(3)
    D, N, B = 128, 700, 24
    bank = torch.nn.functional.normalize(torch.randn((N, D), generator=g), dim=-1)
    labels = torch.randint(0, N, (B,), generator=g)
    q = torch.nn.functional.normalize(torch.randn((B, D), generator=g), dim=-1)
    other = (labels + 7) % N
    q[:8] = torch.nn.functional.normalize(-0.9 * bank[labels[:8]] + 0.9 * bank[other[:8]] + 0.05 * q[:8], dim=-1)
    for tau in (0.01, 0.02, 0.07):
        loss, dq, dtau = run_ours(q, bank, labels, tau, tau_tensor=True, single_pass=True)
        o_loss, o_dq, o_dtau = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, tau, dtype=torch.float64)
        assert torch.isfinite(loss) and torch.isfinite(dq).all(), tau
        assert abs(loss.item() - o_loss.item()) <= 2e-3 * max(abs(o_loss.item()), 1.0), (tau, loss, o_loss)
        assert rel_l2(dq, o_dq) < 4e-3, (tau, rel_l2(dq, o_dq))
        assert abs(dtau.item() - o_dtau.item()) <= 1e-2 * max(abs(o_dtau.item()), 1e-2)


def test_cold_temperature_and_unnormalised_queries_take_the_exact_path():
    """Outside the envelope (tau = 0.004, or |q| = 3) the default must stay correct (two-pass, running maximum)."""
    g = torch.Generator().manual_seed(5)
    D, N, B = 128, 900, 16
    bank = torch.nn.functional.normalize(torch.randn((N, D), generator=g), dim=-1)
    labels = torch.randint(0, N, (B,), generator=g)
    q = torch.nn.functional.normalize(torch.randn((B, D), generator=g), dim=-1)
    loss, dq, _ = run_ours(q, bank, labels, 0.004)  # default path selection
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, 0.004, dtype=torch.float64)
    assert abs(loss.item() - o_loss.item()) <= 2e-3 * max(abs(o_loss.item()), 1.0)
    assert rel_l2(dq, o_dq) < 4e-3
    loss, dq, _ = run_ours(3.0 * q, bank, labels, 0.02, single_pass=False)
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(bf16_round(3.0 * q), bf16_round(bank), labels, 0.02, dtype=torch.float64)
    assert abs(loss.item() - o_loss.item()) <= 2e-3 * max(abs(o_loss.item()), 1.0)
    assert rel_l2(dq, o_dq) < 4e-3


def test_peaked_softmax_has_no_cancellation():
    """p_t -> 1 (1 - p_t ~ 1e-3): dq = -(1 - p_t) bank_t + sum_{n != t} p_n bank_n must keep its relative accuracy
    (the one-hot never passes through the bf16 rounding of P)."""
    g = torch.Generator().manual_seed(4)
    D, N, B = 256, 3000, 64
    bank = torch.nn.functional.normalize(torch.randn((N, D), generator=g), dim=-1)
    labels = torch.randint(0, N, (B,), generator=g)
    z = torch.randn((B, D), generator=g)
    # mean loss 4e-4 (1 - p_t ~ 1e-4 .. 1e-3): both paths; mean loss 4.5e-6: only the fused path resolves it -- the
    # two-pass path (like the fp32 reference) forms p_t from s_t - lse, which cancels below fp32 resolution of lse
    for noise, paths, tol in ((1.9, (True, False), 4e-3), (1.55, (True,), 2e-2)):
        q = torch.nn.functional.normalize(bank[labels] + noise * z / D ** 0.5, dim=-1)
        o_loss, o_dq, _ = obank.infonce_loss_and_grads(bf16_round(q), bf16_round(bank), labels, 0.02, dtype=torch.float64)
        assert 1e-7 < o_loss.item() < 5e-3, o_loss.item()
        for sp in paths:
            loss, dq, _ = run_ours(q, bank, labels, 0.02, single_pass=sp)
            assert rel_l2(dq, o_dq) < tol, (noise, sp, rel_l2(dq, o_dq))


@pytest.mark.parametrize("B,N,D,tau", [(512, 9000, 640, 0.02), (64, 2000, 128, 0.005), (256, 3000, 768, 0.02)])
def test_step_primitive_equals_autograd(B, N, D, tau):
    """spn.bank_ce_step (one C-ABI call, no autograd engine) returns the loss / gradients of the autograd path."""
    q, bank, labels = synth.make_bank_case(B, N, D, 1, "S", 77)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    loss, dq, dtau = spn.bank_ce_step(q.to(DEV), db, labels.to(DEV), tau, grad_scale=3.0, want_dtau=True)
    qd = q.to(DEV).requires_grad_(True)
    t = torch.tensor(tau, device=DEV, requires_grad=True)
    ref = spn.bank_infonce(qd, db, labels.to(DEV), t)
    (3.0 * ref).backward()
    # (two launches of the same kernels: the fp32 sums are accumulated with atomics / TMA reduces in a run-dependent order)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(abs(ref.item()), 1e-3)
    assert rel_l2(dq.cpu(), qd.grad.cpu()) < 1e-4
    assert abs(dtau.item() - t.grad.item()) <= 1e-3 * max(abs(t.grad.item()), 1e-2)
    o_loss, o_dq, _ = obank.infonce_loss_and_grads(q, bank, labels, tau, grad_out=3.0)
    assert abs(loss.item() - o_loss.item()) <= 1e-2 * max(abs(o_loss.item()), 1e-3)
    assert rel_l2(dq.cpu(), o_dq) < 1e-2


@pytest.mark.parametrize("name", ["part_small", "part_640"])
def test_part_infonce_against_reference_fixture(golden, name):
    """Sampled-negatives loss through spn_part_ce: the compat call seeded like the reference run draws the
    same negatives (clip4cir/models.py:109-113) and reproduces its loss / gradient within the bf16 tolerance."""
    import random
    g = golden(name)
    q, bank, labels = synth.make_bank_case(int(g["B"]), int(g["N"]), int(g["D"]), 1, str(g["mode"]), int(g["seed"]))
    qd = q.to(DEV).requires_grad_(True)
    random.seed(int(g["seed"]))
    loss = spn.compat.part_infonce_loss(qd, spn.DeviceBank.from_tensor(bank, DEV), labels, int(g["k"]), float(g["tau"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-2 * max(abs(float(g["loss"])), 1e-3), (loss.item(), float(g["loss"]))
    assert rel_l2(qd.grad.cpu(), torch.from_numpy(g["dq"])) < 1e-2
    # kernel arithmetic against the closed form on bf16-rounded operands and the fixture's candidate ids
    cand = torch.from_numpy(g["cand"])
    t = torch.tensor(float(g["tau"]), device=DEV, requires_grad=True)
    q2 = q.to(DEV).requires_grad_(True)
    l2 = spn.part_infonce(q2, spn.DeviceBank.from_tensor(bank, DEV), cand, t)
    (2.0 * l2).backward()
    r_loss, r_dq, r_dtau = obank.part_loss_and_grads(bf16_round(q), bf16_round(bank), cand, float(g["tau"]), grad_out=2.0)
    assert abs(l2.item() - r_loss.item()) <= 1e-5 * max(abs(r_loss.item()), 1.0)
    assert rel_l2(q2.grad.cpu().double(), r_dq) < 1e-5
    assert abs(t.grad.item() - r_dtau.item()) <= 1e-4 * max(abs(r_dtau.item()), 1e-2)


def test_part_infonce_edge_cases():
    """k = 0 (only the positive: loss 0, zero gradient), out-of-range ids masked, fp16 queries, a long candidate list."""
    q, bank, labels = synth.make_bank_case(5, 4000, 128, 1, "U", 3)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    qd = q.to(DEV).requires_grad_(True)
    loss = spn.part_infonce(qd, db, labels.view(-1, 1), 0.05)
    loss.backward()
    assert abs(loss.item()) < 1e-6 and float(qd.grad.abs().max()) < 1e-6
    g = torch.Generator().manual_seed(0)
    cand = torch.cat([labels.view(-1, 1), torch.randint(0, 4000, (5, 3000), generator=g)], dim=1)
    cand_masked = torch.cat([cand, torch.full((5, 2), -1), torch.full((5, 1), 4000)], dim=1)
    a = spn.part_infonce(q.to(DEV, torch.float16), db, cand, 0.05)
    b = spn.part_infonce(q.to(DEV, torch.float16), db, cand_masked, 0.05)
    assert abs(a.item() - b.item()) < 1e-6
    r_loss, _, _ = obank.part_loss_and_grads(bf16_round(q.half().float()), bf16_round(bank), cand, 0.05)
    assert abs(a.item() - r_loss.item()) <= 1e-4 * max(1.0, abs(r_loss.item()))


def test_empty_and_malformed_inputs_fail_loudly():
    """An empty batch / an empty query set / an unsupported feature width is rejected with an exception (the reference would
    return NaN from an empty cross-entropy, clip4cir/models_negplus.py:150-154); nothing is silently skipped or computed on
    the CPU, and an error raised by the C ABI below the PyTorch extension reaches Python as an exception (regression: a
    statically linked libstdc++ in the extension crashed the process while formatting the status code)."""
    from spn4cir_b200 import _torch_ext
    from spn4cir_b200._lib import SpnError
    bank = synth.make_bank(300, 64, 1, seed=5)
    db = spn.DeviceBank.from_tensor(bank, DEV)
    q0 = torch.zeros((0, 64), device=DEV)
    lab0 = torch.zeros((0,), dtype=torch.int64, device=DEV)
    with pytest.raises(SpnError, match="empty batch"):
        spn.bank_infonce(q0, db, lab0, 0.02)
    with pytest.raises(SpnError, match="empty batch"):
        spn.bank_ce_step(q0, db, lab0, 0.02)
    maxnorm, _ = db.stats(False)
    with pytest.raises(RuntimeError, match="empty batch"):
        _torch_ext.ops().bank_ce_step(q0, db.data, lab0, maxnorm, None, 50.0, 0.0, 1, False)
    with pytest.raises(Exception):
        spn.sim_topk(q0, bank.to(DEV).float(), 5)
    # an error from the C ABI (D = 96 is not a multiple of 64) through the extension: status code + message, no crash
    bank96 = torch.randn(200, 96, device=DEV).to(torch.bfloat16)
    q96 = torch.randn(4, 96, device=DEV)
    lab4 = torch.zeros(4, dtype=torch.int64, device=DEV)
    with pytest.raises(RuntimeError, match=r"failed \(-1\).*bad shape"):
        _torch_ext.ops().bank_ce_step(q96, bank96, lab4, maxnorm, None, 50.0, 0.0, 1, False)
    with pytest.raises(SpnError, match="labels for"):
        spn.bank_ce_step(torch.randn(4, 64, device=DEV), db, lab4[:3], 0.02)
    torch.cuda.synchronize()
    # and the device is still usable afterwards
    lab = synth.make_labels(4, 300, seed=6)
    q = synth.make_queries(bank, lab, "S", seed=7).to(DEV)
    loss, dq = spn.bank_ce_step(q, db, lab.to(DEV), 0.02)
    assert torch.isfinite(loss).item() and torch.isfinite(dq).all().item()
