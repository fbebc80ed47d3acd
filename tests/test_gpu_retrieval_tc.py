"""Tensor-core retrieval (spn_sim_topk_tc / spn_sim_rank_tc) against the fp32 tile kernels: the tcgen05 sweep only
prunes, the fp32 chain decides, so values must be the SAME BITS and indices / ranks identical -- including on
galleries full of duplicates and 1-ulp near-duplicates, where the certification has to fall back to its second scan.
The fp32 kernels themselves are pinned to the oracle / the executed reference in test_gpu_retrieval.py."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import spn4cir_b200 as spn
from oracle import retrieval as oretr
from spn4cir_b200 import _lib, synth
from spn4cir_b200.retrieval import _tc_workspace

pytestmark = pytest.mark.gpu
DEV = "cuda"
EPS_REL = 4e-5  # kEpsRel of csrc/sim_tc.cu


def _same_bits(a: torch.Tensor, b: torch.Tensor) -> bool:
    return torch.equal(a.view(torch.int32), b.view(torch.int32))


def _case(Q, G, D, T, seed, noise=6.0):
    q, gallery, _, reference, target, groups = synth.make_retrieval_case(Q, G, D, T, seed, noise=noise)
    return q.to(DEV), gallery.to(DEV), reference, target, groups


@pytest.mark.parametrize("Q,G,D,T,K,seed", [
    (40, 300, 64, 1, 10, 1), (200, 2297, 640, 1, 50, 2), (129, 131, 128, 1, 56, 3), (7, 20, 64, 1, 20, 4),
    (300, 6346, 640, 1, 51, 5), (1, 1000, 256, 1, 1, 6), (513, 1027, 768, 1, 50, 7),
    (33, 50, 256, 32, 10, 8), (130, 701, 256, 32, 50, 9), (16, 9, 256, 32, 5, 10),
])
def test_topk_tc_equals_fp32(Q, G, D, T, K, seed):
    q, g, reference, _, _ = _case(Q, G, D, T, seed)
    v0, i0 = spn.sim_topk(q, g, K, exclude=reference, mode="fp32")
    n2 = torch.zeros(1, dtype=torch.int32, device=DEV)
    v1, i1 = spn.sim_topk(q, g, K, exclude=reference, mode="tc", n_second_out=n2)
    assert torch.equal(i0, i1)
    assert _same_bits(v0, v1)
    if G >= 1024:  # (small galleries: the private per-lane lists overflow into the window, rows take the general path)
        assert int(n2) <= max(1, Q // 50)  # random features: (almost) every row stays on the fast path
    # index-exact mode: same ranking, values are the sweep's S~ (within eps of the exact score)
    v2, i2 = spn.sim_topk(q, g, K, exclude=reference, mode="tc", exact_values=False)
    assert torch.equal(i0, i2)
    fin = torch.isfinite(v0)
    assert torch.equal(fin, torch.isfinite(v2))
    bound = EPS_REL * q.norm(dim=-1, keepdim=True) * g.reshape(-1, D).norm(dim=-1).max()
    assert bool(((v2 - v0).abs()[fin] <= bound.expand_as(v0)[fin]).all())


def test_topk_tc_col_base_and_no_exclude():
    q, g, _, _, _ = _case(100, 900, 128, 1, 11)
    v0, i0 = spn.sim_topk(q, g, 30, col_base=5000, mode="fp32")
    v1, i1 = spn.sim_topk(q, g, 30, col_base=5000, mode="tc")
    assert torch.equal(i0, i1) and _same_bits(v0, v1)
    assert int(i1.min()) >= 5000


def test_topk_tc_fewer_candidates_than_k_pads():
    q, g, *_ = _case(5, 6, 64, 1, 12)
    v, i = spn.sim_topk(q, g, 10, mode="tc")
    assert (i[:, 6:] == -1).all() and (i[:, :6] >= 0).all()
    assert torch.isinf(v[:, 6:]).all()


@pytest.mark.parametrize("T", [1, 32])
def test_sweep_error_is_inside_eps(T):
    """|S~ - S| / (|q| max|g|) of the tcgen05 sweep, read back from the workspace, against the bound the row kernels use."""
    Q, G, D = 300, 1500 if T == 1 else 200, 640 if T == 1 else 256
    q, g, *_ = _case(Q, G, D, T, 13)
    q = q * torch.linspace(0.5, 3.0, Q, device=DEV)[:, None]  # un-normalised queries: eps scales with |q|
    spn.sim_topk(q, g, 10, mode="tc")
    off, pitch, rows = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.call("spn_debug_sim_tc_layout", Q, G, D, T, ctypes.byref(off), ctypes.byref(pitch), ctypes.byref(rows))
    ws = _tc_workspace(Q, G, D, T, q.device)
    approx = ws[off.value: off.value + rows.value * pitch.value * 4].view(torch.float32).view(rows.value, pitch.value)[:Q, :G]
    from spn4cir_b200.compat import inference_sim as sim_matrix
    exact = sim_matrix(q, g)
    gmax = g.reshape(-1, D).norm(dim=-1).max()
    rel = ((approx - exact).abs() / (q.norm(dim=-1, keepdim=True) * gmax)).max().item()
    assert rel < EPS_REL / 4, rel  # the bound has >= 4x head-room on random data


def _adversarial_gallery(G, D, seed):
    """Blocks of exact duplicates and of rows that differ in the last ulp of one coordinate: the top of every ranking
    is a plateau far wider than the 64-candidate window."""
    gen = torch.Generator().manual_seed(seed)
    base = F.normalize(torch.randn(8, D, generator=gen), dim=-1)
    rows = base[torch.randint(0, 8, (G,), generator=gen)].clone()
    ulp = torch.randint(0, 3, (G,), generator=gen)  # 0: exact duplicate, 1 / 2: nudge one coordinate by 1 ulp up / down
    col = torch.randint(0, D, (G,), generator=gen)
    for r in range(G):
        if ulp[r] == 1:
            rows[r, col[r]] = torch.nextafter(rows[r, col[r]], torch.tensor(10.0))
        elif ulp[r] == 2:
            rows[r, col[r]] = torch.nextafter(rows[r, col[r]], torch.tensor(-10.0))
    q = F.normalize(base[torch.randint(0, 8, (97,), generator=gen)] + 0.05 * torch.randn(97, D, generator=gen), dim=-1)
    return q.to(DEV), rows.to(DEV)


@pytest.mark.parametrize("K", [1, 10, 50])
def test_topk_tc_plateaus_take_the_second_scan(K):
    q, g = _adversarial_gallery(700, 128, 20 + K)
    ex = torch.arange(97, dtype=torch.int32) % 700
    v0, i0 = spn.sim_topk(q, g, K, exclude=ex, mode="fp32")
    n2 = torch.zeros(1, dtype=torch.int32, device=DEV)
    v1, i1 = spn.sim_topk(q, g, K, exclude=ex, mode="tc", n_second_out=n2)
    assert int(n2) > 0  # ~88 equal-looking rows per plateau: the window cannot certify
    assert torch.equal(i0, i1) and _same_bits(v0, v1)
    _, i2 = spn.sim_topk(q, g, K, exclude=ex, mode="tc", exact_values=False)
    assert torch.equal(i0, i2)
    # and the fp32 kernel is the oracle's ranking of the exact scores
    from spn4cir_b200.compat import inference_sim as sim_matrix
    sim = sim_matrix(q, g).cpu()
    _, want_i = oretr.topk(sim, K, exclude=ex.tolist())
    assert np.array_equal(i1.cpu().numpy().astype(np.int64), want_i)


def test_topk_tc_twins_are_ordered_by_the_exact_chain():
    """Random gallery where every row has a twin differing by one ulp in one coordinate: S~ cannot order the twins, the
    clusters are rescored (fast path, index-exact mode) and must come out in the fp32 kernel's order."""
    Q, G, D, K = 200, 3000, 256, 50
    gen = torch.Generator().manual_seed(77)
    half = F.normalize(torch.randn(G // 2, D, generator=gen), dim=-1)
    twin = half.clone()
    cols = torch.randint(0, D, (G // 2,), generator=gen)
    rows = torch.arange(G // 2)
    twin[rows, cols] = torch.nextafter(twin[rows, cols], torch.full((G // 2,), 10.0))
    perm = torch.randperm(G, generator=gen)
    g = torch.cat([half, twin])[perm].contiguous().to(DEV)
    q = F.normalize(torch.randn(Q, D, generator=gen), dim=-1).to(DEV)
    v0, i0 = spn.sim_topk(q, g, K, mode="fp32")
    n2 = torch.zeros(1, dtype=torch.int32, device=DEV)
    v1, i1 = spn.sim_topk(q, g, K, mode="tc", exact_values=False, n_second_out=n2)
    assert torch.equal(i0, i1)
    assert int(n2) <= Q // 10
    v2, i2 = spn.sim_topk(q, g, K, mode="tc")
    assert torch.equal(i0, i2) and _same_bits(v0, v2)


@pytest.mark.parametrize("Q,G,D,T,seed", [(50, 400, 64, 1, 31), (300, 2297, 640, 1, 32), (1, 100, 64, 1, 33),
                                          (129, 6346, 640, 1, 34), (40, 333, 256, 32, 35)])
def test_rank_tc_equals_fp32(Q, G, D, T, seed):
    q, g, reference, target, groups = _case(Q, G, D, T, seed, noise=7.0)
    r0, g0 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="fp32")
    r1, g1 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="tc")
    assert torch.equal(r0, r1) and torch.equal(g0, g1)


def test_rank_tc_plateaus():
    q, g = _adversarial_gallery(700, 128, 40)
    target = (torch.arange(97, dtype=torch.int32) * 7) % 700
    ex = (target + 1) % 700
    groups = torch.stack([(target + j) % 700 for j in range(6)], dim=1).to(torch.int32)
    r0, g0 = spn.sim_rank(q, g, target, exclude=ex, group=groups, mode="fp32")
    r1, g1 = spn.sim_rank(q, g, target, exclude=ex, group=groups, mode="tc")
    assert torch.equal(r0, r1) and torch.equal(g0, g1)
    from spn4cir_b200.compat import inference_sim as sim_matrix
    sim = sim_matrix(q, g).cpu()
    want_r, want_g = oretr.rank_of_target(sim, target.tolist(), ex.tolist(), groups.tolist())
    assert np.array_equal(r1.cpu().numpy(), want_r) and np.array_equal(g1.cpu().numpy(), want_g)


def test_rank_tc_sharded_counts_add():
    """Row shards with col_base: target score from the owner (max over shards), counts add -- as in the fp32 path."""
    Q, G, D = 150, 3001, 128
    q, g, reference, target, groups = _case(Q, G, D, 1, 41)
    r_full, g_full = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="fp32")
    bounds = [0, 1000, 2049, G]
    ts = torch.full((Q,), -float("inf"), device=DEV)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ts = torch.maximum(ts, spn.sim_target_score(q, g[lo:hi], target, col_base=lo))
    rank = torch.zeros(Q, dtype=torch.int32, device=DEV)
    grank = torch.zeros(Q, dtype=torch.int32, device=DEV)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        spn.sim_rank(q, g[lo:hi], target, exclude=reference, group=groups, target_score=ts, col_base=lo, rank_out=rank,
                     group_rank_out=grank, mode="tc")
    assert torch.equal(rank, r_full) and torch.equal(grank, g_full)


def test_tc_mode_rejects_unsupported_shapes():
    q, g, *_ = _case(10, 50, 64, 1, 50)
    with pytest.raises(spn.SpnError):
        spn.sim_topk(q, g, 57, mode="tc")
    with pytest.raises(spn.SpnError):
        spn.sim_topk(q[:, :48].contiguous(), g[:, :48].contiguous(), 5, mode="tc")
    # auto quietly keeps the fp32 kernel for them
    v, i = spn.sim_topk(q, g, 57, mode="auto")
    assert i.shape == (10, 57)


def test_query_blocks(monkeypatch):
    """S~ is produced in query blocks when [Q, G] fp32 would exceed the block budget (pair-mining sizes); force 1 MiB
    blocks (128 queries at G = 2000) and compare with the fp32 kernels across the block boundaries."""
    monkeypatch.setenv("SPN_TC_SCORE_BLOCK_MB", "1")
    Q, G, D = 300, 2000, 128
    q, g, reference, target, groups = _case(Q, G, D, 1, 60)
    v0, i0 = spn.sim_topk(q, g, 20, exclude=reference, mode="fp32")
    v1, i1 = spn.sim_topk(q, g, 20, exclude=reference, mode="tc")
    assert torch.equal(i0, i1) and _same_bits(v0, v1)
    r0, g0 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="fp32")
    r1, g1 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="tc")
    assert torch.equal(r0, r1) and torch.equal(g0, g1)
    torch.cuda.synchronize()


def test_pair_mining_shape():
    """N x N similarities with the diagonal excluded (zscir/srm_utils.py:147-150 shape, scaled down): 6000 x 6000 x 768."""
    N, D, K = 6000, 768, 50
    gen = torch.Generator().manual_seed(61)
    x = F.normalize(torch.randn(N, D, generator=gen), dim=-1).to(DEV)
    me = torch.arange(N, dtype=torch.int32)
    v0, i0 = spn.sim_topk(x, x, K, exclude=me, mode="fp32")
    v1, i1 = spn.sim_topk(x, x, K, exclude=me, mode="tc")
    assert torch.equal(i0, i1) and _same_bits(v0, v1)


@pytest.mark.parametrize("Q,G,D,seed", [(300, 2297, 640, 71), (4181, 2297, 640, 72), (130, 129, 64, 73), (257, 5000, 128, 74)])
def test_rank_fused_epilogue_equals_score_matrix_path(Q, G, D, seed, monkeypatch):
    """spn_sim_rank_tc, T = 1: the fused path (the sweep's epilogue counts, no S~ matrix) against the S~ + row kernel
    path (SPN_TC_RANK_FUSED=0) and the fp32 kernels -- incl. the CIRR shape, a ragged last tile, duplicates in the group."""
    q, g, reference, target, groups = _case(Q, G, D, 1, seed, noise=7.0)
    groups = groups.clone()
    groups[::5, 3] = groups[::5, 2]  # a member listed twice counts once
    r0, g0 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="fp32")
    r1, g1 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="tc")
    monkeypatch.setenv("SPN_TC_RANK_FUSED", "0")
    r2, g2 = spn.sim_rank(q, g, target, exclude=reference, group=groups, mode="tc")
    monkeypatch.delenv("SPN_TC_RANK_FUSED")
    assert torch.equal(r0, r1) and torch.equal(g0, g1) and torch.equal(r0, r2) and torch.equal(g0, g2)
    r3, _ = spn.sim_rank(q, g, target, mode="tc")  # no exclusion, no group
    r4, _ = spn.sim_rank(q, g, target, mode="fp32")
    assert torch.equal(r3, r4)
