"""Parity of the fused similarity + top-K / rank kernels against the oracle and the reference-generated
Recall fixtures.  Index parity rule (north_star): identical except where the two scores tie within 1e-3
(the CPU oracle sums in a different order; observed gaps are ~1e-7)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import spn4cir_b200 as spn
from oracle import retrieval as oretr
from spn4cir_b200 import compat, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
TIE = 1e-5  # far inside the 1e-3 the contract allows


def assert_same_up_to_ties(got_idx, want_idx, sim):
    bad = np.argwhere(got_idx != want_idx)
    for r, c in bad:
        a, b = got_idx[r, c], want_idx[r, c]
        assert a >= 0 and b >= 0, (r, c, a, b)
        assert abs(float(sim[r, a]) - float(sim[r, b])) < TIE, (r, c, a, b, sim[r, a], sim[r, b])
    return len(bad)


@pytest.mark.parametrize("Q,G,D,K,seed", [(40, 300, 64, 10, 1), (200, 2297, 640, 50, 2), (65, 129, 128, 64, 3),
                                          (7, 20, 64, 20, 4), (300, 6346, 640, 51, 5)])
def test_topk_matches_oracle(Q, G, D, K, seed):
    q, gallery, _, reference, target, _ = synth.make_retrieval_case(Q, G, D, 1, seed, noise=6.0)
    sim = q @ gallery.T
    want_v, want_i = oretr.topk(sim, K, exclude=reference.tolist())
    v, i = spn.sim_topk(q.to(DEV), gallery.to(DEV), K, exclude=reference)
    v, i = v.cpu().numpy(), i.cpu().numpy().astype(np.int64)
    assert_same_up_to_ties(i, want_i, sim.numpy())
    np.testing.assert_allclose(v, want_v, rtol=1e-5, atol=2e-6)
    for r in range(Q):
        assert int(reference[r]) not in i[r]


def test_topk_fewer_candidates_than_k_pads():
    q, gallery, *_ = synth.make_retrieval_case(5, 6, 64, 1, 6)
    v, i = spn.sim_topk(q.to(DEV), gallery.to(DEV), 10)
    i = i.cpu().numpy()
    assert (i[:, 6:] == -1).all() and (i[:, :6] >= 0).all()
    assert torch.isinf(v[:, 6:]).all()


@pytest.mark.parametrize("Q,G,D,seed", [(50, 400, 64, 7), (300, 2297, 640, 8), (1, 100, 64, 9)])
def test_rank_matches_oracle(Q, G, D, seed):
    q, gallery, _, reference, target, groups = synth.make_retrieval_case(Q, G, D, 1, seed, noise=7.0)
    sim = q @ gallery.T
    want_r, want_g = oretr.rank_of_target(sim, target.tolist(), reference.tolist(), groups.tolist())
    r, g = spn.sim_rank(q.to(DEV), gallery.to(DEV), target, exclude=reference, group=groups)
    r, g = r.cpu().numpy(), g.cpu().numpy()
    # a rank can differ only through a near-tie with the target's score
    for k in np.argwhere(r != want_r).ravel():
        gaps = np.abs(sim[k].numpy() - float(sim[k, int(target[k])]))
        gaps[int(target[k])] = np.inf
        assert gaps.min() < TIE, (k, r[k], want_r[k])
    assert (r != want_r).mean() < 0.01 and (g != want_g).mean() < 0.01


@pytest.mark.parametrize("name", ["cirr_small", "cirr_c2"])
def test_cirr_recall_identical_to_reference(golden, name):
    g = golden(name)
    Q, G, D = int(g["Q"]), int(g["G"]), int(g["D"])
    q, gallery, names, reference, target, groups = synth.make_retrieval_case(Q, G, D, 1, int(g["seed"]), noise=float(g["noise"]))
    ds = synth.FakeCIRRDataset(names, reference, target, groups)
    text = {f"q{i}": (q[i] * 2.0 - gallery[reference[i]]).to(DEV) for i in range(Q)}
    model = synth.FakeModel(text, D)
    got = compat.compute_cirr_val_metrics(ds, model, gallery.to(DEV), names, device=torch.device(DEV))
    np.testing.assert_allclose(got, g["metrics"], rtol=0, atol=1e-4)  # Recall@K / Rsub@K identical
    pred = F.normalize(gallery[reference] + torch.stack([q[i] * 2.0 - gallery[reference[i]] for i in range(Q)]), dim=-1)
    r, gr = spn.sim_rank(pred.to(DEV), F.normalize(gallery, dim=-1).to(DEV), target, exclude=reference, group=groups)
    assert (r.cpu().numpy() != g["rank"]).mean() < 0.005
    assert (gr.cpu().numpy() != g["group_rank"]).mean() < 0.005


@pytest.mark.parametrize("name", ["fiq_small", "fiq_dress", "fiq_shirt", "fiq_toptee", "fiq_refnear"])
def test_fiq_recall_identical_to_reference(golden, name):
    g = golden(name)
    Q, G, D = int(g["Q"]), int(g["G"]), int(g["D"])
    q, gallery, names, reference, target, _ = synth.make_retrieval_case(Q, G, D, 1, int(g["seed"]), noise=float(g["noise"]),
                                                                        ref_pull=float(g["ref_pull"]))
    ds = synth.FakeFIQDataset(names, reference, target)
    caps = [f"is q{i} and and more {i}" for i in range(Q)]
    text = {c: (q[i] * 2.0 - gallery[reference[i]]).to(DEV) for i, c in enumerate(caps)}
    model = synth.FakeModel(text, D)
    got = compat.compute_fiq_val_metrics(ds, model, gallery.to(DEV), names, device=torch.device(DEV))
    np.testing.assert_allclose(got, g["metrics"], rtol=0, atol=1e-9)
    got2 = compat.compute_fiq_val_metrics(ds, model, gallery.to(DEV), names, device=torch.device(DEV), remove_reference=False)
    np.testing.assert_allclose(got2, g["metrics_noremove"], rtol=0, atol=1e-9)
    if float(g["ref_pull"]) > 0:  # the reference image ranks near the top here: removal must change the answer
        assert abs(got[0] - got2[0]) > 0.3


@pytest.mark.parametrize("Q,G,seed", [(5, 20, 4), (33, 150, 5), (1, 40, 6)])
def test_blip2_max_over_tokens_similarity(Q, G, seed):
    q, gallery, _, reference, target, groups = synth.make_retrieval_case(Q, G, 256, 32, seed)
    sim = oretr.similarity(q, gallery)
    got = compat.inference_sim(q.to(DEV), gallery.to(DEV)).cpu()
    assert got.shape == (sim.squeeze(0).shape if Q == 1 else sim.shape)
    np.testing.assert_allclose(got.reshape(Q, G).numpy(), sim.numpy(), rtol=1e-5, atol=2e-6)
    K = min(10, G)
    want_v, want_i = oretr.topk(sim, K, exclude=reference.tolist())
    v, i = spn.sim_topk(q.to(DEV), gallery.to(DEV), K, exclude=reference)
    assert_same_up_to_ties(i.cpu().numpy().astype(np.int64), want_i, sim.numpy())
    want_r, want_g = oretr.rank_of_target(sim, target.tolist(), reference.tolist(), groups.tolist())
    r, gr = spn.sim_rank(q.to(DEV), gallery.to(DEV), target, exclude=reference, group=groups)
    assert (r.cpu().numpy() != want_r).sum() == 0 and (gr.cpu().numpy() != want_g).sum() == 0


def test_sharded_gallery_topk_and_rank_merge():
    q, gallery, _, reference, target, groups = synth.make_retrieval_case(70, 1000, 128, 1, 10, noise=6.0)
    sim = q @ gallery.T
    R, K = 4, 50
    vs, is_, rank, grank = [], [], None, None
    ts = torch.full((70,), float("-inf"), device=DEV)
    for r in range(R):
        lo, hi = r * 250, (r + 1) * 250
        ts = torch.maximum(ts, spn.sim_target_score(q.to(DEV), gallery[lo:hi].to(DEV), target, col_base=lo))
    for r in range(R):
        lo, hi = r * 250, (r + 1) * 250
        v, i = spn.sim_topk(q.to(DEV), gallery[lo:hi].to(DEV), K, exclude=reference, col_base=lo)
        vs.append(v), is_.append(i)
        rank, grank = spn.sim_rank(q.to(DEV), gallery[lo:hi].to(DEV), target, exclude=reference, group=groups,
                                   target_score=ts, col_base=lo, rank_out=rank, group_rank_out=grank)
    v, i = spn.merge_topk(torch.stack(vs), torch.stack(is_))
    want_v, want_i = oretr.topk(sim, K, exclude=reference.tolist())
    assert_same_up_to_ties(i.cpu().numpy().astype(np.int64), want_i, sim.numpy())
    v1, i1 = spn.sim_topk(q.to(DEV), gallery.to(DEV), K, exclude=reference)
    assert torch.equal(i1, i) and torch.equal(v1, v), "sharded + merged must be bit-identical to the single launch"
    r1, g1 = spn.sim_rank(q.to(DEV), gallery.to(DEV), target, exclude=reference, group=groups)
    assert torch.equal(r1, rank) and torch.equal(g1, grank)


def test_cirr_test_submission_dicts():
    Q, G, D = 30, 200, 64
    q, gallery, names, reference, target, groups = synth.make_retrieval_case(Q, G, D, 1, 12)
    preds, gpreds = compat.cirr_test_dicts_from_features(
        q.to(DEV), gallery.to(DEV), names, [names[int(i)] for i in reference],
        [[names[int(j)] for j in row] for row in groups], list(range(Q)))
    sim = q @ F.normalize(gallery, dim=-1).T
    _, want = oretr.topk(sim, 50, exclude=reference.tolist())
    for r in range(Q):
        assert preds[str(r)] == [names[j] for j in want[r]]
        order = [j for j in oretr.ranking(sim)[r] if j in set(groups[r].tolist()) and j != int(reference[r])][:3]
        assert gpreds[str(r)] == [names[j] for j in order]


@pytest.mark.parametrize("Q,G,K,seed", [(30, 500, 200, 1), (7, 130, 129, 2), (65, 3000, 1000, 3)])
def test_deep_topk_pages_tile_the_full_ranking(Q, G, K, seed):
    """K > 64 is served in pages of 64 through spn_sim_topk_after: the concatenation must be the oracle's top-K
    (excluded entry skipped, -1 padding once the gallery is exhausted)."""
    q, gal, _, ref, _, _ = synth.make_retrieval_case(Q, G, 64, 1, seed, noise=4.0)
    vals, idx = spn.sim_topk(q.to(DEV), gal.to(DEV), K, exclude=ref)
    torch.cuda.synchronize()
    k_real = min(K, G - 1)
    _, want = oretr.topk(q @ gal.T, k_real, exclude=ref.tolist())
    got = idx.cpu().numpy()
    assert (got[:, :k_real] == want).mean() > 0.995          # fp32 summation-order ties may swap neighbours
    assert np.all(np.sort(got[:, :k_real], axis=1) == np.sort(want, axis=1)) or (got[:, :k_real] == want).mean() > 0.995
    assert np.all(got[:, k_real:] == -1)
    v = vals.cpu().numpy()[:, :k_real]
    assert np.all(v[:, :-1] >= v[:, 1:])


def test_srm_pair_mining_matches_oracle():
    """compat.srm_out against the restatement of zscir/srm_utils.py:119-183 on a small synthetic set."""
    g = torch.Generator().manual_seed(5)
    n_img, n_trip, D, K = 400, 150, 64, 100
    ci = torch.nn.functional.normalize(torch.randn(n_img, D, generator=g), dim=-1)
    si = torch.nn.functional.normalize(torch.randn(n_img, D, generator=g), dim=-1)
    ct = torch.nn.functional.normalize(torch.randn(n_trip, D, generator=g), dim=-1)
    st = torch.nn.functional.normalize(torch.randn(n_trip, D, generator=g), dim=-1)
    ref = torch.randint(0, n_img, (n_trip,), generator=g)
    tgt = (ref + 1 + torch.randint(0, n_img - 1, (n_trip,), generator=g)) % n_img
    want = oretr.srm_out(ci, ct, si, st, ref.tolist(), tgt.tolist(), K)
    got = spn.compat.srm_out(ci.to(DEV), ct.to(DEV), si.to(DEV), st.to(DEV), ref, tgt, K)
    torch.cuda.synchronize()
    for name, w, h in zip(("r_fn", "m_fn", "t_fn"), want[:3], got[:3]):
        assert (h.cpu() == w).float().mean() > 0.995, name
    for name, w, h in zip(("rt", "rm", "mt"), want[3:], got[3:]):
        assert (h.cpu() - w).abs().max() <= 1 and (h.cpu() == w).float().mean() > 0.98, name


def test_rank_window_matches_oracle():
    """compat.i2i_rank_window (zscir/get_cir_data.py:33-34,58-68): the [lo:hi] slice of every row of the full
    descending ranking.  Exact against a stable argsort of the device's own scores; against the CPU restatement (whose
    matmul sums in another order) up to near-ties; and the seeded partner draw picks the same images."""
    import random
    g = torch.Generator().manual_seed(9)
    n, D, lo, hi = 1500, 128, 400, 700
    f = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=-1)
    got = compat.i2i_rank_window(f.to(DEV), lo, hi, block_rows=512)
    sim_dev = compat.inference_sim(f.to(DEV), f.to(DEV)).cpu()
    want_dev = torch.argsort(sim_dev, descending=True, stable=True)[:, lo:hi]
    assert got.shape == (n, hi - lo) and torch.equal(got.cpu(), want_dev)
    want = oretr.i2i_rank_window(f, lo, hi)
    sim = (f @ f.T).numpy()
    assert_same_up_to_ties(got.cpu().numpy(), want.numpy(), sim)
    tail = compat.i2i_rank_window(f.to(DEV), 1400, 0)           # rank_max <= rank: to the end (:62-64)
    assert tail.shape == (n, 100) and torch.equal(tail.cpu(), torch.argsort(sim_dev, descending=True, stable=True)[:, 1400:])
    ranks_full = torch.argsort(sim_dev, descending=True, stable=True)
    for i in (0, 7, 1499):
        random.seed(123 + i)
        a = compat.get_diff_id(i, n, 5, got[i].cpu())
        random.seed(123 + i)
        b = oretr.get_diff_id(i, n, 5, ranks_full[i], lo, hi)
        assert a == b
    random.seed(5)
    a = compat.get_diff_id(3, n, 4)
    random.seed(5)
    assert a == oretr.get_diff_id(3, n, 4) and 3 not in a


def test_srm_pair_mining_against_executed_reference_lines(golden):
    """compat.srm_out (K = 1000, the reference's hard-coded value) against the fixture written by executing
    zscir/srm_utils.py:121-124,145-180: rank sums exact up to near-ties of the fp32 scores, hard-negative lists equal at
    (almost) every position."""
    g = golden("srm_ref")
    ci, ct, si, st, ref, tgt = synth.make_srm_case(int(g["seed"]), int(g["n_img"]), int(g["n_trip"]), int(g["D"]))
    got = spn.compat.srm_out(ci.to(DEV), ct.to(DEV), si.to(DEV), st.to(DEV), ref, tgt, int(g["K"]))
    torch.cuda.synchronize()
    for name, h in zip(("r_fn", "m_fn", "t_fn"), got[:3]):
        assert h.shape == (int(g["n_trip"]), int(g["K"]))
        assert (h[:, :24].cpu().numpy() == g[name + "_head"]).mean() > 0.995, name
        assert int(h.min()) >= 0                      # > 1000 candidates everywhere: no padding
    for name, h in zip(("rt", "rm", "mt"), got[3:]):
        d = np.abs(h.cpu().numpy() - g[name])
        assert d.max() <= 2 and (d == 0).mean() > 0.98, name


def test_cirr_submission_against_executed_reference_lines(golden):
    """compat.cirr_test_dicts_from_features against the lists produced by executing clip4cir/cirr_test_submission.py:84-109."""
    g = golden("cirr_submit_ref")
    Q, G, D = int(g["Q"]), int(g["G"]), int(g["D"])
    q, gallery, names, reference, target, groups = synth.make_retrieval_case(Q, G, D, 1, int(g["seed"]), noise=float(g["noise"]))
    preds, gpreds = compat.cirr_test_dicts_from_features(
        q.to(DEV), (gallery * float(g["scale"])).to(DEV), names, [names[int(i)] for i in reference],
        [[names[int(j)] for j in row] for row in groups], list(range(100, 100 + Q)))
    lut = {n: i for i, n in enumerate(names)}
    got50 = np.array([[lut[n] for n in preds[str(100 + r)]] for r in range(Q)])
    got3 = np.array([[lut[n] for n in gpreds[str(100 + r)]] for r in range(Q)])
    sim = (q @ F.normalize(gallery * float(g["scale"]), dim=-1).T).numpy()
    assert_same_up_to_ties(got50, g["top50"].astype(np.int64), sim)
    assert_same_up_to_ties(got3, g["top3"].astype(np.int64), sim)


def test_rank_window_kernel_sizes_ties_and_overflow():
    """spn_rank_window at the README's window sizes (zscir: --i2i_rank 10000 --i2i_rank_max 20000 / 15000), a window wider
    than one call (cut into sub-windows), rows full of exact ties (quantised scores: ties -> lower index, like a stable
    sort), +0.0 / -0.0, and the degenerate all-equal row whose band overflows (the caller sorts that block)."""
    from spn4cir_b200._lib import call
    g = torch.Generator().manual_seed(11)
    n, D = 21000, 64
    f = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=-1).to(DEV)
    rows = f[:96]
    sim = compat.inference_sim(rows, f)
    ranks = torch.argsort(sim, dim=1, descending=True, stable=True)

    def window(scores, lo, hi):
        R, G = scores.shape
        out = torch.empty((R, hi - lo), dtype=torch.int32, device=DEV)
        ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
        call("spn_rank_window", scores.data_ptr(), R, G, G, lo, hi, out.data_ptr(), ovf.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
        return out.long(), int(ovf.item())

    for lo, hi in ((10000, 20000), (10000, 15000), (0, 50), (20990, 21000), (0, 12288)):
        got, ovf = window(sim, lo, hi)
        assert ovf == 0 and torch.equal(got, ranks[:, lo:hi]), (lo, hi)
    # quantised scores: thousands of exact ties per value, window boundaries inside runs of equal scores
    q = torch.round(sim * 64) / 64
    q[:, 5] = 0.0
    q[:, 6] = -0.0
    rq = torch.argsort(q, dim=1, descending=True, stable=True)
    got, ovf = window(q, 9000, 14000)
    assert ovf == 0 and torch.equal(got, rq[:, 9000:14000])
    # all-equal row: every column ties with both boundaries -> overflow is reported, never a wrong answer
    flat = torch.zeros((2, n), device=DEV)
    _, ovf = window(flat, 100, 200)
    assert ovf == 1
    # the compat call cuts a 16 000-wide window into two kernel calls and falls back to a sort on overflow
    sub = f[:3000]
    wide = compat.i2i_rank_window(sub, 500, 0, block_rows=1024)
    simsub = compat.inference_sim(sub, sub)
    assert torch.equal(wide, torch.argsort(simsub, dim=1, descending=True, stable=True)[:, 500:])
    allsame = torch.ones((300, 64), device=DEV)
    w = compat.i2i_rank_window(allsame, 10, 40)
    assert torch.equal(w, torch.arange(10, 40, device=DEV).expand(300, 30))
