"""Host-side logic that needs no GPU: id mapping, recall arithmetic, shard geometry, synthetic generators."""
import numpy as np
import pytest
import torch

import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.retrieval import names_to_ids, recall_at


def test_names_to_ids_and_duplicates():
    names = ["a", "b", "c"]
    (ids,) = names_to_ids(names, ["c", "a"])
    assert ids == [2, 0]
    with pytest.raises(spn.SpnError):
        names_to_ids(["a", "a"], ["a"])
    with pytest.raises(KeyError):
        names_to_ids(names, ["zzz"])


def test_recall_arithmetic_matches_reference_form():
    rank = torch.tensor([0, 3, 9, 10, 49, 50, 200])
    r1, r5, r10, r50 = recall_at(rank, (1, 5, 10, 50))
    # the reference: (torch.sum(labels[:, :k]) / len(labels)).item() * 100   (validate.py:148-151)
    for k, got in zip((1, 5, 10, 50), (r1, r5, r10, r50)):
        want = (torch.sum(rank < k) / len(rank)).item() * 100
        assert got == want


def test_synth_is_deterministic_and_normalised():
    q1, b1, l1 = synth.make_bank_case(8, 50, 64, 1, "S", 3)
    q2, b2, l2 = synth.make_bank_case(8, 50, 64, 1, "S", 3)
    assert torch.equal(q1, q2) and torch.equal(b1, b2) and torch.equal(l1, l2)
    assert l1[0] == l1[1], "a duplicate label is forced"
    np.testing.assert_allclose(b1.norm(dim=-1).numpy(), 1.0, rtol=1e-5)
    np.testing.assert_allclose(q1.norm(dim=-1).numpy(), 1.0, rtol=1e-5)
    q, g, names, ref, tgt, groups = synth.make_retrieval_case(5, 30, 64, 32, 1)
    assert g.shape == (30, 32, 64) and len(set(names)) == 30
    assert all(len(set(row.tolist())) == 6 for row in groups)


def test_shard_bounds_cover_the_bank_once():
    from spn4cir_b200.dist import shard_bounds
    for n, r in [(100000, 8), (7, 3), (1000000, 8), (5, 8)]:
        spans = [shard_bounds(n, k, r) for k in range(r)]
        covered = [i for lo, hi in spans for i in range(lo, hi)] if n < 1000 else None
        assert spans[0][0] == 0 and max(hi for _, hi in spans) == n
        assert all(spans[k][1] == spans[k + 1][0] or spans[k + 1][0] == spans[k + 1][1] for k in range(r - 1))
        if covered is not None:
            assert covered == list(range(n))


def test_rank_window_draw_matches_restated_reference():
    """get_diff_id (zscir/get_cir_data.py:58-68): the compat draw from a rank-window row uses the same random.sample call
    on the same list as the restated reference, with and without a window."""
    import random
    from oracle import retrieval as oretr
    from spn4cir_b200 import compat
    g = torch.Generator().manual_seed(1)
    f = torch.nn.functional.normalize(torch.randn(50, 16, generator=g), dim=-1)
    w = oretr.i2i_rank_window(f, 10, 20)
    assert w.shape == (50, 10) and oretr.i2i_rank_window(f, 45, 0).shape == (50, 5)
    full = torch.argsort(f @ f.T, descending=True, stable=True)
    assert torch.equal(w, full[:, 10:20])
    for i in (0, 4, 49):
        random.seed(10 + i)
        a = compat.get_diff_id(i, 50, 3, w[i])
        random.seed(10 + i)
        assert a == oretr.get_diff_id(i, 50, 3, full[i], 10, 20) and set(a) <= set(w[i].tolist())
    random.seed(2)
    a = compat.get_diff_id(4, 50, 3)
    random.seed(2)
    assert a == oretr.get_diff_id(4, 50, 3) and 4 not in a


def test_load_bank_file_contract(tmp_path):
    """the reference's cached bank files (clip4cir/models_negplus.py:78-80,114-125; BLIP-2's three-element
    list blip24cir/models.py:81-89): labeled rows first, unlabeled rows appended, labels point into [0, M)."""
    g = torch.Generator().manual_seed(3)
    refer, target = torch.randn(7, 16, generator=g), torch.nn.functional.normalize(torch.randn(9, 16, generator=g))
    unl = torch.nn.functional.normalize(torch.randn(5, 16, generator=g))
    p2, p3, pu, pb = tmp_path / "b2.pth", tmp_path / "b3.pth", tmp_path / "u.pth", tmp_path / "bare.pth"
    torch.save([refer, target], p2)                                   # models_negplus.py:78
    torch.save([refer, target, torch.zeros(7, 4)], p3)                # blip24cir/models.py:81
    torch.save([unl], pu)                                             # models_negplus.py:114
    torch.save(target, pb)
    b = spn.load_bank_file(str(p2))
    assert torch.equal(b["refer_bank"], refer) and torch.equal(b["target_bank"], target) and b["M"] == 9 and b["query_bank"] is None
    assert spn.load_bank_file(str(p3))["query_bank"].shape == (7, 4)
    assert torch.equal(spn.load_bank_file(str(pb))["target_bank"], target)
    c = spn.load_bank_file(str(p2), unlabeled_path=str(pu))
    assert c["M"] == 9 and torch.equal(c["target_bank"], torch.cat([target, unl]))      # :123-125
    d = spn.load_bank_file(str(p2), unlabeled_path=str(pu), neg_num=2)                  # :117-122
    assert d["target_bank"].shape[0] == 11 and torch.equal(d["target_bank"][9:], unl[:2])
    torch.save([torch.zeros(3, 8)], pu)
    with pytest.raises(spn.SpnError):
        spn.load_bank_file(str(p2), unlabeled_path=str(pu))
    torch.save([refer], p2)
    with pytest.raises(spn.SpnError):
        spn.load_bank_file(str(p2))


def test_device_generators_shard_consistently():
    """bench.py / the full-size tests build the bank chunk by chunk on the device: every world size sees the same rows."""
    cpu = torch.device("cpu")
    full, rb, q, lab = synth.make_device_case(6, 1000, 64, 1, cpu)
    assert rb == 0 and full.shape == (1000, 64) and full.dtype == torch.bfloat16
    parts = [synth.make_device_case(6, 1000, 64, 1, cpu, rank=r, world=4, sharded=True) for r in range(4)]
    assert [p[1] for p in parts] == [0, 250, 500, 750]
    assert torch.equal(torch.cat([p[0] for p in parts]), full)
    assert not torch.equal(parts[0][3], parts[1][3])  # every rank brings its own queries / labels
    with pytest.raises(ValueError):
        synth.make_device_case(6, 1000, 64, 1, cpu, rank=0, world=3, sharded=True)


def test_symm_exchange_mode_rule(monkeypatch):
    """dist.symm_exchange_mode: long sweeps keep the reduce-scatter inside the sweep (push), short sweeps and batches that are not
    a multiple of 128 queries pull; SPN_EXCHANGE forces a form (push only where its tile-per-home-rank constraint holds)."""
    from spn4cir_b200 import dist as sdist
    monkeypatch.delenv("SPN_EXCHANGE", raising=False)
    assert sdist.symm_exchange_mode(1024, 8, 125_000, 768) == "push"       # C5 x 8
    assert sdist.symm_exchange_mode(1024, 2, 500_000, 768) == "push"       # C5 x 2
    assert sdist.symm_exchange_mode(512, 8, 12_500, 640) == "pull"         # C3 x 8
    assert sdist.symm_exchange_mode(256, 8, 12_500, 640) == "pull"         # the reference's batch size
    assert sdist.symm_exchange_mode(96, 2, 500_000, 768) == "pull"         # B % 128 != 0
    assert sdist.symm_exchange_mode(128, 3, 500_000, 768) == "pull"        # odd number of query tiles
    monkeypatch.setenv("SPN_EXCHANGE", "pull")
    assert sdist.symm_exchange_mode(1024, 8, 125_000, 768) == "pull"
    monkeypatch.setenv("SPN_EXCHANGE", "push")
    assert sdist.symm_exchange_mode(512, 8, 12_500, 640) == "push"
    with pytest.raises(ValueError):
        sdist.symm_exchange_mode(96, 2, 5000, 256)
