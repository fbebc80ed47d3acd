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
