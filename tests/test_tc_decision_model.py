"""The decision rules of the tensor-core retrieval path (csrc/sim_tc.cu: row_topk_kernel / row_rank_kernel), restated
in plain Python and fuzzed on the CPU: given ANY approximate scores S~ with |S~ - S| <= eps, the window / cluster /
certification logic must return exactly the ranking of the exact scores S (descending, ties -> lower index).  This
checks the argument the kernels rely on; the GPU tests (tests/test_gpu_retrieval_tc.py) check the kernels themselves."""
import math

import numpy as np
import pytest

W = 64  # candidate window of row_topk_kernel


def exact_topk(S, K, ex=-1):
    order = sorted((i for i in range(len(S)) if i != ex), key=lambda i: (-S[i], i))
    return order[:K]


def general_path(S, A, K, eps, ex):
    """row_topk_general: window of 64 by S~, all rescored, second scan over everything within eps of the K-th exact."""
    cand = sorted((i for i in range(len(S)) if i != ex), key=lambda i: (-A[i], i))
    win, rest = cand[:W], cand[W:]
    lst = sorted(win, key=lambda i: (-S[i], i))
    e_k = S[lst[K - 1]] if len(lst) >= K else -math.inf
    if len(win) == W and not (A[win[-1]] + eps < e_k):
        for i in sorted(rest):                      # the scan walks the row in index order
            if not (A[i] + eps < e_k):
                lst = sorted(lst + [i], key=lambda j: (-S[j], j))
                e_k = S[lst[K - 1]] if len(lst) >= K else -math.inf
    return lst[:K], True


def tc_topk_model(S, A, K, eps, ex=-1):
    cand = sorted((i for i in range(len(S)) if i != ex), key=lambda i: (-A[i], i))
    win = cand[:W]
    a = [A[i] for i in win]
    a_k1 = a[K - 1] if len(win) >= K else -math.inf
    if len(win) == W and not (a[-1] + 2 * eps < a_k1):
        return general_path(S, A, K, eps, ex)
    n = len(win)
    rel = [p < K or not (a[p] + 2 * eps < a_k1) for p in range(n)]
    head = [p == 0 or (a[p - 1] - a[p] > 2 * eps) for p in range(n)]
    cid, cur = [], 0
    for p in range(n):
        if head[p]:
            cur = p
        cid.append(cur)
    nxt_head = [True if p == n - 1 else head[p + 1] for p in range(n)]
    need = [rel[p] and not (head[p] and nxt_head[p]) for p in range(n)]       # only these get the exact chain
    key = [((cid[p] if rel[p] else 10 ** 9), -(S[win[p]] if need[p] else a[p]), win[p]) for p in range(n)]
    order = sorted(range(n), key=lambda p: key[p])
    out = [win[p] for p in order if rel[p]][:K]
    return out, False, sum(need)


def tc_rank_model(S, A, tg, eps, ex=-1):
    ts = S[tg]
    cnt = 0
    for i in range(len(S)):
        if i == ex or i == tg:
            continue
        if A[i] > ts + eps:
            cnt += 1
        elif not (A[i] < ts - eps):                 # undecided by S~: exact chain
            cnt += (S[i] > ts) or (S[i] == ts and i < tg)
    return cnt


def make_rows(rng, G, kind):
    if kind == "gauss":
        S = rng.normal(0, 0.04, G)
    elif kind == "plateau":                          # blocks of equal and nearly equal scores
        base = rng.normal(0, 0.04, 6)
        S = base[rng.integers(0, 6, G)] + rng.integers(-2, 3, G) * 1e-7
    else:                                            # dense: spacing comparable to eps
        S = np.sort(rng.uniform(0, 1e-3, G))[::-1].copy()
        rng.shuffle(S)
    return S.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("kind", ["gauss", "plateau", "dense"])
@pytest.mark.parametrize("K", [1, 10, 50, 56])
def test_topk_model_is_exact_for_any_bounded_error(kind, K):
    rng = np.random.default_rng(1000 * ["gauss", "plateau", "dense"].index(kind) + K)
    eps = 4e-5
    slow = needs = 0
    for trial in range(30):
        G = int(rng.integers(60, 400))
        S = make_rows(rng, G, kind)
        for noise in ("uniform", "adversarial"):
            if noise == "uniform":
                A = S + rng.uniform(-eps, eps, G) * 0.999
            else:                                    # push neighbours across each other as far as the bound allows
                A = S + np.where(rng.integers(0, 2, G) == 1, eps, -eps) * 0.999
            ex = int(rng.integers(-1, G))
            got = tc_topk_model(S, A, K, eps, ex)
            assert got[0] == exact_topk(S, K, ex), (kind, K, trial, noise)
            slow += got[1]
            needs += got[2] if not got[1] else 0
    if kind == "gauss":
        assert slow <= 6                             # random scores almost never need the general path
    elif kind == "plateau":
        assert slow > 0                              # plateaus wider than the window do, and stay exact


@pytest.mark.parametrize("kind", ["gauss", "plateau", "dense"])
def test_rank_model_is_exact_for_any_bounded_error(kind):
    rng = np.random.default_rng(7 + len(kind))
    eps = 4e-5
    for trial in range(60):
        G = int(rng.integers(5, 300))
        S = make_rows(rng, G, kind)
        A = S + rng.uniform(-eps, eps, G) * 0.999
        tg = int(rng.integers(0, G))
        ex = int(rng.integers(-1, G))
        if ex == tg:
            ex = -1
        want = sum(1 for i in range(G) if i != ex and i != tg and (S[i] > S[tg] or (S[i] == S[tg] and i < tg)))
        assert tc_rank_model(S, A, tg, eps, ex) == want
