"""Retrieval scoring over the C ABI: fused similarity + top-K / rank-of-target and the integer metric
tails that replace the reference's numpy string matching.

Reference: ``distances = 1 - q @ g.T; torch.argsort(distances)`` and the Recall arithmetic of
``compute_cirr_val_metrics`` / ``compute_fiq_val_metrics``
(/root/reference/clip4cir/validate.py:19-51,111-156; blip24cir/validate.py:19-43,107-146) and the
max-over-32-tokens similarity of ``Blip2QformerCirAlignPrompt.inference``
(/root/reference/blip24cir/lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py:311-319).

Ordering rule (the reference's argsort is unstable, so it has none): descending score, ties broken
by the lower gallery index.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import OP_SIM_TC, OP_SIM_TOPK, SpnError, call
from .ops import _ptr, _require_cuda, _stream


def _prep_gallery(gallery: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """[G, D] or [G, T, D] -> contiguous fp32 [G*T, D], G, T."""
    _require_cuda(gallery, "gallery")
    if gallery.dim() == 3:
        G, T, D = gallery.shape
    elif gallery.dim() == 2:
        (G, D), T = gallery.shape, 1
    else:
        raise SpnError("retrieval", -1, f"gallery must be [G, D] or [G, T, D], got {tuple(gallery.shape)}")
    return gallery.detach().float().reshape(G * T, D).contiguous(), G, T


def _i32(x, dev) -> Optional[torch.Tensor]:
    if x is None:
        return None
    return torch.as_tensor(x).to(device=dev, dtype=torch.int32).contiguous()


TC_MAX_K = 56          # spn_sim_topk_tc rescoring window: 64 candidates, K <= 56
TC_MIN_SCORES = 1 << 18  # below this many (query, token-row) pairs the fp32 tile kernel is launch-bound anyway


def _use_tc(mode: str, Q: int, G: int, D: int, T: int, k: int = 1) -> bool:
    """``mode``: "fp32" = the fp32 tile kernel, "tc" = tcgen05 sweep + exact rescoring (same results), "auto"."""
    if mode == "fp32":
        return False
    ok = D % 64 == 0 and k <= TC_MAX_K
    if mode == "tc":
        if not ok:
            raise SpnError("retrieval", -1, f"tensor-core retrieval needs D % 64 == 0 and K <= {TC_MAX_K} (D={D}, K={k})")
        return True
    if mode != "auto":
        raise SpnError("retrieval", -7, f"unknown mode {mode!r} (fp32 | tc | auto)")
    return ok and Q * G * T >= TC_MIN_SCORES


_ws_cache = {}


def _tc_workspace(Q: int, G: int, D: int, T: int, dev) -> torch.Tensor:
    """Workspace of the tensor-core path (split operands + one S~ block), kept per device and stream, grown on demand."""
    need = _lib.workspace_bytes(OP_SIM_TC, Q, G, D, T, 0)
    # one workspace per (device, stream): calls issued on different streams may overlap on the GPU
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(dev))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def release_workspaces() -> None:
    """Drop the cached tensor-core workspaces (they grow to the largest problem seen: ~0.6 GB at pair-mining sizes)."""
    _ws_cache.clear()


def sim_topk(query: torch.Tensor, gallery: torch.Tensor, k: int, exclude=None, col_base: int = 0, mode: str = "auto",
             exact_values: bool = True, n_second_out: Optional[torch.Tensor] = None):
    """Top-``k`` gallery entries per query: returns (values [Q, k] fp32, indices [Q, k] int32, global).
    ``mode``: "fp32" (tile kernel), "tc" (``spn_sim_topk_tc``: tcgen05 sweep prunes, exact fp32 rescoring decides --
    identical output) or "auto".  ``exact_values=False`` (tensor-core path only): the indices are still the exact
    ranking, the values are the sweep's S~ (within 4e-5 |q| max|g| of the exact score) -- enough for every consumer that
    only maps indices to names, and ~2x faster at CIRR size; keep True when lists are merged by value.  ``k > 64`` is served in pages of 64 (``spn_sim_topk_after``: each page resumes after
    the last entry of the previous one; -inf / -1 padding once the gallery is exhausted)."""
    _require_cuda(query, "query")
    g, G, T = _prep_gallery(gallery)
    q = query.detach().float().contiguous()
    Q, D = q.shape
    dev = q.device
    ex = _i32(exclude, dev)
    vals = torch.empty((Q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        if _use_tc(mode, Q, G, D, T, k):
            ws = _tc_workspace(Q, G, D, T, dev)
            call("spn_sim_topk_tc", _ptr(q), _ptr(g), Q, G, D, T, k, _ptr(ex), col_base, _ptr(vals), _ptr(idx),
                 1 if exact_values else 0, _ptr(n_second_out), _ptr(ws), ws.numel(), _stream(dev))
            return vals, idx
        if k <= 64:
            ws_bytes = _lib.workspace_bytes(OP_SIM_TOPK, Q, G, D, T, k)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            call("spn_sim_topk", _ptr(q), _ptr(g), Q, G, D, T, k, _ptr(ex), col_base, _ptr(vals), _ptr(idx), _ptr(ws),
                 ws_bytes, _stream(dev))
            return vals, idx
        ws_bytes = _lib.workspace_bytes(OP_SIM_TOPK, Q, G, D, T, 64)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        pv = torch.empty((Q, 64), dtype=torch.float32, device=dev)
        pi = torch.empty((Q, 64), dtype=torch.int32, device=dev)
        cur_v = cur_i = None
        for lo in range(0, k, 64):
            kk = min(64, k - lo)
            if cur_v is None:
                call("spn_sim_topk", _ptr(q), _ptr(g), Q, G, D, T, kk, _ptr(ex), col_base, _ptr(pv), _ptr(pi), _ptr(ws),
                     ws_bytes, _stream(dev))
            else:
                call("spn_sim_topk_after", _ptr(q), _ptr(g), Q, G, D, T, kk, _ptr(ex), col_base, _ptr(cur_v), _ptr(cur_i),
                     _ptr(pv), _ptr(pi), _ptr(ws), ws_bytes, _stream(dev))
            # the kernel writes [Q, kk] densely
            vals[:, lo:lo + kk] = pv.view(-1)[:Q * kk].view(Q, kk)
            idx[:, lo:lo + kk] = pi.view(-1)[:Q * kk].view(Q, kk)
            last_v = vals[:, lo + kk - 1].contiguous()
            last_i = idx[:, lo + kk - 1].contiguous()
            # an exhausted row (padding -inf / -1) keeps an everything-excluding cursor: (-inf, INT_MAX)
            cur_v = last_v
            cur_i = torch.where(last_i >= 0, last_i, torch.full_like(last_i, 2 ** 31 - 1))
    return vals, idx


def merge_topk(vals: torch.Tensor, idx: torch.Tensor):
    """[R, Q, K] per-shard lists -> one [Q, K] list (``spn_merge_topk``)."""
    R, Q, K = vals.shape
    dev = vals.device
    vals = vals.float().contiguous()
    idx = idx.to(torch.int32).contiguous()
    ov = torch.empty((Q, K), dtype=torch.float32, device=dev)
    oi = torch.empty((Q, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("spn_merge_topk", _ptr(vals), _ptr(idx), R, Q, K, _ptr(ov), _ptr(oi), _stream(dev))
    return ov, oi


def sim_target_score(query: torch.Tensor, gallery: torch.Tensor, target, col_base: int = 0) -> torch.Tensor:
    _require_cuda(query, "query")
    g, G, T = _prep_gallery(gallery)
    q = query.detach().float().contiguous()
    Q, D = q.shape
    dev = q.device
    tg = _i32(target, dev)
    out = torch.empty(Q, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("spn_sim_target_score", _ptr(q), _ptr(g), Q, G, D, T, _ptr(tg), col_base, _ptr(out), _stream(dev))
    return out


def sim_rank(query: torch.Tensor, gallery: torch.Tensor, target, exclude=None, group=None, *,
             target_score: Optional[torch.Tensor] = None, col_base: int = 0, rank_out=None, group_rank_out=None,
             mode: str = "auto"):
    """Rank of ``target`` per query (0 = best) after dropping ``exclude``; with ``group`` [Q, M] also the
    rank restricted to the group members.  Returns (rank [Q] int32, group_rank [Q] int32 | None)."""
    _require_cuda(query, "query")
    g, G, T = _prep_gallery(gallery)
    q = query.detach().float().contiguous()
    Q, D = q.shape
    dev = q.device
    tg, ex, gr = _i32(target, dev), _i32(exclude, dev), _i32(group, dev)
    M = 0 if gr is None else gr.shape[1]
    with torch.cuda.device(dev):
        use_tc = _use_tc(mode, Q, G, D, T)
        if target_score is None and not use_tc:  # (the tensor-core kernel scores the target row itself)
            target_score = torch.empty(Q, dtype=torch.float32, device=dev)
            call("spn_sim_target_score", _ptr(q), _ptr(g), Q, G, D, T, _ptr(tg), col_base, _ptr(target_score),
                 _stream(dev))
        rank = rank_out if rank_out is not None else torch.zeros(Q, dtype=torch.int32, device=dev)
        grank = None
        if gr is not None:
            grank = group_rank_out if group_rank_out is not None else torch.zeros(Q, dtype=torch.int32, device=dev)
        if use_tc:
            ws = _tc_workspace(Q, G, D, T, dev)
            call("spn_sim_rank_tc", _ptr(q), _ptr(g), Q, G, D, T, _ptr(tg), _ptr(target_score), _ptr(ex), _ptr(gr), M,
                 col_base, _ptr(rank), _ptr(grank), _ptr(ws), ws.numel(), _stream(dev))
        else:
            call("spn_sim_rank", _ptr(q), _ptr(g), Q, G, D, T, _ptr(tg), _ptr(target_score), _ptr(ex), _ptr(gr), M,
                 col_base, _ptr(rank), _ptr(grank), None, 0, _stream(dev))
    return rank, grank


def recall_at(rank: torch.Tensor, ks: Sequence[int]) -> Tuple[float, ...]:
    """100 * mean(rank < k), computed like the reference: float32 sum / len * 100 (validate.py:148-151)."""
    n = rank.numel()
    return tuple((torch.sum(rank < k).to(torch.float32) / n).item() * 100 for k in ks)


def names_to_ids(index_names: Sequence, *name_lists: Sequence):
    """Map gallery names to their row index so that matching is integer work on the device."""
    lut = {n: i for i, n in enumerate(index_names)}
    if len(lut) != len(index_names):
        raise SpnError("names_to_ids", -7, "gallery names must be unique")
    out = []
    for names in name_lists:
        out.append([lut[n] for n in names])
    return out
