"""Drop-in mirrors of the reference's Python call sites for the hot path.

Same names, argument meaning and error behaviour as the reference, so a maintainer swaps an import:

* ``infonce_loss(query_features, target_features, labels=None, tau=0.01)``
  -- ``CIRPlus.infonce_loss`` (/root/reference/clip4cir/models_negplus.py:150-154 and its 7 copies)
* ``BankLossMixin.bank_large_step(loss, text_feats, indexs, target_indexs, refer_indexs)``
  -- /root/reference/clip4cir/models_negplus.py:130-142
* ``part_infonce_loss(query_feats, target_feats_all, target_indexs, neg_num, tau)`` / ``BankLossMixin.part_infonce_loss``
  -- the sampled-negatives loss /root/reference/clip4cir/models.py:109-128 (``--neg_num > 0``), same RNG calls
* ``srm_out(clip_image_feats, clip_text_feats, srm_image_feats, srm_text_feats, reference_index, target_index, K)``
  -- the pair-mining ranks and hard negatives of ``get_srm_out`` (/root/reference/zscir/srm_utils.py:119-183)
* ``forward_stage2_loss(fusion_feats, target_feats, target_indexs, temp)``
  -- the loss tail of ``Blip2QformerCirAlignPrompt.forward_stage2`` (...align_prompt.py:254-268)
* ``inference_sim(fusion_feats, target_feats)`` -- the sim tail of ``inference`` (...align_prompt.py:311-319)
* ``compute_cirr_val_metrics`` / ``compute_fiq_val_metrics`` / ``generate_cirr_test_dicts``
  -- /root/reference/clip4cir/validate.py:19-51,111-156, clip4cir/cirr_test_submission.py:63-111

The retrieval tails work on integer gallery ids (names are mapped once on the host) and never build the
[Q, G] argsort; ordering rule: descending similarity, ties -> lower gallery index.
"""
from __future__ import annotations

from operator import itemgetter
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

from ._lib import SpnError, call
import random

from . import _torch_ext
from .ops import (DeviceBank, _ptr, _require_cuda, _split_tau, _stream, _use_torch_ext, bank_infonce, cast_to_bf16,
                  part_infonce)
from .retrieval import _prep_gallery, names_to_ids, recall_at, sim_rank, sim_topk


# --------------------------------------------------------------------------------------- bank-CE
def infonce_loss(query_features: torch.Tensor, target_features: Union[DeviceBank, torch.Tensor],
                 labels: Optional[torch.Tensor] = None, tau=0.01, label_smoothing: float = 0.0) -> torch.Tensor:
    """``CIRPlus.infonce_loss``: ``CE((q @ bank.T) / tau, labels)``; ``labels=None`` -> ``arange(B)``.

    The accelerated path treats ``target_features`` as a CONSTANT bank (the reference detaches it at the bank call
    sites, models_negplus.py:138).  The same method is also called by the reference with in-batch targets that carry
    gradient (``bbc_loss`` clip4cir/models.py:160, zscir/models.py:168, the symmetric losses models_negtype.py:84-107):
    those sites are outside the path (they need image-tower gradients) and keep the reference's dense
    formula here, so that rebinding the method never drops a gradient silently."""
    if query_features.dim() == 1:  # clip4cir/models.py:125 scores one query at a time
        query_features = query_features.unsqueeze(0)
    if isinstance(target_features, torch.Tensor) and target_features.requires_grad and torch.is_grad_enabled():
        logits = (query_features @ target_features.T) / tau
        if labels is None:
            labels = torch.arange(query_features.shape[0], dtype=torch.long, device=query_features.device)
        return F.cross_entropy(logits, labels.to(query_features.device), label_smoothing=label_smoothing)
    return bank_infonce(query_features, target_features, labels, tau, label_smoothing=label_smoothing)


def get_neg_id(i: int, N: int, k: int, rng=random) -> List[int]:
    """``CIRPlus.get_neg_id`` (/root/reference/clip4cir/models.py:109-113): k distinct ids from range(N) without i,
    drawn with the SAME calls on the SAME generator (python's global ``random`` unless ``rng`` is given), so a seeded
    run picks the reference's negatives."""
    range_target = list(range(N))
    range_target.remove(i)
    return rng.sample(range_target, k)


def part_infonce_loss(query_feats: torch.Tensor, target_feats_all: Union[DeviceBank, torch.Tensor],
                      target_indexs: torch.Tensor, neg_num: int, tau=0.01, rng=random,
                      candidates: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``CIRPlus.part_infonce_loss`` (/root/reference/clip4cir/models.py:115-128): per query, CE over
    ``[positive, neg_num sampled negatives]`` with label 0, mean over the batch.  The ids are drawn on the host in the
    reference's order (one ``.tolist()`` instead of B ``.item()`` syncs); the arithmetic runs in ``spn_part_ce``.
    ``candidates`` (int64 ``[B, 1 + k]``) overrides the sampling."""
    if candidates is None:
        M = target_feats_all.N if isinstance(target_feats_all, DeviceBank) else target_feats_all.shape[0]
        pos = [int(v) for v in target_indexs.tolist()]
        rows = [[p] + get_neg_id(p, M, neg_num, rng) for p in pos]                       # :120-121,124
        candidates = torch.tensor(rows, dtype=torch.int64)
    return part_infonce(query_feats, target_feats_all, candidates, tau)


class BankLossMixin:
    """State + methods of ``CIRPlus`` that sit on the path: resident banks and ``bank_large_step``.

    Expected attributes (as in the reference class): ``device``, ``tau``, ``plus``; call
    :meth:`set_banks` once instead of keeping CPU banks that are re-uploaded every step.
    """

    label_smoothing = 0.0

    def set_banks(self, refer_bank: torch.Tensor, target_bank: torch.Tensor) -> None:
        self.refer_bank = refer_bank.detach().to(self.device, torch.float32).contiguous()
        self.target_bank = DeviceBank.from_tensor(target_bank, self.device)

    def load_banks(self, bank_path: str, unlabeled_path: Optional[str] = None) -> None:
        """The load branches of ``extract_bank_features`` (/root/reference/clip4cir/models_negplus.py:79-80; BLIP-2's
        2- or 3-element list, blip24cir/models.py:82-89) and ``extract_unlabeled_bank_features`` (:115-125): read the cached
        ``torch.save([refer_bank, target_bank])`` file, append the unlabeled rows after the labeled ones (``self.M`` =
        number of labeled rows) and make both banks resident -- once, instead of the per-step upload of :138."""
        from .ops import load_bank_file
        banks = load_bank_file(bank_path, unlabeled_path=unlabeled_path, neg_num=getattr(self, "neg_num", 0))
        self.M = banks["M"]
        self.query_bank = banks["query_bank"]
        self.set_banks(banks["refer_bank"], banks["target_bank"])

    def combining_function(self, image_features, text_features):  # models_negplus.py:48-49
        return image_features + text_features

    def infonce_loss(self, query_features, target_features, labels=None, tau=0.01):
        return infonce_loss(query_features, target_features, labels, tau, self.label_smoothing)

    def bank_large_step(self, loss: dict, text_feats, indexs, target_indexs, refer_indexs=None):
        """``CIRPlus.bank_large_step`` (/root/reference/clip4cir/models_negplus.py:130-142).  With the reference's own
        combiner (element-wise sum, :48-49) and the full-bank loss the whole method is ONE operator
        (``torch.ops.spn4cir.bank_large_step_loss``): query assembly ``normalize(refer_bank[idx] + text)`` in a prologue
        kernel, the fused sweep, and the backward through the normalisation in one kernel."""
        idx = refer_indexs if getattr(self, "plus", False) else indexs
        if self._fused_large_step_ok():
            tau_t, tau_v = _split_tau(self.tau)
            bank = self.target_bank
            eps = float(self.label_smoothing)
            maxnorm, colsum = bank.stats(eps != 0.0)
            loss["bank_loss"] = _torch_ext.ops().bank_large_step_loss(
                self.refer_bank, idx, text_feats, tau_t, bank.data, target_indexs.to(self.device), maxnorm,
                colsum if eps != 0.0 else None, 1.0 / tau_v, eps, bank.T)
            return
        reference_image_feats = self.refer_bank[idx.to(self.refer_bank.device)].detach()
        query_feats = F.normalize(self.combining_function(reference_image_feats, text_feats))
        target_indexs = target_indexs.to(self.device)
        if getattr(self, "neg_num", 0) > 0:  # clip4cir/models.py:140-141
            loss["bank_loss"] = self.part_infonce_loss(query_feats, self.target_bank, target_indexs)
        else:
            loss["bank_loss"] = self.infonce_loss(query_feats, self.target_bank, target_indexs, tau=self.tau)

    def part_infonce_loss(self, query_feats, target_feats_all, target_indexs):
        return part_infonce_loss(query_feats, target_feats_all, target_indexs, self.neg_num, self.tau)

    def _fused_large_step_ok(self) -> bool:
        """One-operator bank_large_step applies when nothing of the method was customised: the class's own sum combiner,
        the full-bank loss (no sampled negatives), an un-sharded resident bank and a temperature the fused sweep's static
        soft-max reference covers (same rule as ops.bank_infonce)."""
        bank = getattr(self, "target_bank", None)
        if not (_use_torch_ext() and isinstance(bank, DeviceBank) and bank.row_base == 0 and bank.n_total == bank.N):
            return False
        if getattr(self, "neg_num", 0) > 0 or type(self).combining_function is not BankLossMixin.combining_function:
            return False
        if type(self).infonce_loss is not BankLossMixin.infonce_loss:
            return False
        _, tau_v = _split_tau(self.tau)
        if not tau_v > 0.0:
            return False
        return (1.0 / tau_v) * bank.maxnorm_host() * 1.03 <= 109.0 and not (self.label_smoothing != 0.0 and bank.T > 1)


def forward_stage2_loss(fusion_feats: torch.Tensor, target_feats: Union[DeviceBank, torch.Tensor],
                        target_indexs: torch.Tensor, temp) -> Dict[str, torch.Tensor]:
    """BLIP-2 stage-2 loss tail: per sample ``CE(max_k(f_i . bank[n,k,:]) / temp, label_i)``, mean over B.
    ``target_feats``: ``[N, 32, 256]`` tensor or a DeviceBank with ``T == 32``; ``temp`` may be the
    learnable ``nn.Parameter`` (its gradient is produced)."""
    return {"loss_qtc": bank_infonce(fusion_feats, target_feats, target_indexs, temp)}


def inference_sim(fusion_feats: torch.Tensor, target_feats: torch.Tensor) -> torch.Tensor:
    """``sim_i2t`` of ``Blip2QformerCirAlignPrompt.inference``: ``max_k f . gallery[g, k, :]`` -> ``[b, G]``
    (1-D when ``b == 1``, as the reference's ``.squeeze()`` does)."""
    _require_cuda(fusion_feats, "fusion_feats")
    g, G, T = _prep_gallery(target_feats)
    q = fusion_feats.detach().float().contiguous()
    out = torch.empty((q.shape[0], G), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        call("spn_sim_matrix", _ptr(q), _ptr(g), q.shape[0], G, q.shape[1], T, _ptr(out), _stream(q.device))
    return out.squeeze(0) if out.shape[0] == 1 else out


# --------------------------------------------------------------------------------------- retrieval
def _normalize_gallery(index_features: torch.Tensor) -> torch.Tensor:
    if index_features.dim() > 2:  # validate.py:121-122
        index_features = torch.mean(index_features, dim=1)
    return F.normalize(index_features, dim=-1).float()


def cirr_metrics_from_features(predicted_features: torch.Tensor, index_features: torch.Tensor,
                               index_names: Sequence, reference_names: Sequence, target_names: Sequence,
                               group_members: Sequence[Sequence], *, pred_is_sim_gallery=None
                               ) -> Tuple[float, float, float, float, float, float, float]:
    """Metric tail of ``compute_cirr_val_metrics`` (validate.py:120-156) on integer ids."""
    ref_ids, tgt_ids = names_to_ids(index_names, reference_names, target_names)
    lut = {n: i for i, n in enumerate(index_names)}
    grp_ids = [[lut[m] for m in row] for row in group_members]
    gallery = _normalize_gallery(index_features) if pred_is_sim_gallery is None else pred_is_sim_gallery
    rank, grank = sim_rank(predicted_features, gallery, tgt_ids, exclude=ref_ids, group=grp_ids)
    # the reference asserts exactly one hit per row in the ranked list and in the subset (validate.py:144-145)
    for r, t, row in zip(ref_ids, tgt_ids, grp_ids):
        assert t != r, "target equals the removed reference image"
        assert t in row, "target is not a member of its CIRR subset"
    r1, r5, r10, r50 = recall_at(rank, (1, 5, 10, 50))
    g1, g2, g3 = recall_at(grank, (1, 2, 3))
    return g1, g2, g3, r1, r5, r10, r50


def fiq_metrics_from_features(predicted_features: torch.Tensor, index_features: torch.Tensor, index_names: Sequence,
                              target_names: Sequence, refer_names: Optional[Sequence] = None, *,
                              pred_is_sim_gallery=None) -> Tuple[float, float]:
    """Metric tail of ``compute_fiq_val_metrics``.  ``refer_names`` given -> clip4cir/zscir flavour that
    removes the reference (validate.py:38-45); ``None`` -> blip4cir/tgcir/blip2 flavour (no removal)."""
    (tgt_ids,) = names_to_ids(index_names, target_names)
    ref_ids = names_to_ids(index_names, refer_names)[0] if refer_names is not None else None
    gallery = _normalize_gallery(index_features) if pred_is_sim_gallery is None else pred_is_sim_gallery
    rank, _ = sim_rank(predicted_features, gallery, tgt_ids, exclude=ref_ids)
    n = rank.numel()
    if ref_ids is None:
        # blip4cir/validate.py:55-56: (torch.sum(labels[:, :k]) / len(labels)).item() * 100 -- float32 tensor arithmetic, on
        # the host like the reference (its labels tensor lives on the CPU; CUDA's int / int division rounds differently)
        return ((torch.sum(rank < 10).cpu() / n).item() * 100, (torch.sum(rank < 50).cpu() / n).item() * 100)
    # clip4cir/validate.py:38-49: python-int arithmetic, (count / N) * 100
    r10 = (int(torch.sum(rank < 10).item()) / n) * 100
    r50 = (int(torch.sum(rank < 50).item()) / n) * 100
    return r10, r50


def _predict_cirr(model, relative_val_dataset, index_names, index_features, device):
    """Feature generation of ``generate_cirr_val_predictions`` (validate.py:159-213), batch 32, no workers."""
    name_to_feat = dict(zip(index_names, index_features))
    feats, target_names, group_members, reference_names = [], [], [], []
    n = len(relative_val_dataset)
    for s in range(0, n, 32):
        items = [relative_val_dataset[i] for i in range(s, min(n, s + 32))]
        items = [it for it in items if it is not None]
        refs = [it[0] for it in items]
        caps = [it[2] for it in items]
        with torch.no_grad():
            text_features = model.encode_text(caps)
            if text_features.shape[0] == 1:
                ref_feats = itemgetter(*refs)(name_to_feat).unsqueeze(0)
            else:
                ref_feats = torch.stack(itemgetter(*refs)(name_to_feat))
            batch = model.combining_function(ref_feats.to(device), text_features.to(device))
        feats.append(F.normalize(batch, dim=-1))
        reference_names.extend(refs)
        target_names.extend(it[1] for it in items)
        group_members.extend(list(it[3]) for it in items)
    return torch.vstack(feats), reference_names, target_names, group_members


def compute_cirr_val_metrics(relative_val_dataset, model, index_features: torch.Tensor, index_names: List,
                             device=torch.device("cuda")):
    """Signature and return order of the reference (validate.py:111-156):
    (group_recall@1, @2, @3, recall@1, @5, @10, @50)."""
    pred, refs, tgts, groups = _predict_cirr(model, relative_val_dataset, index_names, index_features, device)
    return cirr_metrics_from_features(pred.to(device), index_features.to(device), index_names, refs, tgts, groups)


def _predict_fiq(model, relative_val_dataset, index_names, index_features, device):
    name_to_feat = dict(zip(index_names, index_features))
    feats, target_names, refer_names = [], [], []
    n = len(relative_val_dataset)
    for s in range(0, n, 32):
        items = [relative_val_dataset[i] for i in range(s, min(n, s + 32))]
        items = [it for it in items if it is not None]
        refs = [it[0] for it in items]
        caps = [f"{it[2][0].strip('.?, ')} and {it[2][1].strip('.?, ')}" for it in items]  # validate.py:78-80
        with torch.no_grad():
            text_features = model.encode_text(caps)
            if text_features.shape[0] == 1:
                ref_feats = itemgetter(*refs)(name_to_feat).unsqueeze(0)
            else:
                ref_feats = torch.stack(itemgetter(*refs)(name_to_feat))
            batch = model.combining_function(ref_feats.to(device), text_features.to(device))
        feats.append(F.normalize(batch, dim=-1))
        target_names.extend(it[1] for it in items)
        refer_names.extend(refs)
    return torch.vstack(feats), target_names, refer_names


def compute_fiq_val_metrics(relative_val_dataset, model, index_features: torch.Tensor, index_names: List,
                            device=torch.device("cuda"), remove_reference: bool = True) -> Tuple[float, float]:
    """(recall@10, recall@50) as validate.py:19-51; ``remove_reference=False`` gives the blip4cir/tgcir
    flavour (/root/reference/blip4cir/validate.py:45-56)."""
    pred, tgts, refs = _predict_fiq(model, relative_val_dataset, index_names, index_features, device)
    return fiq_metrics_from_features(pred.to(device), index_features.to(device), index_names, tgts,
                                     refs if remove_reference else None)


def cirr_test_dicts_from_features(predicted_features: torch.Tensor, index_features: torch.Tensor,
                                  index_names: Sequence, reference_names: Sequence,
                                  group_members: Sequence[Sequence], pairs_id: Sequence
                                  ) -> Tuple[Dict[str, list], Dict[str, list]]:
    """Tail of ``generate_cirr_test_dicts`` (cirr_test_submission.py:84-111): top-50 names and top-3
    in-subset names per pair id, from one fused top-K launch each (no argsort)."""
    (ref_ids,) = names_to_ids(index_names, reference_names)
    lut = {n: i for i, n in enumerate(index_names)}
    gallery = _normalize_gallery(index_features)
    k = min(50, len(index_names) - 1)
    _, top = sim_topk(predicted_features, gallery, k, exclude=ref_ids, exact_values=False)  # only the names are used
    top = top.cpu().tolist()
    # subset ranking: the (<= 6)-member groups are scored as tiny per-query galleries on the device
    grp_ids = torch.tensor([[lut[m] for m in row] for row in group_members], dtype=torch.int64,
                           device=gallery.device)
    q = predicted_features.float()
    gs = torch.einsum("qd,qmd->qm", q, gallery[grp_ids])  # [Q, M] gather of at most 6 rows per query
    keep = grp_ids != torch.tensor(ref_ids, device=gallery.device)[:, None]
    gs = gs.masked_fill(~keep, float("-inf"))
    order = torch.argsort(-gs, dim=-1, stable=True)[:, :3]
    sub = torch.gather(grp_ids, 1, order).cpu().tolist()
    pair_preds = {str(int(p)): [index_names[i] for i in row] for p, row in zip(pairs_id, top)}
    pair_group = {str(int(p)): [index_names[i] for i in row] for p, row in zip(pairs_id, sub)}
    return pair_preds, pair_group


# --------------------------------------------------------------------------------------- pair mining
def srm_out(clip_image_feats: torch.Tensor, clip_text_feats: torch.Tensor, srm_image_feats: torch.Tensor,
            srm_text_feats: torch.Tensor, reference_index, target_index, K: int = 1000):
    """``calcu_sims`` + ``get_srm_out`` (/root/reference/zscir/srm_utils.py:119-183) without the four N x N / N x M
    similarity matrices and the five full argsorts: the six rank look-ups (:166-171) are rank-of-target counts
    (``spn_sim_rank``), the K hard negatives per triplet (:176-181) a paged top-K with the row itself excluded (the
    reference masks the diagonal with -10000, :150-151).  Triplet ``i`` owns caption row ``i`` of the text matrices.
    Returns ``(r_fn_indexs, m_fn_indexs, t_fn_indexs [N, K] int64, rt_scores, rm_scores, mt_scores [N] int64)`` --
    the list the reference saves at :183.  Ordering: descending similarity, ties -> lower index."""
    dev = clip_image_feats.device
    _require_cuda(clip_image_feats, "clip_image_feats")
    ref = torch.as_tensor(reference_index).to(dev, torch.int64)
    tgt = torch.as_tensor(target_index).to(dev, torch.int64)
    n = ref.shape[0]
    me = torch.arange(n, device=dev, dtype=torch.int64)
    ci, ct = clip_image_feats.float(), clip_text_feats.float()
    si, st = srm_image_feats.float(), srm_text_feats.float()

    def rank(q, gallery, target):
        return sim_rank(q, gallery, target.to(torch.int32))[0].to(torch.int64)

    rt = rank(si[ref], si, tgt) + rank(si[tgt], si, ref)          # r2t + t2r in the intra-image ranking  (:166-167,172)
    rm = rank(ci[ref], ct, me) + rank(ct, ci, ref)                # r2m (image -> text) + m2r (text -> image) (:168-169,173)
    mt = rank(ct, ci, tgt) + rank(ci[tgt], ct, me)                # m2t + t2m                              (:170-171,174)
    t_fn = sim_topk(si[tgt], si, K, exclude=tgt)[1].to(torch.int64)   # (:176,179)
    m_fn = sim_topk(st, st, K, exclude=me)[1].to(torch.int64)         # (:177,180)
    r_fn = sim_topk(si[ref], si, K, exclude=ref)[1].to(torch.int64)   # (:178,181)
    return r_fn, m_fn, t_fn, rt, rm, mt


RANK_WINDOW_MAX = 12288  # ranks per spn_rank_window call (the band + boundary ties are sorted in 128 KiB of shared memory)


def i2i_rank_window(image_feats: torch.Tensor, rank_lo: int, rank_hi: Optional[int] = None, *, block_rows: int = 2048):
    """Rows ``[rank_lo:rank_hi]`` of ``i2i_ranks = torch.argsort(sims_intra_i2i, descending=True)``
    (/root/reference/zscir/get_cir_data.py:33-34 and the slices of ``get_diff_id`` :60-64; the CIRR twin :106-109,131-141):
    for every image the gallery indices whose similarity rank falls in the window (``rank_hi=None`` or
    ``rank_hi <= rank_lo``: to the end, as :61-64 do).  Neither the N x N matrix nor any full sort exists: blocks of
    ``block_rows`` rows of exact fp32 scores (``spn_sim_matrix``) go through ``spn_rank_window`` (two radix selects for the
    boundary scores, the band in between compacted and sorted in shared memory; descending score, ties -> lower index).
    Windows wider than RANK_WINDOW_MAX are cut into consecutive sub-windows.  Returns int64 ``[N, width]``."""
    _require_cuda(image_feats, "image_feats")
    f = image_feats.detach().float().contiguous()
    n = f.shape[0]
    hi = n if (rank_hi is None or rank_hi <= rank_lo) else min(int(rank_hi), n)
    lo = min(max(int(rank_lo), 0), hi)
    out = torch.empty((n, hi - lo), dtype=torch.int64, device=f.device)
    if hi == lo:
        return out
    overflow = torch.zeros(1, dtype=torch.int32, device=f.device)
    for r0 in range(0, n, block_rows):
        blk = f[r0:r0 + block_rows]
        sim = inference_sim(blk, f)
        if sim.dim() == 1:
            sim = sim[None, :]
        sim = sim.contiguous()
        rows = blk.shape[0]
        for w0 in range(lo, hi, RANK_WINDOW_MAX):
            w1 = min(hi, w0 + RANK_WINDOW_MAX)
            idx = torch.empty((rows, w1 - w0), dtype=torch.int32, device=f.device)
            with torch.cuda.device(f.device):
                call("spn_rank_window", _ptr(sim), rows, n, n, w0, w1, _ptr(idx), _ptr(overflow), _stream(f.device))
            if int(overflow.item()):  # a boundary score shared by thousands of columns (degenerate galleries): sort this block
                idx = torch.sort(sim, dim=1, descending=True, stable=True).indices[:, w0:w1]
            out[r0:r0 + rows, w0 - lo:w1 - lo] = idx
    return out


def get_diff_id(i: int, N: int, k: int, rank_window_row=None, rng=random) -> List[int]:
    """``get_diff_id`` of /root/reference/zscir/get_cir_data.py:58-68: k distinct partner images for image ``i`` --
    from its rank window (a row of :func:`i2i_rank_window`) when one is given (``--i2i_rank >= 0``), else from
    ``range(N)`` without ``i``; the SAME ``random.sample`` call on the same list, so a seeded run draws the reference's
    pairs."""
    if rank_window_row is not None:
        range_target = rank_window_row.tolist() if hasattr(rank_window_row, "tolist") else list(rank_window_row)
    else:
        range_target = list(range(N))
        range_target.remove(i)
    return rng.sample(range_target, k)
