"""Oracle for validation retrieval (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def similarity(q: torch.Tensor, gallery: torch.Tensor) -> torch.Tensor:
    """fp32 similarity; 3-D gallery -> max over tokens as Blip2QformerCirAlignPrompt.inference,
    /root/reference/blip24cir/lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py:311-319."""
    if gallery.dim() == 3:
        sim_t2q = torch.matmul(q.unsqueeze(1).unsqueeze(1), gallery.permute(0, 2, 1)).squeeze(2)  # [b, G, 32]
        return sim_t2q.max(-1).values
    return q @ gallery.t()


def ranking(sim: torch.Tensor) -> np.ndarray:
    """Full ranking [Q, G] under the documented rule: descending similarity, ties -> lower index.
    (The reference sorts `1 - sim` with an unstable argsort, clip4cir/validate.py:126-127.)"""
    return torch.argsort(-sim, dim=-1, stable=True).numpy()


def reference_style_sorted_indices(q: torch.Tensor, gallery: torch.Tensor) -> torch.Tensor:
    """Exactly the reference lines clip4cir/validate.py:126-127."""
    distances = 1 - q @ gallery.T
    return torch.argsort(distances, dim=-1).cpu()


def topk(sim: torch.Tensor, k: int, exclude: Optional[Sequence[int]] = None) -> Tuple[np.ndarray, np.ndarray]:
    order = ranking(sim)
    vals = sim.numpy()
    out_i = np.full((sim.shape[0], k), -1, dtype=np.int64)
    out_v = np.full((sim.shape[0], k), -np.inf, dtype=np.float32)
    for r in range(sim.shape[0]):
        row = order[r]
        if exclude is not None:
            row = row[row != int(exclude[r])]
        row = row[:k]
        out_i[r, : len(row)] = row
        out_v[r, : len(row)] = vals[r, row]
    return out_v, out_i


def rank_of_target(sim: torch.Tensor, target: Sequence[int], exclude: Optional[Sequence[int]] = None,
                   group: Optional[Sequence[Sequence[int]]] = None):
    """Position of the target after dropping `exclude` (0 = first); with `group`, position among the group
    members only (what labels[...]/group_labels[...] encode in clip4cir/validate.py:136-142)."""
    order = ranking(sim)
    Q = sim.shape[0]
    rank = np.zeros(Q, dtype=np.int64)
    grank = np.zeros(Q, dtype=np.int64) if group is not None else None
    for r in range(Q):
        row = order[r]
        if exclude is not None:
            row = row[row != int(exclude[r])]
        rank[r] = int(np.nonzero(row == int(target[r]))[0][0])
        if group is not None:
            sub = row[np.isin(row, np.asarray(group[r]))]
            grank[r] = int(np.nonzero(sub == int(target[r]))[0][0])
    return rank, grank


def cirr_metrics(q: torch.Tensor, index_features: torch.Tensor, index_names: List[str], reference_names: List[str],
                 target_names: List[str], group_members: List[List[str]]):
    """Restatement of compute_cirr_val_metrics' metric tail, /root/reference/clip4cir/validate.py:120-156,
    on names with numpy exactly as the reference does (small sizes only)."""
    if index_features.dim() > 2:                                                     # :121-122
        index_features = torch.mean(index_features, dim=1)
    index_features = F.normalize(index_features, dim=-1).float()                     # :123
    sim = q @ index_features.T
    sorted_indices = torch.from_numpy(ranking(sim))                                  # :126-127 (+ tie rule)
    sorted_index_names = np.array(index_names)[sorted_indices]                       # :128
    reference_mask = torch.tensor(sorted_index_names != np.repeat(np.array(reference_names), len(index_names))
                                  .reshape(len(target_names), -1))                   # :131-132
    sorted_index_names = sorted_index_names[reference_mask].reshape(sorted_index_names.shape[0],
                                                                    sorted_index_names.shape[1] - 1)  # :133-134
    labels = torch.tensor(sorted_index_names == np.repeat(np.array(target_names), len(index_names) - 1)
                          .reshape(len(target_names), -1))                           # :136-137
    gm = np.array(group_members)                                                     # :140
    group_mask = (sorted_index_names[..., None] == gm[:, None, :]).sum(-1).astype(bool)  # :141
    group_labels = labels[group_mask].reshape(labels.shape[0], -1)                   # :142
    assert torch.equal(torch.sum(labels, dim=-1).int(), torch.ones(len(target_names)).int())        # :144
    assert torch.equal(torch.sum(group_labels, dim=-1).int(), torch.ones(len(target_names)).int())  # :145
    rec = [(torch.sum(labels[:, :k]) / len(labels)).item() * 100 for k in (1, 5, 10, 50)]           # :148-151
    grec = [(torch.sum(group_labels[:, :k]) / len(group_labels)).item() * 100 for k in (1, 2, 3)]   # :152-154
    return (*grec, *rec)                                                             # :156


def fiq_metrics(q: torch.Tensor, index_features: torch.Tensor, index_names: List[str], target_names: List[str],
                refer_names: Optional[List[str]]):
    """Restatement of compute_fiq_val_metrics' tail: clip4cir flavour (reference removed,
    /root/reference/clip4cir/validate.py:27-49) when refer_names is given, else the blip4cir/tgcir flavour
    (/root/reference/blip4cir/validate.py:45-56)."""
    index_features = F.normalize(index_features, dim=-1).float()                     # :28
    sim = q @ index_features.T
    sorted_index_names = np.array(index_names)[ranking(sim)]                         # :31-33
    N = sim.shape[0]
    r10 = r50 = 0
    for i in range(N):                                                               # :38-45
        row = sorted_index_names[i]
        if refer_names is not None:
            row = row[row != refer_names[i]]
        if target_names[i] in row[:10]:
            r10 += 1
            r50 += 1
        elif target_names[i] in row[:50]:
            r50 += 1
    if refer_names is None:
        # blip4cir/validate.py:55-56: `(torch.sum(labels[:, :k]) / len(labels)).item() * 100` -- an int64 tensor divided by
        # an int gives a float32 tensor, so this flavour's recalls carry float32 rounding
        return ((torch.tensor(r10) / N).item() * 100, (torch.tensor(r50) / N).item() * 100)
    return (r10 / N) * 100, (r50 / N) * 100                                          # :48-49


def srm_out(clip_image_feats: torch.Tensor, clip_text_feats: torch.Tensor, srm_image_feats: torch.Tensor,
            srm_text_feats: torch.Tensor, reference_index: Sequence[int], target_index: Sequence[int], K: int = 1000):
    """Restatement of calcu_sims + get_srm_out, /root/reference/zscir/srm_utils.py:119-183 (pair mining): the four
    similarity matrices (:121-124), five full argsorts (:147-153), per triplet the six rank look-ups
    `torch.where(ranks[row] == col)[0][0]` (:166-171) summed pairwise into rt / rm / mt scores (:172-174) and the first K
    entries of the diagonal-masked intra-modal rankings as hard negatives (:176-181).
    Triplet `index` is also the row of its caption in the text matrices.  The reference's argsort is not stable; this
    restatement breaks ties by the lower index (the rule of the CUDA path).
    Returns (r_fn, m_fn, t_fn [N, K] int64, rt, rm, mt [N] int64)."""
    sims_cross_i2t = clip_image_feats @ clip_text_feats.T                     # :121
    sims_cross_t2i = sims_cross_i2t.T                                         # :122
    sims_intra_i2i = srm_image_feats @ srm_image_feats.T                      # :123
    sims_intra_t2t = srm_text_feats @ srm_text_feats.T                        # :124

    def order(s):
        return torch.argsort(s, descending=True, stable=True)                 # :147-149,152-153 (ties -> lower index)

    i2i_ranks, i2t_ranks, t2i_ranks = order(sims_intra_i2i), order(sims_cross_i2t), order(sims_cross_t2i)
    sims_intra_i2i = sims_intra_i2i.clone().fill_diagonal_(-10000)            # :150
    sims_intra_t2t = sims_intra_t2t.clone().fill_diagonal_(-10000)            # :151
    i2i_ranks_neg, t2t_ranks_neg = order(sims_intra_i2i), order(sims_intra_t2t)
    N = len(reference_index)
    rt, rm, mt = (torch.ones(N, dtype=torch.long) for _ in range(3))          # :157-159
    m_fn, t_fn, r_fn = (torch.ones((N, K), dtype=torch.long) * -10 for _ in range(3))   # :160-162

    def where(ranks_row, col):
        return torch.where(ranks_row == col)[0][0].item()

    for index in range(N):                                                    # :163
        r, t = int(reference_index[index]), int(target_index[index])         # :164-165
        rt[index] = where(i2i_ranks[r], t) + where(i2i_ranks[t], r)           # :166-167,172
        rm[index] = where(i2t_ranks[r], index) + where(t2i_ranks[index], r)   # :168-169,173
        mt[index] = where(t2i_ranks[index], t) + where(i2t_ranks[t], index)   # :170-171,174
        t_fn[index] = i2i_ranks_neg[t][:K]                                    # :176,179
        m_fn[index] = t2t_ranks_neg[index][:K]                                # :177,180
        r_fn[index] = i2i_ranks_neg[r][:K]                                    # :178,181
    return r_fn, m_fn, t_fn, rt, rm, mt


def i2i_rank_window(image_feats: torch.Tensor, rank_lo: int, rank_hi=None) -> torch.Tensor:
    """Restatement of /root/reference/zscir/get_cir_data.py:33-34 + :60-64 (CIRR twin :106-109,131-141):
    `i2i_ranks = torch.argsort(sims_intra_i2i, descending=True)` with `sims_intra_i2i = feats @ feats.T`
    (zscir/srm_utils.py:123), then `i2i_ranks[i].tolist()[rank:rank_max]` (or `[rank:]` when rank_max <= rank) for every
    row.  The reference's argsort is not stable; ties are broken by the lower index here (the rule of the CUDA path)."""
    sims = image_feats @ image_feats.T
    ranks = torch.argsort(sims, descending=True, stable=True)                  # :34
    if rank_hi is not None and rank_hi > rank_lo:                              # :60-61
        return ranks[:, rank_lo:rank_hi]
    return ranks[:, rank_lo:]                                                  # :62-64


def get_diff_id(i: int, N: int, k: int, i2i_ranks_row=None, rank_lo: int = -1, rank_hi: int = -1, rng=None):
    """get_diff_id, /root/reference/zscir/get_cir_data.py:58-68, with the module globals passed in."""
    import random as _random
    rng = rng or _random
    if rank_lo >= 0:                                                           # :59
        if rank_hi > rank_lo:
            range_target = i2i_ranks_row.tolist()[rank_lo:rank_hi]             # :61
        else:
            range_target = i2i_ranks_row.tolist()[rank_lo:]                    # :63
    else:
        range_target = list(range(N))                                          # :65
        range_target.remove(i)                                                 # :66
    return rng.sample(range_target, k)                                         # :67
