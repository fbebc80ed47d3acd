"""Oracle for the bank contrastive loss (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def infonce_loss(query_features: torch.Tensor, target_features: torch.Tensor, labels: Optional[torch.Tensor] = None,
                 tau: float = 0.01, label_smoothing: float = 0.0) -> torch.Tensor:
    """Follows CIRPlus.infonce_loss, /root/reference/clip4cir/models_negplus.py:150-154 (criterion :33;
    label_smoothing variant /root/reference/zscir/models_bank.py:35)."""
    logits = torch.matmul(query_features, target_features.t()) / tau          # :151
    if labels is None:                                                        # :152-153
        labels = torch.arange(query_features.shape[0], dtype=torch.long)
    return F.cross_entropy(logits, labels, label_smoothing=label_smoothing)   # :154, reduction='mean'


def infonce_loss_and_grads(q: torch.Tensor, bank: torch.Tensor, labels: torch.Tensor, tau: float,
                           label_smoothing: float = 0.0, grad_out: float = 1.0, dtype=torch.float32
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """loss, dL/dq, dL/dtau by autograd through the restated forward, as train_negplus.py:121 does
    (`scaler.scale(loss).backward()` -> scalar upstream gradient `grad_out`)."""
    qq = q.detach().to(dtype).requires_grad_(True)
    tt = torch.tensor(float(tau), dtype=dtype, requires_grad=True)
    logits = torch.matmul(qq, bank.detach().to(dtype).t()) / tt
    loss = F.cross_entropy(logits, labels, label_smoothing=label_smoothing)
    loss.backward(torch.tensor(float(grad_out), dtype=dtype))
    return loss.detach(), qq.grad.detach(), tt.grad.detach()


def shard_stats(q: torch.Tensor, bank_shard: torch.Tensor, labels: torch.Tensor, tau: float, row_base: int = 0):
    """(m, l, s_t, s_sum) of one bank shard in natural-log units: the quantities spn_bank_ce_fwd returns
    (SURVEY.md 8e merge identity), computed densely in float64."""
    s = (q.double() @ bank_shard.double().t()) / tau
    if bank_shard.dim() == 3:
        raise ValueError("use blip2 helpers for 3-D banks")
    m = s.max(dim=1).values
    l = torch.exp(s - m[:, None]).sum(dim=1)
    local = labels - row_base
    inside = (local >= 0) & (local < bank_shard.shape[0])
    st = torch.full_like(m, float("-inf"))
    st[inside] = s[inside, local[inside]]
    return m, l, st, s.sum(dim=1)


def blip2_stage2_loss_loop(fusion_feats: torch.Tensor, target_feats: torch.Tensor, target_indexs: torch.Tensor,
                           temp: torch.Tensor) -> torch.Tensor:
    """Line-faithful restatement of the loss loop of Blip2QformerCirAlignPrompt.forward_stage2,
    /root/reference/blip24cir/lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py:254-268.
    fusion_feats [B, 256] (normalised, :251-253), target_feats [N, 32, 256], temp 0-d (init_stage2 :106)."""
    bs = target_indexs.shape[0]                                               # :227
    loss_qtc = torch.tensor(0, dtype=torch.float)                             # :254
    for i in range(bs):                                                       # :256
        sim_t2q_ = torch.matmul(fusion_feats[i:i + 1].unsqueeze(1).unsqueeze(1),
                                target_feats.permute(0, 2, 1)).squeeze()      # :257-259  -> [N, 32]
        sim_q2t, _ = sim_t2q_.max(-1)                                         # :260
        ground_truth_i = target_indexs[i]                                     # :261
        sim_q2t = sim_q2t / temp                                              # :262
        loss_qtc_i = F.cross_entropy(sim_q2t, ground_truth_i)                 # :263
        loss_qtc = loss_qtc + loss_qtc_i                                      # :264
    return loss_qtc / bs                                                      # :265


def blip2_stage2_loss(fusion_feats: torch.Tensor, target_feats: torch.Tensor, target_indexs: torch.Tensor,
                      temp) -> torch.Tensor:
    """Closed form of the loop above: CE(einsum('bd,ntd->bnt').amax(-1) / temp, labels)."""
    sim = torch.einsum("bd,ntd->bnt", fusion_feats, target_feats).amax(dim=-1)
    return F.cross_entropy(sim / temp, target_indexs)


def blip2_loss_and_grads(f: torch.Tensor, bank3: torch.Tensor, labels: torch.Tensor, temp: float, dtype=torch.float32):
    ff = f.detach().to(dtype).requires_grad_(True)
    tt = torch.tensor(float(temp), dtype=dtype, requires_grad=True)
    loss = blip2_stage2_loss(ff, bank3.detach().to(dtype), labels, tt)
    loss.backward()
    return loss.detach(), ff.grad.detach(), tt.grad.detach()


def bank_large_step_query(refer_bank: torch.Tensor, idx: torch.Tensor, text_feats: torch.Tensor) -> torch.Tensor:
    """Query assembly of CIRPlus.bank_large_step, /root/reference/clip4cir/models_negplus.py:133-137:
    normalize(refer_bank[idx] + text_feats)."""
    return F.normalize(refer_bank[idx] + text_feats)


def part_infonce_loss(query_feats: torch.Tensor, target_feats_all: torch.Tensor, target_indexs: torch.Tensor,
                      neg_num: int, tau: float, rng=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Line-faithful restatement of CIRPlus.part_infonce_loss + get_neg_id,
    /root/reference/clip4cir/models.py:109-128 (same loop at /root/reference/zscir/models_bank.py:103-122).
    Returns (loss, candidate ids [N, 1 + neg_num]); rng defaults to python's global `random`, as in the reference."""
    import random as _random
    rng = rng or _random
    M = target_feats_all.shape[0]                                             # :116
    N = query_feats.shape[0]                                                  # :117
    target_loss = torch.tensor(0, dtype=torch.float)                          # :118
    cands = []
    for i in range(N):                                                        # :119
        pos_id = target_indexs[i].item()                                      # :120
        range_target = list(range(M))                                         # :110  (get_neg_id)
        range_target.remove(pos_id)                                           # :111
        neg_ids = rng.sample(range_target, neg_num)                           # :112
        neg_feats = target_feats_all[neg_ids]                                 # :122
        pos_feats = target_feats_all[pos_id].unsqueeze(0)                     # :123
        target_feats = torch.cat([pos_feats, neg_feats])                      # :124
        target_loss = target_loss + infonce_loss(query_feats[i], target_feats, torch.tensor(0, dtype=torch.long), tau=tau)  # :125-126
        cands.append([pos_id] + list(neg_ids))
    return target_loss / N, torch.tensor(cands, dtype=torch.long)             # :127-128


def part_loss_and_grads(q: torch.Tensor, bank: torch.Tensor, cand: torch.Tensor, tau: float, grad_out: float = 1.0,
                        dtype=torch.float64):
    """Closed form of the loop above for given candidate ids: mean_b CE(<q_b, bank[cand_b]> / tau, 0) + autograd."""
    qq = q.detach().to(dtype).requires_grad_(True)
    tt = torch.tensor(float(tau), dtype=dtype, requires_grad=True)
    logits = torch.einsum("bd,bkd->bk", qq, bank.detach().to(dtype)[cand]) / tt
    loss = F.cross_entropy(logits, torch.zeros(q.shape[0], dtype=torch.long))
    loss.backward(torch.tensor(float(grad_out), dtype=dtype))
    return loss.detach(), qq.grad.detach(), tt.grad.detach()
