"""Row-sharding of the bank / gallery across the ranks of one NVSwitch box.

One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.  Each rank
scores ALL queries of the step against its own bank rows; one small exchange merges the partials:

  bank-CE forward : all_gather of [4, B_total] statistics (m, l, s_t, s_sum) -> spn_merge_stats
  bank-CE backward: reduce_scatter(sum) of the fp32 dq partials [B_total, D] -> each rank keeps its rows
  retrieval       : all_gather of [Q, K] (value, index) lists -> spn_merge_topk; ranks are summed counts.

Queries are data-parallel (each rank brings B_local queries, all-gathered as bf16 before scoring), so
per-rank work stays B_total * N / R and the job's throughput scales with the rank count.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of an n-row bank owned by `rank`: ceil(n / world) rows per rank, last ranks may be short."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def merge_stats_across_ranks(stats: torch.Tensor, group=None, merge_fn=None) -> torch.Tensor:
    """stats [4, B] local partial (m, l, s_t, s_sum) -> global [4, B] after one all_gather + device merge.
    `merge_fn([R, 4, B]) -> [4, B]` defaults to the CUDA merge kernel (tests inject a CPU stand-in for gloo)."""
    rank, world = _world(group)
    if world == 1:
        return stats
    gathered = _all_gather_stacked(stats, world, group)
    if merge_fn is None:
        from .ops import merge_stats as merge_fn
    return merge_fn(gathered)


def _all_gather_stacked(x: torch.Tensor, world: int, group) -> torch.Tensor:
    """all_gather of equally shaped tensors, returned as [world, *x.shape] (dim-0 concatenation, viewed)."""
    x = x.contiguous()
    flat = torch.empty((world * x.shape[0], *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=group)
    return flat.view(world, *x.shape)


class _ShardedBankCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_local, tau_t, bank, labels_local, inv_tau, eps, group):
        from .ops import _fwd_stats, _loss_from_stats, cast_to_bf16
        rank, world = _world(group)
        B = q_local.shape[0]
        q16 = cast_to_bf16(q_local)
        if world > 1:
            q_all = torch.empty((world * B, q16.shape[1]), dtype=q16.dtype, device=q16.device)
            dist.all_gather_into_tensor(q_all, q16, group=group)
            lab_all = torch.empty(world * B, dtype=labels_local.dtype, device=labels_local.device)
            dist.all_gather_into_tensor(lab_all, labels_local, group=group)
        else:
            q_all, lab_all = q16, labels_local
        stats = _fwd_stats(q_all, bank, lab_all, inv_tau, eps != 0.0)
        stats = merge_stats_across_ranks(stats, group)
        lse, loss_q, _ = _loss_from_stats(stats, bank.n_total, eps)
        ctx.bank, ctx.inv_tau, ctx.eps, ctx.group = bank, inv_tau, eps, group
        ctx.q_dtype, ctx.B, ctx.rank, ctx.world = q_local.dtype, B, rank, world
        ctx.tau_needs_grad = tau_t is not None and tau_t.requires_grad
        ctx.save_for_backward(q_all, lab_all, lse)
        # each rank reports the mean over ITS OWN queries (what a DDP replica of the reference would log)
        return loss_q[rank * B:(rank + 1) * B].mean()

    @staticmethod
    def backward(ctx, g):
        from .ops import _bwd
        q_all, lab_all, lse = ctx.saved_tensors
        Bt, D = q_all.shape
        dev = q_all.device
        B, world, rank = ctx.B, ctx.world, ctx.rank
        g32 = g.detach().to(torch.float32).contiguous()
        dq = torch.zeros((Bt, D), dtype=torch.float32, device=dev)
        ddot = torch.zeros(Bt, dtype=torch.float32, device=dev) if ctx.tau_needs_grad else None
        _bwd(q_all, ctx.bank, lab_all, ctx.inv_tau, lse, ctx.inv_tau / B, None, ctx.eps, dq, ddot)
        if world > 1:
            mine = torch.empty((B, D), dtype=torch.float32, device=dev)
            dist.reduce_scatter_tensor(mine, dq, op=dist.ReduceOp.SUM, group=ctx.group)
            if ddot is not None:
                dist.all_reduce(ddot, group=ctx.group)
        else:
            mine = dq
        mine = mine * g32
        dtau = None
        if ddot is not None:
            dtau = (-(ctx.inv_tau / B)) * ddot[rank * B:(rank + 1) * B].sum() * g32
        return mine.to(ctx.q_dtype), dtau, None, None, None, None, None


def bank_infonce_sharded(query_features: torch.Tensor, bank, labels: torch.Tensor, tau=0.01, *,
                         label_smoothing: float = 0.0, group=None) -> torch.Tensor:
    """Bank-CE with the bank row-sharded over `group` (bank: this rank's DeviceBank shard with row_base /
    n_total set, labels: GLOBAL entry indices of this rank's queries).  Returns the mean loss of the local
    queries; gradients w.r.t. the local queries are exact (all shards contribute through one
    reduce_scatter)."""
    from .ops import _prep_labels, _split_tau
    tau_t, tau_v = _split_tau(tau)
    lab = _prep_labels(labels, query_features.shape[0], query_features.device)
    return _ShardedBankCE.apply(query_features, tau_t, bank, lab, 1.0 / tau_v, float(label_smoothing), group)


def sim_topk_sharded(query: torch.Tensor, gallery_shard: torch.Tensor, k: int, col_base: int, exclude=None,
                     group=None):
    """Top-k over a row-sharded gallery: local fused top-k, all_gather of the [Q, k] lists, device merge."""
    from .retrieval import merge_topk, sim_topk
    vals, idx = sim_topk(query, gallery_shard, k, exclude=exclude, col_base=col_base)
    rank, world = _world(group)
    if world == 1:
        return vals, idx
    return merge_topk(_all_gather_stacked(vals, world, group), _all_gather_stacked(idx, world, group))


def sim_rank_sharded(query: torch.Tensor, gallery_shard: torch.Tensor, target, col_base: int, exclude=None,
                     group_members=None, group=None):
    """Rank of the target over a row-sharded gallery: the owner's target score (max all-reduce), local
    counts, sum all-reduce."""
    from .retrieval import sim_rank, sim_target_score
    ts = sim_target_score(query, gallery_shard, target, col_base=col_base)
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX, group=group)
    r, gr = sim_rank(query, gallery_shard, target, exclude=exclude, group=group_members, target_score=ts,
                     col_base=col_base)
    if world > 1:
        dist.all_reduce(r, group=group)
        if gr is not None:
            dist.all_reduce(gr, group=group)
    return r, gr
