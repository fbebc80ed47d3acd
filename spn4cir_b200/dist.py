"""Row-sharding of the bank / gallery across the ranks of one NVSwitch box.

One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.  Each rank
scores ALL queries of the step against its own bank rows; one small exchange merges the partials:

  bank-CE (fused single pass): MIN all-reduce of the static soft-max reference [B_total]; after the sweep a small
                   SUM all_reduce of (l, s_t, es, owned) [4, B_total], the owner folds the target-row terms into its
                   partial, and ONE reduce_scatter(sum) of o_acc [B_total, D] fp32 brings every rank its own rows
  bank-CE statistics only  : all_gather of [4, B_total] (m, l, s_t, s_sum) -> spn_merge_stats
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


def global_maxnorm(bank, group=None) -> torch.Tensor:
    """max row norm over ALL shards (device float [1]); one MAX all-reduce per bank, cached on the DeviceBank."""
    if getattr(bank, "_maxnorm_global", None) is None:
        mx = bank.stats()[0].clone()
        if _world(group)[1] > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        bank._maxnorm_global = mx
    return bank._maxnorm_global


class _CudaShardCompute:
    """The per-rank compute of the sharded step: the C-ABI calls of the fused path (spn_cast_to_bf16, spn_bank_ce_prep,
    spn_bank_ce_fused, spn_shard_pack, spn_shard_fold, spn_bank_ce_finalize).  ``tests/test_dist_gloo.py`` swaps in a CPU
    emulation of the same sums so that the exchange below runs over gloo without a GPU."""

    def cast(self, q_local):
        from .ops import cast_to_bf16
        return cast_to_bf16(q_local)

    def maxnorm(self, bank, group):
        return global_maxnorm(bank, group)

    def colsum(self, bank, group):
        return global_colsum(bank, group)

    def fused_ok(self, bank, inv_tau, group) -> bool:
        """The fused sweep's static soft-max reference is provably safe while inv_tau |q| max|bank| <= 109 nats (same
        guard as ops.bank_infonce; one cached host read of the GLOBAL max row norm).  Colder temperatures (a learnable
        temp drifting below ~0.0092) take the two-pass sharded path (exact running maximum)."""
        if getattr(bank, "_maxnorm_global_host", None) is None:
            bank._maxnorm_global_host = float(global_maxnorm(bank, group).item())
        return inv_tau * bank._maxnorm_global_host * 1.03 <= 109.0

    def prep(self, q_all, lab_all, bank, inv_tau, maxnorm):
        """-> ref [Bt]: this shard's static soft-max reference (owner rows clamped around the target logit)."""
        from ._lib import call
        from .ops import _TORCH2SPN, _ptr, _stream
        dev = q_all.device
        ref = torch.empty(q_all.shape[0], dtype=torch.float32, device=dev)
        q16b = torch.empty_like(q_all)
        with torch.cuda.device(dev):
            call("spn_bank_ce_prep", _ptr(q_all), _TORCH2SPN[q_all.dtype], _ptr(bank.data), _ptr(lab_all), inv_tau,
                 _ptr(maxnorm), q_all.shape[0], bank.N, q_all.shape[1], bank.T, bank.row_base, _ptr(q16b), _ptr(ref),
                 _stream(dev))
        return ref

    def fused(self, q_all, lab_all, ref, bank, inv_tau, tau_grad):
        """-> o_acc [Bt, D], (l_acc, s_t, es_acc | None) [Bt] each, lbl_row [Bt] (LOCAL target row, -1 if not owned)."""
        from ._lib import call
        from .ops import _ptr, _stream, fused_scratch
        Bt, D = q_all.shape
        dev = q_all.device
        o_acc = torch.empty((Bt, D), dtype=torch.float32, device=dev)
        part = torch.empty((3, Bt), dtype=torch.float32, device=dev)  # l_acc, s_t, es_acc
        lbl_row = torch.empty(Bt, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            ws, ws_bytes = fused_scratch(Bt, bank, dev)
            call("spn_bank_ce_fused", _ptr(q_all), _ptr(bank.data), _ptr(lab_all), _ptr(ref), inv_tau, Bt, bank.N, D,
                 bank.T, bank.row_base, 1, _ptr(o_acc), _ptr(part[0]), _ptr(part[1]), _ptr(lbl_row),
                 _ptr(part[2]) if tau_grad else None, None, _ptr(ws), ws_bytes, _stream(dev))
        return o_acc, part[0], part[1], (part[2] if tau_grad else None), lbl_row

    def pack(self, l_acc, s_t, es_acc, lbl_row):
        """-> vec [4, Bt] = (l_acc, owner's s_t else 0, es_acc or 0, owned): what the sum all-reduce carries."""
        from ._lib import call
        from .ops import _ptr, _stream
        Bt = l_acc.shape[0]
        dev = l_acc.device
        vec = torch.empty((4, Bt), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("spn_shard_pack", _ptr(l_acc), _ptr(s_t), _ptr(es_acc), _ptr(lbl_row), Bt, _ptr(vec), _stream(dev))
        return vec

    def fold(self, o_acc, vec, ref, lbl_row, bank, eps):
        """o_acc[b] += (-l + eps L) bank[target row] on the rank that owns label b (vec = all-reduced); -> s_t [Bt]."""
        from ._lib import call
        from .ops import _ptr, _stream
        Bt, D = o_acc.shape
        dev = o_acc.device
        s_t = torch.empty(Bt, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("spn_shard_fold", _ptr(o_acc), _ptr(vec), _ptr(ref), _ptr(lbl_row), _ptr(bank.data), Bt, D, eps, _ptr(s_t),
                 _stream(dev))
        return s_t

    def finalize(self, o_mine, l_mine, st_mine, ref_mine, es_mine, q16, colsum, bank, inv_tau, B, eps):
        """This rank's B queries: (loss_mean, dq_unit [B, D], dlogit_dot | None) from the globally reduced sums."""
        from ._lib import call
        from .ops import _ptr, _stream
        D = o_mine.shape[1]
        dev = o_mine.device
        out = torch.empty((2, B), dtype=torch.float32, device=dev)  # lse, loss_per_query
        mean = torch.empty((), dtype=torch.float32, device=dev)
        dq = torch.empty((B, D), dtype=torch.float32, device=dev)
        ddot = torch.empty(B, dtype=torch.float32, device=dev) if es_mine is not None else None
        with torch.cuda.device(dev):
            call("spn_bank_ce_finalize", _ptr(o_mine), _ptr(l_mine), _ptr(st_mine), None, _ptr(ref_mine), _ptr(es_mine),
                 _ptr(q16), _ptr(colsum), inv_tau, inv_tau / B, eps, bank.n_total, B, D, _ptr(out[0]), _ptr(out[1]),
                 _ptr(mean), _ptr(dq), _ptr(ddot), _stream(dev))
        return mean, dq, ddot


def global_colsum(bank, group=None) -> torch.Tensor:
    """column sums over ALL shards (device float [D], label smoothing only); one SUM all-reduce per bank, cached."""
    if getattr(bank, "_colsum_global", None) is None:
        cs = bank.stats(True)[1].clone()
        if _world(group)[1] > 1:
            dist.all_reduce(cs, op=dist.ReduceOp.SUM, group=group)
        bank._colsum_global = cs
    return bank._colsum_global


def _reduce_scatter_sum(out: torch.Tensor, inp: torch.Tensor, rank: int, group) -> None:
    """out = rows [rank * B, (rank + 1) * B) of the sum of `inp` over the group.  NCCL: one reduce_scatter; gloo (CPU tests)
    has no reduce_scatter: all_reduce + slice."""
    if dist.get_backend(group) == "gloo":
        tmp = inp.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out.copy_(tmp[rank * out.shape[0]:(rank + 1) * out.shape[0]])
    else:
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)


_checked_batch = set()


def _check_uniform_batch(B: int, dev, group) -> None:
    """all_gather_into_tensor / reduce_scatter_tensor need the same B on every rank (a ragged last batch would hang or
    corrupt the exchange): verified once per (group, B) with one tiny all-reduce."""
    key = (id(group), B)
    if key in _checked_batch:
        return
    t = torch.tensor([B, -B], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    lo, hi = -int(t[1].item()), int(t[0].item())
    if lo != hi:
        from ._lib import SpnError
        raise SpnError("bank_infonce_sharded", -1,
                       f"every rank must bring the same number of queries per step (got between {lo} and {hi}): pad or drop the ragged last batch")
    _checked_batch.add(key)


def _sharded_sweep(q_local, bank, labels_local, inv_tau, eps, tau_grad, group, compute=None):
    """The sharded training step without autograd: every rank sweeps its shard for ALL queries of the step; the exchange
    is  all_gather(q bf16), all_gather(labels)  ->  MIN all_reduce(ref [Bt])  ->  [sweep]  ->  SUM all_reduce(vec [4, Bt])
    ->  [fold target rows]  ->  reduce_scatter(o_acc [Bt, D] fp32);  each rank finalises its own rows.
    Returns (loss_mean_of_local_queries, dq_unit [B, D] fp32, dlogit_dot | None)."""
    if compute is None and (not _CudaShardCompute().fused_ok(bank, inv_tau, group) or (eps != 0.0 and bank.T > 1)):
        # cold temperature, or label smoothing over a multi-token bank (the fused finalize cannot form sum_n s_n from column
        # sums there): forward statistics with an exact running maximum + recompute backward, merged across the ranks
        return _sharded_two_pass(q_local, bank, labels_local, inv_tau, eps, tau_grad, group)
    if compute is None and symm_path_ok(bank, q_local.shape[0], group) and _symm_setup_ok(bank, q_local.shape[0], group):
        return _sharded_sweep_symm(q_local, bank, labels_local, inv_tau, eps, tau_grad, group)
    compute = compute or _CudaShardCompute()
    rank, world = _world(group)
    B, D = q_local.shape
    dev = q_local.device
    q16 = compute.cast(q_local)
    if world > 1:
        _check_uniform_batch(B, dev, group)
        q_all = torch.empty((world * B, D), dtype=q16.dtype, device=dev)
        dist.all_gather_into_tensor(q_all, q16, group=group)
        lab_all = torch.empty(world * B, dtype=labels_local.dtype, device=dev)
        dist.all_gather_into_tensor(lab_all, labels_local, group=group)
    else:
        q_all, lab_all = q16, labels_local
    maxnorm = compute.maxnorm(bank, group)
    colsum = compute.colsum(bank, group) if eps != 0.0 else None
    ref = compute.prep(q_all, lab_all, bank, inv_tau, maxnorm)
    if world > 1:  # owner's clamped reference wins over the other ranks' Cauchy-Schwarz bound
        dist.all_reduce(ref, op=dist.ReduceOp.MIN, group=group)
    o_acc, l_acc, s_t, es_acc, lbl_row = compute.fused(q_all, lab_all, ref, bank, inv_tau, tau_grad)
    vec = compute.pack(l_acc, s_t, es_acc, lbl_row)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    st_all = compute.fold(o_acc, vec, ref, lbl_row, bank, eps)
    if world > 1:
        o_mine = torch.empty((B, D), dtype=o_acc.dtype, device=dev)
        _reduce_scatter_sum(o_mine, o_acc, rank, group)
    else:
        o_mine = o_acc
    sl = slice(rank * B, (rank + 1) * B)
    return compute.finalize(o_mine, vec[0, sl], st_all[sl], ref[sl], vec[2, sl] if tau_grad else None, q16, colsum, bank,
                            inv_tau, B, eps)


# ---------------------------------------------------------------------------------------------------------------------
# The exchange WITHOUT NCCL: peer-mapped symmetric memory + kernels that read / write the other ranks' HBM over NVLink.
#   home prologue (peer STORES of the bf16 query rows = the all-gather; peer LOADS of the target rows from the owner's
#   shard = no MIN all-reduce, no second reduce-scatter)  ->  barrier  ->  fused sweep whose TMA reduce-add epilogue
#   targets the home rank's window (= the reduce-scatter, inside the kernel)  ->  barrier  ->  finalize.
# Short sweeps / batches that are not a multiple of 128 take the "pull" form instead (symm_exchange_mode): the sweep accumulates
# into this rank's own peer-readable buffers, and after the barrier the home rank LOADS its rows from every rank and sums them
# in rank order (shard_pull_kernel) -- fewer NVLink bytes when many CTA pairs share a query tile, no atomics across ranks.
# torch.distributed._symmetric_memory only allocates the buffers and exchanges their handles (plumbing).
def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


def _ptr_array(ptrs):
    import ctypes
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


_symm_probe = {}


def symmetric_memory_available(dev: torch.device, group=None) -> bool:
    """True when torch's symmetric-memory allocator can serve EVERY rank of the group (CUDA device, NCCL group up, a probe
    allocation succeeds everywhere).  Collective on first use per group (one tiny MIN all-reduce), cached afterwards: all
    ranks take the same exchange path, or a one-sided failure would dead-lock the rendezvous."""
    if not (dist.is_available() and dist.is_initialized()) or dev.type != "cuda":
        return False
    key = id(group)
    if key not in _symm_probe:
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem
            if dist.get_backend(group) != "nccl":
                ok = 0
            else:
                symm_mem.empty(256, dtype=torch.uint8, device=dev)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        _symm_probe[key] = bool(flag.item())
    return _symm_probe[key]


def make_bank_symmetric(bank, group=None) -> None:
    """Move this rank's bank shard into peer-mapped symmetric memory (one collective rendezvous, once per bank): every
    rank can then LOAD target rows straight from the owner's shard.  All shards get the same capacity
    (ceil(n_total / world) entries); ``bank.data`` becomes a view of the symmetric buffer."""
    if getattr(bank, "peer_ptrs", None) is not None:
        return
    import torch.distributed._symmetric_memory as symm_mem
    rank, world = _world(group)
    per = (bank.n_total + world - 1) // world
    if bank.row_base != rank * per or bank.N > per:
        from ._lib import SpnError
        raise SpnError("make_bank_symmetric", -1, f"rank {rank}: shard [{bank.row_base}, +{bank.N}) is not the uniform "
                       f"ceil(n_total / world) = {per} split DeviceBank.from_tensor(shard_rank, shard_world) produces")
    dev = bank.data.device
    buf = symm_mem.empty(per * bank.T * bank.D, dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
    view = buf[: bank.N * bank.T * bank.D].view(bank.N * bank.T, bank.D)
    view.copy_(bank.data)
    if bank.N < per:
        buf[bank.N * bank.T * bank.D:].zero_()
    bank.data = view
    bank._symm = (buf, hdl)
    bank.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
    bank.rows_per_rank = per
    torch.cuda.synchronize(dev)
    dist.barrier(group)


class SymmExchange:
    """Windows of one (bank shard, B) pair in symmetric memory: the query window qall [world * B, D] bf16 with its ref /
    label metadata, this rank's accumulation windows o [B, D], l [B], es [B] fp32, the barrier flags, and (pull exchange) this
    rank's partial sums for all queries o_all [world * B, D], l_all, es_all."""

    def __init__(self, bank, B: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = _world(group)
        self.group = group
        self.B, self.D, self.Bt = B, bank.D, self.world * B
        dev = bank.data.device
        Bt, D = self.Bt, self.D
        off, total = {}, 0
        for name, nbytes in (("qall", Bt * D * 2), ("ref", Bt * 4), ("lab", Bt * 8), ("o", B * D * 4), ("l", B * 4),
                             ("es", B * 4), ("flags", 256), ("oall", Bt * D * 4), ("lall", Bt * 4), ("esall", Bt * 4)):
            off[name] = total
            total += _align(nbytes)
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = {name: _ptr_array([b + o for b in base]) for name, o in off.items()}

        def view(name, nbytes, dtype, shape):
            return self.buf[off[name]: off[name] + nbytes].view(dtype).view(shape)

        self.qall = view("qall", Bt * D * 2, torch.bfloat16, (Bt, D))
        self.ref_all = view("ref", Bt * 4, torch.float32, (Bt,))
        self.lab_all = view("lab", Bt * 8, torch.int64, (Bt,))
        self.o_win = view("o", B * D * 4, torch.float32, (B, D))
        self.l_win = view("l", B * 4, torch.float32, (B,))
        self.es_win = view("es", B * 4, torch.float32, (B,))
        # "pull" exchange: this rank's partial sums for ALL queries, read by the home ranks
        self.o_all = view("oall", Bt * D * 4, torch.float32, (Bt, D))
        self.l_all = view("lall", Bt * 4, torch.float32, (Bt,))
        self.es_all = view("esall", Bt * 4, torch.float32, (Bt,))
        self.bank_ptrs = _ptr_array(bank.peer_ptrs)
        self.epoch = 0
        # local scratch of the step (not shared): target rows, s_t, ref of the local queries; s_t / lbl_row of the sweep
        self.tgt = torch.empty((B, D), dtype=torch.float32, device=dev)
        self.s_t = torch.empty(B, dtype=torch.float32, device=dev)
        self.ref = torch.empty(B, dtype=torch.float32, device=dev)
        self.sw_st = torch.empty(Bt, dtype=torch.float32, device=dev)
        self.sw_lbl = torch.empty(Bt, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's flags are zero before the first device-side barrier

    def barrier(self, stream) -> None:
        from ._lib import call
        self.epoch += 1
        call("spn_peer_barrier", self.ptrs["flags"], self.rank, self.world, self.epoch & 0xFFFFFFFF, stream)


_exchanges = []  # one token per SymmExchange ever created (tests check that the path was taken); the objects hang off their bank


def _symm_exchange(bank, B: int, group) -> "SymmExchange":
    """The windows of (bank, B, group), created on first use (collective) and kept ON the bank object -- never keyed by
    id(bank): a recycled id must not resurrect another bank's peer pointers."""
    cache = bank.__dict__.setdefault("_symm_exchanges", {})
    key = (B, None if group is None else id(group))
    x = cache.get(key)
    if x is None:
        make_bank_symmetric(bank, group)
        x = cache[key] = SymmExchange(bank, B, group)
        _exchanges.append((B, bank.N, bank.D))
    return x


def _symm_setup_ok(bank, B: int, group) -> bool:
    """Create the windows of (bank, B) on first use and AGREE across the ranks that everybody succeeded (one MIN all-reduce,
    once): if any rank could not map its peers (no P2P, allocation failure) every rank takes the NCCL exchange instead."""
    cache = bank.__dict__.setdefault("_symm_setup", {})
    key = (B, None if group is None else id(group))
    if key not in cache:
        ok = 1
        try:
            _symm_exchange(bank, B, group)
        except Exception as exc:  # noqa: BLE001 -- any failure means "not available here"
            import warnings
            warnings.warn(f"spn4cir_b200: symmetric-memory exchange unavailable ({exc!r}); using the NCCL exchange", RuntimeWarning)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=bank.data.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        cache[key] = bool(flag.item())
        if not cache[key]:
            _symm_probe[id(group)] = False
    return cache[key]


def release_exchanges(bank=None) -> None:
    """Drop the symmetric-memory windows cached on `bank` (collective: call on every rank)."""
    if bank is not None:
        bank.__dict__.pop("_symm_exchanges", None)


def _push_shape_ok(B: int, world: int) -> bool:
    """push: a 128-query tile must have ONE home rank and the sweep runs on CTA pairs (an even number of tiles)"""
    return B % 128 == 0 and (world * B // 128) % 2 == 0


def symm_path_ok(bank, B: int, group=None) -> bool:
    """The NCCL-free exchange needs symmetric memory, world <= 8 and the uniform ceil(n_total / world) sharding (the push
    form additionally B % 128 == 0 and an even number of query tiles; other batch sizes pull)."""
    import os
    rank, world = _world(group)
    if os.environ.get("SPN_EXCHANGE", "symm") == "nccl" or world < 2 or world > 8:
        return False
    per = (bank.n_total + world - 1) // world
    return bank.row_base == rank * per and bank.N <= per and symmetric_memory_available(bank.data.device, group)


PULL_BELOW_SWEEP_S = 0.7e-3  # sweeps estimated shorter than this pull (see symm_exchange_mode)


def symm_exchange_mode(B: int, world: int, n_local: int = 0, D: int = 0, T: int = 1) -> str:
    """"push" or "pull" -- how the partial sums reach the home rank (SPN_EXCHANGE=push / pull forces one).
    push: every CTA pair's partial O block [256, D] is reduce-added into the home window by the sweep's TMA epilogue --
          74 * 256 * D * 4 bytes per rank and step whatever the batch, all of it at the END of the sweep (every CTA pair
          finishes at the same time): a ~50 us NVLink tail.
    pull: the sweep accumulates locally, the home rank loads and sums its rows from every rank (B * D * 4 bytes per peer,
          a deterministic cross-rank sum) -- two more launches, no batch-size constraint.
    Measured on 8 GPUs (profiles/r02_notes_variance_and_exchange.md): equal at C5 (2.6 ms sweeps), pull 3.5 - 4 % faster at
    C3 (0.12 ms sweeps).  Rule: short sweeps pull, long sweeps keep the reduce-scatter inside the sweep."""
    import os
    forced = os.environ.get("SPN_EXCHANGE", "symm")
    if forced == "push" and not _push_shape_ok(B, world):
        raise ValueError(f"SPN_EXCHANGE=push needs B % 128 == 0 and an even number of 128-query tiles (B={B}, world={world})")
    if forced in ("push", "pull"):
        return forced
    if not _push_shape_ok(B, world):
        return "pull"
    per_pair = 4.0 if T == 1 else 2.0 * T + 2.0
    sweep_s = per_pair * world * B * float(n_local) * D / 1.2e15  # ~ sustained tensor rate of the sweep
    return "pull" if sweep_s < PULL_BELOW_SWEEP_S else "push"


def _sharded_sweep_symm(q_local, bank, labels_local, inv_tau, eps, tau_grad, group, stage_events=None, mode=None):
    """One sharded step over NVLink-mapped windows, no NCCL call.  push: 3 kernels + 2 device-side barriers; pull: 5 + 2.
    Returns (loss_mean_of_local_queries, dq_unit [B, D] fp32, dlogit_dot | None).
    ``stage_events``: optional list that receives the timing events (start, after the home prologue, barrier 1, the sweep,
    barrier 2, [pull,] finalize) -- bench.py's exchange breakdown."""
    from ._lib import call
    from .ops import _TORCH2SPN, _ptr, _stream, fused_scratch
    B, D = q_local.shape
    dev = q_local.device
    x = _symm_exchange(bank, B, group)
    if mode is None:
        mode = symm_exchange_mode(B, x.world, bank.N, D, bank.T)
    qc = q_local.detach()
    if qc.dtype not in _TORCH2SPN:
        qc = qc.float()
    qc = qc.contiguous()
    maxnorm = global_maxnorm(bank, group)
    colsum = global_colsum(bank, group) if eps != 0.0 else None
    out = torch.empty((2, B), dtype=torch.float32, device=dev)  # lse, loss_per_query
    mean = torch.empty((), dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    ddot = torch.empty(B, dtype=torch.float32, device=dev) if tau_grad else None
    def mark():
        if stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stage_events.append(ev)

    with torch.cuda.device(dev):
        st = _stream(dev)
        mark()
        call("spn_shard_home_prologue", _ptr(qc), _TORCH2SPN[qc.dtype], _ptr(labels_local), x.bank_ptrs, bank.rows_per_rank,
             bank.n_total, bank.T, inv_tau, _ptr(maxnorm), B, D, x.rank, x.world, x.ptrs["qall"], x.ptrs["ref"], x.ptrs["lab"],
             _ptr(x.s_t), _ptr(x.ref), _ptr(x.tgt), _ptr(x.o_win), _ptr(x.l_win), _ptr(x.es_win) if tau_grad else None, st)
        mark()
        x.barrier(st)  # every rank's queries / refs / labels have landed, every window is zeroed
        mark()
        ws, ws_bytes = fused_scratch(x.Bt, bank, dev)
        if mode == "pull":
            call("spn_bank_ce_fused", _ptr(x.qall), _ptr(bank.data), _ptr(x.lab_all), _ptr(x.ref_all), inv_tau, x.Bt, bank.N, D,
                 bank.T, bank.row_base, 1, _ptr(x.o_all), _ptr(x.l_all), _ptr(x.sw_st), _ptr(x.sw_lbl),
                 _ptr(x.es_all) if tau_grad else None, None, _ptr(ws), ws_bytes, st)
            mark()
            x.barrier(st)  # every rank's partial sums are complete in its own buffers
            mark()
            call("spn_shard_pull", x.ptrs["oall"], x.ptrs["lall"], x.ptrs["esall"] if tau_grad else None, x.rank, x.world, B, D,
                 _ptr(x.o_win), _ptr(x.l_win), _ptr(x.es_win) if tau_grad else None, st)
        else:
            call("spn_bank_ce_fused_peers", _ptr(x.qall), _ptr(bank.data), _ptr(x.lab_all), _ptr(x.ref_all), inv_tau, x.Bt, bank.N,
                 D, bank.T, bank.row_base, B, x.world, x.ptrs["o"], x.ptrs["l"], x.ptrs["es"] if tau_grad else None,
                 _ptr(x.sw_st), _ptr(x.sw_lbl), _ptr(ws), ws_bytes, st)
            mark()
            x.barrier(st)  # every rank's partial sums have been reduce-added into the home windows
        mark()
        call("spn_bank_ce_finalize", _ptr(x.o_win), _ptr(x.l_win), _ptr(x.s_t), _ptr(x.tgt), _ptr(x.ref),
             _ptr(x.es_win) if tau_grad else None, _ptr(x.qall[x.rank * B:(x.rank + 1) * B]), _ptr(colsum), inv_tau,
             inv_tau / B, eps, bank.n_total, B, D, _ptr(out[0]), _ptr(out[1]), _ptr(mean), _ptr(dq), _ptr(ddot), st)
        mark()
    return mean, dq, ddot


def _sharded_two_pass(q_local, bank, labels_local, inv_tau, eps, tau_grad, group):
    """The sharded step through the two-pass kernels (any temperature, any T / label-smoothing combination): forward
    statistics (m, l, s_t, s_sum) of ALL queries on this shard -> all_gather + spn_merge_stats -> global lse -> recompute
    backward on this shard -> reduce_scatter of dq.  Slower than the fused sweep (S is computed twice); the fallback."""
    from .ops import _bwd, _fwd_stats, _loss_from_stats, cast_to_bf16
    rank, world = _world(group)
    B, D = q_local.shape
    dev = q_local.device
    q16 = cast_to_bf16(q_local)
    if world > 1:
        _check_uniform_batch(B, dev, group)
        q_all = torch.empty((world * B, D), dtype=q16.dtype, device=dev)
        dist.all_gather_into_tensor(q_all, q16, group=group)
        lab_all = torch.empty(world * B, dtype=labels_local.dtype, device=dev)
        dist.all_gather_into_tensor(lab_all, labels_local, group=group)
    else:
        q_all, lab_all = q16, labels_local
    stats = merge_stats_across_ranks(_fwd_stats(q_all, bank, lab_all, inv_tau, eps != 0.0), group)
    lse, loss_q, _ = _loss_from_stats(stats, bank.n_total, eps)
    dq_part = torch.zeros((world * B, D), dtype=torch.float32, device=dev)
    ddot = torch.zeros(world * B, dtype=torch.float32, device=dev) if tau_grad else None
    _bwd(q_all, bank, lab_all, inv_tau, lse, inv_tau / B, None, eps, dq_part, ddot)
    sl = slice(rank * B, (rank + 1) * B)
    if world > 1:
        dq = torch.empty((B, D), dtype=torch.float32, device=dev)
        _reduce_scatter_sum(dq, dq_part, rank, group)
        if tau_grad:
            dist.all_reduce(ddot, group=group)
    else:
        dq = dq_part
    return loss_q[sl].mean(), dq, (ddot[sl].contiguous() if tau_grad else None)


class _ShardedBankCE(torch.autograd.Function):
    """Single-pass bank-CE over a row-sharded bank (autograd wrapper of :func:`_sharded_sweep`)."""

    @staticmethod
    def forward(ctx, q_local, tau_t, bank, labels_local, inv_tau, eps, group):
        tau_grad = tau_t is not None and tau_t.requires_grad
        loss, dq_unit, ddot = _sharded_sweep(q_local, bank, labels_local, inv_tau, eps, tau_grad, group)
        ctx.inv_tau, ctx.B, ctx.q_dtype, ctx.tau_grad = inv_tau, q_local.shape[0], q_local.dtype, tau_grad
        ctx.save_for_backward(dq_unit, ddot if tau_grad else None)
        return loss  # mean over THIS rank's queries (what a DDP replica of the reference would log)

    @staticmethod
    def backward(ctx, g):
        dq_unit, ddot = ctx.saved_tensors
        g32 = g.detach().to(torch.float32)
        dq = (dq_unit * g32).to(ctx.q_dtype)
        dtau = (-(ctx.inv_tau / ctx.B)) * ddot.sum() * g32 if ctx.tau_grad else None
        return dq, dtau, None, None, None, None, None


def _check_tau(tau_v: float, fn: str) -> None:
    import math
    if not (tau_v > 0.0 and math.isfinite(tau_v)):
        from ._lib import SpnError
        raise SpnError(fn, -7, f"tau must be positive and finite, got {tau_v}")


def bank_infonce_sharded(query_features: torch.Tensor, bank, labels: torch.Tensor, tau=0.01, *,
                         label_smoothing: float = 0.0, group=None) -> torch.Tensor:
    """Bank-CE with the bank row-sharded over `group` (bank: this rank's DeviceBank shard with row_base /
    n_total set, labels: GLOBAL entry indices of this rank's queries; every rank brings the same number of queries).
    Returns the mean loss of the local queries; gradients w.r.t. the local queries are exact (all shards contribute
    through one reduce_scatter).  Label smoothing (zscir/models_bank.py:35) uses the all-reduced column sums of the bank.
    A label no rank owns gives loss = +inf (as the single-GPU path does)."""
    from .ops import _prep_labels, _split_tau
    tau_t, tau_v = _split_tau(tau)
    _check_tau(tau_v, "bank_infonce_sharded")
    lab = _prep_labels(labels, query_features.shape[0], query_features.device)
    return _ShardedBankCE.apply(query_features, tau_t, bank, lab, 1.0 / tau_v, float(label_smoothing), group)


def bank_ce_step_sharded(query_features: torch.Tensor, bank, labels: torch.Tensor, tau: float = 0.01, *,
                         label_smoothing: float = 0.0, group=None):
    """``(loss, dq)`` of this rank's queries against the row-sharded bank, without the autograd engine (the sharded
    counterpart of :func:`spn4cir_b200.bank_ce_step`): same kernels and the same exchange as
    :func:`bank_infonce_sharded`."""
    from .ops import _prep_labels
    _check_tau(float(tau), "bank_ce_step_sharded")
    lab = _prep_labels(labels, query_features.shape[0], query_features.device)
    loss, dq, _ = _sharded_sweep(query_features, bank, lab, 1.0 / float(tau), float(label_smoothing), False, group)
    return loss, dq


def sim_topk_sharded(query: torch.Tensor, gallery_shard: torch.Tensor, k: int, col_base: int, exclude=None,
                     group=None):
    """Top-k over a row-sharded gallery: local fused top-k, all_gather of the [Q, k] lists, device merge."""
    from .retrieval import merge_topk, sim_topk
    if k > 64:
        from ._lib import SpnError
        raise SpnError("sim_topk_sharded", -1, f"k = {k}: the device merge of per-shard lists handles k <= 64 (page deeper lists per shard with sim_topk)")
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
