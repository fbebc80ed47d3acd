"""world_size-2 gloo test (CPU) of the sharding plumbing: shard geometry, the all_gather + merge exchange of
forward statistics, and the gather + merge of per-shard top-K lists.  The CUDA merge kernels are replaced
by torch stand-ins here (there is no GPU); the same exchange code runs over NCCL on the B200 box."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bank_ce as obank
from oracle import retrieval as oretr
from spn4cir_b200 import synth
from spn4cir_b200.dist import merge_stats_across_ranks, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def cpu_merge(parts):  # [R, 4, B] -> [4, B]; same arithmetic as spn_merge_stats
    m, l, st, ss = parts[:, 0], parts[:, 1], parts[:, 2], parts[:, 3]
    M = m.max(dim=0).values
    L = (l * torch.exp(m - M)).sum(dim=0)
    return torch.stack([M, L, st.max(dim=0).values, ss.sum(dim=0)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, bank, labels = synth.make_bank_case(12, 501, 64, 1, "S", 31)
        tau = 0.02
        lo, hi = shard_bounds(bank.shape[0], rank, world)
        m, l, st, ss = obank.shard_stats(q, bank[lo:hi], labels, tau, row_base=lo)
        stats = torch.stack([m, l, st, ss])
        merged = merge_stats_across_ranks(stats, merge_fn=cpu_merge)
        loss = (merged[0] + torch.log(merged[1]) - merged[2]).mean()
        dense = obank.infonce_loss(q.double(), bank.double(), labels, tau)
        assert abs(loss.item() - dense.item()) < 1e-9 * max(1.0, abs(dense.item())), (loss, dense)

        # top-K: local lists with global indices, gathered and merged
        qq, gal, _, ref, tgt, _ = synth.make_retrieval_case(9, 203, 64, 1, 5)
        lo, hi = shard_bounds(gal.shape[0], rank, world)
        sim = qq @ gal[lo:hi].T
        v, i = oretr.topk(sim, 7)
        v, i = torch.from_numpy(v), torch.from_numpy(i) + lo
        gv = [torch.empty_like(v) for _ in range(world)]
        gi = [torch.empty_like(i) for _ in range(world)]
        dist.all_gather(gv, v)
        dist.all_gather(gi, i)
        allv, alli = torch.cat(gv, dim=1), torch.cat(gi, dim=1)
        order = torch.argsort(-allv, dim=1, stable=True)[:, :7]
        got = torch.gather(alli, 1, order)
        _, want = oretr.topk(qq @ gal.T, 7)
        assert np.array_equal(got.numpy(), want)
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_over_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


# ---------------------------------------------------------------------------------------------------------------------
# The fused sharded training step (spn4cir_b200.dist._sharded_sweep: what `bench.py --gpus N` and
# bank_infonce_sharded run): the exchange -- all_gather of the queries, MIN all-reduce of the static soft-max reference,
# all_reduce of the row sums, the owner's fold of the target-row terms, ONE reduce_scatter of the additive partials,
# per-rank finalize -- is product code and runs here
# over gloo; the per-rank CUDA calls are replaced by a CPU emulation of the SAME sums (include/spn_b200.h:
# spn_bank_ce_prep / spn_bank_ce_fused / spn_bank_ce_finalize) in fp64.
class CpuShardCompute:
    def cast(self, q_local):
        return q_local.to(torch.bfloat16).double()  # the kernels score the bf16-rounded query


    def maxnorm(self, bank, group):
        mx = bank.data.norm(dim=-1).max().reshape(1).clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        return mx

    def colsum(self, bank, group):
        cs = bank.data.sum(0)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM, group=group)
        return cs

    def prep(self, q_all, lab_all, bank, inv_tau, maxnorm):
        # prep_query_kernel: Cauchy-Schwarz bound, clamped around the target logit by the rank that owns the label
        r = inv_tau * q_all.norm(dim=-1) * maxnorm * 1.01 + 1e-3
        g = lab_all - bank.row_base
        own = (g >= 0) & (g < bank.N)
        s_t = inv_tau * (q_all[own] * bank.data[g[own]]).sum(-1)
        r[own] = torch.minimum(r[own], torch.maximum(s_t + 30.0, r[own] - 64.0))
        return r

    def fused(self, q_all, lab_all, ref, bank, inv_tau, tau_grad):
        s = inv_tau * q_all @ bank.data.T                                   # [Bt, N_shard]
        g = lab_all - bank.row_base
        own = (g >= 0) & (g < bank.N)
        p = torch.exp(s - ref[:, None])
        rows = torch.nonzero(own).squeeze(1)
        p[rows, g[own]] = 0.0                                               # the label column is excluded from P~
        o_acc = p @ bank.data
        l_acc = p.sum(-1)
        es = (p * s).sum(-1)
        s_t = torch.full_like(l_acc, -float("inf"))
        s_t[own] = s[rows, g[own]]
        lbl_row = torch.where(own, g, torch.full_like(g, -1))
        return o_acc, l_acc, s_t, (es if tau_grad else None), lbl_row

    def pack(self, l_acc, s_t, es_acc, lbl_row):                            # shard_pack_kernel
        own = lbl_row >= 0
        z = torch.zeros_like(l_acc)
        return torch.stack([l_acc, torch.where(own, s_t, z), es_acc if es_acc is not None else z, own.to(l_acc.dtype)])

    def fold(self, o_acc, vec, ref, lbl_row, bank, eps):                    # shard_fold_kernel
        own = lbl_row >= 0
        L = vec[0] + torch.exp(vec[1] - ref)
        coef = -vec[0] + eps * L
        o_acc[own] += coef[own, None] * bank.data[lbl_row[own]]
        return torch.where(vec[3] > 0.5, vec[1], torch.full_like(vec[1], -float("inf")))

    def finalize(self, o_mine, l_mine, st_mine, ref_mine, es_mine, q16, colsum, bank, inv_tau, B, eps):
        # fused_finalize_kernel with the target-row terms already folded into o (bank_t == NULL)
        L = l_mine + torch.exp(st_mine - ref_mine)
        lse = ref_mine + torch.log(L)
        ssum_over_n = inv_tau * (q16 @ colsum) / bank.n_total if eps != 0.0 else 0.0
        loss = (lse - (1.0 - eps) * st_mine - eps * ssum_over_n).mean()
        dq = o_mine / L[:, None]
        if eps != 0.0:
            dq = dq - (eps / bank.n_total) * colsum[None, :]
        ddot = None
        if es_mine is not None:
            ddot = (es_mine - l_mine * st_mine) / L + eps * st_mine - eps * ssum_over_n
        return loss, (inv_tau / B) * dq, ddot


class _Shard:  # the attributes of DeviceBank the exchange reads
    def __init__(self, data, row_base, n_total):
        self.data, self.N, self.T, self.row_base, self.n_total = data, data.shape[0], 1, row_base, n_total


def _worker_fused(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spn4cir_b200.dist import _sharded_sweep
        B, N, D, tau = 6, 333, 64, 0.02
        _, bank, _ = synth.make_bank_case(B, N, D, 1, "S", 41)
        bank = bank.to(torch.bfloat16).double()                              # the resident bank is bf16
        labels = synth.make_labels(B, N, seed=50 + rank)                     # every rank brings its own queries
        q = synth.make_queries(bank.float(), labels, "S", seed=60 + rank)
        lo, hi = shard_bounds(N, rank, world)
        shard = _Shard(bank[lo:hi].contiguous(), lo, N)
        # dense fp64 oracle on this rank's queries (bf16-rounded, as the kernels see them)
        for eps in (0.0, 0.1):  # plain CE and the label-smoothing variant (zscir/models_bank.py:35)
            qd = q.to(torch.bfloat16).double().requires_grad_(True)
            logits = qd @ bank.T / tau
            dense = torch.nn.functional.cross_entropy(logits, labels, label_smoothing=eps)
            dense.backward()
            loss, dq, ddot = _sharded_sweep(q, shard, labels, 1.0 / tau, eps, True, None, compute=CpuShardCompute())
            assert abs(loss.item() - dense.item()) < 1e-9 * max(1.0, abs(dense.item())), (rank, eps, loss.item(), dense.item())
            err = (dq - qd.grad).abs().max().item() / qd.grad.abs().max().item()
            assert err < 1e-9, (rank, eps, err)
            # d loss / d tau = -(inv_tau / B) sum_b dlogit_dot[b]
            tt = torch.tensor(tau, dtype=torch.float64, requires_grad=True)
            torch.nn.functional.cross_entropy(qd.detach() @ bank.T / tt, labels, label_smoothing=eps).backward()
            dtau = -(1.0 / tau / B) * ddot.sum().item()
            assert abs(dtau - tt.grad.item()) < 1e-8 * max(1.0, abs(tt.grad.item())), (rank, eps, dtau, tt.grad.item())
        # a label that no rank owns: loss = +inf (what the single-GPU path returns), never a plausible finite value
        bad = labels.clone()
        bad[0] = N + 5
        loss_bad, _, _ = _sharded_sweep(q, shard, bad, 1.0 / tau, 0.0, False, None, compute=CpuShardCompute())
        assert torch.isinf(loss_bad) and loss_bad > 0, loss_bad
        labels_elsewhere = ((labels < lo) | (labels >= hi)).sum().item()
        out.put(("ok", rank, labels_elsewhere))
    finally:
        dist.destroy_process_group()


def test_two_rank_fused_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fused, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert [g[0] for g in got] == ["ok", "ok"]
    assert sum(g[2] for g in got) > 0  # some targets live on the other rank: the owner-only terms really travelled
