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
