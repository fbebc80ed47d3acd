"""NVLink byte counters around K sharded steps (run under torchrun, N >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/nvlink_counters.py --config C5 --steps 200 --out gpurun_out/nvlink_n2.json

Rank 0 reads `nvidia-smi nvlink -gt d -i <gpu>` (cumulative data Tx / Rx KiB per link) before and after K steps of
bank_ce_step_sharded, once with the symmetric-memory exchange (peer stores / TMA reduce-add inside the kernels) and once with
SPN_EXCHANGE=nccl, and prints the bytes per step next to the bytes the design predicts.  Evidence only: nothing here is a
timed number.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import spn4cir_b200 as spn
from spn4cir_b200 import synth
from spn4cir_b200.dist import bank_ce_step_sharded, release_exchanges

CONFIGS = {"C5": dict(B=1024, N=1_000_000, D=768), "C3": dict(B=512, N=100_000, D=640)}


def read_counters(gpu: int):
    """(tx_kib, rx_kib, raw) summed over the links of one GPU; (None, None, raw) when the driver prints another format."""
    res = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True)
    raw = res.stdout + res.stderr
    tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", raw)]
    rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", raw)]
    if not tx or not rx:
        return None, None, raw
    return sum(tx), sum(rx), raw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default="gpurun_out/nvlink.json")
    args = ap.parse_args()
    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    c = CONFIGS[args.config]
    B, N, D, tau = c["B"], c["N"], c["D"], 0.01
    rows, row_base, q, lab = synth.make_device_case(B, N, D, 1, dev, rank=rank, world=world, sharded=True)
    out = {"config": args.config, "world": world, "B_per_rank": B, "N": N, "D": D, "steps": args.steps, "gpu": local}
    for mode in ("symm", "nccl"):
        os.environ["SPN_EXCHANGE"] = mode
        bank = spn.DeviceBank(rows, 1, row_base=row_base, n_total=N)
        for _ in range(5):
            bank_ce_step_sharded(q, bank, lab, tau)
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        t0, r0, raw0 = read_counters(local) if rank == 0 else (None, None, "")
        dist.barrier()
        for _ in range(args.steps):
            loss, dq = bank_ce_step_sharded(q, bank, lab, tau)
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        t1, r1, raw1 = read_counters(local) if rank == 0 else (None, None, "")
        dist.barrier()
        if rank == 0:
            rec = {"loss": float(loss)}
            if t0 is None or t1 is None:
                rec["unparsed"] = raw1[-1500:]
            else:
                rec["tx_bytes_per_step"] = (t1 - t0) * 1024 / args.steps
                rec["rx_bytes_per_step"] = (r1 - r0) * 1024 / args.steps
            out[mode] = rec
        release_exchanges(bank)
        del bank
    if rank == 0:
        Bt = B * world
        pair_tiles = Bt // 256
        pairs = 74
        chunks = max(1.0, pairs / pair_tiles)  # bank chunks (= partial O blocks) per pair tile on one rank
        remote = (world - 1) / world
        out["predicted"] = {
            "symm_tx_o_partials": remote * pair_tiles * chunks * 256 * D * 4,
            "symm_tx_q_gather": (world - 1) * B * D * 2,
            "nccl_reduce_scatter_o_min": remote * Bt * D * 4,
            "nccl_all_gather_q_min": (world - 1) * B * D * 2,
            "note": "symm: every CTA pair's partial O block of a remote home tile is reduce-added over NVLink by the sweep's TMA "
                    "epilogue (chunks per pair tile = 74 / pair tiles, not pre-reduced locally) + bf16 q rows stored to every peer; "
                    "nccl: ring/NVLS reduce-scatter of the locally reduced fp32 o [B_total, D] + all-gather of bf16 q (lower bounds; "
                    "protocol overhead and the small vectors come on top)"}
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
