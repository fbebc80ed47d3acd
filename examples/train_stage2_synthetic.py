#!/usr/bin/env python
"""Stage-2 training loop on synthetic features with the B200 path -- the shape of clip4cir/train_negplus.py:105-123.

    python examples/train_stage2_synthetic.py                       # one GPU, whole bank resident
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_stage2_synthetic.py --sharded

A toy "text tower" (one linear layer) stands in for CLIP's encode_text; everything else is the reference's step:
autocast -> model.bank_large_step(...) -> loss.item() -> GradScaler.scale(loss).backward() -> optimizer step.  With
--sharded the target bank is row-sharded over the ranks and every rank trains on its own batch (bank_infonce_sharded).
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import spn4cir_b200 as spn  # noqa: E402
from spn4cir_b200.compat import BankLossMixin  # noqa: E402


class ToyCIR(BankLossMixin, nn.Module):
    """CIRPlus with the backbones replaced by a linear text projection (the bank methods come from the mixin)."""

    def __init__(self, dim, device, tau=0.02, plus=True):
        nn.Module.__init__(self)
        self.text_proj = nn.Linear(dim, dim)
        self.device, self.tau, self.plus = device, tau, plus
        self.to(device)

    def encode_text(self, x):
        return self.text_proj(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bank", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=640)
    ap.add_argument("--sharded", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(0)
    refer_bank = F.normalize(torch.randn(4096, args.dim, generator=g), dim=-1)
    target_bank = F.normalize(torch.randn(args.bank, args.dim, generator=g), dim=-1)
    model = ToyCIR(args.dim, dev)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    scaler = torch.amp.GradScaler("cuda")
    if args.sharded and world > 1:
        from spn4cir_b200.dist import bank_infonce_sharded
        model.refer_bank = refer_bank.to(dev)
        shard = spn.DeviceBank.from_tensor(target_bank, dev, shard_rank=rank, shard_world=world)
    else:
        model.set_banks(refer_bank, target_bank)  # upload + bf16 cast ONCE (the reference re-uploads the bank every step)
    gen = torch.Generator().manual_seed(100 + rank)
    for step in range(args.steps):
        captions = torch.randn(args.batch, args.dim, generator=gen).to(dev)            # stand-in for tokenised captions
        refer_idx = torch.randint(0, 4096, (args.batch,), generator=gen).to(dev)
        target_idx = torch.randint(0, args.bank, (args.batch,), generator=gen).to(dev)
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            text_feats = model.encode_text(captions)
            loss = {}
            if args.sharded and world > 1:
                q = F.normalize(model.refer_bank[refer_idx] + text_feats.float())
                loss["bank_loss"] = bank_infonce_sharded(q, shard, target_idx, model.tau)
            else:
                model.bank_large_step(loss, text_feats, refer_idx, target_idx, refer_idx)
        total = sum(loss.values())
        scaler.scale(total).backward()
        scaler.step(opt)
        scaler.update()
        if rank == 0 and (step % 10 == 0 or step == args.steps - 1):
            print(f"step {step:3d}  bank_loss {total.item():.4f}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
