"""A short stage-2 training loop (the shape of clip4cir/train_negplus.py:105-123: autocast -> bank_large_step ->
GradScaler.scale(loss).backward() -> optimizer step) through the drop-in mixin, against the same loop with the reference's
dense fp32 formula (models_negplus.py:130-154) in eager torch on the GPU: the loss trajectories must stay together."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import spn4cir_b200 as spn
from spn4cir_b200.compat import BankLossMixin

pytestmark = pytest.mark.gpu
DEV = "cuda"


class Toy(BankLossMixin, nn.Module):
    def __init__(self, dim, tau, plus):
        nn.Module.__init__(self)
        self.text_proj = nn.Linear(dim, dim)
        self.device, self.tau, self.plus = torch.device(DEV), tau, plus
        self.to(DEV)


class ToyReference(nn.Module):
    """the reference's own lines, eager"""

    def __init__(self, dim, tau, plus, refer_bank, target_bank):
        super().__init__()
        self.text_proj = nn.Linear(dim, dim)
        self.tau, self.plus = tau, plus
        self.refer_bank, self.target_bank = refer_bank.to(DEV), target_bank.to(DEV)  # (the reference keeps them on the CPU)
        self.crossentropy_criterion = nn.CrossEntropyLoss()
        self.to(DEV)

    def bank_large_step(self, loss, text_feats, indexs, target_indexs, refer_indexs=None):
        idx = refer_indexs if self.plus else indexs
        reference_image_feats = self.refer_bank[idx].detach().to(DEV)                    # :133-135
        query_feats = F.normalize(reference_image_feats + text_feats)                     # :136-137
        target_feats_all = self.target_bank.detach().to(DEV)                              # :138
        logits = (query_feats @ target_feats_all.T) / self.tau                            # :151
        loss["bank_loss"] = self.crossentropy_criterion(logits, target_indexs.to(DEV))    # :154


@pytest.mark.parametrize("plus", [True, False])
def test_training_trajectories_stay_together(plus):
    dim, N, R, B, tau, steps = 128, 6000, 500, 64, 0.03, 25
    g = torch.Generator().manual_seed(7)
    refer_bank = F.normalize(torch.randn(R, dim, generator=g), dim=-1)
    target_bank = F.normalize(torch.randn(N, dim, generator=g), dim=-1)
    torch.manual_seed(3)
    ours = Toy(dim, tau, plus)
    ref = ToyReference(dim, tau, plus, refer_bank, target_bank)
    ref.text_proj.load_state_dict(ours.text_proj.state_dict())
    ours.set_banks(refer_bank, target_bank)
    opts = [torch.optim.SGD(m.parameters(), lr=0.05) for m in (ours, ref)]
    scalers = [torch.amp.GradScaler("cuda", init_scale=1024.0) for _ in range(2)]
    data = torch.Generator().manual_seed(11)
    hist = [[], []]
    for _ in range(steps):
        cap = torch.randn(B, dim, generator=data).to(DEV)
        ix = torch.randint(0, R, (B,), generator=data).to(DEV)
        rix = torch.randint(0, R, (B,), generator=data).to(DEV)
        # learnable signal: the target is the entry most similar to (reference image + a fixed linear map of the caption)
        with torch.no_grad():
            want = F.normalize(refer_bank.to(DEV)[rix if plus else ix] + 0.5 * cap.roll(1, dims=1), dim=-1)
            tix = (want @ target_bank.to(DEV).T).argmax(-1)
        for k, (m, opt, sc) in enumerate(zip((ours, ref), opts, scalers)):
            opt.zero_grad(set_to_none=True)
            loss = {}
            if k == 0:
                with torch.autocast("cuda", dtype=torch.float16):
                    m.bank_large_step(loss, m.text_proj(cap), ix, tix, rix)
            else:  # the reference arm in plain fp32 (its CPU numerics)
                m.bank_large_step(loss, m.text_proj(cap), ix, tix, rix)
            sc.scale(loss["bank_loss"]).backward()
            sc.step(opt)
            sc.update()
            hist[k].append(loss["bank_loss"].item())
    a, b = torch.tensor(hist[0]), torch.tensor(hist[1])
    assert b[-5:].mean() < 0.9 * b[:5].mean(), "the reference loop did not learn: the test would be vacuous"
    assert torch.allclose(a, b, rtol=2e-2, atol=2e-2), (a - b).abs().max()
    w_rel = (ours.text_proj.weight - ref.text_proj.weight).norm() / ref.text_proj.weight.norm()
    assert w_rel < 2e-2, w_rel
