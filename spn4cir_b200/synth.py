"""Deterministic synthetic inputs of the hot path: seeded CPU generators, fp32.

Shared by tests, bench.py and the golden-vector script so that every arm sees identical tensors.
No dataset or checkpoint is needed (none is reachable offline).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

# the BASELINE.json configs (C1..C5)
CONFIGS: Dict[str, dict] = {
    "C1": dict(B=128, N=32768, D=640, T=1, tau=0.02),
    "C2": dict(Q=4181, G=2297, D=640, T=1, K=50),
    "C3": dict(B=512, N=100000, D=640, T=1, tau=0.02),
    "C4": dict(B=256, N=200000, D=256, T=32, tau=0.05),
    "C5": dict(B=1024, N=1000000, D=768, T=1, tau=0.02),
}
FIQ_SHAPES = [(2017, 3817), (2038, 6346), (1961, 5373)]  # (Q, G): dress / shirt / toptee, public statistics


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def make_bank(N: int, D: int, T: int = 1, seed: int = 1234) -> torch.Tensor:
    """Row-normalised bank [N, D] (or [N, T, D]) like extract_bank_features produces."""
    shape = (N, D) if T == 1 else (N, T, D)
    return F.normalize(torch.randn(shape, generator=_gen(seed)), dim=-1)


def make_labels(B: int, M: int, seed: int = 1235) -> torch.Tensor:
    """labels in [0, M) with at least one forced duplicate (duplicates are legal in a batch)."""
    lab = torch.randint(0, M, (B,), generator=_gen(seed), dtype=torch.int64)
    if B >= 2:
        lab[1] = lab[0]
    return lab


def make_queries(bank: torch.Tensor, labels: torch.Tensor, mode: str = "S", seed: int = 1236) -> torch.Tensor:
    """'U': uniform random unit vectors; 'S': signal = normalize(bank[label] + 3.2 * noise / sqrt(D))."""
    B = labels.shape[0]
    D = bank.shape[-1]
    noise = torch.randn((B, D), generator=_gen(seed))
    if mode == "U":
        return F.normalize(noise, dim=-1)
    tgt = bank[labels] if bank.dim() == 2 else bank[labels].mean(dim=1)
    return F.normalize(tgt + 3.2 * noise / D ** 0.5, dim=-1)


def make_bank_case(B: int, N: int, D: int, T: int = 1, mode: str = "S", seed: int = 0, label_frac: float = 1.0):
    bank = make_bank(N, D, T, seed=1234 + seed)
    labels = make_labels(B, max(1, int(N * label_frac)), seed=1235 + seed)
    q = make_queries(bank, labels, mode, seed=1236 + seed)
    return q, bank, labels


def make_retrieval_case(Q: int, G: int, D: int, T: int = 1, seed: int = 0, group_size: int = 6,
                        noise: float = 3.0, ref_pull: float = 0.0):
    """CIRR-shaped set: gallery [G, D] (or [G, T, D]); per query a group of `group_size` distinct gallery
    members (reference = member 0, target = member 1); q = normalize(gallery[target] + noise * randn / sqrt(D))
    (noise 3 -> easy, ~8 at D=640 -> target rank spread over the top few dozen).  ``ref_pull`` > 0 adds
    ref_pull * gallery[reference] to the query direction: the reference image then ranks near the top, so metrics that
    remove it (clip4cir/validate.py:38-45) differ from those that do not (blip4cir/validate.py:45-56)."""
    g = _gen(4321 + seed)
    shape = (G, D) if T == 1 else (G, T, D)
    gallery = F.normalize(torch.randn(shape, generator=g), dim=-1)
    groups = torch.stack([torch.randperm(G, generator=g)[:group_size] for _ in range(Q)])  # [Q, group_size]
    reference, target = groups[:, 0].clone(), groups[:, 1].clone()
    base = gallery[target] if T == 1 else gallery[target].mean(dim=1)
    noise_t = noise * torch.randn((Q, D), generator=g) / D ** 0.5
    if ref_pull:
        base = base + ref_pull * (gallery[reference] if T == 1 else gallery[reference].mean(dim=1))
    q = F.normalize(base + noise_t, dim=-1)
    names = [f"img{i:07d}" for i in range(G)]
    return q, gallery, names, reference, target, groups


class FakeCIRRDataset:
    """Duck-typed `CIRDataset('cirr', 'val', 'relative')`: yields (reference_name, target_name, caption,
    group_members) like /root/reference/clip4cir/data_utils.py's relative-val branch."""

    def __init__(self, names: List[str], reference, target, groups):
        self.names, self.reference, self.target, self.groups = names, reference, target, groups

    def __len__(self):
        return len(self.reference)

    def __getitem__(self, i):
        return (self.names[int(self.reference[i])], self.names[int(self.target[i])], f"q{i}",
                [self.names[int(j)] for j in self.groups[i]])


class FakeFIQDataset:
    """Duck-typed FashionIQ relative-val dataset: (reference_name, target_name, [cap1, cap2])."""

    dress_types = ["synthetic"]

    def __init__(self, names: List[str], reference, target):
        self.names, self.reference, self.target = names, reference, target

    def __len__(self):
        return len(self.reference)

    def __getitem__(self, i):
        return (self.names[int(self.reference[i])], self.names[int(self.target[i])], [f"is q{i}.", f"and more {i}?"])


class FakeModel:
    """Duck-typed CIRPlus for validation: `encode_text(captions)` returns a fixed text feature per caption
    such that combining_function(reference_feature, text) reproduces the prepared query direction."""

    def __init__(self, text_by_caption: Dict[str, torch.Tensor], output_dim: int):
        self.text_by_caption, self.output_dim = text_by_caption, output_dim

    def encode_text(self, captions):
        return torch.stack([self.text_by_caption[c] for c in captions])

    @staticmethod
    def combining_function(a, b):
        return a + b


def make_srm_case(seed: int = 21, n_img: int = 1100, n_trip: int = 1050, D: int = 64):
    """Seeded pair-mining inputs (zscir/srm_utils.py:119-183 shape): CLIP / SRM image features [n_img, D], caption features
    [n_trip, D] (triplet i owns caption row i), reference / target image ids per triplet.  K = 1000 is hard-coded in the
    reference (:155), so both sides need more than 1000 rows."""
    g = _gen(seed)
    ci = F.normalize(torch.randn(n_img, D, generator=g), dim=-1)
    si = F.normalize(torch.randn(n_img, D, generator=g), dim=-1)
    ct = F.normalize(torch.randn(n_trip, D, generator=g), dim=-1)
    st = F.normalize(torch.randn(n_trip, D, generator=g), dim=-1)
    ref = torch.randint(0, n_img, (n_trip,), generator=g)
    tgt = (ref + 1 + torch.randint(0, n_img - 1, (n_trip,), generator=g)) % n_img
    return ci, ct, si, st, ref, tgt


def make_bank_step_case(seed: int, B: int, R: int, N: int, D: int):
    """Inputs of CIRPlus.bank_large_step (clip4cir/models_negplus.py:130-142): row-normalised refer / target banks (built as
    :76-77 do), un-normalised text features, the three index vectors."""
    g = _gen(seed)
    refer_bank = F.normalize(torch.randn(R, D, generator=g), dim=-1)
    target_bank = F.normalize(torch.randn(N, D, generator=g), dim=-1)
    text = 0.7 * torch.randn(B, D, generator=g)
    indexs = torch.randint(0, R, (B,), generator=g)
    refer_indexs = torch.randint(0, R, (B,), generator=g)
    target_indexs = torch.randint(0, N, (B,), generator=g)
    return refer_bank, target_bank, text, indexs, refer_indexs, target_indexs


# ------------------------------------------------------------------------------------------------------------------
# Full-size synthetic banks generated ON THE DEVICE (bench.py and the full-size property tests): a 1 M x 768 bank is
# 3 GB of fp32 on the host, seconds of CPU time per process; on the GPU it is milliseconds.  The bank is defined chunk by
# chunk (fixed chunk count, one seed per chunk), so every rank of an N-GPU job can build exactly its own rows and any rank
# can regenerate the rows its labels point at -- the global bank is the same for every world size.
DEVICE_CHUNKS = 8


def device_bank_chunk(c: int, N: int, D: int, T: int, dev, seed: int = 1234) -> Tuple[int, int, torch.Tensor]:
    """Chunk c of DEVICE_CHUNKS: (lo, hi, rows) with rows = normalize(randn) fp32 [(hi - lo) * T, D] on `dev`."""
    per = (N + DEVICE_CHUNKS - 1) // DEVICE_CHUNKS
    lo, hi = min(N, c * per), min(N, (c + 1) * per)
    g = torch.Generator(device=dev).manual_seed(seed + 7919 * c)
    rows = torch.randn(((hi - lo) * T, D), generator=g, device=dev, dtype=torch.float32)
    return lo, hi, F.normalize(rows, dim=-1)


def make_device_case(B: int, N: int, D: int, T: int, dev, *, rank: int = 0, world: int = 1, sharded: bool = False,
                     seed: int = 0, noise: float = 3.2):
    """(bank_rows_bf16 [n_local * T, D], row_base, q fp32 [B, D], labels int64 [B]) on `dev`.  Labels are uniform over
    the WHOLE bank (they mostly live on other ranks when sharded); q = normalize(bank[label] (mean over tokens) +
    noise * randn / sqrt(D)) like make_queries('S').  `sharded`: keep only the chunks of this rank (world must divide
    DEVICE_CHUNKS)."""
    if sharded and DEVICE_CHUNKS % world != 0:
        raise ValueError(f"world size {world} must divide {DEVICE_CHUNKS}")
    g = torch.Generator(device=dev).manual_seed(99991 + 31 * seed + rank)
    labels = torch.randint(0, N, (B,), generator=g, device=dev, dtype=torch.int64)
    if B >= 2:
        labels[1] = labels[0]
    tgt = torch.zeros((B, D), dtype=torch.float32, device=dev)
    keep, row_base = [], None
    cpr = DEVICE_CHUNKS // world if sharded else DEVICE_CHUNKS
    for c in range(DEVICE_CHUNKS):
        lo, hi, rows = device_bank_chunk(c, N, D, T, dev, seed=1234 + seed)
        if hi <= lo:
            continue
        m = (labels >= lo) & (labels < hi)
        if bool(m.any()):
            sel = (labels[m] - lo)
            tgt[m] = rows.view(hi - lo, T, D)[sel].mean(dim=1)
        if (not sharded) or (rank * cpr <= c < (rank + 1) * cpr):
            if row_base is None:
                row_base = lo
            keep.append(rows.to(torch.bfloat16))
        del rows
    q = F.normalize(tgt + noise * torch.randn((B, D), generator=g, device=dev) / D ** 0.5, dim=-1)
    return torch.cat(keep), row_base, q, labels
