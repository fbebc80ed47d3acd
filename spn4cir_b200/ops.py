"""Torch-facing operators over the C ABI: bank-CE (forward statistics / loss / backward) as an
``autograd.Function`` and the device-resident bank.

torch is plumbing here (device memory, streams, autograd graph); every arithmetic step of the path
runs in libspn_b200's kernels.  Reference semantics: ``CIRPlus.infonce_loss``
(/root/reference/clip4cir/models_negplus.py:150-154 and siblings) and the BLIP-2 ``forward_stage2``
tail (/root/reference/blip24cir/lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py:254-268).
"""
from __future__ import annotations

import math
import os
from typing import Optional, Union

import torch

from . import _lib, _torch_ext
from ._lib import DT_BF16, DT_F16, DT_F32, OP_BANK_CE_FWD, SpnError, call

_TORCH2SPN = {torch.float32: DT_F32, torch.float16: DT_F16, torch.bfloat16: DT_BF16}


def _use_torch_ext() -> bool:
    """The torch.library operators (C++ autograd node, ~4x less host time per step) are the default entry into the fused
    step; SPN_TORCH_EXT=0 selects the ctypes + Python autograd.Function route (same kernels) for A/B tests."""
    return os.environ.get("SPN_TORCH_EXT", "1") != "0"


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device: torch.device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise SpnError("spn4cir_b200", -5, f"{name} must be a CUDA tensor: this path has no CPU fallback")


def cast_to_bf16(x: torch.Tensor) -> torch.Tensor:
    """Contiguous bf16 copy of a CUDA tensor through ``spn_cast_to_bf16`` (no-op if already bf16)."""
    _require_cuda(x, "tensor")
    x = x.detach()
    if x.dtype == torch.bfloat16 and x.is_contiguous():
        return x
    if x.dtype not in _TORCH2SPN:
        x = x.float()
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        call("spn_cast_to_bf16", _ptr(x), _TORCH2SPN[x.dtype], _ptr(out), x.numel(), _stream(x.device))
    return out


class DeviceBank:
    """A static memory bank resident in HBM as bf16 rows.

    Replaces the per-step ``self.target_bank.detach().to(self.device)``
    (/root/reference/clip4cir/models_negplus.py:138, blip24cir/models.py:119): the bank is uploaded and cast
    once; every step reads it in place.  ``data`` is ``[N * T, D]`` bf16; ``row_base`` / ``n_total``
    describe the shard when the bank is row-sharded across ranks (whole entries per shard).
    """

    def __init__(self, data_bf16: torch.Tensor, tokens_per_entry: int = 1, row_base: int = 0,
                 n_total: Optional[int] = None):
        _require_cuda(data_bf16, "bank")
        if data_bf16.dtype != torch.bfloat16 or data_bf16.dim() != 2 or not data_bf16.is_contiguous():
            raise SpnError("DeviceBank", -3, "expects a contiguous 2-D bf16 CUDA tensor [N*T, D]")
        if data_bf16.shape[0] % tokens_per_entry != 0:
            raise SpnError("DeviceBank", -1, "row count is not a multiple of tokens_per_entry")
        self.data = data_bf16
        self.T = int(tokens_per_entry)
        self.N = data_bf16.shape[0] // self.T
        self.D = data_bf16.shape[1]
        self.row_base = int(row_base)
        self.n_total = int(n_total) if n_total is not None else self.N
        self._maxnorm = None  # device float [1]: max row norm (spn_bank_stats), computed once on first use
        self._colsum = None   # device float [D]: column sums (label smoothing only)

    def stats(self, want_colsum: bool = False):
        """(maxnorm [1], colsum [D] | None) of this shard, computed once by ``spn_bank_stats``."""
        if self._maxnorm is None or (want_colsum and self._colsum is None):
            dev = self.data.device
            mx = torch.empty(1, dtype=torch.float32, device=dev)
            cs = torch.empty(self.D, dtype=torch.float32, device=dev) if want_colsum else None
            with torch.cuda.device(dev):
                call("spn_bank_stats", _ptr(self.data), self.data.shape[0], self.D, _ptr(mx), _ptr(cs), _stream(dev))
            self._maxnorm = mx
            if want_colsum:
                self._colsum = cs
        return self._maxnorm, self._colsum

    def maxnorm_host(self) -> float:
        """max row norm as a python float (one device read per bank, cached)."""
        if getattr(self, "_maxnorm_host", None) is None:
            self._maxnorm_host = float(self.stats()[0].item())
        return self._maxnorm_host

    @property
    def device(self) -> torch.device:
        return self.data.device

    # ---- tensor-like no-ops, so that reference code written for a CPU tensor bank keeps working unchanged when the
    # attribute holds a resident bank: `self.target_bank.detach().to(self.device)` (clip4cir/models_negplus.py:138,
    # blip4cir/models.py:102, tgcir/models.py:279, blip24cir/models.py:119) and the gather `target_feats_all[target_indexs]`
    # (models_negplus.py:140; its result is unused there)
    def detach(self) -> "DeviceBank":
        return self

    def to(self, device=None, *args, **kwargs) -> "DeviceBank":
        if device is not None and not isinstance(device, torch.dtype) and torch.device(device).type == "cuda" and \
                torch.device(device).index not in (None, self.data.device.index):
            raise SpnError("DeviceBank", -7, f"the bank is resident on {self.data.device}; build another DeviceBank for {device}")
        return self

    @property
    def shape(self):
        return (self.N, self.D) if self.T == 1 else (self.N, self.T, self.D)

    def __len__(self) -> int:
        return self.N

    def __getitem__(self, idx) -> torch.Tensor:
        """rows of the resident bank as fp32 (``[len, D]``, or ``[len, T, D]`` for multi-token banks)"""
        idx = torch.as_tensor(idx, device=self.data.device)
        rows = self.data.view(self.N, self.T, self.D)[idx].float()
        return rows.squeeze(-2) if self.T == 1 else rows

    @classmethod
    def from_tensor(cls, bank: torch.Tensor, device: Union[str, torch.device] = "cuda", *, shard_rank: int = 0,
                    shard_world: int = 1) -> "DeviceBank":
        """``bank``: ``[N, D]`` or ``[N, T, D]`` float tensor (CPU or CUDA).  With ``shard_world > 1`` only
        rows ``[r * ceil(N/R), ...)`` are uploaded."""
        if bank.dim() == 3:
            n, t, d = bank.shape
        elif bank.dim() == 2:
            (n, d), t = bank.shape, 1
        else:
            raise SpnError("DeviceBank", -1, f"bank must be [N, D] or [N, T, D], got {tuple(bank.shape)}")
        per = (n + shard_world - 1) // shard_world
        lo, hi = min(n, shard_rank * per), min(n, (shard_rank + 1) * per)
        if hi <= lo:
            raise SpnError("DeviceBank", -1, f"rank {shard_rank} of {shard_world} owns no rows of a {n}-entry bank")
        part = bank[lo:hi].detach().reshape((hi - lo) * t, d)
        dev = torch.device(device)
        if part.dtype not in _TORCH2SPN:
            part = part.float()
        part = part.to(dev, non_blocking=True).contiguous()
        return cls(cast_to_bf16(part), t, row_base=lo, n_total=n)

    @classmethod
    def from_file(cls, path: str, device: Union[str, torch.device] = "cuda", index: int = 1,
                  unlabeled_path: Optional[str] = None, neg_num: int = 0, **kw) -> "DeviceBank":
        """Load the reference's cached bank file(s) and make the target bank resident: see :func:`load_bank_file`.
        ``index`` selects the list element (1 = target bank); ``unlabeled_path`` appends the unlabeled rows after the
        labeled ones (labels keep pointing into the first ``bank.M`` rows)."""
        banks = load_bank_file(path, unlabeled_path=unlabeled_path, neg_num=neg_num)
        src = banks["target_bank"] if index == 1 else banks["items"][index]
        bank = cls.from_tensor(src, device, **kw)
        bank.M = banks["M"]
        return bank


def load_bank_file(path: str, unlabeled_path: Optional[str] = None, neg_num: int = 0) -> dict:
    """The on-disk bank contract of the reference, read on the host:

    * ``torch.save([refer_bank, target_bank], bank_path)`` (/root/reference/clip4cir/models_negplus.py:78, loaded :80) --
      BLIP-2 stores three elements ``[refer_bank, target_bank, query_bank]`` (/root/reference/blip24cir/models.py:81,
      loaded :83-89); a bare tensor is accepted as the target bank;
    * ``torch.save([unlabeled_target_bank], path)`` (:114, loaded :116), optionally truncated to the first ``neg_num``
      rows (:117-122), concatenated AFTER the labeled rows (:123-125): ``M`` = number of labeled rows, labels only
      point into ``[0, M)``.

    Returns ``{"refer_bank", "target_bank", "query_bank", "M", "items"}`` (CPU tensors, not cast)."""
    obj = torch.load(path, map_location="cpu")
    items = list(obj) if isinstance(obj, (list, tuple)) else [None, obj]
    if len(items) not in (2, 3):
        raise SpnError("load_bank_file", -1, f"{path}: expected [refer_bank, target_bank(, query_bank)], got {len(items)} items")
    refer_bank, target_bank = items[0], items[1]
    query_bank = items[2] if len(items) == 3 else None
    M = int(target_bank.shape[0])
    if unlabeled_path is not None:
        un = torch.load(unlabeled_path, map_location="cpu")
        un = un[0] if isinstance(un, (list, tuple)) else un
        if neg_num > 0:
            un = un[:neg_num]
        if un.shape[1:] != target_bank.shape[1:]:
            raise SpnError("load_bank_file", -1, f"unlabeled rows {tuple(un.shape)} do not match the bank {tuple(target_bank.shape)}")
        target_bank = torch.cat([target_bank, un.to(target_bank.dtype)])
    return {"refer_bank": refer_bank, "target_bank": target_bank, "query_bank": query_bank, "M": M, "items": items}


def _fwd_stats(q_bf16: torch.Tensor, bank: DeviceBank, labels: torch.Tensor, inv_tau: float, want_ssum: bool):
    B, D = q_bf16.shape
    dev = q_bf16.device
    stats = torch.empty((4, B), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = _lib.workspace_bytes(OP_BANK_CE_FWD, B, bank.N, D, bank.T)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        call("spn_bank_ce_fwd", _ptr(q_bf16), _ptr(bank.data), _ptr(labels), inv_tau, B, bank.N, D, bank.T,
             bank.row_base, _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]) if want_ssum else None,
             _ptr(ws), ws_bytes, _stream(dev))
    return stats


def _loss_from_stats(stats: torch.Tensor, n_total: int, eps: float):
    """stats [4, B] (m, l, s_t, s_sum) -> (lse [B], loss_per_query [B], loss_mean [])."""
    B = stats.shape[1]
    dev = stats.device
    out = torch.empty((2, B), dtype=torch.float32, device=dev)
    mean = torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("spn_bank_ce_loss", _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]) if eps != 0.0 else None,
             B, n_total, eps, _ptr(out[0]), _ptr(out[1]), _ptr(mean), _stream(dev))
    return out[0], out[1], mean


def merge_stats(parts: torch.Tensor) -> torch.Tensor:
    """parts [P, 4, B] (m, l, s_t, s_sum per partial) -> merged [4, B] (``spn_merge_stats``)."""
    P, four, B = parts.shape
    assert four == 4
    dev = parts.device
    flat = parts.permute(1, 0, 2).contiguous()  # [4, P, B]
    out = torch.empty((4, B), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("spn_merge_stats", _ptr(flat[0]), _ptr(flat[1]), _ptr(flat[2]), _ptr(flat[3]), P, B, _ptr(out[0]),
             _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _stream(dev))
    return out


def _bwd(q_bf16, bank: DeviceBank, labels, inv_tau, lse, gscale, gscale_dev, eps, dq, dlogit_dot):
    B, D = q_bf16.shape
    dev = q_bf16.device
    with torch.cuda.device(dev):
        call("spn_bank_ce_bwd", _ptr(q_bf16), _ptr(bank.data), _ptr(labels), inv_tau, _ptr(lse), gscale,
             _ptr(gscale_dev), eps, bank.n_total, B, bank.N, D, bank.T, bank.row_base, _ptr(dq), _ptr(dlogit_dot),
             None, 0, _stream(dev))


class _BankCE(torch.autograd.Function):
    """loss = mean_b CE(inv_tau * sim(q_b, bank), label_b); grads for q and (optionally) tau."""

    @staticmethod
    def forward(ctx, q, tau_t, bank: DeviceBank, labels, inv_tau: float, eps: float):
        q_bf16 = cast_to_bf16(q)
        stats = _fwd_stats(q_bf16, bank, labels, inv_tau, eps != 0.0)
        lse, _, loss = _loss_from_stats(stats, bank.n_total, eps)
        ctx.bank, ctx.inv_tau, ctx.eps = bank, inv_tau, eps
        ctx.q_dtype = q.dtype
        ctx.tau_needs_grad = tau_t is not None and tau_t.requires_grad
        ctx.save_for_backward(q_bf16, labels, lse)
        return loss

    @staticmethod
    def backward(ctx, g):
        q_bf16, labels, lse = ctx.saved_tensors
        B, D = q_bf16.shape
        dev = q_bf16.device
        g32 = g.detach().to(torch.float32).contiguous()
        dq = torch.zeros((B, D), dtype=torch.float32, device=dev)
        ddot = torch.zeros(B, dtype=torch.float32, device=dev) if ctx.tau_needs_grad else None
        _bwd(q_bf16, ctx.bank, labels, ctx.inv_tau, lse, ctx.inv_tau / B, g32, ctx.eps, dq, ddot)
        dtau = None
        if ctx.tau_needs_grad:
            # dL/dtau = -(g / B) * inv_tau * sum_b sum_n dlogit[b,n] * s[b,n]
            dtau = (-(ctx.inv_tau / B)) * ddot.sum() * g32
        return dq.to(ctx.q_dtype), dtau, None, None, None, None


_STEP_WS = {}


def fused_scratch(B: int, bank: DeviceBank, dev: torch.device):
    """(buffer | None, bytes): scratch of ``spn_bank_ce_fused`` for the two-pass kernel (P~ in bf16, [B_pad, N_pad]);
    0 bytes when the shape runs single-pass."""
    key = ("fused", B, bank.N, bank.D, bank.T)
    if key not in _STEP_WS:
        _STEP_WS[key] = _lib.workspace_bytes(_lib.OP_BANK_CE_FUSED, B, bank.N, bank.D, bank.T)
    n = _STEP_WS[key]
    return (_scratch_buffer(n, dev) if n else None), n


_SCRATCH = {}


def _scratch_buffer(nbytes: int, dev: torch.device) -> torch.Tensor:
    """One scratch buffer per (device, stream), grown on demand and reused by every step (the P~ scratch is 2 GB at
    1024 x 1 M: it is write-before-read inside one launch, so successive steps on a stream may share it)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(dev))
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _SCRATCH.pop(key, None)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _SCRATCH[key] = buf
    return buf


def release_scratch() -> None:
    """Drop the cached step scratch buffers."""
    _SCRATCH.clear()


def _fused_sweep(q: torch.Tensor, bank: DeviceBank, labels: torch.Tensor, inv_tau: float, want_es: bool,
                 ref: Optional[torch.Tensor] = None):
    """prep (cast + reference) and one fused sweep over this shard.  Returns a dict of device buffers."""
    B, D = q.shape
    dev = q.device
    qc = q.detach()
    if qc.dtype not in _TORCH2SPN:
        qc = qc.float()
    qc = qc.contiguous()
    maxnorm, _ = bank.stats()
    q16 = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
    my_ref = torch.empty(B, dtype=torch.float32, device=dev)
    o_acc = torch.empty((B, D), dtype=torch.float32, device=dev)
    bank_t = torch.empty((B, D), dtype=torch.float32, device=dev)
    vec = torch.empty((3, B), dtype=torch.float32, device=dev)  # l_acc, s_t, es_acc
    lbl_row = torch.empty(B, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = _stream(dev)
        call("spn_bank_ce_prep", _ptr(qc), _TORCH2SPN[qc.dtype], _ptr(bank.data), _ptr(labels), inv_tau, _ptr(maxnorm), B,
             bank.N, D, bank.T, bank.row_base, _ptr(q16), _ptr(my_ref), st)
        if ref is None:
            ref = my_ref
        ws, ws_bytes = fused_scratch(B, bank, dev)
        call("spn_bank_ce_fused", _ptr(q16), _ptr(bank.data), _ptr(labels), _ptr(ref), inv_tau, B, bank.N, D, bank.T,
             bank.row_base, 1, _ptr(o_acc), _ptr(vec[0]), _ptr(vec[1]), _ptr(lbl_row), _ptr(vec[2]) if want_es else None,
             _ptr(bank_t), _ptr(ws), ws_bytes, st)
    return dict(q16=q16, ref=ref, my_ref=my_ref, o_acc=o_acc, bank_t=bank_t, l_acc=vec[0], s_t=vec[1],
                es_acc=vec[2] if want_es else None, lbl_row=lbl_row)


def _fused_finalize(sw: dict, bank: DeviceBank, inv_tau: float, b_total: int, eps: float, want_dq: bool):
    B, D = sw["o_acc"].shape
    dev = sw["o_acc"].device
    colsum = bank.stats(True)[1] if eps != 0.0 else None
    out = torch.empty((2, B), dtype=torch.float32, device=dev)  # lse, loss_per_query
    mean = torch.empty((), dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev) if want_dq else None
    ddot = torch.empty(B, dtype=torch.float32, device=dev) if sw["es_acc"] is not None else None
    with torch.cuda.device(dev):
        call("spn_bank_ce_finalize", _ptr(sw["o_acc"]), _ptr(sw["l_acc"]), _ptr(sw["s_t"]), _ptr(sw["bank_t"]),
             _ptr(sw["ref"]), _ptr(sw["es_acc"]), _ptr(sw["q16"]), _ptr(colsum), inv_tau, inv_tau / b_total, eps,
             bank.n_total, B, D, _ptr(out[0]), _ptr(out[1]), _ptr(mean), _ptr(dq), _ptr(ddot), _stream(dev))
    return out[0], out[1], mean, dq, ddot


def _fused_step(q: torch.Tensor, bank: DeviceBank, labels: torch.Tensor, inv_tau: float, eps: float, want_ddot: bool):
    """The whole single-GPU training step in ONE library call (``spn_bank_ce_step``: prologue -> fused sweep ->
    finalize, three launches).  Returns (lse, loss_per_query, loss_mean, dq_unit, dlogit_dot | None)."""
    B, D = q.shape
    dev = q.device
    qc = q.detach()
    if qc.dtype not in _TORCH2SPN:
        qc = qc.float()
    qc = qc.contiguous()
    maxnorm, colsum = bank.stats(eps != 0.0)
    key = (B, bank.N, D, bank.T)
    if key not in _STEP_WS:
        _STEP_WS[key] = _lib.workspace_bytes(_lib.OP_BANK_CE_STEP, B, bank.N, D, bank.T)
    ws_bytes = _STEP_WS[key]
    ws = _scratch_buffer(ws_bytes, dev)
    out = torch.empty((3, B), dtype=torch.float32, device=dev)  # lse, loss_per_query, dlogit_dot
    mean = torch.empty((), dtype=torch.float32, device=dev)
    dq = torch.empty((B, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("spn_bank_ce_step", _ptr(qc), _TORCH2SPN[qc.dtype], _ptr(bank.data), _ptr(labels), inv_tau, _ptr(maxnorm),
             _ptr(colsum) if eps != 0.0 else None, eps, B, bank.N, D, bank.T, _ptr(out[0]), _ptr(out[1]), _ptr(mean),
             _ptr(dq), _ptr(out[2]) if want_ddot else None, _ptr(ws), ws_bytes, _stream(dev))
    return out[0], out[1], mean, dq, out[2] if want_ddot else None


class _BankCEFused(torch.autograd.Function):
    """Single-pass training path: forward already produces the gradient for an upstream gradient of 1
    (spn_bank_ce_prep -> spn_bank_ce_fused -> spn_bank_ce_finalize); backward only scales it."""

    @staticmethod
    def forward(ctx, q, tau_t, bank: DeviceBank, labels, inv_tau: float, eps: float):
        tau_grad = tau_t is not None and tau_t.requires_grad
        if bank.row_base == 0 and bank.n_total == bank.N:
            _, _, loss, dq_unit, ddot = _fused_step(q, bank, labels, inv_tau, eps, tau_grad)
        else:  # one shard of a larger bank, driven piecewise (the exchange happens in dist.py)
            sw = _fused_sweep(q, bank, labels, inv_tau, tau_grad)
            _, _, loss, dq_unit, ddot = _fused_finalize(sw, bank, inv_tau, q.shape[0], eps, True)
        ctx.inv_tau, ctx.B, ctx.q_dtype, ctx.tau_grad = inv_tau, q.shape[0], q.dtype, tau_grad
        ctx.save_for_backward(dq_unit, ddot if tau_grad else None)
        return loss

    @staticmethod
    def backward(ctx, g):
        dq_unit, ddot = ctx.saved_tensors
        g32 = g.detach().to(torch.float32)
        dq = (dq_unit * g32).to(ctx.q_dtype)
        dtau = None
        if ctx.tau_grad:
            dtau = (-(ctx.inv_tau / ctx.B)) * ddot.sum() * g32
        return dq, dtau, None, None, None, None


_warned_raw_bank = False


def _warn_raw_bank(bank: torch.Tensor) -> None:
    """A raw tensor bank is uploaded, cast to bf16 and scanned for its max row norm on EVERY call (the reference's per-step
    `.to(device)`, models_negplus.py:138, made worse): fine for tests, a bug in a training loop.  Said once."""
    global _warned_raw_bank
    if not _warned_raw_bank and bank.numel() >= (1 << 22):
        import warnings
        warnings.warn("spn4cir_b200: a raw tensor was passed as the memory bank; it is re-uploaded and re-cast on every call. "
                      "Build a DeviceBank once (DeviceBank.from_tensor / from_file, BankLossMixin.set_banks / load_banks).",
                      RuntimeWarning, stacklevel=3)
        _warned_raw_bank = True


def _prep_labels(labels: Optional[torch.Tensor], B: int, dev: torch.device) -> torch.Tensor:
    if B <= 0:
        raise SpnError("bank_ce", -1, "empty batch (B = 0): the loss of zero queries is undefined (the reference's "
                       "F.cross_entropy returns NaN); nothing was launched")
    if labels is not None and labels.numel() != B:
        raise SpnError("bank_ce", -1, f"{labels.numel()} labels for {B} queries")
    if labels is None:  # models_negplus.py:152-153
        return torch.arange(B, dtype=torch.int64, device=dev)
    return labels.to(device=dev, dtype=torch.int64).contiguous()


def _split_tau(tau):
    if isinstance(tau, torch.Tensor):
        val = float(tau.detach().float().item())  # learnable temperature (blip4cir/models.py:29): one host read
        return tau, val
    return None, float(tau)


def bank_infonce(query_features: torch.Tensor, bank: Union[DeviceBank, torch.Tensor],
                 labels: Optional[torch.Tensor] = None, tau=0.01, *, label_smoothing: float = 0.0,
                 single_pass: Optional[bool] = None) -> torch.Tensor:
    """``CrossEntropy((q @ bank.T) / tau, labels)`` without materialising the logits.

    ``bank``: a :class:`DeviceBank` (preferred: resident, cast once) or a CUDA float tensor
    ``[N, D]`` / ``[N, T, D]`` (cast on every call).  With ``T > 1`` the logit of an entry is the max
    over its T token rows (BLIP-2).  Differentiable w.r.t. ``query_features`` and ``tau`` (if it is a
    tensor requiring grad); the bank is a constant.

    ``single_pass``: ``True`` = fused sweep (statistics + gradient in one pass over the bank, static soft-max
    reference), ``False`` = forward-statistics kernel + recompute backward; default: fused iff a gradient is needed.
    """
    _require_cuda(query_features, "query_features")
    if query_features.dim() != 2:
        raise SpnError("bank_infonce", -1, f"query_features must be [B, D], got {tuple(query_features.shape)}")
    if not isinstance(bank, DeviceBank):
        _warn_raw_bank(bank)
        bank = DeviceBank.from_tensor(bank, query_features.device)
    if bank.D != query_features.shape[1]:
        raise SpnError("bank_infonce", -1, f"feature dims differ: q {query_features.shape[1]} vs bank {bank.D}")
    tau_t, tau_v = _split_tau(tau)
    if not (tau_v > 0.0 and math.isfinite(tau_v)):
        raise SpnError("bank_infonce", -7, f"tau must be positive and finite, got {tau_v}")
    lab = _prep_labels(labels, query_features.shape[0], query_features.device)
    needs_grad = torch.is_grad_enabled() and (query_features.requires_grad or (tau_t is not None and tau_t.requires_grad))
    if single_pass is None:
        # training: one fused sweep; inference: forward statistics only.  The fused sweep's static reference is
        # provably safe while inv_tau * |q| * max|bank| <= 109 nats (unit-norm q: tau >= 0.0092, every setting the
        # reference uses); colder temperatures take the two-pass path (exact running maximum).
        single_pass = needs_grad and (1.0 / tau_v) * bank.maxnorm_host() * 1.03 <= 109.0
    if label_smoothing != 0.0 and bank.T > 1:
        # the fused finalize forms sum_n s_n from the bank's column sums, which for a multi-token bank would sum every
        # token row instead of the per-entry max: the two-pass path accumulates it on the entry max
        single_pass = False
    if single_pass and _use_torch_ext() and bank.row_base == 0 and bank.n_total == bank.N:
        # the training hot path: one torch.library operator with a C++ autograd node (csrc_torch/spn_torch.cpp)
        eps = float(label_smoothing)
        maxnorm, colsum = bank.stats(eps != 0.0)
        return _torch_ext.ops().bank_infonce(query_features, tau_t, bank.data, lab, maxnorm, colsum if eps != 0.0 else None,
                                             1.0 / tau_v, eps, bank.T)
    fn = _BankCEFused if single_pass else _BankCE
    return fn.apply(query_features, tau_t, bank, lab, 1.0 / tau_v, float(label_smoothing))


def bank_ce_step(query_features: torch.Tensor, bank: DeviceBank, labels: Optional[torch.Tensor] = None, tau: float = 0.01,
                 *, label_smoothing: float = 0.0, grad_scale: float = 1.0, want_dtau: bool = False):
    """The training-step primitive without the autograd engine: ``(loss, dq)`` or ``(loss, dq, dtau)``.

    Same arithmetic as ``loss = bank_infonce(q, bank, labels, tau); (grad_scale * loss).backward()`` -- i.e. the forward of
    /root/reference/clip4cir/models_negplus.py:150-154 and its autograd backward
    (/root/reference/clip4cir/train_negplus.py:110-121) -- as ONE call into the C ABI (``spn_bank_ce_step``), for trainers
    that own their backward (and for ``bench.py``, where PyTorch's autograd thread hand-off would otherwise be the
    bottleneck: ~100-200 us of host time per step against ~130 us of GPU time).  ``bank`` must be an un-sharded
    :class:`DeviceBank`; ``tau`` a python float.
    """
    _require_cuda(query_features, "query_features")
    if not isinstance(bank, DeviceBank) or bank.row_base != 0 or bank.n_total != bank.N:
        raise SpnError("bank_ce_step", -7, "needs an un-sharded DeviceBank (use dist.bank_infonce_sharded for shards)")
    if query_features.dim() != 2 or query_features.shape[1] != bank.D:
        raise SpnError("bank_ce_step", -1, f"query_features must be [B, {bank.D}], got {tuple(query_features.shape)}")
    tau_v = float(tau)
    if not (tau_v > 0.0 and math.isfinite(tau_v)):
        raise SpnError("bank_ce_step", -7, f"tau must be positive and finite, got {tau_v}")
    inv_tau = 1.0 / tau_v
    B = query_features.shape[0]
    lab = _prep_labels(labels, B, query_features.device)
    eps = float(label_smoothing)
    # the fused sweep's static soft-max reference is provably safe; multi-token banks with label smoothing take the
    # two-pass path (see bank_infonce)
    if inv_tau * bank.maxnorm_host() * 1.03 <= 109.0 and not (eps != 0.0 and bank.T > 1):
        if _use_torch_ext():
            maxnorm, colsum = bank.stats(eps != 0.0)
            loss, dq, ddot = _torch_ext.ops().bank_ce_step(query_features, bank.data, lab, maxnorm,
                                                           colsum if eps != 0.0 else None, inv_tau, eps, bank.T, want_dtau)
        else:
            _, _, loss, dq, ddot = _fused_step(query_features, bank, lab, inv_tau, eps, want_dtau)
    else:  # colder temperatures: forward statistics (exact running maximum) + recompute backward
        q16 = cast_to_bf16(query_features)
        stats = _fwd_stats(q16, bank, lab, inv_tau, eps != 0.0)
        lse, _, loss = _loss_from_stats(stats, bank.n_total, eps)
        dq = torch.zeros((B, bank.D), dtype=torch.float32, device=query_features.device)
        ddot = torch.zeros(B, dtype=torch.float32, device=query_features.device) if want_dtau else None
        _bwd(q16, bank, lab, inv_tau, lse, inv_tau / B * grad_scale, None, eps, dq, ddot)
        if want_dtau:
            return loss, dq, (-(inv_tau / B) * grad_scale) * ddot.sum()
        return loss, dq
    if grad_scale != 1.0:
        dq = dq * grad_scale
    if want_dtau:
        return loss, dq, (-(inv_tau / B) * grad_scale) * ddot.sum()
    return loss, dq


class _PartCE(torch.autograd.Function):
    """mean_b CE(inv_tau * <q_b, bank[cand[b]]>, label 0): forward already holds the gradient for an upstream 1."""

    @staticmethod
    def forward(ctx, q, tau_t, bank: DeviceBank, cand, inv_tau: float):
        B, D = q.shape
        dev = q.device
        qc = q.detach()
        if qc.dtype not in _TORCH2SPN:
            qc = qc.float()
        qc = qc.contiguous()
        tau_grad = tau_t is not None and tau_t.requires_grad
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dq = torch.empty((B, D), dtype=torch.float32, device=dev)
        ddot = torch.empty(B, dtype=torch.float32, device=dev) if tau_grad else None
        with torch.cuda.device(dev):
            call("spn_part_ce", _ptr(qc), _TORCH2SPN[qc.dtype], _ptr(bank.data), _ptr(cand), inv_tau, B, cand.shape[1], D,
                 bank.data.shape[0], None, _ptr(loss), _ptr(dq), _ptr(ddot), _stream(dev))
        ctx.inv_tau, ctx.B, ctx.q_dtype, ctx.tau_grad = inv_tau, B, q.dtype, tau_grad
        ctx.save_for_backward(dq, ddot)
        return loss

    @staticmethod
    def backward(ctx, g):
        dq_unit, ddot = ctx.saved_tensors
        g32 = g.detach().to(torch.float32)
        dtau = (-(ctx.inv_tau / ctx.B)) * ddot.sum() * g32 if ctx.tau_grad else None
        return (dq_unit * g32).to(ctx.q_dtype), dtau, None, None, None


def part_infonce(query_features: torch.Tensor, bank: Union[DeviceBank, torch.Tensor], candidates: torch.Tensor,
                 tau=0.01) -> torch.Tensor:
    """Sampled-negatives loss: ``candidates`` int64 ``[B, 1 + k]`` bank rows per query, column 0 the positive
    (/root/reference/clip4cir/models.py:115-128 with the ids of ``get_neg_id`` :109-113).  Differentiable w.r.t. the
    queries and ``tau`` (if a tensor requiring grad)."""
    _require_cuda(query_features, "query_features")
    if not isinstance(bank, DeviceBank):
        bank = DeviceBank.from_tensor(bank, query_features.device)
    if bank.T != 1 or bank.row_base != 0:
        raise SpnError("part_infonce", -7, "needs an un-sharded single-token bank")
    if query_features.dim() != 2 or query_features.shape[1] != bank.D:
        raise SpnError("part_infonce", -1, f"query_features must be [B, {bank.D}], got {tuple(query_features.shape)}")
    cand = candidates.to(device=query_features.device, dtype=torch.int64).contiguous()
    if cand.dim() != 2 or cand.shape[0] != query_features.shape[0]:
        raise SpnError("part_infonce", -1, f"candidates must be [B, 1 + k], got {tuple(cand.shape)}")
    tau_t, tau_v = _split_tau(tau)
    if not (tau_v > 0.0 and math.isfinite(tau_v)):
        raise SpnError("part_infonce", -7, f"tau must be positive and finite, got {tau_v}")
    return _PartCE.apply(query_features, tau_t, bank, cand, 1.0 / tau_v)


def bank_ce_stats(query_features: torch.Tensor, bank: DeviceBank, labels: torch.Tensor, tau: float,
                  want_ssum: bool = False) -> torch.Tensor:
    """Forward statistics only: [4, B] = (m, l, s_t, s_sum) in natural-log units (no autograd)."""
    q_bf16 = cast_to_bf16(query_features)
    lab = _prep_labels(labels, q_bf16.shape[0], q_bf16.device)
    return _fwd_stats(q_bf16, bank, lab, 1.0 / float(tau), want_ssum)


def query_assemble(refer_bank: torch.Tensor, refer_idx: torch.Tensor, text_feats: torch.Tensor):
    """normalize(refer_bank[idx] + text) -> (q_bf16, q_f32, inv_norm)  (models_negplus.py:133-137)."""
    _require_cuda(refer_bank, "refer_bank")
    B, D = text_feats.shape
    dev = refer_bank.device
    rb = refer_bank.detach().float().contiguous()
    tx = text_feats.detach().float().contiguous()
    idx = refer_idx.to(device=dev, dtype=torch.int64).contiguous()
    q16 = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
    q32 = torch.empty((B, D), dtype=torch.float32, device=dev)
    inv = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("spn_query_assemble", _ptr(rb), _ptr(idx), _ptr(tx), B, D, rb.shape[0], _ptr(q16), _ptr(q32), _ptr(inv),
             _stream(dev))
    return q16, q32, inv
