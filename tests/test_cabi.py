"""C-ABI checks that need no GPU: the library builds/loads, exports every symbol include/spn_b200.h declares,
answers shape queries, and FAILS LOUDLY (never falls back) when asked to compute without a B200."""
import ctypes
import re

import pytest
import torch

import spn4cir_b200 as spn
from spn4cir_b200 import _lib


def declared_symbols():
    text = _lib.HEADER_PATH.read_text() + _lib.DEBUG_HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = spn.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spn_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes signature table out of sync with the headers"
    product = re.sub(r"/\*.*?\*/", "", _lib.HEADER_PATH.read_text(), flags=re.S)
    assert "spn_debug_" not in product, "bring-up entry points belong in spn_b200_debug.h"


def test_version_and_error_string():
    lib = spn.load_library()
    assert lib.spn_version() == 2
    assert isinstance(lib.spn_last_error(), bytes)


def test_workspace_query_and_shape_errors():
    # answers without a device (plans against 148 SMs / 227 KiB when no GPU is visible)
    n = _lib.workspace_bytes(_lib.OP_BANK_CE_FWD, 512, 100000, 640, 1)
    assert n > 0 and n % 256 == 0
    assert _lib.workspace_bytes(_lib.OP_SIM_TOPK, 4181, 2297, 640, 1, 50) % 256 == 0
    with pytest.raises(spn.SpnError) as e:
        _lib.workspace_bytes(_lib.OP_BANK_CE_FWD, 512, 100000, 100, 1)  # D not a multiple of 64
    assert e.value.code == -1 and "multiple of 64" in str(e.value)
    with pytest.raises(spn.SpnError):
        _lib.workspace_bytes(_lib.OP_BANK_CE_FWD, 512, 100000, 640, 7)  # T unsupported
    with pytest.raises(spn.SpnError):
        _lib.workspace_bytes(99, 1, 1, 64, 1)


def test_null_pointer_is_rejected_before_any_launch():
    with pytest.raises(spn.SpnError) as e:
        _lib.call("spn_bank_ce_fwd", None, None, None, 50.0, 1, 1, 64, 1, 0, None, None, None, None, None, 0, None)
    assert e.value.code == -7


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_without_gpu_fails_loudly():
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    with pytest.raises(spn.SpnError) as e:
        _lib.call("spn_merge_stats", p, p, p, None, 1, 4, p, p, p, None, None)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    with pytest.raises(spn.SpnError):
        spn.bank_infonce(torch.randn(4, 64), torch.randn(10, 64))
    with pytest.raises(spn.SpnError):
        spn.sim_topk(torch.randn(4, 64), torch.randn(10, 64), 3)


def test_product_package_never_imports_oracle():
    import pathlib
    for py in pathlib.Path(spn.__file__).parent.glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{py.name} routes through the oracle"


def test_sim_tc_workspace_layout_and_policy():
    """Host side of the tensor-core retrieval path: workspace sizing, the S~ block layout, the auto-mode policy."""
    Q, G, D = 4181, 2297, 640
    n = _lib.workspace_bytes(_lib.OP_SIM_TC, Q, G, D, 1)
    off, pitch, rows = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.call("spn_debug_sim_tc_layout", Q, G, D, 1, ctypes.byref(off), ctypes.byref(pitch), ctypes.byref(rows))
    assert n % 256 == 0 and off.value % 256 == 0
    assert pitch.value == 2300 and pitch.value % 4 == 0           # 16-byte row pitch for the TMA store
    assert rows.value % 128 == 0 and rows.value >= Q              # one block at CIRR size
    split = 2 * Q * D * 2 + 2 * G * D * 2                          # bf16 hi + lo of both operands
    assert n >= split + rows.value * pitch.value * 4 and off.value >= split
    # pair-mining size: S~ is produced in query blocks of <= 512 MiB
    _lib.call("spn_debug_sim_tc_layout", 47000, 47000, 768, 1, ctypes.byref(off), ctypes.byref(pitch), ctypes.byref(rows))
    assert rows.value % 128 == 0 and rows.value * pitch.value * 4 <= 512 << 20 and rows.value < 47000
    with pytest.raises(spn.SpnError):
        _lib.workspace_bytes(_lib.OP_SIM_TC, 10, 10, 100, 1)     # D % 64 != 0
    from spn4cir_b200.retrieval import _use_tc
    assert _use_tc("auto", 4181, 2297, 640, 1, 50) and not _use_tc("auto", 50, 500, 128, 1, 10)   # too small: launch-bound
    assert not _use_tc("auto", 4181, 2297, 640, 1, 64) and not _use_tc("auto", 4181, 2297, 600, 1, 50)
    assert not _use_tc("fp32", 4181, 2297, 640, 1, 50)
    with pytest.raises(spn.SpnError):
        _use_tc("tc", 4181, 2297, 640, 1, 57)
    with pytest.raises(spn.SpnError):
        _use_tc("bf16", 1, 1, 64, 1)


def test_split_error_bound_used_by_the_row_kernels():
    """csrc/sim_tc.cu prunes with S~ = qh.gh + qh.gl + ql.gh (bf16 hi / lo parts, fp32 accumulation) and treats entries
    closer than eps = 4e-5 |q| max|g| to a decision boundary as undecided.  Restate the split on the CPU (numpy, fp64
    accumulation: isolates the dropped terms) and check the bound that analysis gives: 3 * 2^-18 |q| |g|."""
    import numpy as np
    g = torch.Generator().manual_seed(5)
    q = torch.randn(64, 640, generator=g) * torch.logspace(-2, 2, 64)[:, None]     # norms over four decades
    x = torch.randn(300, 640, generator=g)
    x[::7] *= 50.0

    def split(t):
        hi = t.to(torch.bfloat16).to(torch.float32)
        lo = (t - hi).to(torch.bfloat16).to(torch.float32)
        assert float(((t - hi - lo).abs() / t.abs().clamp_min(1e-30)).max()) <= 2.0 ** -17   # x = hi + lo + O(2^-18 |x|)
        return hi.double().numpy(), lo.double().numpy()

    qh, ql = split(q)
    xh, xl = split(x)
    approx = qh @ xh.T + qh @ xl.T + ql @ xh.T
    exact = q.double().numpy() @ x.double().numpy().T
    scale = np.linalg.norm(q.double().numpy(), axis=1)[:, None] * np.linalg.norm(x.double().numpy(), axis=1).max()
    rel = np.abs(approx - exact) / scale
    assert rel.max() <= 3 * 2.0 ** -18, rel.max()
    assert rel.max() <= 4e-5 / 3        # head-room left for the tensor core's fp32 accumulation (measured on the GPU test)
