"""C-ABI checks that need no GPU: the library builds/loads, exports every symbol include/spn_b200.h declares,
answers shape queries, and FAILS LOUDLY (never falls back) when asked to compute without a B200."""
import ctypes
import re

import pytest
import torch

import spn4cir_b200 as spn
from spn4cir_b200 import _lib


def declared_symbols():
    text = _lib.HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = spn.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spn_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes signature table out of sync with the header"


def test_version_and_error_string():
    lib = spn.load_library()
    assert lib.spn_version() == 1
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
