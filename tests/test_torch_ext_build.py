"""The PyTorch extension (spn4cir_b200/csrc_torch/spn_torch.cpp) without a GPU: it builds in-tree, loads, registers the
torch.ops.spn4cir operators with the documented schemas, and fails loudly (no CPU fallback) on host tensors."""
import pytest
import torch

import spn4cir_b200 as spn
from spn4cir_b200 import _torch_ext


def test_extension_builds_loads_and_registers_ops():
    path = _torch_ext.build_torch_extension()
    assert path.exists() and path.parent == _torch_ext.PKG_DIR, "the extension must live in-tree"
    ops = _torch_ext.ops()
    for name in ("bank_ce_step", "bank_infonce", "bank_large_step_loss"):
        assert hasattr(ops, name)
    schema = str(torch.ops.spn4cir.bank_infonce.default._schema)
    assert "Tensor q" in schema and "Tensor? tau" in schema and "float inv_tau" in schema and "int T" in schema


def test_extension_has_no_cpu_fallback():
    ops = _torch_ext.ops()
    q, bank = torch.randn(4, 64), torch.randn(10, 64).to(torch.bfloat16)
    lab, mx = torch.zeros(4, dtype=torch.int64), torch.ones(1)
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.bank_infonce(q, None, bank, lab, mx, None, 50.0, 0.0, 1)
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        ops.bank_ce_step(q, bank, lab, mx, None, 50.0, 0.0, 1, False)


def test_extension_errors_with_numbers_raise_instead_of_crashing():
    """Regression: with $CXX pointing at a toolchain that only ships a static libstdc++ (this image's /opt/gcc) the extension
    carried a second copy of the iostream machinery and formatting a NUMBER into an error message crashed the process.  The
    build now prefers the system g++ and rejects such a link; a shape error (its message contains numbers) must raise."""
    assert not _torch_ext._embeds_libstdcxx(_torch_ext.build_torch_extension())
    ops = _torch_ext.ops()
    q, bank = torch.randn(4, 64), torch.randn(10, 32).to(torch.bfloat16)
    lab, mx = torch.zeros(4, dtype=torch.int64), torch.ones(1)
    with pytest.raises(RuntimeError, match="last sizes 64 / 32"):
        ops.bank_ce_step(q, bank, lab, mx, None, 50.0, 0.0, 1, False)
    with pytest.raises(RuntimeError, match="last sizes 64 / 32"):
        ops.bank_infonce(q, None, bank, lab, mx, None, 50.0, 0.0, 1)


def test_compat_infonce_keeps_gradient_of_in_batch_targets():
    """`infonce_loss` is also called by the reference with in-batch targets that carry gradient
    (clip4cir/models.py:160, models_negtype.py:84-107): rebinding the method must not drop that gradient."""
    from spn4cir_b200 import compat
    g = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(6, 32, generator=g)).requires_grad_(True)
    t = torch.nn.functional.normalize(torch.randn(6, 32, generator=g)).requires_grad_(True)
    loss = compat.infonce_loss(q, t, None, tau=0.05)
    loss.backward()
    ref = torch.nn.functional.cross_entropy((q.detach() @ t.detach().T) / 0.05, torch.arange(6))
    assert abs(loss.item() - ref.item()) < 1e-6 and t.grad is not None and t.grad.abs().sum() > 0
    # 1-D query (clip4cir/models.py:125 scores one query at a time)
    one = compat.infonce_loss(q[0].detach(), t, torch.tensor([0]), tau=0.05)
    assert one.dim() == 0
