"""Hardware bring-up of the tcgen05 encodings (see csrc/umma_selftest.cu): one-CTA GEMMs against torch fp32."""
import pytest
import torch

from spn4cir_b200 import _lib

pytestmark = pytest.mark.gpu


def run_gemm(a, b, N, K, mode):
    c = torch.empty((128, N), dtype=torch.float32, device="cuda")
    _lib.call("spn_debug_umma_gemm", a.data_ptr(), b.data_ptr(), c.data_ptr(), N, K, mode,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("N,K", [(128, 64), (128, 128), (64, 128), (128, 256), (64, 256)])
def test_umma_gemm(mode, N, K):
    if (mode & 2) and K != 128:
        pytest.skip("MN-major probe is defined for K == 128")
    g = torch.Generator().manual_seed(100 * mode + N + K)
    a = torch.randn((128, K), generator=g).to(torch.bfloat16).cuda()
    if mode & 2:
        b = torch.randn((K, N), generator=g).to(torch.bfloat16).cuda()   # MN-major: C = A @ B
        ref = a.float() @ b.float()
    else:
        b = torch.randn((N, K), generator=g).to(torch.bfloat16).cuda()   # K-major:  C = A @ B^T
        ref = a.float() @ b.float().T
    c = run_gemm(a, b, N, K, mode)
    err = (c - ref).abs().max().item()
    assert err < 2e-3 * K ** 0.5, f"mode={mode} N={N} K={K}: max abs err {err}"


@pytest.mark.parametrize("mode", [6, 7])
@pytest.mark.parametrize("N", [128, 256])
def test_umma_gemm_wide_mn_major(mode, N):
    """One MN-major instruction spanning several 64-column swizzle atoms (descriptor LBO = distance between atoms):
    the shape the second pass of the fused bank-CE kernel issues."""
    g = torch.Generator().manual_seed(7 * mode + N)
    a = torch.randn((128, 128), generator=g).to(torch.bfloat16).cuda()
    b = torch.randn((128, N), generator=g).to(torch.bfloat16).cuda()
    c = run_gemm(a, b, N, 128, mode)
    err = (c - a.float() @ b.float()).abs().max().item()
    assert err < 2e-3 * 128 ** 0.5, f"mode={mode} N={N}: max abs err {err}"
