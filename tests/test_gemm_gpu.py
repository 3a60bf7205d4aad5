"""-m gpu: the tcgen05/TMEM/TMA GEMM behind the C ABI vs a plain torch fp32 reference of the same op
(floating-point kernel -> torch fp32 reference; inputs are fp16-representable so the only difference is
fp32 accumulation order: tolerance 2e-3 relative to the output scale)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def _gemm(A, B, bias, mode, out):
    import torch
    from dream2real_b200 import _native as N
    M, K = A.shape
    Nn = B.shape[0]
    N.check(N.lib().d2r_gemm_f16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, Nn, K,
                                 bias.data_ptr() if bias is not None else None, mode, out.data_ptr(), out.stride(0),
                                 N.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 64, 128), (256, 256, 768), (200, 2304, 768), (50 * 7 + 3, 768, 3072), (1, 512, 768),
                                   (38400 + 17, 256, 128), (25600, 768, 768)])   # the last two take the 256-wide tile path
def test_gemm_plain_f32_out(M, N, K):
    import torch
    torch.manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = _gemm(A, B, bias, 3, torch.empty(M, N, device="cuda"))
    ref = A.float() @ B.float().t() + bias
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_gemm_epilogues():
    import torch
    torch.manual_seed(0)
    M, N, K = 300, 384, 256
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t() + bias
    o16 = _gemm(A, B, bias, 0, torch.empty(M, N, device="cuda", dtype=torch.half))
    assert (o16.float() - ref).abs().max().item() < 2e-2
    og = _gemm(A, B, bias, 1, torch.empty(M, N, device="cuda", dtype=torch.half))
    refg = ref * torch.sigmoid(1.702 * ref)
    assert (og.float() - refg).abs().max().item() < 2e-2
    x0 = torch.randn(M, N, device="cuda")
    x = _gemm(A, B, bias, 2, x0.clone())
    assert (x - (x0 + ref)).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    nb = _gemm(A, B, None, 3, torch.empty(M, N, device="cuda"))
    assert (nb - (ref - bias)).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


def test_gemm_strided_operands_and_errors():
    import torch
    from dream2real_b200 import _native as N
    torch.manual_seed(1)
    big = (torch.randn(260, 1024, device="cuda") * 0.3).half()
    A = big[:, 128:128 + 512]            # lda = 1024, K = 512
    B = (torch.randn(128, 512, device="cuda") * 0.1).half()
    out = _gemm(A, B, None, 3, torch.empty(260, 128, device="cuda"))
    ref = A.float() @ B.float().t()
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    rc = N.lib().d2r_gemm_f16(A.data_ptr(), 1024, B.data_ptr(), 512, 260, 128, 500, None, 3, out.data_ptr(), 128, N.stream_ptr())
    assert rc != 0 and b"multiple of 64" in N.lib().d2r_last_error()
