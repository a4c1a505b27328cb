"""tcgen05 / TMEM building blocks (csrc/scnet_tc.cu): bf16 x bf16 -> fp32 GEMM against torch on bf16-rounded inputs.
The accumulation is fp32 in both, so only summation order differs: tolerance 1e-4 relative to ||row||."""
import pytest

from relativepose_b200.scnet_engine import h16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (128, 128, 64, 128), (256, 128, 256, 64), (384, 256, 1024, 128)])
def test_tc_gemm_matches_torch(M, N, K, bn):
    import torch
    from relativepose_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.rp_tc_gemm_test(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, bn, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = A.to(h16()).float() @ B.to(h16()).float().t()
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    print("M=%d N=%d K=%d bn=%d  max err %.3e (|ref|max %.2f)" % (M, N, K, bn, err, scale))
    assert err <= 1e-4 * scale + 1e-3
