"""The tcgen05 GEMM building block vs torch.matmul on the same bf16 operands (fp32 accumulate)."""
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def run(M, N, K, bias=False, relu=False, k_split=0, lda=None, ldb=None):
    import _dgm_lib
    lib = _dgm_lib.lib()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    lda, ldb = lda or K, ldb or K
    A = torch.zeros(M, lda).bfloat16().cuda()
    B = torch.zeros(N, ldb).bfloat16().cuda()
    A[:, :K] = torch.randn(M, K, generator=g).bfloat16().cuda()
    B[:, :K] = torch.randn(N, K, generator=g).bfloat16().cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    ldc = (N + 3) // 4 * 4
    C = torch.zeros(M, ldc, device="cuda")
    rc = lib.dgl_gemm_bf16(M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, b.data_ptr() if bias else None,
                           int(relu), C.data_ptr(), ldc, k_split, _dgm_lib.stream_ptr())
    _dgm_lib.check(rc, "dgl_gemm_bf16")
    ref = A[:, :K].float() @ B[:, :K].float().t()
    if bias:
        ref = ref + b
    if relu:
        ref = ref.clamp_min(0)
    return C[:, :N], ref


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (128, 32, 64), (300, 256, 96), (1000, 16, 256), (257, 96, 352),
                                   (4096, 256, 256), (128, 256, 16), (100_000, 256, 256)])
def test_gemm_matches_torch(M, N, K):
    C, ref = run(M, N, K)
    assert util.rel_err(C, ref) < 2e-5


def test_gemm_epilogue_and_strides():
    C, ref = run(500, 256, 256, bias=True, relu=True, lda=264, ldb=272)
    assert util.rel_err(C, ref) < 2e-5
    C, ref = run(256, 352, 100_032, k_split=2048)     # dW-shaped: long K, split-K reduction
    assert util.rel_err(C, ref) < 1e-4
