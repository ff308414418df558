"""GPU: the tcgen05 / TMEM / shared-memory-descriptor plumbing in isolation
(rs_tc_selftest: one CTA, D[128,N] = A[128,K] B[N,K]^T), plain bf16 and the bf16x3
split the recurrent and projection kernels use to reach fp32-grade products."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(32, 64), (32, 256), (64, 128), (256, 128), (16, 256)])
def test_tcgen05_selftest_matches_fp64_matmul(pkg, cuda, N, K):
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    Ad, Bd = torch.from_numpy(A).to(cuda), torch.from_numpy(B).to(cuda)
    scale = np.abs(want).max()
    for split, tol in ((0, 2e-2), (1, 3e-5)):
        D = torch.full((128, N), float("nan"), dtype=torch.float32, device=cuda)
        pkg._lib.diag_call("rs_tc_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split,
                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        err = np.abs(D.cpu().numpy() - want).max() / scale
        print("tcgen05 selftest N=%d K=%d split=%d: rel err %.2e" % (N, K, split, err))
        assert err < tol


@pytest.mark.parametrize("M,N,K", [(300, 200, 120), (1024, 3072, 768), (31936, 80, 768), (128, 128, 64), (4000, 768, 3072)])
def test_tcgen05_gemm_against_fp64(pkg, cuda, M, N, K):
    """The persistent TMA + tcgen05 GEMM used for every batched projection: tails in M, N
    and K (TMA zero fill), bias, plain bf16 and bf16x3."""
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T + bias
    Ad, Bd, bd = (torch.from_numpy(a).to(cuda) for a in (A, B, bias))
    scratch = torch.empty(2 * (M * K + N * K) * 2 + 64, dtype=torch.uint8, device=cuda)
    for products, tol in ((1, 2e-2), (3, 2e-5)):
        C = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
        pkg._lib.diag_call("rs_gemm_tc_test", Ad.data_ptr(), Bd.data_ptr(), bd.data_ptr(), C.data_ptr(), M, N, K, products,
                      scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = C.cpu().numpy()
        err = np.abs(got - want).max() / np.abs(want).max()
        print("tcgen05 gemm %dx%dx%d products=%d: rel err %.2e" % (M, N, K, products, err))
        assert not np.isnan(got).any()
        assert err < tol


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (768, 3072, 4096), (120, 768, 1000), (768, 80, 520), (200, 328, 72)])
def test_tcgen05_gemm_mn_major_against_fp64(pkg, cuda, M, N, K):
    """The MN-major ("TN") form used by the weight-gradient GEMMs: C = A^T B with A stored [K][M] and B stored [K][N]
    (row index = the contraction index), operand tiles brought in as 64-column TMA boxes and described to
    tcgen05.mma as MN-major SWIZZLE_128B tiles.  Tails in M, N (boxes zero-filled) and K; bias; bf16 and bf16x3."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M)).astype(np.float32)
    B = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64) + bias
    Ad, Bd, bd = (torch.from_numpy(a).to(cuda) for a in (A, B, bias))
    scratch = torch.empty(2 * (M * K + N * K) * 2 + 64, dtype=torch.uint8, device=cuda)
    for products, tol in ((1, 2e-2), (3, 2e-5)):
        C = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
        pkg._lib.diag_call("rs_gemm_tc_test", Ad.data_ptr(), Bd.data_ptr(), bd.data_ptr(), C.data_ptr(), M, N, K, 16 + products,
                      scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = C.cpu().numpy()
        err = np.abs(got - want).max() / np.abs(want).max()
        print("tcgen05 MN-major gemm %dx%dx%d products=%d: rel err %.2e" % (M, N, K, products, err))
        assert not np.isnan(got).any()
        assert err < tol
