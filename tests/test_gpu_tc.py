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
        pkg._lib.call("rs_tc_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split,
                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        err = np.abs(D.cpu().numpy() - want).max() / scale
        print("tcgen05 selftest N=%d K=%d split=%d: rel err %.2e" % (N, K, split, err))
        assert err < tol
