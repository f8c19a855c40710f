"""Pins the tcgen05 operand layout / descriptor encodings on real hardware (test hook of the C ABI)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (16, 256), (32, 32), (64, 256), (48, 128), (256, 16), (128, 64)])
@pytest.mark.parametrize("nsplit", [1, 2])
def test_umma_selftest(K, N, nsplit):
    from mobody_b200 import _ffi
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    D = torch.full((128, N), float("nan"), device="cuda")
    Ad, Bd = A.cuda(), B.cuda()                       # keep the device copies alive across the launch
    _ffi.check(_ffi.lib().mobody_selftest_umma(_ffi.ptr(Ad), _ffi.ptr(Bd), K, N, nsplit, _ffi.ptr(D), _ffi.stream_ptr()))
    torch.cuda.synchronize()
    want = (A.double() @ B.double().T).numpy()
    got = D.cpu().numpy()
    scale = np.abs(A.numpy()) @ np.abs(B.numpy()).T          # error bound scale per element
    err = np.max(np.abs(got - want) / scale)
    assert np.isfinite(got).all()
    assert err < (2e-2 if nsplit == 1 else 1e-4), err        # bf16: 2^-8 per product; hi+lo split: ~2^-16
    if nsplit == 1:   # exact check against the bf16-rounded product accumulated in fp64
        Ab, Bb = A.bfloat16().double(), B.bfloat16().double()
        assert np.max(np.abs(got - (Ab @ Bb.T).numpy()) / scale) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 256, 64), (128, 16, 256), (77, 35, 100), (1000, 256, 40), (129, 8, 8), (256, 200, 1000)])
@pytest.mark.parametrize("a_src,b_src", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_tf32x3_gemm_tiles(M, N, K, a_src, b_src):
    """The train step's tcgen05 GEMM tile (kind::tf32, hi+lo split, 3 MMAs per K step) for every operand-source combination,
    ragged M / N / K, aligned and unaligned leading dimensions: fp32-class accuracy against an fp64 product."""
    from mobody_b200 import _ffi
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(K, N, generator=g)
    As = (A if a_src == 0 else A.t()).contiguous().cuda()          # [M][K] or [K][M]
    Bs = (B.t() if b_src == 0 else B).contiguous().cuda()          # [N][K] or [K][N]
    lda, ldb = As.shape[1], Bs.shape[1]
    C = torch.full((M, N), float("nan"), device="cuda")
    _ffi.check(_ffi.lib().mobody_selftest_gemm(_ffi.ptr(As), _ffi.ptr(Bs), M, N, K, a_src, b_src, lda, ldb, _ffi.ptr(C), _ffi.stream_ptr()))
    torch.cuda.synchronize()
    got = C.cpu().numpy()
    want = (A.double() @ B.double()).numpy()
    scale = np.abs(A.numpy()) @ np.abs(B.numpy())
    assert np.isfinite(got).all()
    err = np.max(np.abs(got - want) / scale)
    assert err < 5e-6, err                                            # 3xTF32 ~ 2^-21 per product (single-pass TF32 would be ~5e-4)
