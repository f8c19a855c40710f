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
