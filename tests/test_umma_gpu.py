"""Hardware check of the hand-written tcgen05 building blocks (csrc/umma.cuh) used by the tensor-core decoder:
bf16x3 (hi*hi + hi*lo + lo*hi) GEMMs with fp32 TMEM accumulation against a float64 torch matmul.
Tolerance: 1e-4 of the output scale (bf16x3 carries ~2^-16 relative error per product)."""
import pytest
import torch

from remixfusion_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N,mode", [(16, 16, 0), (32, 32, 0), (96, 32, 0), (96, 64, 0), (80, 16, 0), (128, 64, 0),
                                       (96, 32, 1), (32, 16, 1), (80, 64, 1), (128, 64, 1),
                                       (16, 64, 2), (16, 32, 2), (32, 16, 2), (64, 32, 2), (64, 16, 2), (32, 32, 2),
                                       (16, 16, 3), (32, 32, 3), (64, 16, 3), (112, 32, 3), (112, 64, 3)])
def test_umma_bf16x3_gemm(rf_lib, cuda, K, N, mode):
    torch.manual_seed(K * 100 + N + mode)
    A = torch.randn(128, K, device=cuda)
    B = torch.randn(*{0: (N, K), 1: (128, N), 2: (K, N), 3: (N, K)}[mode], device=cuda)
    D = torch.full((128, N), float("nan"), device=cuda)
    abi.check(rf_lib.rf_umma_selftest(abi.dptr(A), abi.dptr(B), abi.dptr(D), K, N, mode, abi.stream_ptr()), "rf_umma_selftest")
    if mode in (0, 3):
        ref, got = A.double() @ B.double().t(), D.double()
    elif mode == 2:
        ref, got = A.double() @ B.double(), D.double()
    else:
        ref, got = A.double().t() @ B.double(), D.double()[:K]
    err = (got - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item()
