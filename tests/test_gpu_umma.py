"""tcgen05 building-block self test (descriptors, swizzled operands, TMEM loads)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,N,K', [(128, 16, 64), (128, 48, 128), (256, 48, 256), (384, 96, 128),
                                   (128, 128, 64), (512, 32, 192)])
def test_umma_gemm_matches_fp16_matmul(M, N, K):
    from slotformer_b200 import engine
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    W = torch.randn((M, K), device='cuda', generator=g)
    X = torch.randn((N, K), device='cuda', generator=g)
    out = engine.umma_gemm(W, X)
    ref = X.half().double() @ W.half().double().t()
    err = (out.double() - ref).abs().max().item()
    assert err < 1e-3 * (K ** 0.5), (err, out[:2, :4], ref[:2, :4])
