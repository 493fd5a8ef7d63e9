"""tcgen05 GEMM (sei_gemm_bf16_tn) against a plain PyTorch fp32 reference of the same contraction.
Inputs are bf16 (exactly representable), products accumulate in fp32 in both; tolerance is the bf16
output rounding (2^-8 relative) for bf16 outputs and fp32 accumulation-order noise for fp32 outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


SHAPES = [(128, 32, 64), (256, 128, 128), (1000, 96, 200), (4096, 256, 512), (130, 40, 72), (8192, 512, 2048),
          (257, 300, 1032), (128, 8, 8), (64, 256, 64)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_matches_fp32_reference(dev, M, N, K, out_dtype):
    from sei_b200 import ops, last_kernel
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = a.float() @ b.float().t() + bias
    for tile_n in (0, 32, 64, 128, 256):
        d = ops.gemm_bf16_tn(a, b, bias, out_dtype=out_dtype, tile_n=tile_n)
        assert last_kernel() in ("gemm_bf16_tn_kernel", "gemm_bf16_tn_2cta_kernel")
        assert d.dtype == out_dtype and d.shape == (M, N)
        err = (d.float() - ref).abs().max() / ref.abs().max()
        tol = 2e-5 if out_dtype == torch.float32 else 6e-3
        assert float(err) < tol, (tile_n, float(err))
    d0 = ops.gemm_bf16_tn(a, b, None, out_dtype=torch.float32)
    assert float((d0 - (ref - bias)).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (300, 264, 72), (8192, 2048, 1024), (4100, 8192, 520),
                                   (257, 256, 8)])
def test_gemm_cta_pair_kernel(dev, M, N, K):
    """large bf16-output products run on the cta_group::2 kernel (two CTAs share a 256 x 256 tile); ragged edges are
    clipped by the tensor maps"""
    from sei_b200 import ops, last_kernel
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = a.float() @ b.float().t() + bias
    d = ops.gemm_bf16_tn(a, b, bias, out_dtype=torch.bfloat16)
    assert last_kernel() == "gemm_bf16_tn_2cta_kernel"
    assert float((d.float() - ref).abs().max() / ref.abs().max()) < 6e-3
    d2 = ops.gemm_bf16_tn(a, b, bias, out_dtype=torch.bfloat16)          # persistent barriers / TMEM reused correctly
    assert torch.equal(d, d2)


def test_gemm_strided_operands(dev):
    from sei_b200 import ops
    torch.manual_seed(1)
    big_a = torch.randn(300, 264, device=dev).bfloat16()
    big_b = torch.randn(70, 264, device=dev).bfloat16()
    a, b = big_a[:, :200], big_b[:, :200]          # row pitch 264 (multiple of 8), K = 200
    ref = a.float() @ b.float().t()
    d = ops.gemm_bf16_tn(a, b, None, out_dtype=torch.float32)
    assert float((d - ref).abs().max() / ref.abs().max()) < 2e-5
    from sei_b200 import SeiError
    with pytest.raises(SeiError, match="multiples of 8"):
        ops.gemm_bf16_tn(torch.randn(8, 12, device=dev).bfloat16(), torch.randn(8, 12, device=dev).bfloat16())


def test_gemm_split_k_weight_gradient_shapes(dev):
    """few output tiles, very long K (K = pixels): the fp32-output path splits K across CTAs (atomic accumulation)"""
    from sei_b200 import ops
    torch.manual_seed(2)
    for M, N, K in [(128, 32, 65536), (32, 128, 131072), (512, 128, 16384), (3, 288, 20000)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        ref = a.float() @ b.float().t()
        d = ops.gemm_bf16_tn(a, b, None, out_dtype=torch.float32)
        assert float((d - ref).abs().max() / ref.abs().max()) < 5e-5, (M, N, K)
        bias = torch.randn(N, device=dev)
        d = ops.gemm_bf16_tn(a, b, bias, out_dtype=torch.float32)
        assert float((d - ref - bias).abs().max() / ref.abs().max()) < 5e-5, (M, N, K)


@pytest.mark.parametrize("K,M,N", [(4096, 128, 64), (65536, 128, 32), (10000, 512, 128), (2048, 2048, 512), (777, 136, 72),
                                   (8192, 256, 1024), (5000, 384, 264), (64, 256, 256), (100000, 1024, 256), (8192, 8192, 2048),
                                   (3000, 136, 512)])
def test_gemm_atb_mn_major(dev, K, M, N):
    """D = A^T B with both operands read in place (MN-major UMMA descriptors): the weight-gradient shape"""
    from sei_b200 import ops, last_kernel
    torch.manual_seed(K + M + N)
    a = torch.randn(K, M, device=dev).bfloat16()
    b = torch.randn(K, N, device=dev).bfloat16()
    ref = a.float().t() @ b.float()
    d = ops.gemm_bf16_atb(a, b)
    # wide weight matrices take the CTA-pair kernel (cta_group::2, 256 x 256 tiles per cluster)
    assert last_kernel() == ("gemm_bf16_mn_2cta_kernel" if M >= 256 and N >= 256 else "gemm_bf16_mn_kernel")
    assert float((d - ref).abs().max() / ref.abs().max()) < 5e-5, float((d - ref).abs().max() / ref.abs().max())
    d0 = torch.randn(M, N, device=dev)
    acc = ops.gemm_bf16_atb(a, b, out=d0.clone())                     # in-place accumulation (weight gradients)
    assert float((acc - d0 - ref).abs().max() / ref.abs().max()) < 5e-5
    d2 = ops.gemm_bf16_atb(a, b)                                      # barriers / TMEM reused correctly by a second launch
    assert float((d2 - ref).abs().max() / ref.abs().max()) < 5e-5


def test_gemm_throughput_smoke(dev):
    """not a benchmark: just exercises a deep-K, many-tile launch (scale-4 ConvBlock shape) for hangs"""
    from sei_b200 import ops
    a = torch.randn(8192, 8192, device=dev).bfloat16()
    b = torch.randn(4096, 8192, device=dev).bfloat16()
    d = ops.gemm_bf16_tn(a, b, None)
    ref = (a[:64].float() @ b.float().t())
    assert float((d[:64].float() - ref).abs().max() / ref.abs().max()) < 6e-3


@pytest.mark.parametrize("M,N,K", [(512, 256, 64), (300, 128, 32), (8192, 2048, 512), (130, 64, 16), (2048, 512, 128)])
def test_gemm_with_gelu_backward_epilogue(dev, M, N, K):
    """(A B^T) * gelu'(H) in the epilogue (single-CTA and CTA-pair kernels) vs the unfused fp32 formulation"""
    from sei_b200 import ops
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    h = (torch.randn(M, N, device=dev) * 2).bfloat16()
    d = ops.gemm_bf16_tn_gelu_bwd(a, b, h)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).sum().backward()
    ref = (a.float() @ b.float().t()) * hf.grad
    assert float((d.float() - ref).abs().max() / ref.abs().max()) < 8e-3
