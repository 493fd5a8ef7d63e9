"""The hand-written kernels of the restoration CNN beyond the GEMMs (csrc/bgemm.cu, gemm.cu:bgemm_tc_kernel, cnn_elem.cu).

Ideal (Fourier-domain) resamplers as explicit operators (models/resample.py) against fixtures produced by running the
reference's IdealUpsample / IdealDownsample (tests/golden/make_golden.py gen_resample, float64): outputs and the
autograd vector-Jacobian products.
  CPU:  the four operator matrices reproduce the reference to 1e-12 (forward and transposed).
  GPU:  the batched tensor-core products (tcgen05 and mma.sync kernels) against fp32 matmul of the same bf16 operands
        (accumulation order only: 8e-3 covers the bf16 rounding of the result), and the resamplers on bf16
        channels-last activations against the float64 fixtures within the bf16 tolerance 2e-2 (max |a-b| / max |b|).
Channel LayerNorm, column sums, depthwise 7x7, the 3x3 output convolution and GELU (GPU): against the PyTorch fp32
formulation of the same op on the same bf16 inputs -- forward / input gradients within bf16 rounding of the result
(4e-3 .. 6e-3), parameter gradients 2e-3 (fp32 sums)."""
import numpy as np
import pytest
import torch

from util import rel_err

N_CASES = 6


@pytest.mark.parametrize("kind", ["down", "up"])
def test_resample_operators_match_reference(golden, kind):
    from models import resample
    g = golden("resample")
    for i in range(N_CASES):
        x, y = torch.from_numpy(g[f"{kind}{i}_x"]), g[f"{kind}{i}_y"]
        got = resample.apply_dense(kind, x, 2)
        assert rel_err(got.numpy(), y) < 1e-12, (kind, i)
        Gr, Gi, P, Q = resample.operator(kind, x.shape[-2], x.shape[-1], 2)
        gy = torch.from_numpy(g[f"{kind}{i}_gy"])
        gx = Gr.t() @ gy @ P + Gi.t() @ gy @ Q
        assert rel_err(gx.numpy(), g[f"{kind}{i}_gx"]) < 1e-12, (kind, i)


def _emulated_bgemm(A, x, out, M, K, N, tile_rows, batches, b_inner, x_b, k_inner, x_k, d_b, m_inner, d_m):
    """sei_bgemm_bf16's addressing (include/sei_b200.h) restated with as_strided views, for CPU tensors of any dtype:
    entry b = bo * b_inner + bi, row k = ko * k_inner + ki of X, row m = mo * m_inner + mi of D, columns contiguous"""
    bo = batches // b_inner
    xs = torch.as_strided(x, (bo, b_inner, K // k_inner, k_inner, N), (x_b[0], x_b[1], x_k[0], x_k[1], 1)).reshape(batches, K, N)
    res = torch.matmul(A[:M, :K].to(x.dtype), xs)
    torch.as_strided(out, (bo, b_inner, M // m_inner, m_inner, N), (d_b[0], d_b[1], d_m[0], d_m[1], 1)).copy_(
        res.reshape(bo, b_inner, M // m_inner, m_inner, N))
    return out


@pytest.mark.parametrize("kind,B,H,W,C", [("down", 8, 32, 32, 8), ("up", 8, 16, 16, 16), ("down", 4, 64, 64, 8), ("up", 6, 8, 8, 8),
                                          ("down", 3, 16, 32, 8), ("up", 2, 64, 64, 8)])
def test_resampler_products_layout_and_batching(monkeypatch, kind, B, H, W, C):
    """models/resample.py's four batched products (contiguous two-term intermediate, interleaved columns of the height
    operator, block-diagonal batching of small operators) reproduce the dense operator: run on the CPU in float64 with the
    addressing of sei_bgemm_bf16 emulated by strided views -- forward and transposed, 1e-12"""
    from models import resample
    from sei_b200 import ops
    monkeypatch.setattr(ops, "bgemm_bf16", _emulated_bgemm)
    monkeypatch.setattr(ops, "bgemm_tile_rows", lambda M, Kpad: 128)      # the library asks the device; only pads A's rows
    torch.manual_seed(B * H + W)
    pk = resample._packed(kind, H, W, 2, "cpu", dtype=torch.float64)
    x = torch.randn(B, H, W, C, dtype=torch.float64)
    y = resample._forward_cl(x, pk)
    ref = resample.apply_dense(kind, x.permute(0, 3, 1, 2), 2)
    assert rel_err(y.permute(0, 3, 1, 2).numpy(), ref.numpy()) < 1e-12
    g = torch.randn_like(y)
    gx = resample._backward_cl(g, pk, H, W)
    Gr, Gi, P, Q = resample.operator(kind, H, W, 2)
    gl = g.permute(0, 3, 1, 2)
    gref = Gr.t() @ gl @ P + Gi.t() @ gl @ Q
    assert rel_err(gx.permute(0, 3, 1, 2).numpy(), gref.numpy()) < 1e-12
    # the batching factors: rows x P <= 128, columns x P <= 512, P | batches, powers of two only
    for name, batches in (("A1", B * H), ("A2", B), ("A2T", B), ("A1T", B * H)):
        prod = pk[name]
        M, K = prod.A.shape
        P_ = prod._pack_factor(batches)
        assert batches % P_ == 0 and (P_ == 1 or (P_ * M <= 128 and P_ * K <= 512 and P_ <= 8))
        if M >= 128 or M & (M - 1) or K & (K - 1):
            assert P_ == 1
        elif batches % 2 == 0 and 2 * M <= 128 and 2 * K <= 512:
            assert P_ >= 2
    assert resample._Product(torch.zeros(32, 32, dtype=torch.float64), "cpu")._pack_factor(32) == 4
    assert resample._Product(torch.zeros(16, 64, dtype=torch.float64), "cpu")._pack_factor(1024) == 8
    assert resample._Product(torch.zeros(24, 32, dtype=torch.float64), "cpu")._pack_factor(32) == 1


def test_torch_formulation_of_the_resamplers_matches_reference(golden):
    """tests/torch_formulation.py (the library formulation the CPU structure tests install) against the reference's
    IdealUpsample / IdealDownsample fixtures; the package itself has no library path: a CPU tensor raises"""
    import models.convolutional as mc
    import sei_b200
    import torch_formulation
    g = golden("resample")
    for i in range(N_CASES):
        for kind in ("down", "up"):
            x = torch.from_numpy(g[f"{kind}{i}_x"]).float()
            assert rel_err(torch_formulation.ideal_resample(x, kind, 2).numpy(), g[f"{kind}{i}_y"]) < 2e-5, (kind, i)
    with pytest.raises(sei_b200.SeiError):
        mc.IdealDownsample(2)(torch.rand(1, 8, 16, 16))
    with pytest.raises(sei_b200.SeiError):
        mc.LayerNorm(8)(torch.rand(1, 8, 4, 4))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,N,batches", [(256, 256, 128, 37), (128, 512, 200, 3), (512, 128, 64, 9), (48, 24, 8, 50),
                                           (16, 32, 1032, 2), (100, 70, 72, 5), (32, 64, 64, 300)])
def test_bgemm_matches_matmul(dev, M, K, N, batches):
    from sei_b200 import ops, launch_count
    torch.manual_seed(M + K + N)
    A = torch.randn(M, K, device=dev) / K ** 0.5
    x = torch.randn(batches, K, N, device=dev).bfloat16()
    kpad = -(-K // 64) * 64
    tile = ops.bgemm_tile_rows(M, kpad)
    assert tile in (16, 32, 64, 128, 256)
    Ap = torch.zeros(-(-M // tile) * tile, kpad, device=dev, dtype=torch.bfloat16)
    Ap[:M, :K] = A.bfloat16()
    out = torch.full((batches, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    n0 = launch_count()
    ops.bgemm_bf16(Ap, x, out, M, K, N, tile, batches, 1, (K * N, 0), K, (0, N), (M * N, 0), M, (0, N))
    assert launch_count() == n0 + 1
    from sei_b200 import last_kernel
    assert last_kernel() in ("bgemm_tc_kernel", "bgemm_kernel")
    ref = Ap[:M, :K].float() @ x.float()
    assert rel_err(out.float().cpu().numpy(), ref.cpu().numpy()) < 8e-3


@pytest.mark.gpu
def test_bgemm_split_rows(dev):
    """split batch / row addressing: the two-term intermediate [B, 2, H, W', C] written and read in place"""
    from sei_b200 import ops
    torch.manual_seed(5)
    B, H, W, Wo, C = 3, 20, 24, 12, 16
    A1 = torch.randn(2 * Wo, W, device=dev) / W ** 0.5
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    kpad = 64
    tile = ops.bgemm_tile_rows(2 * Wo, kpad)
    Ap = torch.zeros(-(-2 * Wo // tile) * tile, kpad, device=dev, dtype=torch.bfloat16)
    Ap[:2 * Wo, :W] = A1.bfloat16()
    y = torch.zeros(B, 2, H, Wo, C, device=dev, dtype=torch.bfloat16)
    ops.bgemm_bf16(Ap, x, y, 2 * Wo, W, C, tile, B * H, H, (H * W * C, W * C), W, (0, C),
                   (2 * H * Wo * C, Wo * C), Wo, (H * Wo * C, C))
    ref = torch.einsum("mw,bhwc->bhmc", Ap[:2 * Wo, :W].float(), x.float())        # m = (t, w')
    ref = ref.view(B, H, 2, Wo, C).permute(0, 2, 1, 3, 4)
    assert rel_err(y.float().cpu().numpy(), ref.cpu().numpy()) < 8e-3
    # read it back with split K rows: gx[b, h, w, c] = sum_{t, w'} A1[(t, w'), w] y[b, t, h, w', c]
    A1T = A1.t().contiguous()
    kpad2 = 64
    tile2 = ops.bgemm_tile_rows(W, kpad2)
    Ap2 = torch.zeros(-(-W // tile2) * tile2, kpad2, device=dev, dtype=torch.bfloat16)
    Ap2[:W, :2 * Wo] = A1T.bfloat16()
    gx = torch.zeros(B, H, W, C, device=dev, dtype=torch.bfloat16)
    ops.bgemm_bf16(Ap2, y, gx, W, 2 * Wo, C, tile2, B * H, H, (2 * H * Wo * C, Wo * C), Wo, (H * Wo * C, C),
                   (H * W * C, W * C), W, (0, C))
    ref2 = torch.einsum("wm,bhmc->bhwc", Ap2[:W, :2 * Wo].float(), y.float().permute(0, 2, 1, 3, 4).reshape(B, H, 2 * Wo, C))
    assert rel_err(gx.float().cpu().numpy(), ref2.cpu().numpy()) < 8e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C,kind", [(3, 16, 32, 16, "down"), (2, 64, 64, 8, "down"), (2, 32, 16, 64, "up"),
                                          (1, 128, 128, 32, "down"), (1, 64, 64, 128, "up"), (2, 256, 256, 8, "down"),
                                          (1, 512, 512, 8, "down"), (1, 256, 256, 8, "up"), (1, 1024, 1024, 8, "down"),
                                          # small operators: 2 / 4 / 8 batch entries per UMMA through a block-diagonal operator
                                          (32, 16, 16, 64, "up"), (8, 32, 32, 128, "down"), (4, 16, 16, 256, "up"),
                                          (32, 8, 8, 64, "down"), (6, 32, 32, 32, "down"), (3, 16, 16, 64, "up")])
def test_ideal_resample_tcgen05_path(dev, B, H, W, C, kind):
    """power-of-two shapes take the tcgen05 kernel (5-D tensor maps for the split rows): forward and transposed
    operator against the dense fp32 formulation of the same operator (models/resample.apply_dense)"""
    from models import resample
    from sei_b200 import last_kernel
    torch.manual_seed(H + W + C)
    x = torch.randn(B, C, H, W, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = resample.ideal_resample(x, kind, 2)
    assert last_kernel() == "bgemm_tc_kernel"
    ref_in = x.detach().float().requires_grad_(True)
    ref = resample.apply_dense(kind, ref_in, 2)
    assert rel_err(y.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) < 2e-2
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    assert last_kernel() == "bgemm_tc_kernel"
    (gref,) = torch.autograd.grad(ref, ref_in, gy.float())
    assert rel_err(gx.float().cpu().numpy(), gref.cpu().numpy()) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["down", "up"])
def test_ideal_resample_gpu_matches_reference(golden, dev, kind):
    from models import resample
    from sei_b200 import launch_count
    g = golden("resample")
    for i in range(N_CASES):
        x64, y64 = g[f"{kind}{i}_x"], g[f"{kind}{i}_y"]
        gy64, gx64 = g[f"{kind}{i}_gy"], g[f"{kind}{i}_gx"]
        nc = x64.shape[1]
        idx = [c % nc for c in range(8)]                       # 8 channels (the op needs C % 8 == 0): fixture channels repeated
        x = torch.from_numpy(x64[:, idx]).to(dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        assert resample.supported(x)
        n0 = launch_count()
        y = resample.ideal_resample(x, kind, 2)
        assert launch_count() == n0 + 2
        assert rel_err(y.detach().float().cpu().numpy(), y64[:, idx]) < 2e-2, (kind, i)
        gy = torch.from_numpy(gy64[:, idx]).to(dev).bfloat16()
        (gx,) = torch.autograd.grad(y, x, gy)
        assert launch_count() == n0 + 4
        assert rel_err(gx.float().cpu().numpy(), gx64[:, idx]) < 2e-2, (kind, i)


@pytest.mark.gpu
@pytest.mark.parametrize("T,C", [(1000, 32), (4096 + 3, 128), (777, 512), (300, 2048), (65, 8192), (50, 8), (129, 64), (40, 256),
                                 (5001, 3), (700, 12), (33, 1), (100, 31)])
def test_layer_norm_cl_matches_torch(dev, T, C):
    """hand-written channel LayerNorm (csrc/cnn_elem.cu) vs torch.nn.functional.layer_norm in fp32 on the same bf16
    input: forward within bf16 rounding of the result (6e-3), dx likewise, dgamma / dbeta 2e-3 (fp32 sums)."""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(T + C)
    x = (torch.randn(T, C, device=dev) * 1.7 + 0.3).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, device=dev)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, device=dev)).requires_grad_(True)
    gy = torch.randn(T, C, device=dev).bfloat16()
    assert ops.ln_any_supported(x)
    xa = x.clone().requires_grad_(True)
    y = ops.layer_norm_cl(xa, gamma, beta, 1e-6)
    y.backward(gy)
    xr = x.float().requires_grad_(True)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), g2, b2, 1e-6)
    yr.backward(gy.float())
    assert rel_err(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) < 6e-3
    assert rel_err(xa.grad.float().cpu().numpy(), xr.grad.cpu().numpy()) < 6e-3
    assert rel_err(gamma.grad.cpu().numpy(), g2.grad.cpu().numpy()) < 2e-3
    assert rel_err(beta.grad.cpu().numpy(), b2.grad.cpu().numpy()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("T,C", [(5000, 32), (333, 128), (2049, 512), (77, 8192), (10, 8)])
def test_colsum_matches_torch(dev, T, C):
    from sei_b200 import ops
    torch.manual_seed(T)
    x = torch.randn(T, C, device=dev).bfloat16()
    got = ops.colsum_bf16(x)
    ref = x.double().sum(0)
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,Cin,Cout,need_gx", [(2, 17, 23, 32, 3, True), (1, 8, 8, 8, 3, True), (3, 16, 16, 64, 4, True),
                                                     (2, 9, 33, 16, 1, False)])
def test_conv3x3_small_matches_torch(dev, B, H, W, Cin, Cout, need_gx):
    """direct 3x3 / few-output-channel convolution (the network's out_conv) vs F.conv2d in fp32 on the same bf16 input:
    forward within bf16 rounding of the result, gradients 6e-3 / 2e-3."""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(B * H + Cin)
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).requires_grad_(True)
    b = (0.1 * torch.randn(Cout, device=dev)).requires_grad_(True)
    gy = torch.randn(B, H, W, Cout, device=dev).bfloat16()
    assert ops.conv3x3_small_supported(x, Cout)
    xa = x.clone().requires_grad_(need_gx)
    y = ops.conv3x3_small(xa, w, b)
    assert y.shape == (B, H, W, 4) and (Cout == 4 or float(y[..., Cout:].abs().max()) == 0.0)
    y[..., :Cout].backward(gy)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, w2, b2, padding=1)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    assert rel_err(y[..., :Cout].detach().float().cpu().numpy(), yr.detach().permute(0, 2, 3, 1).cpu().numpy()) < 6e-3
    if need_gx:
        assert rel_err(xa.grad.float().cpu().numpy(), xr.grad.permute(0, 2, 3, 1).cpu().numpy()) < 6e-3
    assert rel_err(w.grad.cpu().numpy(), w2.grad.cpu().numpy()) < 2e-3
    assert rel_err(b.grad.cpu().numpy(), b2.grad.cpu().numpy()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C", [(2, 19, 21, 32), (1, 16, 16, 128), (1, 7, 40, 8), (1, 5, 9, 256), (2, 33, 8, 16),
                                     (2, 40, 37, 64), (1, 33, 50, 192), (3, 16, 16, 512)])
def test_dwconv7_matches_torch(dev, B, H, W, C):
    """depthwise 7x7 (ConvBlock.conv1) vs F.conv2d(groups=C) in fp32 on the same bf16 input"""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(B + H + C)
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    w = (torch.randn(C, 1, 7, 7, device=dev) / 7).requires_grad_(True)
    b = (0.1 * torch.randn(C, device=dev)).requires_grad_(True)
    gy = torch.randn(B, H, W, C, device=dev).bfloat16()
    assert ops.dwconv7_supported(x)
    xa = x.clone().requires_grad_(True)
    y = ops.dwconv7(xa, w, b)
    y.backward(gy)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, w2, b2, padding=3, groups=C)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    assert rel_err(y.detach().float().cpu().numpy(), yr.detach().permute(0, 2, 3, 1).cpu().numpy()) < 6e-3
    assert rel_err(xa.grad.float().cpu().numpy(), xr.grad.permute(0, 2, 3, 1).cpu().numpy()) < 6e-3
    assert rel_err(w.grad.cpu().numpy(), w2.grad.cpu().numpy()) < 2e-3
    assert rel_err(b.grad.cpu().numpy(), b2.grad.cpu().numpy()) < 2e-3


@pytest.mark.gpu
def test_gelu_matches_torch(dev):
    """erf-form GELU with the A&S 7.1.26 erf (|err| < 1.5e-7) vs torch's exact GELU in fp32 on the same bf16 input"""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(3)
    x = torch.cat([torch.randn(4096 * 8, device=dev) * 3, torch.linspace(-12, 12, 4096, device=dev)]).bfloat16()
    gy = torch.randn_like(x)
    xa = x.clone().requires_grad_(True)
    y = ops.gelu(xa)
    y.backward(gy)
    xr = x.float().requires_grad_(True)
    yr = F.gelu(xr)
    yr.backward(gy.float())
    # element-wise: one bf16 rounding of the exact value (2^-8 relative) plus the erf approximation's 1.5e-7 * |x|
    assert bool(torch.all((y.detach().float() - yr.detach()).abs() <= 2.0 ** -8 * yr.detach().abs() + 2e-6))
    assert rel_err(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) < 4e-3
    assert rel_err(xa.grad.float().cpu().numpy(), xr.grad.cpu().numpy()) < 6e-3
    # channels-last dense layout goes through unchanged
    x4 = torch.randn(2, 16, 5, 7, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    assert ops.gelu_supported(x4)
    y4 = ops.gelu(x4)
    assert y4.stride() == x4.stride()
    assert rel_err(y4.float().cpu().numpy(), F.gelu(x4.float()).cpu().numpy()) < 4e-3


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(256, 128, 512), (1000, 32, 128), (4096, 512, 2048), (130, 40, 72), (300, 264, 72),
                                   (2048, 64, 256)])
def test_gemm_residual_epilogue(dev, M, N, K):
    """D = A B^T + bias + s R with the addition in the GEMM epilogue (single-CTA TMA-store path, direct-store path for
    narrow outputs, CTA-pair kernel) vs the fp32 formulation"""
    from sei_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).bfloat16()
    for s in (1.0, 2.0):
        ref = a.float() @ b.float().t() + bias + s * r.float()
        d = ops.gemm_bf16_tn_residual(a, b, bias, r, s)
        assert d.dtype == torch.bfloat16 and d.shape == (M, N)
        assert rel_err(d.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3
    d0 = ops.gemm_bf16_tn_residual(a, b, None, r, 1.0)
    assert rel_err(d0.float().cpu().numpy(), (a.float() @ b.float().t() + r.float()).cpu().numpy()) < 6e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C", [(2, 19, 21, 32), (1, 16, 16, 128), (1, 7, 40, 8), (2, 40, 37, 64)])
def test_dwconv7_residual_store(dev, B, H, W, C):
    """y = dwconv7(x) + s * res with the addition in the kernel's store"""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(C)
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    res = torch.randn(B, H, W, C, device=dev).bfloat16()
    w = torch.randn(C, 49, device=dev) / 7
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.view(C, 1, 7, 7), None, padding=3, groups=C).permute(0, 2, 3, 1)
    for s in (1.0, 2.0):
        y = ops._dwconv7_raw(x, w.t().contiguous(), None, res=res, res_scale=s)
        assert rel_err(y.float().cpu().numpy(), (ref + s * res.float()).cpu().numpy()) < 6e-3


@pytest.mark.gpu
@pytest.mark.parametrize("T,C", [(1000, 128), (4096, 512), (333, 2048), (77, 8), (5000, 32)])
def test_gelu_bwd_colsum(dev, T, C):
    """gy * gelu'(h) and its column sums in one pass vs the two separate kernels (bit-identical products; sums within
    fp32 reordering)"""
    from sei_b200 import ops
    torch.manual_seed(T + C)
    h = (2 * torch.randn(T, C, device=dev)).bfloat16()
    gy = torch.randn(T, C, device=dev).bfloat16()
    gx, gb = ops.gelu_bwd_colsum(h, gy)
    ha = h.clone().requires_grad_(True)
    ops.gelu(ha).backward(gy)
    assert torch.equal(gx, ha.grad)
    ref = ha.grad.float().sum(0)
    assert rel_err(gb.cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("C,res_scale", [(32, 1.0), (128, 2.0), (512, 1.0)])
def test_convblock_single_node_matches_op_by_op(dev, C, res_scale, monkeypatch):
    """ConvBlock as one autograd node (additions fused into the GEMM epilogue / depthwise store, bias gradient inside
    the GELU backward) against the op-by-op autograd graph of the same kernels"""
    import models.convolutional as mc
    torch.manual_seed(C)
    monkeypatch.setattr(mc, "_GELU_EPILOGUE_MIN_C", 128)       # C = 32: separate GELU kernels; 128, 512: GEMM epilogues
    blk = mc.ConvBlock(C).to(dev)
    x = torch.randn(2, C, 24, 16, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, C, 24, 16, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    outs = []
    for node in (True, False):
        monkeypatch.setattr(mc, "_CONVBLOCK_NODE", node)
        blk.zero_grad(set_to_none=True)
        xa = x.clone().requires_grad_(True)
        y = blk(xa, res_scale=res_scale)
        y.backward(gy)
        outs.append((y.detach().float(), xa.grad.float(), {n: p.grad.detach().float().clone() for n, p in blk.named_parameters()}))
    (y1, gx1, g1), (y0, gx0, g0) = outs
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 8e-3          # one bf16 rounding instead of two
    assert rel_err(gx1.cpu().numpy(), gx0.cpu().numpy()) < 8e-3
    for n in g0:       # (the stored gelu' is rounded to bf16 on the epilogue path: 2^-9 relative per element)
        assert rel_err(g1[n].cpu().numpy(), g0[n].cpu().numpy()) < 6e-3, n


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(256, 128, 32), (1000, 512, 128), (4096, 2048, 512), (130, 72, 40), (300, 264, 72)])
def test_gemm_gelu_dual_and_multiplier_epilogues(dev, M, N, K):
    """conv2 + GELU as one kernel (gelu and gelu' written from the GEMM epilogue) and the multiplier epilogue of conv3's
    input gradient, vs the fp32 formulation on the same bf16 operands"""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = 0.3 * torch.randn(N, device=dev)
    act, der = ops.gemm_bf16_tn_gelu_dual(a, b, bias)
    h = (a.float() @ b.float().t() + bias).bfloat16().float().requires_grad_(True)     # the rounded pre-activation
    ref_act = F.gelu(h)
    ref_act.sum().backward()
    assert rel_err(act.float().cpu().numpy(), ref_act.detach().cpu().numpy()) < 8e-3
    assert rel_err(der.float().cpu().numpy(), h.grad.cpu().numpy()) < 8e-3
    # away from roundings of h that flip with the accumulation order, the values are those of the separate GELU kernel
    h_sei = ops.gemm_bf16_tn(a, b, bias)
    assert torch.equal(act, ops.gelu_raw(h_sei))
    mult = torch.randn(M, N, device=dev).bfloat16()
    d = ops.gemm_bf16_tn_mul(a, b, mult)
    ref = (a.float() @ b.float().t()) * mult.float()
    assert rel_err(d.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 24, 40), (3, 19, 21), (2, 64, 48)])
@pytest.mark.parametrize("kind", ["in", "out"])
def test_conv3x3_implicit_gemm_matches_torch(dev, B, H, W, kind):
    """the 3x3 'same' edge convolutions as implicit GEMMs on tcgen05 (nine TMA windows with zero fill as K-chunks) vs
    F.conv2d in fp32 on the same bf16 operands: forward, input gradient, weight and bias gradients"""
    import torch.nn.functional as F
    import models.convolutional as mc
    torch.manual_seed(B * 100 + H + W)
    cin, cout = (3, 32) if kind == "in" else (32, 3)
    conv = mc._conv(cin, cout, 3, padding="same").to(dev)
    x = torch.randn(B, cin, H, W, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(B, cout, H, W, device=dev).bfloat16()
    xa = x.clone().requires_grad_(True)
    y = conv(xa)
    y.backward(gy)
    wq = conv.weight.detach().bfloat16().float().requires_grad_(True)           # the tensor cores read bf16 weights
    bq = conv.bias.detach().clone().requires_grad_(True)
    xr = x.float().requires_grad_(True)
    yr = F.conv2d(xr, wq, bq, padding=1)
    yr.backward(gy.float())
    assert y.shape == yr.shape
    assert rel_err(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) < 6e-3
    assert rel_err(xa.grad.float().cpu().numpy(), xr.grad.cpu().numpy()) < 8e-3
    assert rel_err(conv.weight.grad.cpu().numpy(), wq.grad.cpu().numpy()) < 3e-3
    assert rel_err(conv.bias.grad.cpu().numpy(), bq.grad.cpu().numpy()) < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("C", [32, 512])
def test_convblock_mlp_recompute_is_bit_identical(dev, C, monkeypatch):
    """recomputing gelu / gelu' in the backward pass (memory option) runs the same kernels on the same inputs: outputs and
    input gradients are bit-identical to the default path, parameter gradients equal up to fp32 summation order"""
    import models.convolutional as mc
    torch.manual_seed(C + 1)
    blk = mc.ConvBlock(C).to(dev)
    x = torch.randn(2, C, 16, 24, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, C, 16, 24, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    res = []
    for rec in (False, True):
        monkeypatch.setattr(mc, "_RECOMPUTE_MLP", rec)
        blk.zero_grad(set_to_none=True)
        xa = x.clone().requires_grad_(True)
        y = blk(xa)
        y.backward(gy)
        res.append((y.detach().clone(), xa.grad.clone(), {n: p.grad.detach().clone() for n, p in blk.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for n in res[0][2]:          # (split-K weight gradients add their slices with atomics: equal up to the order of fp32 sums)
        assert rel_err(res[1][2][n].float().cpu().numpy(), res[0][2][n].float().cpu().numpy()) < 1e-5, n


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,period", [(1024, 128, 32, 256), (4096, 512, 128, 64), (300, 40, 72, 25), (2048, 2048, 512, 16)])
def test_gemm_rowscaled_bias(dev, M, N, K, period):
    """D = A B^T + bias[n] * row_scale[m % period] (Downsample's convolution behind its resampler) vs fp32"""
    from sei_b200 import ops
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias, rs = torch.randn(N, device=dev), torch.rand(period, device=dev) + 0.5
    d = ops.gemm_bf16_tn_rowscaled_bias(a, b, bias, rs)
    ref = a.float() @ b.float().t() + rs.repeat(-(-M // period))[:M, None] * bias[None, :]
    assert rel_err(d.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (300, 264, 72), (4100, 2048, 520), (8192, 512, 2048)])
def test_gemm_nn_reads_the_weight_in_place(dev, M, N, K):
    """D = A B with B [K, N] row-major read in place as an MN-major operand of the CTA-pair kernel (the input gradient of a
    pointwise convolution without a transposed weight copy), plain and with the multiplier epilogue, vs fp32"""
    from sei_b200 import ops, last_kernel
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
    mult = torch.randn(M, N, device=dev).bfloat16()
    ref = a.float() @ b.float()
    d = ops.gemm_bf16_nn(a, b)
    assert last_kernel() == "gemm_bf16_tn_2cta_kernel"
    assert rel_err(d.float().cpu().numpy(), ref.cpu().numpy()) < 6e-3
    assert torch.equal(d, ops.gemm_bf16_tn(a, b.t().contiguous()))          # same products, same accumulation order
    dm = ops.gemm_bf16_nn(a, b, mult)
    assert rel_err(dm.float().cpu().numpy(), (ref * mult.float()).cpu().numpy()) < 6e-3
