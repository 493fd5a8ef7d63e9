"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every comparison goes through the
C ABI of libsei_b200.so (via the sei_b200 ctypes binding and the reference-facing modules) and
checks against (a) golden vectors produced by running the reference itself and (b) the CPU
oracle, within the north-star tolerance: 1e-5 relative (max |a-b| / max |b|) in fp32."""
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402
from util import rel_err  # noqa: E402

TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sei_b200
    lib = sei_b200._lib.load()
    assert lib.sei_device_info(None, None, None, None) == 0, lib.sei_last_error()
    return torch.device("cuda:0")


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def npy(t):
    return t.detach().float().cpu().numpy()


def base_args(**kw):
    a = dict(task="deblurring", noise_level=5, physics_v2=True, kernel="Gaussian_R2", sr_factor=None,
             physics_true_adjoint=False, partial_sure=True, sure_margin=None, partial_sure_sr=False,
             Loss__crop_training_pairs=False, Loss__crop_size=48, ProposedLoss__stop_gradient=True,
             ProposedLoss__sure_alternative=None, ProposedLoss__alpha_tradeoff=1.0,
             ProposedLoss__transforms="Scaling_Transforms", ScalingTransform__kind="padded",
             ScalingTransform__antialias=False, method="proposed", sure_cropped_div=True,
             sure_averaged_cst=None)
    a.update(kw)
    return Namespace(**a)


# ------------------------------------------------------------------------------------ blur
@pytest.mark.parametrize("ci", range(7))
@pytest.mark.parametrize("path", [0, 1])
def test_blur_golden(golden, dev, ci, path):
    from sei_b200 import ops, last_kernel
    g = golden("blur")
    kern = orc.named_kernel(str(g[f"c{ci}_kernel_name"]))
    x, gy = cu(g[f"c{ci}_x"], dev), cu(g[f"c{ci}_gy"], dev)
    y = npy(ops.blur_circular(x, kern, path=path))
    if path == 1:
        assert last_kernel() == "blur_direct_kernel"
    for ref in ("v2_A_f32", "v1_A_f32", "v2_A_f64"):
        assert rel_err(y, g[f"c{ci}_{ref}"]) < TOL, ref
    xt = npy(ops.blur_circular(gy, kern, adjoint=True, path=path))
    for ref in ("v2_At_f32", "v2_vjp_f32", "v1_At_f32", "v2_At_f64"):
        assert rel_err(xt, g[f"c{ci}_{ref}"]) < TOL, ref


@pytest.mark.parametrize("kname,shape", [("Gaussian_R2", (4, 3, 256, 256)), ("Box_R3", (2, 3, 256, 256)),
                                         ("Gaussian_R3", (1, 3, 64, 512)), ("Gaussian_R1", (2, 3, 48, 48)),
                                         ("Box_R2", (1, 1, 40, 8)), ("Gaussian_R2", (1, 2, 100, 36)),
                                         ("Box_R4", (2, 1, 512, 512))])
def test_blur_tiled_vs_oracle(dev, kname, shape):
    from sei_b200 import ops, last_kernel
    rng = np.random.default_rng(5)
    x = rng.random(shape, dtype=np.float32)
    n = rng.standard_normal(shape).astype(np.float32)
    kern = orc.named_kernel(kname)
    xd, nd = cu(x, dev), cu(n, dev)
    y = npy(ops.blur_circular(xd, kern, path=2))
    assert last_kernel() == "blur_band_kernel"
    ref = orc.blur_circular(x.astype(np.float64), kern)
    assert rel_err(y, ref) < TOL
    assert rel_err(npy(ops.blur_circular(xd, kern, adjoint=True, path=2)), orc.blur_circular(x.astype(np.float64), kern, adjoint=True)) < TOL
    # fused noise epilogue == deepinv GaussianNoise applied to A(x)
    sigma = float(np.float32(5 / 255))
    yn = npy(ops.blur_circular(xd, kern, noise=nd, sigma=sigma, path=2))
    assert last_kernel() == "blur_band_kernel<noise>"
    assert rel_err(yn, orc.add_noise(ref, n.astype(np.float64), sigma)) < TOL
    # both kernels agree
    assert rel_err(npy(ops.blur_circular(xd, kern, path=1)), y) < 2e-6


def test_blur_nonseparable_and_even_kernels(dev):
    from sei_b200 import ops, last_kernel
    rng = np.random.default_rng(6)
    x = rng.random((2, 2, 24, 20), dtype=np.float32)
    for shape in [(3, 3), (4, 5), (6, 6), (1, 5), (13, 13)]:
        k = rng.random(shape)
        k /= k.sum()
        y = npy(ops.blur_circular(cu(x, dev), k))
        assert last_kernel() == "blur_direct_kernel"
        assert rel_err(y, orc.blur_circular(x.astype(np.float64), k)) < TOL
        yt = npy(ops.blur_circular(cu(x, dev), k, adjoint=True))
        assert rel_err(yt, orc.blur_circular(x.astype(np.float64), k, adjoint=True)) < TOL
    from sei_b200 import SeiError
    with pytest.raises(SeiError, match="smaller"):
        ops.blur_circular(cu(x[:, :, :8, :8], dev), orc.named_kernel("Gaussian_R2"))


@pytest.mark.parametrize("fname", ["g5", "box7", "rand4x5", "rand3x3", "row1x5"])
def test_v1_paddings_golden(golden, dev, fname):
    from physics.blur import conv, conv_transpose
    g = golden("blur_paddings")
    x, f = cu(g["x"], dev), torch.from_numpy(g[f"{fname}_filter"])
    for padding in ["valid", "circular", "replicate", "reflect"]:
        y = conv(x, f, padding)
        assert tuple(y.shape) == g[f"{fname}_{padding}_A"].shape
        assert rel_err(npy(y), g[f"{fname}_{padding}_A"]) < TOL
        xt = conv_transpose(cu(g[f"{fname}_{padding}_gy"], dev), f, padding)
        assert tuple(xt.shape) == g[f"{fname}_{padding}_At"].shape
        assert rel_err(npy(xt), g[f"{fname}_{padding}_At"]) < TOL
    assert rel_err(npy(conv_transpose(cu(g[f"{fname}_zero_gy"], dev), f, "zero")), g[f"{fname}_zero_At"]) < TOL


# ------------------------------------------------------------------------------------ SR
@pytest.mark.parametrize("ci", range(8))
@pytest.mark.parametrize("path", [0, 1])
def test_down_golden(golden, dev, ci, path):
    from sei_b200 import ops
    g = golden("downsampling")
    rate = int(g[f"c{ci}_rate"])
    x, gy = cu(g[f"c{ci}_x"], dev), cu(g[f"c{ci}_gy"], dev)
    y = ops.down_aa(x, rate, path=path)
    assert tuple(y.shape) == g[f"c{ci}_A_f32"].shape
    for tag in ("f32", "f64"):
        assert rel_err(npy(y), g[f"c{ci}_A_{tag}"]) < TOL
        assert rel_err(npy(ops.down_aa_transpose(gy, rate, tuple(x.shape[-2:]), path=path)), g[f"c{ci}_vjp_{tag}"]) < TOL
        assert rel_err(npy(ops.up_bicubic(gy, rate)), g[f"c{ci}_At_plain_{tag}"]) < TOL


@pytest.mark.parametrize("rate,shape", [(2, (2, 3, 512, 512)), (4, (1, 3, 1024, 1024)), (3, (1, 2, 96, 96)),
                                        (2, (3, 3, 96, 96)), (4, (2, 1, 100, 192)), (2, (1, 1, 37, 64)),
                                        (2, (1, 3, 256, 2048))])
def test_down_tiled_vs_oracle(dev, rate, shape):
    from sei_b200 import ops, last_kernel
    rng = np.random.default_rng(7)
    x = rng.random(shape, dtype=np.float32)
    xd = cu(x, dev)
    y = ops.down_aa(xd, rate, path=2)
    assert last_kernel() == "down_rows_kernel"
    ref = orc.down_aa(x.astype(np.float64), rate)
    assert rel_err(npy(y), ref) < TOL
    gy = rng.standard_normal(ref.shape).astype(np.float32)
    gx = ops.down_aa_transpose(cu(gy, dev), rate, shape[-2:], path=2)
    assert last_kernel() in ("down_t_rows_kernel", "down_t_band_kernel")
    assert rel_err(npy(gx), orc.down_aa_vjp(gy.astype(np.float64), rate, shape[-2:])) < TOL
    n = rng.standard_normal(ref.shape).astype(np.float32)
    yn = ops.down_aa(xd, rate, noise=cu(n, dev), sigma=0.02, path=2)
    assert rel_err(npy(yn), orc.add_noise(ref, n.astype(np.float64), 0.02)) < TOL
    assert rel_err(npy(ops.down_aa(xd, rate, path=1)), npy(y)) < 2e-6
    assert rel_err(npy(ops.down_aa_transpose(cu(gy, dev), rate, shape[-2:], path=1)), npy(gx)) < 2e-6


# ------------------------------------------------------------------------------------ scale transform
@pytest.mark.parametrize("ci", range(4))
@pytest.mark.parametrize("path", [0, 1])
def test_scale_transform_golden(golden, dev, ci, path):
    from sei_b200 import ops
    g = golden("transform")
    x, rate, center = cu(g[f"c{ci}_x"], dev), cu(g[f"c{ci}_rate"], dev), cu(g[f"c{ci}_center"], dev)
    out = npy(ops.scale_transform(x, rate, center, path=path))
    assert rel_err(out, g[f"c{ci}_T_f32"]) < TOL
    assert rel_err(out, g[f"c{ci}_T_f64"]) < 5 * TOL   # fp32 grid rounding vs an fp64 grid


@pytest.mark.parametrize("B,C,S", [(4, 3, 256), (2, 3, 48), (3, 1, 512), (2, 2, 128), (1, 1, 8)])
def test_scale_transform_tiled_vs_oracle(dev, B, C, S):
    from sei_b200 import ops, last_kernel
    rng = np.random.default_rng(8)
    x = rng.random((B, C, S, S), dtype=np.float32)
    rate = rng.choice(np.array([0.75, 0.5], np.float32), size=B)
    rate[0] = 0.5
    center = (2 * rng.random((B, 1, 1, 2), dtype=np.float32) - 1).astype(np.float32)
    center[0, 0, 0] = (1.0, -1.0)     # extreme centre: half of the output is reflected content
    out = npy(ops.scale_transform(cu(x, dev), cu(rate, dev), cu(center, dev), path=2))
    assert last_kernel() == "scale_rows_kernel"
    ref = orc.scale_transform(x, rate, center)          # fp32 oracle: same grid rounding as the reference
    assert rel_err(out, ref) < TOL
    assert rel_err(npy(ops.scale_transform(cu(x, dev), cu(rate, dev), cu(center, dev), path=1)), ref) < TOL
    # rates below 0.5 (outside what ScalingTransform samples) take the in-kernel global-gather branch
    rate2 = np.full(B, 0.3, np.float32)
    out2 = npy(ops.scale_transform(cu(x, dev), cu(rate2, dev), cu(center, dev), path=2))
    assert rel_err(out2, orc.scale_transform(x, rate2, center)) < TOL
    rate3 = np.full(B, 1.7, np.float32)               # zoom-in also works
    out3 = npy(ops.scale_transform(cu(x, dev), cu(rate3, dev), cu(center, dev), path=2))
    assert rel_err(out3, orc.scale_transform(x, rate3, center)) < TOL


def test_scale_params_and_module(golden, dev):
    import transforms
    from sei_b200 import draws
    g = golden("transform")
    with draws.inject([g["params_draw0_rand"], g["params_draw1_rand"]]):
        rate, center = transforms.sample_downsampling_parameters(16, dev, torch.float32, [0.75, 0.5])
    assert np.array_equal(npy(rate), g["params_rate"]) and np.array_equal(npy(center), g["params_center"])
    T = transforms.ScalingTransform(kind="padded", antialias=False)
    with draws.inject([g["module_draw1_rand"], g["module_draw2_rand"]]):
        y = T(cu(g["module_x"], dev))
    assert rel_err(npy(y), g["module_T"]) < TOL
    # without injection: draws come from the device generator, rates only from {0.75, 0.5}
    torch.manual_seed(0)
    rate, center = transforms.sample_downsampling_parameters(4096, dev, torch.float32, [0.75, 0.5])
    assert set(np.unique(npy(rate))) == {0.5, 0.75} and abs(float((rate == 0.5).float().mean()) - 0.5) < 0.05
    assert float(center.min()) >= -1 and float(center.max()) <= 1 and center.shape == (4096, 1, 1, 2)


def test_scale_transform_backward_vs_oracle(dev):
    from sei_b200 import ops
    rng = np.random.default_rng(13)
    for B, C, S in [(3, 2, 40), (2, 3, 256)]:
        g = rng.standard_normal((B, C, S, S)).astype(np.float32)
        rate = rng.choice(np.array([0.75, 0.5], np.float32), size=B)
        center = (2 * rng.random((B, 1, 1, 2), dtype=np.float32) - 1).astype(np.float32)
        gx = ops.scale_transform_backward(cu(g, dev), cu(rate, dev), cu(center, dev))
        assert rel_err(npy(gx), orc.scale_transform_vjp(g, rate, center)) < TOL
        # adjointness with the forward kernel
        x = rng.random((B, C, S, S), dtype=np.float32)
        Tx = ops.scale_transform(cu(x, dev), cu(rate, dev), cu(center, dev))
        lhs, rhs = (Tx.double() * cu(g, dev).double()).sum(), (cu(x, dev).double() * gx.double()).sum()
        assert abs(float(lhs - rhs)) < 1e-5 * abs(float(lhs))


# ------------------------------------------------------------------------------------ fused EI re-measurement
@pytest.mark.parametrize("kname,B,C,S", [("Gaussian_R2", 4, 3, 48), ("Gaussian_R2", 2, 3, 256), ("Box_R3", 3, 3, 256),
                                         ("Gaussian_R1", 2, 1, 64), ("Gaussian_R3", 1, 3, 128), ("Box_R2", 2, 2, 32),
                                         ("Gaussian_R2", 1, 1, 16), ("Gaussian_R2", 1, 3, 1024)])
@pytest.mark.parametrize("fused", [True, False])
def test_ei_remeasure_blur_vs_oracle(dev, kname, B, C, S, fused, monkeypatch):
    from sei_b200 import ops, last_kernel
    rng = np.random.default_rng(9)
    x = rng.random((B, C, S, S), dtype=np.float32)
    rate = rng.choice(np.array([0.75, 0.5], np.float32), size=B)
    rate[0] = 0.5
    center = (2 * rng.random((B, 1, 1, 2), dtype=np.float32) - 1).astype(np.float32)
    n = rng.standard_normal((B, C, S, S)).astype(np.float32)
    kern = orc.named_kernel(kname)
    sigma = float(np.float32(5 / 255))
    monkeypatch.setenv("SEI_EI_FUSED", "1" if fused else "0")
    x2, y = ops.ei_remeasure(cu(x, dev), cu(rate, dev), cu(center, dev), kern, 1, cu(n, dev), sigma)
    assert last_kernel() == ("ei_blur_band_kernel" if (fused and S <= 512) else "blur_band_kernel<noise>")
    x2_ref = orc.scale_transform(x, rate, center)
    assert rel_err(npy(x2), x2_ref) < TOL
    y_ref = orc.add_noise(orc.blur_circular(x2_ref.astype(np.float64), kern), n.astype(np.float64), sigma)
    assert rel_err(npy(y), y_ref) < TOL
    # the unfused kernels give the same pair
    x2u = ops.scale_transform(cu(x, dev), cu(rate, dev), cu(center, dev))
    yu = ops.blur_circular(x2u, kern, noise=cu(n, dev), sigma=sigma)
    assert rel_err(npy(x2), npy(x2u)) < 2e-6 and rel_err(npy(y), npy(yu)) < 2e-6
    # taps computed inside the fused kernel (no workspace) give the same result
    x2n, yn = ops.ei_remeasure(cu(x, dev), cu(rate, dev), cu(center, dev), kern, 1, cu(n, dev), sigma, use_workspace=False)
    assert torch.equal(x2n, x2) and torch.equal(yn, y)
    # no noise
    _, y0 = ops.ei_remeasure(cu(x, dev), cu(rate, dev), cu(center, dev), kern, 1, None, 0.0)
    assert rel_err(npy(y0), orc.blur_circular(x2_ref.astype(np.float64), kern)) < TOL


@pytest.mark.parametrize("r,B,C,S", [(2, 2, 3, 96), (4, 1, 3, 128), (2, 1, 1, 512)])
def test_ei_remeasure_sr_vs_oracle(dev, r, B, C, S):
    from sei_b200 import ops
    rng = np.random.default_rng(10)
    x = rng.random((B, C, S, S), dtype=np.float32)
    rate = rng.choice(np.array([0.75, 0.5], np.float32), size=B)
    center = (2 * rng.random((B, 1, 1, 2), dtype=np.float32) - 1).astype(np.float32)
    n = rng.standard_normal((B, C, S // r, S // r)).astype(np.float32)
    x2, y = ops.ei_remeasure(cu(x, dev), cu(rate, dev), cu(center, dev), None, r, cu(n, dev), 0.02)
    x2_ref = orc.scale_transform(x, rate, center)
    assert rel_err(npy(x2), x2_ref) < TOL
    assert rel_err(npy(y), orc.add_noise(orc.down_aa(x2_ref.astype(np.float64), r), n.astype(np.float64), 0.02)) < TOL


# ------------------------------------------------------------------------------------ reductions
def test_mse_and_sure_reductions(dev):
    from sei_b200 import ops
    rng = np.random.default_rng(11)
    for shape, margin in [((4, 3, 32, 32), 6), ((2, 3, 256, 256), 6), ((1, 1, 9, 13), 0), ((2, 3, 40, 44), 3)]:
        y1, y2, y = (rng.random(shape, dtype=np.float32) for _ in range(3))
        b = rng.standard_normal(shape).astype(np.float32)
        if margin:
            b[:, :, :margin] = 0; b[:, :, -margin:] = 0; b[:, :, :, :margin] = 0; b[:, :, :, -margin:] = 0
        a = cu(y1, dev).requires_grad_(True)
        t = cu(y2, dev).requires_grad_(True)
        m = ops.mse(a, t)
        assert abs(float(m) - orc.mse(y1.astype(np.float64), y2.astype(np.float64))) < 1e-6 * float(m)
        m.backward()
        gref = 2 * (y1.astype(np.float64) - y2) / y1.size
        assert rel_err(npy(a.grad), gref) < TOL and rel_err(npy(t.grad), -gref) < TOL
        for mm, md, avg in [(margin, margin, None), (margin, 0, True)]:
            sigma2, tau = (5 / 255) ** 2, 1e-2
            a1 = cu(y1, dev).requires_grad_(True)
            a2 = cu(y2, dev).requires_grad_(True)
            loss, aux = ops.sure_loss(a1, a2, cu(y, dev), cu(b, dev), mm, md, tau, sigma2, avg)
            ref, mse_r, div_r = orc.sure_loss(y1.astype(np.float64), y2.astype(np.float64), y.astype(np.float64),
                                              b.astype(np.float64), mm, md, tau, sigma2, avg)
            assert abs(float(loss) - ref) < 2e-6 * abs(ref) + 1e-9
            assert abs(float(aux[1]) - mse_r) < 2e-6 * mse_r and abs(float(aux[2]) - div_r) < 2e-5 * abs(div_r) + 1e-7
            (3.0 * loss).backward()
            H, W = shape[-2:]
            mask_m = np.zeros(shape); mask_m[:, :, mm:H - mm, mm:W - mm] = 1
            mask_d = np.zeros(shape); mask_d[:, :, md:H - md, md:W - md] = 1
            d = 3.0 * 2 * sigma2 * b * mask_d / (tau * mask_d.sum())
            g1 = 3.0 * 2 * (y1.astype(np.float64) - y) * mask_m / mask_m.sum() - d
            assert rel_err(npy(a1.grad), g1) < TOL and rel_err(npy(a2.grad), d) < TOL
        # deterministic: bitwise identical on repetition
        assert float(ops.mse(cu(y1, dev), cu(y2, dev))) == float(ops.mse(cu(y1, dev), cu(y2, dev)))


def test_sure_probe(dev):
    from sei_b200 import ops
    rng = np.random.default_rng(12)
    y = rng.random((2, 3, 20, 24), dtype=np.float32)
    for margin in (0, 3):
        draw = rng.standard_normal((2, 3, 20 - 2 * margin, 24 - 2 * margin)).astype(np.float32)
        out, b = ops.sure_perturb(cu(y, dev), cu(draw, dev), margin, 1e-2)
        bref = np.zeros_like(y)
        bref[:, :, margin:20 - margin, margin:24 - margin] = draw
        assert np.array_equal(npy(b), bref)
        assert np.array_equal(npy(out), y + bref * np.float32(1e-2))      # bit-exact with torch's y + b * tau


# ------------------------------------------------------------------------------------ physics objects, autograd
def test_physics_objects_and_autograd(golden, dev):
    import physics
    from sei_b200 import draws
    g = golden("blur")
    for v2 in (True, False):
        phys = physics.get_physics(base_args(physics_v2=v2), device=dev)
        x = cu(g["c6_x"], dev).requires_grad_(True)
        gy = cu(g["c6_gy"], dev)
        y = phys.A(x)
        (y * gy).sum().backward()
        assert rel_err(npy(y), g["c6_v2_A_f32"]) < TOL
        assert rel_err(npy(x.grad), g["c6_v2_vjp_f32"]) < TOL
        assert rel_err(npy(phys.A_adjoint(gy)), g["c6_v2_At_f32"]) < TOL
        n = torch.randn_like(gy)
        with draws.inject([n]):
            yn = phys(x.detach())
        assert np.array_equal(npy(yn), npy(y.detach() + n * phys.noise_model.sigma.to(dev)))
    # A_dagger against the reference's own physics objects (six CG iterations, tests/golden/dagger.npz) and against the
    # oracle's conjugate gradient: deblurring v2 and v1, SR with the plain-upsample and the true adjoint
    gdag = golden("dagger")
    for name, kw in (("deblur_g1", dict(kernel="Gaussian_R1")), ("deblur_box2_v1", dict(kernel="Box_R2", physics_v2=False)),
                     ("sr2_plain", dict(task="sr", kernel=None, sr_factor=2)),
                     ("sr2_true", dict(task="sr", kernel=None, sr_factor=2, physics_true_adjoint=True))):
        phys = physics.get_physics(base_args(**kw), device=dev)
        phys.max_iter, phys.tol = 6, 1e-12
        rec = phys.A_dagger(cu(gdag[f"{name}_y_f32"], dev))
        assert rel_err(npy(rec), gdag[f"{name}_dagger_f32"]) < TOL, name
    kern = orc.named_kernel("Gaussian_R1")
    yb = gdag["deblur_g1_y_f32"]
    rec_o = orc.a_dagger(lambda v: orc.blur_circular(v, kern), lambda v: orc.blur_circular(v, kern, adjoint=True), yb, 6, 1e-12)
    phys = physics.get_physics(base_args(kernel="Gaussian_R1"), device=dev)
    phys.max_iter, phys.tol = 6, 1e-12
    assert rel_err(npy(phys.A_dagger(cu(yb, dev))), rec_o) < TOL
    gd = golden("downsampling")
    for ci in (0, 1):
        rate = int(gd[f"c{ci}_rate"])
        phys = physics.get_physics(base_args(task="sr", kernel=None, sr_factor=rate), device=dev)
        x = cu(gd[f"c{ci}_x"], dev).requires_grad_(True)
        gy = cu(gd[f"c{ci}_gy"], dev)
        (phys.A(x) * gy).sum().backward()
        assert rel_err(npy(x.grad), gd[f"c{ci}_vjp_f32"]) < TOL
        assert rel_err(npy(phys.A_adjoint(gy)), gd[f"c{ci}_At_plain_f32"]) < TOL
        phys_t = physics.get_physics(base_args(task="sr", kernel=None, sr_factor=rate, physics_true_adjoint=True), device=dev)
        assert rel_err(npy(phys_t.A_adjoint(gy)), gd[f"c{ci}_At_true_f32"]) < TOL


def test_randomly_degrade(golden, dev):
    import physics
    g = golden("degrade")
    for name, kw in [("deblur", dict()), ("sr2", dict(task="sr", kernel=None, sr_factor=2))]:
        phys = physics.get_physics(base_args(**kw), device=dev)
        mgr = getattr(phys, "__manager")
        x = cu(g[f"{name}_x"], dev)
        torch.manual_seed(123)
        cpu_state, cuda_state = torch.get_rng_state().clone(), torch.cuda.get_rng_state(dev).clone()
        y1 = mgr.randomly_degrade(x, seed=42)
        assert torch.equal(cpu_state, torch.get_rng_state()) and torch.equal(cuda_state, torch.cuda.get_rng_state(dev))
        y2 = mgr.randomly_degrade(x, seed=42)
        y3 = mgr.randomly_degrade(x, seed=43)
        assert torch.equal(y1, y2) and not torch.equal(y1, y3)
        # same operator as the reference; the noise values differ (device generator), its level does not
        clean = g[f"{name}_y"] - g[f"{name}_draw0_randn_like"] * g[f"{name}_sigma"]
        resid = npy(y1) - clean
        assert abs(resid.std() - 5 / 255) < 0.1 * 5 / 255 and abs(resid.mean()) < 1e-3


def test_randomly_degrade_batch_equals_per_image_calls(dev):
    """N2: one operator launch for a batch of dataset items = the reference's per-item calls (same seeds, same draws)"""
    import physics
    for kw in [dict(), dict(physics_v2=False), dict(task="sr", kernel=None, sr_factor=2)]:
        phys = physics.get_physics(base_args(**kw), device=dev)
        mgr = getattr(phys, "__manager")
        torch.manual_seed(5)
        x = torch.rand(6, 3, 64, 64, device=dev)
        seeds = [11, 12, 0, 0, 99, 7]
        torch.manual_seed(77)
        state = torch.cuda.get_rng_state(dev).clone()
        yb = mgr.randomly_degrade_batch(x, seeds)
        assert torch.equal(state, torch.cuda.get_rng_state(dev))          # seeded items leave the global stream alone
        ys = torch.cat([mgr.randomly_degrade(x[i:i + 1], seed=s) for i, s in enumerate(seeds)])
        # same draws; the fused epilogue rounds sigma * n + y once (fma), the two-kernel composition twice
        assert yb.shape == ys.shape and float((yb - ys).abs().max()) < 5e-7
        n2, n3 = yb[2] - phys.A(x[2:3])[0], yb[3] - phys.A(x[3:4])[0]                # same seed, same noise
        assert float((n2 - n3).abs().max()) < 5e-7 and abs(float(n2.std()) - 5 / 255) < 0.1 * 5 / 255
        # unseeded items (CSS re-degradation) consume the global stream in item order
        torch.manual_seed(3)
        yb = mgr.randomly_degrade_batch(x, None)
        torch.manual_seed(3)
        ys = torch.cat([mgr.randomly_degrade(x[i:i + 1], seed=None) for i in range(6)])
        assert float((yb - ys).abs().max()) < 5e-7
        with pytest.raises(ValueError):
            mgr.randomly_degrade_batch(x, [1, 2])


# ------------------------------------------------------------------------------------ full loss assembly vs the reference
LOSS_CASES = ["deblur_gauss2_proposed", "deblur_box3_proposed", "deblur_gauss2_v1_proposed", "sr2_proposed",
              "sr4_proposed", "sr2_partial_proposed", "deblur_gauss2_sure", "deblur_gauss2_sure_avgcst",
              "deblur_gauss2_sure_nocrop", "deblur_gauss2_supervised", "sr2_css", "deblur_gauss2_proposed_alpha",
              "cfg1_deblur_gauss2_proposed", "deblur_gauss2_r2r", "sr2_r2r", "deblur_gauss2_nostopgrad", "sr2_nostopgrad", "deblur_gauss2_shifts",
              "deblur_gauss2_normalT", "deblur_gauss2_normalT_aa", "deblur_gauss2_rotations", "deblur_gauss2_rotshift",
              # the reference's default Loss path: random batch crop (with the 4-D padding quirk) before the method loss
              "deblur_gauss2_proposed_crop", "sr2_proposed_crop", "deblur_gauss2_supervised_crop32"]
LOSS_ARGS = {
    "deblur_gauss2_proposed": dict(), "deblur_box3_proposed": dict(kernel="Box_R3"),
    "deblur_gauss2_v1_proposed": dict(physics_v2=False),
    "sr2_proposed": dict(task="sr", kernel=None, sr_factor=2), "sr4_proposed": dict(task="sr", kernel=None, sr_factor=4),
    "sr2_partial_proposed": dict(task="sr", kernel=None, sr_factor=2, partial_sure_sr=True),
    "deblur_gauss2_sure": dict(method="sure"), "deblur_gauss2_sure_avgcst": dict(method="sure", sure_averaged_cst=True),
    "deblur_gauss2_sure_nocrop": dict(method="sure", sure_cropped_div=False),
    "deblur_gauss2_supervised": dict(method="supervised"), "sr2_css": dict(task="sr", kernel=None, sr_factor=2, method="css"),
    "deblur_gauss2_proposed_alpha": dict(ProposedLoss__alpha_tradeoff=0.3),
    "cfg1_deblur_gauss2_proposed": dict(),
    "deblur_gauss2_r2r": dict(ProposedLoss__sure_alternative="r2r"),
    "sr2_r2r": dict(task="sr", kernel=None, sr_factor=2, ProposedLoss__sure_alternative="r2r"),
    "deblur_gauss2_nostopgrad": dict(ProposedLoss__stop_gradient=False),
    "sr2_nostopgrad": dict(task="sr", kernel=None, sr_factor=2, ProposedLoss__stop_gradient=False),
    "deblur_gauss2_shifts": dict(ProposedLoss__transforms="Shifts"),
    "deblur_gauss2_normalT": dict(ScalingTransform__kind="normal"),
    "deblur_gauss2_normalT_aa": dict(ScalingTransform__kind="normal", ScalingTransform__antialias=True),
    "deblur_gauss2_rotations": dict(ProposedLoss__transforms="Rotations"),
    "deblur_gauss2_rotshift": dict(ProposedLoss__transforms="Rotations+Shifts"),
    "deblur_gauss2_proposed_crop": dict(Loss__crop_training_pairs=True),
    "sr2_proposed_crop": dict(task="sr", kernel=None, sr_factor=2, Loss__crop_training_pairs=True),
    "deblur_gauss2_supervised_crop32": dict(method="supervised", Loss__crop_training_pairs=True, Loss__crop_size=32),
}


class Tap(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.outs = []

    def forward(self, y, *args):
        out = self.model(y)
        if out.requires_grad:
            out.retain_grad()
        self.outs.append(out)
        return out


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_and_gradients_match_reference(golden, dev, name):
    """One loss evaluation + backward with the reference's own random tensors injected: loss value,
    the three network outputs, dL/d(network outputs) and dL/d(parameters) against the reference."""
    import losses
    import physics
    from sei_b200 import draws
    from toy_model import ToyModel
    g = golden(f"loss_{name}_f32")
    args = base_args(**LOSS_ARGS[name])
    phys = physics.get_physics(args, device=dev)
    loss_fn = losses.get_loss(args=args, physics=phys)
    rate = int(g["rate"])
    model = Tap(ToyModel(rate=rate).to(dev))
    injected = [g[k] for k in sorted((k for k in g if k.startswith("draw")), key=lambda s: int(s[4:].split("_")[0]))]
    with draws.inject(injected):
        loss = loss_fn(x=cu(g["x"], dev), y=cu(g["y"], dev), model=model)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    for i, o in enumerate(model.outs):
        assert rel_err(npy(o), g[f"model_out{i}"]) < TOL, f"model_out{i}"
        if f"model_out{i}_grad" in g:
            assert rel_err(npy(o.grad), g[f"model_out{i}_grad"]) < 2 * TOL, f"model_out{i}_grad"
    for pname, p in model.model.named_parameters():
        ref = g[f"grad_{pname}"]
        assert np.allclose(npy(p.grad), ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max() + 1e-9), pname


# ------------------------------------------------------------------------------------ size-independent properties at full size
def test_properties_at_benchmark_size(dev):
    """cfg2 / cfg3 shapes: adjointness, DC gain, shift equivariance, tiled == direct."""
    from sei_b200 import ops
    torch.manual_seed(0)
    kern = orc.named_kernel("Gaussian_R2")
    x = torch.rand(32, 3, 256, 256, device=dev)
    v = torch.randn(32, 3, 256, 256, device=dev)
    Ax, Atv = ops.blur_circular(x, kern), ops.blur_circular(v, kern, adjoint=True)
    lhs, rhs = (Ax.double() * v.double()).sum(), (x.double() * Atv.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-5 * abs(float(lhs))
    ones = torch.ones(2, 3, 256, 256, device=dev)
    assert float((ops.blur_circular(ones, kern) - 1).abs().max()) < 1e-6
    assert rel_err(npy(ops.blur_circular(torch.roll(x[:2], (5, -9), (-2, -1)), kern)),
                   npy(torch.roll(Ax[:2], (5, -9), (-2, -1)))) < 1e-6
    assert rel_err(npy(ops.blur_circular(x[:4], kern, path=1)), npy(Ax[:4])) < 2e-6
    # SR x2 at cfg3's per-GPU shape
    xs = torch.rand(8, 3, 512, 512, device=dev)
    vs = torch.randn(8, 3, 256, 256, device=dev)
    Ax, Atv = ops.down_aa(xs, 2), ops.down_aa_transpose(vs, 2, (512, 512))
    lhs, rhs = (Ax.double() * vs.double()).sum(), (xs.double() * Atv.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-5 * abs(float(lhs))
    assert float((ops.down_aa(torch.ones(1, 3, 512, 512, device=dev), 2) - 1).abs().max()) < 1e-6
    assert float((ops.down_aa(torch.ones(1, 3, 1024, 1024, device=dev), 4) - 1).abs().max()) < 1e-6
    # scale transform: constants are preserved (bicubic weights sum to 1), output bounded by bicubic overshoot
    rate = torch.tensor([0.75, 0.5] * 16, device=dev)
    center = 2 * torch.rand(32, 1, 1, 2, device=dev) - 1
    assert float((ops.scale_transform(torch.full_like(x, 0.37), rate, center) - 0.37).abs().max()) < 1e-6
    t = ops.scale_transform(x, rate, center)
    assert rel_err(npy(ops.scale_transform(x[:4], rate[:4], center[:4], path=1)), npy(t[:4])) < 2e-6
    x2, y = ops.ei_remeasure(x, rate, center, kern, 1, v, 0.02)
    assert rel_err(npy(x2[:4]), npy(t[:4])) < 2e-6
    assert rel_err(npy(y[:4]), npy(ops.blur_circular(t[:4], kern, noise=v[:4], sigma=0.02))) < 2e-6


def test_cuda_graph_capture(dev):
    """every entry point is asynchronous on the caller's stream and capturable"""
    from sei_b200 import ops
    kern = orc.named_kernel("Gaussian_R2")
    x = torch.rand(4, 3, 64, 64, device=dev)
    rate = torch.tensor([0.75, 0.5, 0.5, 0.75], device=dev)
    center = torch.zeros(4, 1, 1, 2, device=dev)
    n = torch.randn_like(x)
    eager = ops.ei_remeasure(x, rate, center, kern, 1, n, 0.02)
    m_eager = ops.mse(eager[0], eager[1])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ops.mse(*ops.ei_remeasure(x, rate, center, kern, 1, n, 0.02))
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        x2, y = ops.ei_remeasure(x, rate, center, kern, 1, n, 0.02)
        m = ops.mse(x2, y)
    x.copy_(torch.rand_like(x))
    graph.replay()
    torch.cuda.synchronize()
    ref = ops.ei_remeasure(x, rate, center, kern, 1, n, 0.02)
    assert torch.equal(x2, ref[0]) and torch.equal(y, ref[1])
    assert float(m) == float(ops.mse(ref[0], ref[1])) and float(m) != float(m_eager)


@pytest.mark.parametrize("rate", [0.75, 0.5])
@pytest.mark.parametrize("aa", [False, True])
def test_normal_transform_vs_reference(golden, dev, rate, aa):
    """ScalingTransform(kind="normal"): sei_resize_bicubic_f32 against fixtures produced by the reference's
    normal_downsampling_transform and against the oracle, 1e-5 relative"""
    import transforms
    g = golden("normal_transform")
    for i in range(3):
        x = g[f"x{i}"].astype(np.float32)
        got = transforms.normal_downsampling_transform(cu(x, dev), rate, "bicubic", aa).cpu().numpy()
        ref = g[f"y32_{i}_r{int(rate * 100)}_aa{int(aa)}"]
        assert got.shape == ref.shape
        assert rel_err(got, ref) < 1e-5, (i, rate, aa)
        assert rel_err(got, orc.resize_bicubic(x, rate, aa)) < 1e-5


def test_antialiased_padded_transform_vs_reference(golden, dev):
    """ScalingTransform(kind="padded", antialias=True): sei_resize_bicubic_f32 (pre-filter) + sei_scale_transform_src_f32
    against fixtures produced by the reference and against the oracle"""
    import transforms
    from sei_b200 import last_kernel
    g = golden("transform_aa")
    for ci in range(4):
        x, rate, center = (g[f"c{ci}_{k}"].astype(np.float32) for k in ("x", "rate", "center"))
        got = transforms.padded_downsampling_transform(cu(x, dev), cu(rate, dev), cu(center, dev), "bicubic", "reflection",
                                                       True).cpu().numpy()
        assert last_kernel() == "scale_direct_kernel"
        assert got.shape == x.shape
        assert rel_err(got, g[f"c{ci}_T_f32"]) < 1e-5, ci
        assert rel_err(got, orc.scale_transform_antialiased(x, rate, center)) < 1e-5, ci
    x = torch.rand(2, 3, 32, 32, device=dev)
    with pytest.raises(RuntimeError, match="equal size"):          # the reference fails the same way (torch.stack)
        transforms.padded_downsampling_transform(x, torch.tensor([0.75, 0.5], device=dev), torch.zeros(2, 1, 1, 2, device=dev),
                                                 "bicubic", "reflection", True)
    # gradient through the anti-aliased transform (hand-written transposes of both steps) against torch's CPU autograd
    # of the same composition: F.interpolate(antialias=True) then grid_sample on the reference's grid
    import torch.nn.functional as F
    rate, center = torch.full((2,), 0.5), torch.tensor([[0.2, -0.3], [-0.5, 0.4]]).view(2, 1, 1, 2)
    xa = x.clone().requires_grad_(True)
    gy = torch.rand(2, 3, 32, 32, device=dev)
    transforms.padded_downsampling_transform(xa, rate.to(dev), center.to(dev), "bicubic", "reflection", True).backward(gy)
    xr = x.cpu().clone().requires_grad_(True)
    small = F.interpolate(xr, scale_factor=0.5, mode="bicubic", antialias=True)
    grid = transforms.get_downsampling_grid((2, 3, 32, 32), rate, center, torch.float32, "cpu")
    F.grid_sample(small, grid, mode="bicubic", padding_mode="reflection", align_corners=True).backward(gy.cpu())
    assert rel_err(npy(xa.grad), xr.grad.numpy()) < 1e-5
    # module form: EI re-measurement goes through the unfused composition
    import physics
    phys = physics.get_physics(base_args(), device=dev)
    t = transforms.ScalingTransform(kind="padded", antialias=True)
    with draws_inject_equal_rates(dev):
        x2, y2 = t.fused_remeasure(x[:1], phys, apply_noise=False)
    assert x2.shape == (1, 3, 32, 32) and y2.shape == (1, 3, 32, 32)
    assert rel_err(npy(y2), npy(phys.A(x2))) < 1e-6


def draws_inject_equal_rates(dev):
    from sei_b200 import draws
    return draws.inject([np.array([0.7], dtype=np.float32), np.array([[0.25, 0.5]], dtype=np.float32)])


def test_rotate_vs_torchvision(golden, dev):
    """deepinv Rotate (torchvision rotate defaults): sei_rotate_nearest_f32 picks the same source pixels as torchvision's CPU
    fixtures and as the oracle (ties in the rounding of a source coordinate excepted: at most 1e-4 of the pixels)"""
    from sei_b200 import draws, last_kernel, ops
    from sei_b200.linear_physics import Rotate
    g = golden("rotate")
    bad = bad_orc = tot = 0
    for i in range(4):
        x = g[f"x{i}"]
        for a in g["angles"]:
            got = ops.rotate_nearest(cu(x, dev), float(a)).cpu().numpy()
            bad += int((got != g[f"y{i}_a{a}"]).sum())
            bad_orc += int((got != orc.rotate_nearest(x, float(a))).sum())
            tot += got.size
    assert last_kernel() == "rotate_nearest_kernel"
    assert bad <= 1e-4 * tot and bad_orc == 0, (bad, bad_orc, tot)
    with draws.inject([g["module_draw0_randperm"]]):
        y = Rotate()(cu(g["module_x"], dev))
    assert int((npy(y) != g["module_y"]).sum()) <= 2
    # backward: the transpose of the pixel selection (adjointness against the forward kernel: <R x, y> = <x, R^T y>)
    xa = cu(g["module_x"], dev).requires_grad_(True)
    with draws.inject([g["module_draw0_randperm"]]):
        ya = Rotate()(xa)
    w = torch.rand_like(ya)
    ya.backward(w)
    lhs, rhs = float((ya.detach().double() * w.double()).sum()), float((xa.detach().double() * xa.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs)


def test_normal_scaling_transform_module(dev):
    import transforms
    from sei_b200 import draws, last_kernel
    t = transforms.ScalingTransform(kind="normal", antialias=True)
    x = torch.rand(2, 3, 64, 64, device=dev)
    with draws.inject([np.array(0.7, dtype=np.float32)]):          # floor(2 * 0.7) = 1 -> rate 0.5
        y = t(x)
    assert y.shape == (2, 3, 32, 32) and last_kernel() == "resize_bicubic_kernel"
    with draws.inject([np.array(0.2, dtype=np.float32)]):          # rate 0.75
        assert t(x).shape == (2, 3, 48, 48)
    # gradient through the resize (--no-ProposedLoss__stop_gradient) against torch's CPU autograd of F.interpolate
    import torch.nn.functional as F
    for aa, u in ((True, 0.7), (False, 0.2)):
        tt = transforms.ScalingTransform(kind="normal", antialias=aa)
        xa = x.clone().requires_grad_(True)
        with draws.inject([np.array(u, dtype=np.float32)]):
            ya = tt(xa)
        gy = torch.rand_like(ya)
        ya.backward(gy)
        xr = x.cpu().clone().requires_grad_(True)
        yr = F.interpolate(xr, scale_factor=0.5 if u > 0.5 else 0.75, mode="bicubic", antialias=aa)
        yr.backward(gy.cpu())
        assert rel_err(npy(ya), yr.detach().numpy()) < 1e-5 and rel_err(npy(xa.grad), xr.grad.numpy()) < 1e-5, aa
