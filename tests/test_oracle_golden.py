"""Pin the CPU oracle (oracle/) against fixtures produced by running the reference itself
(tests/golden/make_golden.py).  fp64 instantiation vs the reference in fp64: 1e-11;
fp32 instantiation vs the reference in fp32: 2e-6 (accumulation-order noise only)."""
import numpy as np
import pytest

from oracle import oracle as orc
from util import rel_err

TOL64 = 1e-11
TOL32 = 2e-6
KERNELS = ["Gaussian_R1", "Gaussian_R2", "Gaussian_R3", "Box_R2", "Box_R3", "Box_R4"]


def test_named_kernels(golden):
    g = golden("kernels")
    for name in KERNELS:
        k = orc.named_kernel(name)
        assert k.shape == g[name].shape
        assert rel_err(k, g[name]) < 1e-14
        assert abs(k.sum() - 1) < 1e-14
    with pytest.raises(ValueError):
        orc.named_kernel("Gaussian_R9")


@pytest.mark.parametrize("ci", range(7))
def test_blur_circular(golden, ci):
    g = golden("blur")
    kern = orc.named_kernel(str(g[f"c{ci}_kernel_name"]))
    x, gy = g[f"c{ci}_x"], g[f"c{ci}_gy"]
    assert rel_err(orc.blur_circular(x, kern), g[f"c{ci}_v2_A_f64"]) < TOL64
    assert rel_err(orc.blur_circular(gy, kern, adjoint=True), g[f"c{ci}_v2_At_f64"]) < TOL64
    assert rel_err(orc.blur_circular(gy, kern, adjoint=True), g[f"c{ci}_v2_vjp_f64"]) < TOL64
    x32, gy32 = x.astype(np.float32), gy.astype(np.float32)
    for ref in ("v2_A_f32", "v1_A_f32"):
        assert rel_err(orc.blur_circular(x32, kern), g[f"c{ci}_{ref}"]) < TOL32
    for ref in ("v2_At_f32", "v2_vjp_f32", "v1_At_f32"):
        assert rel_err(orc.blur_circular(gy32, kern, adjoint=True), g[f"c{ci}_{ref}"]) < TOL32


@pytest.mark.parametrize("fname", ["g5", "box7", "rand4x5", "rand3x3", "row1x5"])
def test_v1_paddings(golden, fname):
    g = golden("blur_paddings")
    x, f = g["x"], g[f"{fname}_filter"]
    for padding in ["valid", "circular", "replicate", "reflect"]:
        y = orc.conv_v1(x, f, padding)
        assert y.shape == g[f"{fname}_{padding}_A"].shape
        assert rel_err(y, g[f"{fname}_{padding}_A"]) < TOL32
        xt = orc.conv_transpose_v1(g[f"{fname}_{padding}_gy"], f, padding)
        assert xt.shape == g[f"{fname}_{padding}_At"].shape
        assert rel_err(xt, g[f"{fname}_{padding}_At"]) < TOL32
    assert rel_err(orc.conv_transpose_v1(g[f"{fname}_zero_gy"], f, "zero"), g[f"{fname}_zero_At"]) < TOL32


@pytest.mark.parametrize("ci", range(8))
def test_downsampling(golden, ci):
    g = golden("downsampling")
    rate = int(g[f"c{ci}_rate"])
    x, gy = g[f"c{ci}_x"], g[f"c{ci}_gy"]
    assert orc.down_aa(x, rate).shape == g[f"c{ci}_A_f64"].shape
    assert rel_err(orc.down_aa(x, rate), g[f"c{ci}_A_f64"]) < TOL64
    assert rel_err(orc.down_aa_vjp(gy, rate, x.shape[-2:]), g[f"c{ci}_vjp_f64"]) < TOL64
    assert rel_err(orc.up_bicubic(gy, rate), g[f"c{ci}_At_plain_f64"]) < TOL64
    x32, gy32 = x.astype(np.float32), gy.astype(np.float32)
    assert rel_err(orc.down_aa(x32, rate), g[f"c{ci}_A_f32"]) < TOL32
    assert rel_err(orc.down_aa_vjp(gy32, rate, x.shape[-2:]), g[f"c{ci}_vjp_f32"]) < TOL32
    assert rel_err(orc.up_bicubic(gy32, rate), g[f"c{ci}_At_plain_f32"]) < TOL32
    if f"c{ci}_At_true_f64" in g:
        # reference quirk: adjoint_function is called without dtype (downsampling/__init__.py:30),
        # so the "true adjoint" is evaluated in fp32 even for fp64 inputs
        assert g[f"c{ci}_At_true_f64"].dtype == np.float32
        assert rel_err(orc.down_aa_vjp(gy32, rate, x.shape[-2:]), g[f"c{ci}_At_true_f64"]) < TOL32
        assert rel_err(orc.down_aa_vjp(gy32, rate, x.shape[-2:]), g[f"c{ci}_At_true_f32"]) < TOL32


@pytest.mark.parametrize("ci", range(4))
def test_scale_transform(golden, ci):
    g = golden("transform")
    x, rate, center = g[f"c{ci}_x"], g[f"c{ci}_rate"], g[f"c{ci}_center"]
    B, _, S, _ = x.shape
    assert rel_err(orc.scale_grid(B, S, rate, center, np.float64), g[f"c{ci}_grid_f64"]) < 1e-14
    assert rel_err(orc.scale_transform(x, rate, center), g[f"c{ci}_T_f64"]) < TOL64
    # fp32: the grid must match bit for bit (same rounding sequence as the reference)
    g32 = orc.scale_grid(B, S, rate.astype(np.float32), center.astype(np.float32), np.float32)
    assert np.array_equal(g32, g[f"c{ci}_grid_f32"])
    out32 = orc.scale_transform(x.astype(np.float32), rate.astype(np.float32), center.astype(np.float32))
    assert rel_err(out32, g[f"c{ci}_T_f32"]) < TOL32


def test_transform_parameter_sampling(golden):
    g = golden("transform")
    rate, center = orc.sample_params_from_uniforms(g["params_draw0_rand"], g["params_draw1_rand"])
    assert np.array_equal(rate, g["params_rate"])
    assert np.array_equal(center, g["params_center"])
    rate, center = orc.sample_params_from_uniforms(g["module_draw1_rand"], g["module_draw2_rand"])
    assert rel_err(orc.scale_transform(g["module_x"], rate, center), g["module_T"]) < TOL32


def _toy_model_np(params, rate):
    w, c = params["param_w"], params["param_c"]

    def model(y):
        u = y
        if rate != 1:
            u = np.repeat(np.repeat(u, rate, axis=-2), rate, axis=-1)
        dt = y.dtype.type
        return dt(w[0]) * u + dt(w[1]) * np.roll(u, (1, 2), (-2, -1)) + dt(w[2]) * u * u + dt(c)

    return model


LOSS_CASES = [
    ("deblur_gauss2_proposed", "deblurring", "Gaussian_R2", 1, 6, {}),
    ("deblur_box3_proposed", "deblurring", "Box_R3", 1, 3, {}),
    ("deblur_gauss2_v1_proposed", "deblurring", "Gaussian_R2", 1, 6, {}),
    ("sr2_proposed", "sr", None, 2, 0, {}),
    ("sr4_proposed", "sr", None, 4, 0, {}),
    ("sr2_partial_proposed", "sr", None, 2, 2, {}),
    ("deblur_gauss2_proposed_alpha", "deblurring", "Gaussian_R2", 1, 6, {"alpha": 0.3}),
    ("cfg1_deblur_gauss2_proposed", "deblurring", "Gaussian_R2", 1, 6, {}),
    ("deblur_gauss2_normalT", "deblurring", "Gaussian_R2", 1, 6, {"kind": "normal"}),
    ("deblur_gauss2_normalT_aa", "deblurring", "Gaussian_R2", 1, 6, {"kind": "normal", "antialias": True}),
]


@pytest.mark.parametrize("case", LOSS_CASES, ids=[c[0] for c in LOSS_CASES])
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_proposed_loss(golden, case, tag):
    name, task, kname, rate, margin, kw = case
    if tag == "f64" and "v1" in name:
        pytest.skip("the reference's v1 Blur only runs in fp32")
    g = golden(f"loss_{name}_{tag}")
    kern = orc.named_kernel(kname) if kname else None
    phys = orc.OraclePhysics(task, kernel=kern, rate=rate, sigma=float(np.float32(5 / 255)))
    if kw.get("kind") == "normal":       # one scalar rate draw, then the noise
        draws = dict(b=None, u_rate=g["draw1_rand"], u_center=None, noise=g["draw2_randn_like"])
    else:
        draws = dict(b=None, u_rate=g["draw1_rand"], u_center=g["draw2_rand"], noise=g["draw3_randn_like"])
    y = g["y"]
    b0 = g["draw0_randn"] if "draw0_randn" in g else g["draw0_randn_like"]
    b = np.zeros_like(y)
    if margin:
        b[:, :, margin:-margin, margin:-margin] = b0
    else:
        b = b0
    draws["b"] = b
    out = orc.proposed_loss(phys, _toy_model_np(g, rate), y, draws, margin, **kw)
    tol = 1e-10 if tag == "f64" else 2e-5
    assert abs(out["loss"] - float(g["loss"])) <= tol * abs(float(g["loss"]))
    etol = TOL64 if tag == "f64" else 5e-6
    assert rel_err(out["x_net"], g["model_out0"]) < etol
    assert rel_err(out["x_net2"], g["model_out1"]) < etol
    assert rel_err(out["x3"], g["model_out2"]) < etol


@pytest.mark.parametrize("name,kw", [("deblur_gauss2_sure", {}), ("deblur_gauss2_sure_avgcst", {"averaged_cst": True}),
                                     ("deblur_gauss2_sure_nocrop", {"cropped_div": False})])
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_sure_loss(golden, name, kw, tag):
    g = golden(f"loss_{name}_{tag}")
    kern = orc.named_kernel("Gaussian_R2")
    phys = orc.OraclePhysics("deblurring", kernel=kern, sigma=float(np.float32(5 / 255)))
    model = _toy_model_np(g, 1)
    y, margin = g["y"], 6
    cropped = kw.get("cropped_div", True)
    if cropped:
        b = np.zeros_like(y)
        b[:, :, margin:-margin, margin:-margin] = g["draw0_randn"]
    else:
        b = g["draw0_randn_like"]
    y1 = phys.A(model(y))
    y2 = phys.A(model(y + b * y.dtype.type(1e-2)))
    loss, _, _ = orc.sure_loss(y1, y2, y, b, margin, margin if cropped else 0, 1e-2, (5 / 255) ** 2,
                               kw.get("averaged_cst"))
    tol = 1e-10 if tag == "f64" else 2e-5
    assert abs(loss - float(g["loss"])) <= tol * abs(float(g["loss"]))


def test_supervised_and_css(golden):
    for name, rate in (("deblur_gauss2_supervised", 1), ("sr2_css", 2)):
        for tag, tol in (("f64", 1e-12), ("f32", 2e-6)):
            g = golden(f"loss_{name}_{tag}")
            x_net = _toy_model_np(g, rate)(g["y"])
            assert abs(orc.mse(x_net, g["x"]) - float(g["loss"])) <= tol * float(g["loss"])


def test_noise_and_degrade(golden):
    g = golden("degrade")
    kern = orc.named_kernel("Gaussian_R2")
    y = orc.add_noise(orc.blur_circular(g["deblur_x"], kern), g["deblur_draw0_randn_like"], float(g["deblur_sigma"]))
    assert rel_err(y, g["deblur_y"]) < TOL32
    y = orc.add_noise(orc.down_aa(g["sr2_x"], 2), g["sr2_draw0_randn_like"], float(g["sr2_sigma"]))
    assert rel_err(y, g["sr2_y"]) < TOL32


@pytest.mark.parametrize("name,task,kname,rate", [("deblur_gauss2_r2r", "deblurring", "Gaussian_R2", 1), ("sr2_r2r", "sr", None, 2)])
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_r2r_loss(golden, name, task, kname, rate, tag):
    g = golden(f"loss_{name}_{tag}")
    phys = orc.OraclePhysics(task, kernel=orc.named_kernel(kname) if kname else None, rate=rate)
    draws = dict(pert=g["draw0_randn_like"], eps1=g["draw1_randn_like"], u_rate=g["draw2_rand"], u_center=g["draw3_rand"],
                 eps2=g["draw4_randn_like"])
    out = orc.r2r_ei_loss(phys, _toy_model_np(g, rate), g["y"], draws, eta=5 / 255)
    tol = 1e-10 if tag == "f64" else 2e-5
    assert abs(out["loss"] - float(g["loss"])) <= tol * abs(float(g["loss"]))
    etol = TOL64 if tag == "f64" else 5e-6
    for ours, ref in ((out["out0"], "model_out0"), (out["x1"], "model_out1"), (out["x3"], "model_out2")):
        assert rel_err(ours, g[ref]) < etol


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_scale_transform_vjp_through_no_stop_gradient_loss(golden, tag):
    """dL/dx_net of the proposed loss with stop_gradient off contains T^T: pins orc.scale_transform_vjp"""
    g = golden(f"loss_deblur_gauss2_nostopgrad_{tag}")
    dt = g["y"].dtype.type
    kern = orc.named_kernel("Gaussian_R2")
    phys = orc.OraclePhysics("deblurring", kernel=kern, sigma=float(np.float32(5 / 255)))
    fwd, _ = orc.pointwise_model(g["param_w"], g["param_c"], 1)
    y, m = g["y"], 6
    b = np.zeros_like(y)
    b[:, :, m:-m, m:-m] = g["draw0_randn"]
    draws = dict(b=b, u_rate=g["draw1_rand"], u_center=g["draw2_rand"], noise=g["draw3_randn_like"])
    out = orc.proposed_loss(phys, fwd, y, draws, m)
    assert abs(out["loss"] - float(g["loss"])) <= (1e-10 if tag == "f64" else 2e-5) * abs(float(g["loss"]))
    B, C, H, W = y.shape
    n_int = B * C * (H - 2 * m) * (W - 2 * m)
    mask = np.zeros_like(y)
    mask[:, :, m:-m, m:-m] = 1
    g_y2 = dt(2 * (5 / 255) ** 2 / (1e-2 * n_int)) * b * mask
    g_y1 = dt(2.0 / n_int) * (out["y1"] - y) * mask - g_y2
    g_x3 = dt(2.0 / out["x3"].size) * (out["x3"] - out["x2"])
    rate, center = orc.sample_params_from_uniforms(draws["u_rate"], draws["u_center"])
    g_x2 = phys.A_vjp(fwd.input_vjp(out["y_ei"], g_x3)) - g_x3          # through physics(x2) -> model, and the target
    g_xnet = phys.A_vjp(g_y1) + orc.scale_transform_vjp(g_x2, rate, center)
    tol = 1e-9 if tag == "f64" else 2e-5
    assert rel_err(g_xnet, g["model_out0_grad"]) < tol
    assert rel_err(g_x3, g["model_out2_grad"]) < tol


def test_oracle_antialiased_scale_transform_matches_reference(golden):
    """padded_downsampling_transform(antialiased=True) (reference src/transforms.py:44-83): pre-filter + grid_sample of the
    smaller image; equal rates only -- the reference's torch.stack raises for mixed rates, and so does the oracle"""
    g = golden("transform_aa")
    assert "stack expects each tensor to be equal size" in str(g["mixed_rates_error"])
    for ci in range(4):
        x, rate, center = g[f"c{ci}_x"], g[f"c{ci}_rate"], g[f"c{ci}_center"]
        assert rel_err(orc.scale_transform_antialiased(x, rate, center), g[f"c{ci}_T_f64"]) < 1e-11
        f = np.float32
        assert rel_err(orc.scale_transform_antialiased(x.astype(f), rate.astype(f), center.astype(f)), g[f"c{ci}_T_f32"]) < 5e-6
    with pytest.raises(RuntimeError, match="equal size"):
        orc.scale_transform_antialiased(np.zeros((2, 1, 16, 16)), np.array([0.75, 0.5]), np.zeros((2, 2)))


def test_oracle_rotate_matches_torchvision(golden):
    """deepinv Rotate = torchvision rotate(x, angle) with its defaults; fixtures made by torchvision itself.  Index work:
    identical pixels, except source coordinates on a rounding tie (bounded at 1e-4 of the pixels; 0 on these fixtures)"""
    g = golden("rotate")
    bad = tot = 0
    for i in range(4):
        for a in g["angles"]:
            ref, got = g[f"y{i}_a{a}"], orc.rotate_nearest(g[f"x{i}"], float(a))
            assert got.shape == ref.shape
            bad += int((ref != got).sum())
            tot += ref.size
    assert bad <= 1e-4 * tot, (bad, tot)
    x = g["x0"]
    assert np.array_equal(orc.rotate_nearest(x, 180.0), x[..., ::-1, ::-1])        # exact for the half turn


def test_oracle_resize_bicubic_matches_reference(golden):
    """normal_downsampling_transform (reference src/transforms.py:112-124), both rates, with and without antialiasing"""
    g = golden("normal_transform")
    for i in range(3):
        x = g[f"x{i}"]
        for rate in (0.75, 0.5):
            for aa in (0, 1):
                tag = f"{i}_r{int(rate * 100)}_aa{aa}"
                assert rel_err(orc.resize_bicubic(x, rate, bool(aa)), g[f"y64_{tag}"]) < 1e-11, tag
                # the reference's float32 run rounds its source coordinates / weights in fp32 (scale 4/3): 3e-6
                assert rel_err(orc.resize_bicubic(x.astype(np.float32), rate, bool(aa)), g[f"y32_{tag}"]) < 5e-6, tag


@pytest.mark.parametrize("name", ["deblur_g1", "deblur_box2_v1", "sr2_plain", "sr2_true"])
def test_a_dagger(golden, name):
    """A_dagger of the reference's physics objects (six CG iterations) against the oracle's conjugate gradient over the
    oracle's operators: the non-overcomplete branch (A A^T z = y, x = A^T z) with each kind of `A_adjoint`."""
    g = golden("dagger")
    if name.startswith("deblur"):
        kern = orc.named_kernel("Gaussian_R1" if name == "deblur_g1" else "Box_R2")
        A = lambda v: orc.blur_circular(v, kern)
        At = lambda v: orc.blur_circular(v, kern, adjoint=True)
    else:
        A = lambda v: orc.down_aa(v, 2)
        At = (lambda v: orc.up_bicubic(v, 2)) if name == "sr2_plain" else (lambda v: orc.down_aa_vjp(v, 2, (2 * v.shape[-2], 2 * v.shape[-1])))
    for tag, tol in (("f64", 1e-9), ("f32", 1e-5)):
        if f"{name}_y_{tag}" not in g:
            continue
        y = g[f"{name}_y_{tag}"]
        if name == "sr2_true" and tag == "f64":
            continue            # the reference's true adjoint evaluates in fp32 whatever the input
        rec = orc.a_dagger(A, At, y, max_iter=6, tol=1e-12)
        assert rec.dtype == y.dtype
        assert rel_err(rec, g[f"{name}_dagger_{tag}"]) < tol, (name, tag)
