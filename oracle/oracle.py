"""numpy/ctypes front end of the CPU ORACLE (oracle/libsei_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product
package (scale-equivariant-imaging_b200/) never does and fails loudly without its CUDA
library instead of falling back to anything here.

Every function takes and returns C-contiguous numpy arrays; dtype float32 selects the
fp32 instantiation of the C restatement (mimics the reference run in fp32), float64 the
fp64 one.  The reference file:line each function follows is cited in
oracle/sei_oracle_impl.h next to its C body.

Parity pin: tests/test_oracle_golden.py compares every function with the fixtures in
tests/golden/, which were produced by running the reference itself
(tests/golden/make_golden.py).  The pieces of the path that live in the reference's
un-vendored dependency deepinv v0.2.0 (GaussianNoise, EILoss, SupLoss, mse) are restated
from its published behaviour and are pinned only through the reference's own call sites
(losses, physics factory) as executed with tests/golden/deepinv_shim: parity for those
is "unpinned" against upstream deepinv.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsei_oracle.so")
_lib = None

PAD_MODES = {"valid": 0, "circular": 1, "replicate": 2, "reflect": 3, "zero": 4}


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("sei_oracle.c", "sei_oracle_impl.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libsei_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        for sfx in ("f32", "f64"):
            getattr(_lib, f"orc_sure_loss_{sfx}").restype = C.c_double
            getattr(_lib, f"orc_mse_{sfx}").restype = C.c_double
    return _lib


def set_threads(n):
    """number of OpenMP threads the oracle's loops use (overrides OMP_NUM_THREADS); returns the effective count"""
    return int(lib().orc_set_threads(int(n)))


def _sfx(a):
    if a.dtype == np.float32:
        return "f32", C.c_float
    if a.dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"oracle supports float32/float64, got {a.dtype}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def named_kernel(name):
    """physics/kernels.py get_kernel -> (k, k) float64."""
    buf = np.zeros(64 * 64, dtype=np.float64)
    k = lib().orc_named_kernel(name.encode(), _p(buf))
    if k < 0:
        raise ValueError(f"Unsupported kernel: {name}")
    return buf[: k * k].reshape(k, k).copy()


def blur_circular(x, kernel, adjoint=False):
    """BlurV2.A / Blur(circular).A (adjoint=False) or their transpose (adjoint=True). x: (B,C,H,W)."""
    x = _c(x)
    sfx, _ = _sfx(x)
    h = _c(kernel, x.dtype).reshape(kernel.shape[-2], kernel.shape[-1])
    B, Cc, H, W = x.shape
    if H < h.shape[0] or W < h.shape[1]:
        raise ValueError("image smaller than the blur kernel")
    y = np.empty_like(x)
    getattr(lib(), f"orc_blur_circular_{sfx}")(_p(x), _p(y), C.c_long(B * Cc), H, W, _p(h), h.shape[0], h.shape[1], int(adjoint))
    return y


def conv_v1(x, filt, padding):
    x = _c(x)
    sfx, _ = _sfx(x)
    f = _c(filt, x.dtype).reshape(filt.shape[-2], filt.shape[-1])
    B, Cc, H, W = x.shape
    ho, wo = C.c_int(), C.c_int()
    mode = PAD_MODES[padding]
    getattr(lib(), f"orc_conv_v1_out_size_{sfx}")(H, W, f.shape[0], f.shape[1], mode, C.byref(ho), C.byref(wo))
    y = np.empty((B, Cc, ho.value, wo.value), dtype=x.dtype)
    getattr(lib(), f"orc_conv_v1_{sfx}")(_p(x), _p(y), C.c_long(B * Cc), H, W, _p(f), f.shape[0], f.shape[1], mode)
    return y


def conv_transpose_v1(y, filt, padding):
    y = _c(y)
    sfx, _ = _sfx(y)
    f = _c(filt, y.dtype).reshape(filt.shape[-2], filt.shape[-1])
    B, Cc, H, W = y.shape
    ho, wo = C.c_int(), C.c_int()
    mode = PAD_MODES[padding]
    getattr(lib(), f"orc_conv_transpose_v1_out_size_{sfx}")(H, W, f.shape[0], f.shape[1], mode, C.byref(ho), C.byref(wo))
    x = np.empty((B, Cc, ho.value, wo.value), dtype=y.dtype)
    getattr(lib(), f"orc_conv_transpose_v1_{sfx}")(_p(y), _p(x), C.c_long(B * Cc), H, W, _p(f), f.shape[0], f.shape[1], mode)
    return x


def down_aa(x, rate):
    """Downsampling.A: antialiased bicubic decimation by `rate`."""
    x = _c(x)
    sfx, _ = _sfx(x)
    B, Cc, H, W = x.shape
    fn = getattr(lib(), f"orc_down_out_size_{sfx}")
    y = np.empty((B, Cc, fn(H, rate), fn(W, rate)), dtype=x.dtype)
    getattr(lib(), f"orc_down_aa_{sfx}")(_p(x), _p(y), C.c_long(B * Cc), H, W, rate)
    return y


def down_aa_vjp(gy, rate, in_hw):
    """Transpose of down_aa (autograd backward of A / true adjoint). in_hw = (H, W) of x."""
    gy = _c(gy)
    sfx, _ = _sfx(gy)
    B, Cc = gy.shape[:2]
    H, W = in_hw
    gx = np.empty((B, Cc, H, W), dtype=gy.dtype)
    getattr(lib(), f"orc_down_aa_vjp_{sfx}")(_p(gy), _p(gx), C.c_long(B * Cc), H, W, rate)
    return gx


def up_bicubic(y, rate):
    """Downsampling.A_adjoint with true_adjoint=False: plain bicubic upsample."""
    y = _c(y)
    sfx, _ = _sfx(y)
    B, Cc, h, w = y.shape
    x = np.empty((B, Cc, h * rate, w * rate), dtype=y.dtype)
    getattr(lib(), f"orc_up_bicubic_{sfx}")(_p(y), _p(x), C.c_long(B * Cc), h, w, rate)
    return x


def _resize_axis_matrix(n_in, scale_factor, antialias):
    """dense (n_out, n_in) weight matrix of F.interpolate(mode='bicubic') along one axis, in float64.
    scale = 1 / scale_factor, n_out = floor(n_in * scale_factor)  (ATen upsample_bicubic2d / _upsample_bicubic2d_aa)."""
    import math
    n_out = int(math.floor(n_in * float(scale_factor)))
    scale = 1.0 / float(scale_factor)
    M = np.zeros((n_out, n_in), dtype=np.float64)
    for i in range(n_out):
        if not antialias:
            real = scale * (i + 0.5) - 0.5              # area_pixel_compute_source_index(cubic=True): no clamping
            i0 = math.floor(real)
            t = real - i0
            A = -0.75                                   # get_cubic_upsample_coefficients
            c = [((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A,
                 ((A + 2) * t - (A + 3)) * t * t + 1,
                 ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1,
                 ((A * (2 - t) - 5 * A) * (2 - t) + 8 * A) * (2 - t) - 4 * A]
            for k in range(4):
                M[i, min(max(i0 - 1 + k, 0), n_in - 1)] += c[k]
        else:
            support = 2.0 * scale if scale >= 1.0 else 2.0
            invscale = 1.0 / scale if scale >= 1.0 else 1.0
            center = scale * (i + 0.5)
            lo = max(int(center - support + 0.5), 0)
            hi = min(int(center + support + 0.5), n_in)
            a = -0.5

            def cubic(v):
                v = abs(v)
                if v < 1.0:
                    return ((a + 2) * v - (a + 3)) * v * v + 1
                if v < 2.0:
                    return (((v - 5) * v + 8) * v - 4) * a
                return 0.0

            w = np.array([cubic((j + lo - center + 0.5) * invscale) for j in range(hi - lo)])
            M[i, lo:hi] = w / w.sum()
    return M


def resize_bicubic(x, scale_factor, antialias):
    """normal_downsampling_transform (reference src/transforms.py:112-124): F.interpolate(x_i, scale_factor=rate,
    mode='bicubic', antialias=antialiased) for every image, restated with dense per-axis weight matrices."""
    x = np.asarray(x)
    My = _resize_axis_matrix(x.shape[-2], scale_factor, antialias)
    Mx = _resize_axis_matrix(x.shape[-1], scale_factor, antialias)
    return np.einsum("ih,bchw,jw->bcij", My, x.astype(np.float64), Mx).astype(x.dtype)


def rotate_nearest(x, angle):
    """deepinv.transform.Rotate (third-party, deepinv v0.2.0 transform/rotate.py; absent from /root/reference, used at
    src/losses/__init__.py:86-91) = torchvision.transforms.functional.rotate(x, angle) with its defaults, restated from
    torchvision 0.26's tensor path: _get_inverse_affine_matrix(center=[0,0], -angle) -> _gen_affine_grid (linspace base grid,
    theta^T / [w/2, h/2], bmm) -> F.grid_sample(mode="nearest", padding_mode="zeros", align_corners=False), all fp32.
    Pinned against torchvision itself (tests/golden/rotate.npz): equal up to ~1e-6 of the pixels, whose source coordinate
    sits on a rounding tie and follows the library's bmm summation order."""
    import math
    f = np.float32
    x = np.asarray(x)
    H, W = x.shape[-2:]
    rot = math.radians(-float(angle))
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    theta = np.array([d, -b, 0.0, -c, a, 0.0], dtype=f).reshape(2, 3)
    resc = (theta.T / np.array([0.5 * W, 0.5 * H], dtype=f)).astype(f)          # (3, 2)

    def linspace(start, end, steps):                                            # ATen: from both ends towards the middle
        start, end = f(start), f(end)
        step = f((end - start) / f(steps - 1)) if steps > 1 else f(0)
        i = np.arange(steps)
        lo = (start + step * i.astype(f)).astype(f)
        hi = (end - step * (steps - 1 - i).astype(f)).astype(f)
        return np.where(i < steps // 2, lo, hi).astype(f)

    bx = np.broadcast_to(linspace(-W * 0.5 + 0.5, W * 0.5 + 0.5 - 1, W)[None, :], (H, W))
    by = np.broadcast_to(linspace(-H * 0.5 + 0.5, H * 0.5 + 0.5 - 1, H)[:, None], (H, W))
    gx = ((bx * resc[0, 0]).astype(f) + (by * resc[1, 0]).astype(f)).astype(f) + resc[2, 0]
    gy = ((bx * resc[0, 1]).astype(f) + (by * resc[1, 1]).astype(f)).astype(f) + resc[2, 1]
    ix = np.rint((((gx + f(1)) * f(W)).astype(f) - f(1)) / f(2)).astype(np.int64)
    iy = np.rint((((gy + f(1)) * f(H)).astype(f) - f(1)) / f(2)).astype(np.int64)
    ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    out = np.zeros_like(x)
    out[..., ok] = x[..., iy[ok], ix[ok]]
    return out


def scale_grid(B, S, rate, center, dtype):
    rate = _c(rate, dtype).reshape(B)
    center = _c(center, dtype).reshape(B, 2)
    grid = np.empty((B, S, S, 2), dtype=dtype)
    sfx, _ = _sfx(grid)
    getattr(lib(), f"orc_scale_grid_{sfx}")(_p(grid), B, S, _p(rate), _p(center))
    return grid


def scale_transform(x, rate, center):
    """padded_downsampling_transform(x, rate, center, 'bicubic', 'reflection', antialiased=False)."""
    x = _c(x)
    sfx, _ = _sfx(x)
    B, Cc, S, S2 = x.shape
    if S != S2:
        raise ValueError("the scale transform is defined for square images only")
    rate = _c(rate, x.dtype).reshape(B)
    center = _c(center, x.dtype).reshape(B, 2)
    out = np.empty_like(x)
    getattr(lib(), f"orc_scale_transform_{sfx}")(_p(x), _p(out), B, Cc, S, _p(rate), _p(center))
    return out


def scale_transform_antialiased(x, rate, center):
    """padded_downsampling_transform(..., antialiased=True) (src/transforms.py:60-83): alias_free_interpolate (:44-57,
    F.interpolate(scale_factor=rate_i, antialias=True) per image + torch.stack, which needs equal rates) and then
    grid_sample of the smaller image with the grid of the original shape"""
    x = _c(x)
    sfx, _ = _sfx(x)
    B, Cc, S, S2 = x.shape
    if S != S2:
        raise ValueError("the scale transform is defined for square images only")
    rate = _c(rate, x.dtype).reshape(B)
    center = _c(center, x.dtype).reshape(B, 2)
    if len(set(float(r) for r in rate)) != 1:
        raise RuntimeError("stack expects each tensor to be equal size")
    return scale_transform_from(resize_bicubic(x, float(rate[0]), True), S, rate, center)


def scale_transform_from(x_src, S, rate, center):
    """grid_sample step of the scale transform on a source of another size: x_src (B, C, Ss, Ss) -> (B, C, S, S) with the
    grid of an S x S image (src/transforms.py:69-82 when the anti-aliasing pre-filter has shrunk the image)"""
    x_src = _c(x_src)
    sfx, _ = _sfx(x_src)
    B, Cc, Ss, Ss2 = x_src.shape
    if Ss != Ss2:
        raise ValueError("the scale transform is defined for square images only")
    rate = _c(rate, x_src.dtype).reshape(B)
    center = _c(center, x_src.dtype).reshape(B, 2)
    out = np.empty((B, Cc, S, S), dtype=x_src.dtype)
    getattr(lib(), f"orc_scale_transform_src_{sfx}")(_p(x_src), _p(out), B, Cc, Ss, S, _p(rate), _p(center))
    return out


def scale_transform_vjp(gout, rate, center):
    """transpose of scale_transform w.r.t. its image argument (autograd through grid_sample)"""
    gout = _c(gout)
    sfx, _ = _sfx(gout)
    B, Cc, S, _ = gout.shape
    rate = _c(rate, gout.dtype).reshape(B)
    center = _c(center, gout.dtype).reshape(B, 2)
    gx = np.empty_like(gout)
    getattr(lib(), f"orc_scale_transform_vjp_{sfx}")(_p(gout), _p(gx), B, Cc, S, _p(rate), _p(center))
    return gx


def sample_params_from_uniforms(u_rate, u_center, rates=(0.75, 0.5)):
    """sample_from + sample_downsampling_parameters (src/transforms.py:5-24) given the two
    uniform draws (shape (B,) then (B, 2)) the reference takes from torch.rand, in order."""
    dt = u_rate.dtype
    values = np.asarray(rates, dtype=dt)
    idx = np.floor(dt.type(len(rates)) * u_rate).astype(np.int32)
    rate = values[idx]
    center = (dt.type(2) * u_center - dt.type(1)).reshape(-1, 1, 1, 2)
    return rate, center


def add_noise(y, n, sigma):
    y = _c(y)
    sfx, ct = _sfx(y)
    n = _c(n, y.dtype)
    out = np.empty_like(y)
    getattr(lib(), f"orc_add_noise_{sfx}")(_p(y), _p(n), _p(out), C.c_long(y.size), ct(sigma))
    return out


def sure_loss(y1, y2, y, b, margin_mse, margin_div, tau, sigma2, averaged_cst):
    y1 = _c(y1)
    sfx, _ = _sfx(y1)
    y2, y, b = _c(y2, y1.dtype), _c(y, y1.dtype), _c(b, y1.dtype)
    B, Cc, H, W = y1.shape
    mse, div = C.c_double(), C.c_double()
    loss = getattr(lib(), f"orc_sure_loss_{sfx}")(
        _p(y1), _p(y2), _p(y), _p(b), B, Cc, H, W, int(margin_mse), int(margin_div),
        C.c_double(tau), C.c_double(sigma2), int(bool(averaged_cst)), C.byref(mse), C.byref(div))
    return loss, mse.value, div.value


def mse(a, b):
    a = _c(a)
    sfx, _ = _sfx(a)
    b = _c(b, a.dtype)
    return getattr(lib(), f"orc_mse_{sfx}")(_p(a), _p(b), C.c_long(a.size))


# ---------------------------------------------------------------------------------------
# Loss assembly (src/losses/__init__.py:67-142, src/losses/sure.py, deepinv EILoss) with the
# network supplied as a callable and every random tensor supplied by the caller.
# ---------------------------------------------------------------------------------------
class OraclePhysics:
    """A (and its transpose) of either task, in the oracle."""

    def __init__(self, task, kernel=None, rate=None, sigma=5 / 255):
        self.task, self.kernel, self.rate, self.sigma = task, kernel, rate, sigma

    def A(self, x):
        if self.task == "deblurring":
            return blur_circular(x, self.kernel)
        return down_aa(x, self.rate)

    def A_vjp(self, gy, in_hw=None):
        if self.task == "deblurring":
            return blur_circular(gy, self.kernel, adjoint=True)
        return down_aa_vjp(gy, self.rate, in_hw)


def conjugate_gradient(op, b, max_iter, tol):
    """deepinv.optim.utils.conjugate_gradient (v0.2.0; un-vendored dependency, restated in
    tests/golden/deepinv_shim/deepinv/physics/forward.py): CG from x0 = 0 in b's dtype, stop once |r| < tol."""
    x = np.zeros_like(b)
    r = b
    p = r
    rsold = (r * r).sum(dtype=b.dtype)
    for _ in range(int(max_iter)):
        Ap = op(p)
        alpha = rsold / (p * Ap).sum(dtype=b.dtype)
        x = x + p * alpha
        r = r + Ap * (-alpha)
        rsnew = (r * r).sum(dtype=b.dtype)
        if np.sqrt(rsnew) < tol:
            break
        p = r + p * (rsnew / rsold)
        rsold = rsnew
    return x


def a_dagger(A, At, y, max_iter=50, tol=1e-3):
    """deepinv LinearPhysics.A_dagger (call sites: reference demo/test.py:122, src/models/__init__.py:28): normal
    equations A^T A x = A^T y when A^T y is smaller than y, else A A^T z = y and x = A^T z.  `At` is the physics
    object's A_adjoint -- for the reference's default SR operator the plain bicubic upsample, not the transpose
    (src/physics/downsampling/__init__.py:21-35)."""
    Aty = At(y)
    if Aty.size < y.size:
        return conjugate_gradient(lambda v: At(A(v)), Aty, max_iter, tol)
    return At(conjugate_gradient(lambda v: A(At(v)), y, max_iter, tol))


def proposed_loss(physics, model, y, draws, margin, cropped_div=True, averaged_cst=None,
                  alpha=1.0, tau=1e-2, sure_sigma=5 / 255, kind="padded", antialias=False):
    """ProposedLoss.forward for transforms="Scaling_Transforms", stop_gradient=True.
    draws = dict(b=..., u_rate=..., u_center=..., noise=...) in the reference's draw order
    (SURVEY.md section 3.1).  Returns the loss and the intermediates the tests compare.
    kind="normal" (src/transforms.py:127-145): draws["u_rate"] is the single scalar draw, no centres."""
    # SureGaussianLoss gets sigma = noise_level/255 as a Python float (src/losses/__init__.py:104),
    # while the noise model holds it as a float32 Parameter (deepinv GaussianNoise)
    sigma2 = sure_sigma ** 2
    x_net = model(y)
    y1 = physics.A(x_net)
    b = draws["b"]
    x_net2 = model(y + b * y.dtype.type(tau))
    y2 = physics.A(x_net2)
    margin_div = margin if cropped_div else 0
    l_sure, mse_v, div_v = sure_loss(y1, y2, y, b, margin, margin_div, tau, sigma2, averaged_cst)
    if kind == "normal":
        rates = [0.75, 0.5]
        x2 = resize_bicubic(x_net, rates[int(np.floor(len(rates) * float(draws["u_rate"])))], antialias)
    else:
        rate, center = sample_params_from_uniforms(draws["u_rate"], draws["u_center"])
        x2 = scale_transform(x_net, rate, center)
    y_ei = add_noise(physics.A(x2), draws["noise"], physics.sigma)
    x3 = model(y_ei)
    l_ei = alpha * mse(x3, x2)
    return dict(loss=l_sure + l_ei, loss_sure=l_sure, loss_ei=l_ei, mse=mse_v, div=div_v,
                x_net=x_net, x_net2=x_net2, x2=x2, y_ei=y_ei, x3=x3, y1=y1, y2=y2)


def pointwise_model(w, c, rate=1):
    """The 4-parameter stand-in network of tests/toy_model.py in numpy, plus its parameter gradient."""

    def up(v):
        return np.repeat(np.repeat(v, rate, axis=-2), rate, axis=-1) if rate != 1 else v

    def fwd(v):
        u = up(v)
        dt = u.dtype.type
        return dt(w[0]) * u + dt(w[1]) * np.roll(u, (1, 2), (-2, -1)) + dt(w[2]) * u * u + dt(c)

    def param_grad(v, g):
        u = up(v).astype(np.float64)
        g = g.astype(np.float64)
        return np.array([(g * u).sum(), (g * np.roll(u, (1, 2), (-2, -1))).sum(), (g * u * u).sum()]), g.sum()

    def input_vjp(v, g):
        """d<g, fwd(v)>/dv (rate 1 only: nearest upsampling is not needed by the tests that use this)"""
        assert rate == 1
        dt = v.dtype.type
        return dt(w[0]) * g + dt(w[1]) * np.roll(g, (-1, -2), (-2, -1)) + dt(2) * dt(w[2]) * v * g

    fwd.input_vjp = input_vjp
    return fwd, param_grad


def proposed_step(physics, w, c, y, draws, margin, rate=1, alpha=1.0, tau=1e-2, sure_sigma=5 / 255):
    """One 'proposed' training step of the reference (loss forward AND backward to the network
    parameters; demo/train.py:258-266 with src/losses) for the stand-in network: the CPU
    restatement that bench.py times as cpu_baseline / --impl reference."""
    fwd, pgrad = pointwise_model(w, c, rate)
    out = proposed_loss(physics, fwd, y, draws, margin, alpha=alpha, tau=tau, sure_sigma=sure_sigma)
    dt = y.dtype.type
    B, C, H, W = y.shape
    n_int = B * C * (H - 2 * margin) * (W - 2 * margin)
    mask = np.zeros_like(y)
    mask[:, :, margin:H - margin, margin:W - margin] = 1
    k = dt(2.0 * sure_sigma ** 2 / (tau * n_int))
    g_y2 = k * draws["b"] * mask
    g_y1 = dt(2.0 / n_int) * (out["y1"] - y) * mask - g_y2
    in_hw = out["x_net"].shape[-2:]
    g_xnet = physics.A_vjp(g_y1, in_hw)
    g_xnet2 = physics.A_vjp(g_y2, in_hw)
    g_x3 = dt(2.0 * alpha / out["x3"].size) * (out["x3"] - out["x2"])
    gw = np.zeros(3)
    gc = 0.0
    for v, g in ((y, g_xnet), (y + draws["b"] * dt(tau), g_xnet2), (out["y_ei"], g_x3)):
        a, b_ = pgrad(v, g)
        gw += a
        gc += b_
    out.update(grad_w=gw, grad_c=gc, g_xnet=g_xnet, g_xnet2=g_xnet2, g_x3=g_x3)
    return out


def r2r_ei_loss(physics, model, y, draws, eta, alpha_r2r=0.5):
    """R2REILoss.forward (src/losses/r2r.py:26-57) with every random tensor supplied:
    draws = dict(pert=, eps1=, u_rate=, u_center=, eps2=) in the reference's draw order; eta = sigma."""
    dt = y.dtype.type
    pert = draws["pert"] * dt(eta)
    y_plus, y_minus = y + pert * dt(alpha_r2r), y - pert / dt(alpha_r2r)
    out0 = model(y_plus)
    l_r2r = mse(physics.A(out0), y_minus)
    x1 = model(y + dt(0.5) * dt(eta) * draws["eps1"])
    rate, center = sample_params_from_uniforms(draws["u_rate"], draws["u_center"])
    x2 = scale_transform(x1, rate, center)
    y2 = physics.A(x2)
    x3 = model(y2 + dt(1.5) * dt(eta) * draws["eps2"])
    l_ei = mse(x3, x2)
    return dict(loss=l_r2r + l_ei, out0=out0, x1=x1, x2=x2, x3=x3)
