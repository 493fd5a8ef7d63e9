"""CPU stand-ins for the raw operators of sei_b200.ops, built on the oracle, for tests of the HOST logic only (draw order,
dispatch between fused and unfused paths, loss assembly, autograd wiring of the mirrors).  Installed with
`fake_ops.install(monkeypatch)`; the autograd Functions of sei_b200.ops stay the real ones and reach these through the
module's globals, so their forward / backward pairing is exercised too.  Test infrastructure: never imported by the package."""
import numpy as np
import torch

from oracle import oracle as orc


def _n(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _noisy(y, noise, sigma):
    return y if noise is None else orc.add_noise(y, _n(noise), float(sigma))


def blur_circular(x, kernel_host, adjoint=False, noise=None, sigma=0.0, path=0):
    return _t(_noisy(orc.blur_circular(_n(x), np.asarray(kernel_host), adjoint=bool(adjoint)), noise, sigma))


def blur_padded(x, filter_host, padding, transpose=False):
    f = np.asarray(filter_host, dtype=np.float32)
    return _t((orc.conv_transpose_v1 if transpose else orc.conv_v1)(_n(x), f, padding))


def down_aa(x, rate, noise=None, sigma=0.0, path=0):
    return _t(_noisy(orc.down_aa(_n(x), int(rate)), noise, sigma))


def down_aa_transpose(gy, rate, in_hw, path=0):
    return _t(orc.down_aa_vjp(_n(gy), int(rate), tuple(in_hw)))


def up_bicubic(y, rate):
    return _t(orc.up_bicubic(_n(y), int(rate)))


def resize_bicubic(x, scale_factor, antialias):
    return _t(orc.resize_bicubic(_n(x), float(scale_factor), bool(antialias)))


def rotate_nearest(x, angle):
    return _t(orc.rotate_nearest(_n(x), float(angle)))


def roll(x, shift_h, shift_w):
    return torch.roll(x.detach(), (int(shift_h), int(shift_w)), (-2, -1))


def scale_transform(x, rate, center, path=0):
    return _t(orc.scale_transform(_n(x), _n(rate), _n(center)))


def scale_transform_from(x_src, out_size, rate, center):
    return _t(orc.scale_transform_from(_n(x_src), int(out_size), _n(rate), _n(center)))


def scale_transform_backward(g, rate, center):
    return _t(orc.scale_transform_vjp(_n(g), _n(rate), _n(center)))


def ei_remeasure(x_net, rate, center, kernel_host, rate_sr, noise, sigma, use_workspace=True):
    x2 = orc.scale_transform(_n(x_net), _n(rate), _n(center))
    y = orc.blur_circular(x2, np.asarray(kernel_host)) if kernel_host is not None else orc.down_aa(x2, int(rate_sr))
    return _t(x2), _t(_noisy(y, noise, sigma))


def add_noise(y, noise, sigma):
    return _t(orc.add_noise(_n(y), _n(noise), float(sigma)))


def sure_perturb(y, draw, margin, tau):
    b = torch.zeros_like(y)
    if margin:
        b[..., margin:-margin, margin:-margin] = draw
    else:
        b.copy_(draw)
    return y.detach() + b * tau, b


def _interior(t, m):
    return t[..., m:-m, m:-m] if m else t


def mse(a, b):
    return ((a - b) ** 2).mean()


def mc_div(y1, y2, b, margin, tau):
    return _interior(b * (y2 - y1) / tau, margin).mean()


def sure_loss(y1, y2, y, b, margin_mse, margin_div, tau, sigma2, averaged_cst):
    """SureGaussianLoss.forward's arithmetic (reference src/losses/sure.py:60-76) in torch, so autograd differentiates it"""
    mse_v = (_interior(y1 - y, margin_mse) ** 2).mean()
    div_v = mc_div(y1, y2, b, margin_div, tau)
    loss = mse_v + 2 * sigma2 * div_v - (sigma2 if averaged_cst else sigma2 / y.shape[0])
    return loss, torch.stack([loss.detach(), mse_v.detach(), div_v.detach()])


NAMES = ["blur_circular", "blur_padded", "down_aa", "down_aa_transpose", "up_bicubic", "resize_bicubic", "rotate_nearest",
         "roll", "scale_transform", "scale_transform_from", "scale_transform_backward", "ei_remeasure", "add_noise",
         "sure_perturb", "mse", "mc_div", "sure_loss", "crop_batch"]


def crop_batch(x, tops, lefts, height, width):
    xn = _n(x)
    B, C, H, W = xn.shape
    out = np.zeros((B, C, int(height), int(width)), dtype=xn.dtype)
    for b in range(B):
        t, l = int(tops[b]), int(lefts[b])
        r1, c1 = min(H, t + height), min(W, l + width)
        out[b, :, : max(0, r1 - t), : max(0, c1 - l)] = xn[b, :, t:r1, l:c1]
    return _t(out)


def crop_batch(x, tops, lefts, height, width):
    xn = _n(x)
    B, C, H, W = xn.shape
    out = np.zeros((B, C, int(height), int(width)), dtype=xn.dtype)
    for b in range(B):
        t, l = int(tops[b]), int(lefts[b])
        r1, c1 = min(H, t + int(height)), min(W, l + int(width))
        out[b, :, : max(0, r1 - t), : max(0, c1 - l)] = xn[b, :, t:r1, l:c1]
    return _t(out)


def install(monkeypatch):
    from sei_b200 import ops
    for name in NAMES:
        assert hasattr(ops, name), name
        monkeypatch.setattr(ops, name, globals()[name])
