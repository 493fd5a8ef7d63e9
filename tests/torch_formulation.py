"""PyTorch (library) formulation of the restoration CNN's operator hooks -- TEST INFRASTRUCTURE ONLY.

models/convolutional.py routes every non-GEMM operator through five module-level hooks that call the hand-written
kernels and raise SeiError for anything they do not take (CPU tensors, fp32 activations).  The structure tests check
the module tree against the reference's fixtures at fp32 accuracy on the CPU; for that they install the plain torch
formulation below (the same arithmetic the reference's src/models/convolutional.py issues: F.layer_norm over the
channels :21-30, depthwise F.conv2d :36-38, nn.GELU :41, rfft2 / fftshift / mask or zero-pad / irfft2 :54-133).
bench.py's `--impl reference --force-port` leg uses it too.  Nothing under scale-equivariant-imaging_b200/ imports it."""
from math import ceil

import torch
import torch.nn.functional as F


def layer_norm(rows, ln):
    return F.layer_norm(rows, ln.normalized_shape, ln.weight.to(rows.dtype), ln.bias.to(rows.dtype), ln.eps)


def dwconv7(xl, conv):
    x = xl.permute(0, 3, 1, 2)
    y = F.conv2d(x, conv.weight.to(x.dtype), None if conv.bias is None else conv.bias.to(x.dtype), padding=3, groups=x.shape[1])
    return y.permute(0, 2, 3, 1)


def gelu(x):
    return F.gelu(x)


def ideal_resample(x, kind, rate):
    dtype = x.dtype
    s = (x.shape[-2], x.shape[-1])
    X = torch.fft.fftshift(torch.fft.rfft2(x.float(), dim=(-2, -1)), dim=(-2, -1))
    if kind == "up":
        r = rate
        hs, ws = X.shape[-2], X.shape[-1]
        X2 = torch.zeros((X.shape[0], X.shape[1], hs * r, ws * r), device=X.device, dtype=X.dtype)
        mv, mh = (hs * (r - 1)) // 2, (ws * (r - 1)) // 2
        mt, mb = (mv + 1, mv) if hs % 2 == 1 else (mv, mv)
        ml, mr = (mh + 1, mh) if ws % 2 == 1 else (mh, mh)
        X2[:, :, mt:-mb, ml:-mr] = X
        return torch.fft.irfft2(X2, dim=(-2, -1), s=(s[0] * r, s[1] * r)).to(dtype)
    hcsh = ceil(X.shape[-2] / (2 * rate))
    hcsw = ceil(X.shape[-1] / (2 * rate))
    otf = torch.zeros_like(X)
    otf[:, :, hcsh:-hcsh, hcsw:-hcsw] = 1
    out = torch.fft.irfft2(otf * X, dim=(-2, -1), s=s)
    return out[:, :, ::rate, ::rate].to(dtype)


def install(mc, setattr_fn=setattr):
    """route models.convolutional's operator hooks and GEMMs through torch in fp32 (setattr_fn: monkeypatch.setattr in tests)"""
    setattr_fn(mc, "COMPUTE_DTYPE", torch.float32)
    setattr_fn(mc, "COMMUTE_DOWNSAMPLE", False)      # reference order: convolution, then the resampler
    setattr_fn(mc, "_gemm_tn", lambda a, b, bias, out_dtype: (a @ b.t() + (bias if bias is not None else 0)).to(out_dtype))
    setattr_fn(mc, "_gemm_atb", lambda a, b, out=None: (a.t() @ b).float() if out is None else out.add_((a.t() @ b).float()))
    setattr_fn(mc, "_op_layer_norm", layer_norm)
    setattr_fn(mc, "_op_dwconv7", dwconv7)
    setattr_fn(mc, "_op_gelu", gelu)
    setattr_fn(mc, "_op_ideal_resample", ideal_resample)
    setattr_fn(mc, "_op_conv3x3_small", lambda xl, conv: None)
    setattr_fn(mc, "_colsum", lambda gy: torch.sum(gy, 0, dtype=torch.float32))
