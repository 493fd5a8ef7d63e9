"""Tensor-level wrappers over the C ABI (include/sei_b200.h) and the autograd Functions whose
backward passes call the hand-written transpose kernels.  torch is used for device memory and
streams only; every arithmetic step below runs in libsei_b200.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SeiError, check

PATH_AUTO, PATH_DIRECT, PATH_TILED = 0, 1, 2
PAD_MODES = {"valid": 0, "circular": 1, "replicate": 2, "reflect": 3, "zero": 4}


def _t(x, name):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not x.is_cuda:
        raise SeiError(f"{name} is on {x.device}: the sei_b200 operators run on CUDA (sm_100a) only; "
                       "there is no CPU fallback")
    if x.dtype != torch.float32:
        raise SeiError(f"{name} has dtype {x.dtype}: the sei_b200 operators compute in float32")
    return x.contiguous()


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def _stream(x):
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _host64(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a, C.c_void_p(a.ctypes.data)


def kernel_to_host(kernel):
    """(1,1,kh,kw) / (kh,kw) torch or numpy kernel -> contiguous float64 numpy (kh, kw)."""
    if isinstance(kernel, torch.Tensor):
        kernel = kernel.detach().to("cpu", torch.float64).numpy()
    k = np.asarray(kernel, dtype=np.float64)
    k = k.reshape(k.shape[-2], k.shape[-1])
    return np.ascontiguousarray(k)


_workspaces = {}


def _workspace(x):
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        n = int(_lib.load().sei_reduce_workspace_bytes())
        ws = torch.zeros(n, dtype=torch.uint8, device=x.device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------ raw operators
def blur_circular(x, kernel_host, adjoint=False, noise=None, sigma=0.0, path=PATH_AUTO):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    n = _t(noise, "noise") if noise is not None else None
    if n is not None and n.shape != x.shape:
        raise SeiError("noise must have the shape of the measurement")
    k, kp = _host64(kernel_host)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_blur_circular_f32(_ptr(x), _ptr(y), B * Cc, H, W, kp, k.shape[0], k.shape[1],
                                                int(bool(adjoint)), _ptr(n), float(sigma), path, _stream(x)))
    return y


def blur_padded(x, filter_host, padding, transpose=False):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    f, fp = _host64(filter_host)
    mode = PAD_MODES[padding]
    ho, wo = C.c_int(), C.c_int()
    lib = _lib.load()
    check(lib.sei_blur_padded_f32(None, None, B * Cc, H, W, fp, f.shape[0], f.shape[1], mode, int(transpose),
                                  C.byref(ho), C.byref(wo), None))
    out = torch.empty((B, Cc, ho.value, wo.value), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.sei_blur_padded_f32(_ptr(x), _ptr(out), B * Cc, H, W, fp, f.shape[0], f.shape[1], mode,
                                      int(transpose), C.byref(ho), C.byref(wo), _stream(x)))
    return out


def down_out_size(n, rate):
    return int(np.floor(n * (1.0 / rate)))


def down_aa(x, rate, noise=None, sigma=0.0, path=PATH_AUTO):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc, down_out_size(H, rate), down_out_size(W, rate)), dtype=x.dtype, device=x.device)
    n = _t(noise, "noise") if noise is not None else None
    if n is not None and n.shape != y.shape:
        raise SeiError("noise must have the shape of the measurement")
    with torch.cuda.device(x.device):
        check(_lib.load().sei_down_aa_f32(_ptr(x), _ptr(y), B * Cc, H, W, int(rate), _ptr(n), float(sigma), path,
                                          _stream(x)))
    return y


def down_aa_transpose(gy, rate, in_hw, path=PATH_AUTO):
    gy = _t(gy, "gy")
    B, Cc, Ho, Wo = gy.shape
    H, W = in_hw
    if (Ho, Wo) != (down_out_size(H, rate), down_out_size(W, rate)):
        raise SeiError(f"gradient of shape {tuple(gy.shape)} does not match an input of {H}x{W} at rate {rate}")
    gx = torch.empty((B, Cc, H, W), dtype=gy.dtype, device=gy.device)
    with torch.cuda.device(gy.device):
        check(_lib.load().sei_down_aa_transpose_f32(_ptr(gy), _ptr(gx), B * Cc, H, W, int(rate), path, _stream(gy)))
    return gx


def up_bicubic(y, rate):
    y = _t(y, "y")
    B, Cc, h, w = y.shape
    x = torch.empty((B, Cc, h * rate, w * rate), dtype=y.dtype, device=y.device)
    with torch.cuda.device(y.device):
        check(_lib.load().sei_up_bicubic_f32(_ptr(y), _ptr(x), B * Cc, h, w, int(rate), _stream(y)))
    return x


def resize_bicubic(x, scale_factor, antialias):
    """F.interpolate(x, scale_factor=scale_factor, mode='bicubic', antialias=antialias) (sei_resize_bicubic_f32), with the
    hand-written transpose as its backward (sei_resize_bicubic_backward_f32)"""
    if x.requires_grad and torch.is_grad_enabled():
        return _ResizeBicubic.apply(x, float(scale_factor), bool(antialias))
    return _resize_bicubic_raw(x, scale_factor, antialias)


def _resize_bicubic_raw(x, scale_factor, antialias):
    import math
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    Ho, Wo = int(math.floor(H * float(scale_factor))), int(math.floor(W * float(scale_factor)))
    if Ho < 1 or Wo < 1:
        raise SeiError(f"resize_bicubic: scale factor {scale_factor} leaves no pixels of a {H}x{W} image")
    y = torch.empty((B, Cc, Ho, Wo), dtype=x.dtype, device=x.device)
    s = 1.0 / float(scale_factor)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_resize_bicubic_f32(_ptr(x), _ptr(y), B * Cc, H, W, Ho, Wo, s, s, int(bool(antialias)), _stream(x)))
    return y


def rotate_rescaled_theta(angle, H, W):
    """The 3 x 2 fp32 matrix torchvision builds for rotate(x, angle): _get_inverse_affine_matrix(center=[0, 0], -angle,
    translate=[0, 0], scale=1, shear=[0, 0]) -> theta (2 x 3, fp32) -> theta^T / [0.5 W, 0.5 H]."""
    import math
    rot = math.radians(-float(angle))
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    theta = np.array([d, -b, 0.0, -c, a, 0.0], dtype=np.float32).reshape(2, 3)
    return np.ascontiguousarray(theta.T / np.array([0.5 * W, 0.5 * H], dtype=np.float32), dtype=np.float32)


def rotate_nearest(x, angle):
    """torchvision.transforms.functional.rotate(x, angle) with its defaults (nearest, same size, zero fill), the call
    deepinv's Rotate makes (sei_rotate_nearest_f32); backward: every output gradient returns to the pixel it was read
    from (sei_rotate_nearest_backward_f32)"""
    if x.requires_grad and torch.is_grad_enabled():
        return _RotateNearest.apply(x, float(angle))
    return _rotate_nearest_raw(x, angle, False)


def _rotate_nearest_raw(x, angle, backward):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    r = rotate_rescaled_theta(angle, H, W)
    fn = _lib.load().sei_rotate_nearest_backward_f32 if backward else _lib.load().sei_rotate_nearest_f32
    with torch.cuda.device(x.device):
        check(fn(_ptr(x), _ptr(y), B * Cc, H, W, C.c_void_p(r.ctypes.data), _stream(x)))
    return y


class _RotateNearest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, angle):
        ctx.angle = angle
        return _rotate_nearest_raw(x, angle, False)

    @staticmethod
    def backward(ctx, g):
        return _rotate_nearest_raw(g.contiguous(), ctx.angle, True), None


class _ResizeBicubic(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale_factor, antialias):
        ctx.args = (tuple(x.shape), scale_factor, antialias)
        return _resize_bicubic_raw(x, scale_factor, antialias)

    @staticmethod
    def backward(ctx, g):
        (B, Cc, H, W), scale_factor, antialias = ctx.args
        g = _t(g, "g")
        gx = torch.empty((B, Cc, H, W), dtype=g.dtype, device=g.device)
        s = 1.0 / float(scale_factor)
        with torch.cuda.device(g.device):
            check(_lib.load().sei_resize_bicubic_backward_f32(_ptr(g), _ptr(gx), B * Cc, H, W, g.shape[2], g.shape[3], s, s,
                                                              int(antialias), _stream(g)))
        return gx, None, None


def crop_batch(x, tops, lefts, height, width):
    """out[b, c] = x[b, c, tops[b] : tops[b] + height, lefts[b] : lefts[b] + width], zero outside the image
    (sei_crop_batch_f32: one launch for the whole batch, offsets in device memory); no autograd (dataset path)"""
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    tops = torch.as_tensor(tops, dtype=torch.int32).to(x.device).contiguous()
    lefts = torch.as_tensor(lefts, dtype=torch.int32).to(x.device).contiguous()
    if tops.numel() != B or lefts.numel() != B:
        raise SeiError("crop_batch: one (top, left) offset per image is required")
    out = torch.empty((B, Cc, int(height), int(width)), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_crop_batch_f32(_ptr(x), _ptr(out), B, Cc, H, W, int(height), int(width), _ptr(tops), _ptr(lefts),
                                             _stream(x)))
    return out


def scale_params(u_rate, u_center, rates):
    u_rate, u_center = _t(u_rate, "u_rate"), _t(u_center, "u_center")
    B = u_rate.numel()
    rate = torch.empty((B,), dtype=torch.float32, device=u_rate.device)
    center = torch.empty((B, 1, 1, 2), dtype=torch.float32, device=u_rate.device)
    r = np.ascontiguousarray(np.asarray(rates, dtype=np.float32))
    with torch.cuda.device(u_rate.device):
        check(_lib.load().sei_scale_params_f32(_ptr(u_rate), _ptr(u_center), B, C.c_void_p(r.ctypes.data), len(r),
                                               _ptr(rate), _ptr(center), _stream(u_rate)))
    return rate, center


def scale_transform(x, rate, center, path=PATH_AUTO):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    if H != W:
        raise SeiError("the scale transform is defined for square images only (like the reference's grid)")
    rate = _t(rate, "downsampling_rate").reshape(-1)
    center = _t(center, "center").reshape(-1)
    if rate.numel() != B or center.numel() != 2 * B:
        raise SeiError("downsampling_rate must have B entries and center B x 2")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_scale_transform_f32(_ptr(x), _ptr(out), B, Cc, H, _ptr(rate), _ptr(center), path,
                                                  _stream(x)))
    return out


def scale_transform_from(x_src, out_size, rate, center):
    """grid_sample step of the scale transform reading a source of another (square) size: x_src [B, C, Ssrc, Ssrc]
    -> [B, C, out_size, out_size] (sei_scale_transform_src_f32; the anti-aliased variant); backward by
    sei_scale_transform_src_backward_f32"""
    if x_src.requires_grad and torch.is_grad_enabled():
        return _ScaleTransformFrom.apply(x_src, int(out_size), rate, center)
    return _scale_transform_from_raw(x_src, out_size, rate, center)


class _ScaleTransformFrom(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_src, out_size, rate, center):
        ctx.save_for_backward(rate, center)
        ctx.src_shape = tuple(x_src.shape)
        return _scale_transform_from_raw(x_src, out_size, rate, center)

    @staticmethod
    def backward(ctx, g):
        rate, center = ctx.saved_tensors
        B, Cc, Ss, _ = ctx.src_shape
        g = _t(g, "g")
        rate_c, center_c = _t(rate, "downsampling_rate").reshape(-1), _t(center, "center").reshape(-1)
        gx = torch.empty(ctx.src_shape, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(_lib.load().sei_scale_transform_src_backward_f32(_ptr(g), _ptr(gx), B, Cc, Ss, g.shape[-1], _ptr(rate_c),
                                                                   _ptr(center_c), _stream(g)))
        return gx, None, None, None


def _scale_transform_from_raw(x_src, out_size, rate, center):
    x_src = _t(x_src, "x")
    B, Cc, H, W = x_src.shape
    if H != W:
        raise SeiError("the scale transform is defined for square images only (like the reference's grid)")
    rate = _t(rate, "downsampling_rate").reshape(-1)
    center = _t(center, "center").reshape(-1)
    if rate.numel() != B or center.numel() != 2 * B:
        raise SeiError("downsampling_rate must have B entries and center B x 2")
    out = torch.empty((B, Cc, out_size, out_size), dtype=x_src.dtype, device=x_src.device)
    with torch.cuda.device(x_src.device):
        check(_lib.load().sei_scale_transform_src_f32(_ptr(x_src), _ptr(out), B, Cc, H, out_size, _ptr(rate), _ptr(center),
                                                      _stream(x_src)))
    return out


_ei_workspaces = {}


def _ei_workspace(x, B, S):
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream, B, S)
    ws = _ei_workspaces.get(key)
    if ws is None:
        ws = torch.empty(int(_lib.load().sei_ei_workspace_bytes(B, S)), dtype=torch.uint8, device=x.device)
        _ei_workspaces[key] = ws
    return ws


def ei_remeasure(x_net, rate, center, kernel_host, rate_sr, noise, sigma, use_workspace=True):
    x_net = _t(x_net, "x_net")
    B, Cc, S, S2 = x_net.shape
    if S != S2:
        raise SeiError("the scale transform is defined for square images only")
    rate = _t(rate, "downsampling_rate").reshape(-1)
    center = _t(center, "center").reshape(-1)
    x2 = torch.empty_like(x_net)
    So = S if rate_sr == 1 else down_out_size(S, rate_sr)
    y = torch.empty((B, Cc, So, So), dtype=x_net.dtype, device=x_net.device)
    n = _t(noise, "noise") if noise is not None else None
    if kernel_host is not None:
        k, kp = _host64(kernel_host)
        kh, kw = k.shape
    else:
        kp, kh, kw = None, 0, 0
    ws = _ei_workspace(x_net, B, S) if use_workspace else None
    with torch.cuda.device(x_net.device):
        check(_lib.load().sei_ei_remeasure_f32(_ptr(x_net), _ptr(x2), _ptr(y), B, Cc, S, _ptr(rate), _ptr(center),
                                               kp, kh, kw, int(rate_sr), _ptr(n), float(sigma), _ptr(ws),
                                               _stream(x_net)))
    return x2, y


def roll(x, shift_h, shift_w):
    x = _t(x, "x")
    B, Cc, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_roll_f32(_ptr(x), _ptr(out), B * Cc, H, W, int(shift_h), int(shift_w), _stream(x)))
    return out


class _Roll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, shift_h, shift_w):
        ctx.shift = (shift_h, shift_w)
        return roll(x, shift_h, shift_w)

    @staticmethod
    def backward(ctx, g):
        return roll(g, -ctx.shift[0], -ctx.shift[1]), None, None


def add_noise(y, noise, sigma):
    y, noise = _t(y, "y"), _t(noise, "noise")
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        check(_lib.load().sei_add_noise_f32(_ptr(y), _ptr(noise), y.numel(), float(sigma), _ptr(out), _stream(y)))
    return out


def sure_perturb(y, draw, margin, tau):
    y, draw = _t(y, "y"), _t(draw, "draw")
    B, Cc, H, W = y.shape
    out, b = torch.empty_like(y), torch.empty_like(y)
    with torch.cuda.device(y.device):
        check(_lib.load().sei_sure_perturb_f32(_ptr(y), _ptr(draw), B, Cc, H, W, int(margin), float(tau), _ptr(out),
                                               _ptr(b), _stream(y)))
    return out, b


# tensor-core work issued through this module since import (2*M*N*K per product): bench.py reports the step's
# achieved TFLOP/s from it
_FLOPS = [0.0]


def flop_count():
    return _FLOPS[0]


def gemm_bf16_tn(a, b, bias=None, out_dtype=torch.bfloat16, tile_n=0):
    """D[M,N] = a[M,K] @ b[N,K]^T (+ bias[N]) on the tcgen05 tensor cores; a, b bf16 row-major (last dim
    contiguous, row pitch a multiple of 8 elements); fp32 accumulation; D bf16 or fp32."""
    for t, name in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise SeiError(f"gemm_bf16_tn: {name} must be a 2-D CUDA bf16 tensor with a contiguous last dimension")
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise SeiError(f"gemm_bf16_tn: inner dimensions differ ({K} vs {K2})")
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise SeiError("gemm_bf16_tn: out_dtype must be bfloat16 or float32")
    if bias is not None:
        bias = _t(bias, "bias")
    d = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn(_ptr(a), _ptr(b), _ptr(d), _ptr(bias), M, N, K, a.stride(0), b.stride(0), N,
                                           int(out_dtype == torch.float32), int(tile_n), _stream(a)))
    return d


def gemm_nn_supported(M, N):
    """shapes sei_gemm_bf16_nn takes (the CTA-pair kernel)"""
    return M >= 256 and N >= 256 and N % 8 == 0


def gemm_bf16_nn(a, bkn, mult=None):
    """D[M,N] (bf16) = a[M,K] @ bkn[K,N] (* mult[M,N]) with bkn read in place (MN-major operand of the CTA-pair kernel,
    sei_gemm_bf16_nn): the input gradient of a pointwise convolution straight from its weight matrix"""
    _check_2d_bf16("gemm_bf16_nn", a=a, bkn=bkn)
    M, K = a.shape
    K2, N = bkn.shape
    if K != K2 or (mult is not None and tuple(mult.shape) != (M, N)):
        raise SeiError("gemm_bf16_nn: shape mismatch")
    if mult is not None:
        _check_2d_bf16("gemm_bf16_nn", mult=mult)
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_nn(_ptr(a), _ptr(bkn), _ptr(mult), _ptr(d), M, N, K, a.stride(0), bkn.stride(0), N,
                                           0 if mult is None else mult.stride(0), _stream(a)))
    return d


def gemm_bf16_tn_rowscaled_bias(a, b, bias, row_scale):
    """D[M,N] (bf16) = a @ b^T + bias[n] * row_scale[m % len(row_scale)] (sei_gemm_bf16_tn_rowscaled_bias)"""
    _check_2d_bf16("gemm_bf16_tn_rowscaled_bias", a=a, b=b)
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise SeiError("gemm_bf16_tn_rowscaled_bias: inner dimensions differ")
    bias, row_scale = _t(bias, "bias"), _t(row_scale, "row_scale").reshape(-1)
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn_rowscaled_bias(_ptr(a), _ptr(b), _ptr(d), _ptr(bias), _ptr(row_scale),
                                                          row_scale.numel(), M, N, K, a.stride(0), b.stride(0), N, _stream(a)))
    return d


def gemm_bf16_tn_residual(a, b, bias, res, res_scale=1.0):
    """D[M,N] (bf16) = a[M,K] @ b[N,K]^T + bias[N] + res_scale * res[M,N]: a pointwise convolution added to a tensor of its
    output's shape in the GEMM epilogue (sei_gemm_bf16_tn_residual)"""
    for t, name in ((a, "a"), (b, "b"), (res, "res")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise SeiError(f"gemm_bf16_tn_residual: {name} must be a 2-D CUDA bf16 tensor with a contiguous last dimension")
    M, K = a.shape
    N, K2 = b.shape
    if K != K2 or tuple(res.shape) != (M, N):
        raise SeiError("gemm_bf16_tn_residual: shape mismatch")
    if bias is not None:
        bias = _t(bias, "bias")
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn_residual(_ptr(a), _ptr(b), _ptr(d), _ptr(bias), _ptr(res), float(res_scale), M, N, K,
                                                    a.stride(0), b.stride(0), N, res.stride(0), _stream(a)))
    return d


def _check_2d_bf16(fn, **tensors):
    for name, t in tensors.items():
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise SeiError(f"{fn}: {name} must be a 2-D CUDA bf16 tensor with a contiguous last dimension")


def gemm_bf16_tn_gelu_dual(a, b, bias):
    """(gelu(h), gelu'(h)) with h = a[M,K] @ b[N,K]^T + bias, both bf16 [M, N], written from the GEMM epilogue
    (sei_gemm_bf16_tn_gelu_dual): the pre-activation never reaches memory"""
    _check_2d_bf16("gemm_bf16_tn_gelu_dual", a=a, b=b)
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise SeiError("gemm_bf16_tn_gelu_dual: inner dimensions differ")
    if bias is not None:
        bias = _t(bias, "bias")
    act = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    der = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn_gelu_dual(_ptr(a), _ptr(b), _ptr(bias), _ptr(act), _ptr(der), M, N, K,
                                                     a.stride(0), b.stride(0), N, _stream(a)))
    return act, der


def gemm_bf16_tn_mul(a, b, mult):
    """D[M,N] (bf16) = (a[M,K] @ b[N,K]^T) * mult[M,N] with the product applied in the GEMM epilogue (sei_gemm_bf16_tn_mul)"""
    _check_2d_bf16("gemm_bf16_tn_mul", a=a, b=b, mult=mult)
    M, K = a.shape
    N, K2 = b.shape
    if K != K2 or tuple(mult.shape) != (M, N):
        raise SeiError("gemm_bf16_tn_mul: shape mismatch")
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn_mul(_ptr(a), _ptr(b), _ptr(mult), _ptr(d), M, N, K, a.stride(0), b.stride(0), N,
                                               mult.stride(0), _stream(a)))
    return d


def gemm_bf16_tn_gelu_bwd(a, b, h):
    """D[M,N] (bf16) = (a[M,K] @ b[N,K]^T) * gelu'(h[M,N]): dgrad of the layer after a GELU with the GELU backward in the
    GEMM epilogue (sei_gemm_bf16_tn_gelu_bwd)"""
    for t, name in ((a, "a"), (b, "b"), (h, "h")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise SeiError(f"gemm_bf16_tn_gelu_bwd: {name} must be a 2-D CUDA bf16 tensor with a contiguous last dimension")
    M, K = a.shape
    N, K2 = b.shape
    if K != K2 or tuple(h.shape) != (M, N):
        raise SeiError("gemm_bf16_tn_gelu_bwd: shape mismatch")
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _FLOPS[0] += 2.0 * M * N * K
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_tn_gelu_bwd(_ptr(a), _ptr(b), _ptr(d), _ptr(h), M, N, K, a.stride(0), b.stride(0), N,
                                                    h.stride(0), _stream(a)))
    return d


def gelu_raw(x):
    """gelu(x) without autograd bookkeeping (dense bf16 CUDA tensor)"""
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_gelu_bf16(_ptr(x), None, _ptr(y), x.numel(), _stream(x)))
    return y


def gemm_bf16_atb(a, b, out=None):
    """D[M,N] (fp32) = a[K,M]^T @ b[K,N] on the tcgen05 tensor cores, both operands read in place (MN-major).
    out: contiguous fp32 [M, N] tensor to ACCUMULATE into (D += a^T b) instead of allocating a result."""
    for t, name in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise SeiError(f"gemm_bf16_atb: {name} must be a 2-D CUDA bf16 tensor with a contiguous last dimension")
    K, M = a.shape
    K2, N = b.shape
    if K != K2:
        raise SeiError(f"gemm_bf16_atb: contraction dimensions differ ({K} vs {K2})")
    _FLOPS[0] += 2.0 * M * N * K
    if out is not None:
        if out.dtype != torch.float32 or tuple(out.shape) != (M, N) or not out.is_contiguous() or out.device != a.device:
            raise SeiError("gemm_bf16_atb: out must be a contiguous float32 [M, N] tensor on the operands' device")
        with torch.cuda.device(a.device):
            check(_lib.load().sei_gemm_bf16_atb_accumulate(_ptr(a), _ptr(b), _ptr(out), K, M, N, a.stride(0), b.stride(0),
                                                           _stream(a)))
        return out
    d = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.load().sei_gemm_bf16_atb(_ptr(a), _ptr(b), _ptr(d), K, M, N, a.stride(0), b.stride(0), _stream(a)))
    return d


def bgemm_tile_rows(M, Kpad):
    """rows of the operator kept resident per CTA by sei_bgemm_bf16 for an (M, Kpad) operator"""
    return int(_lib.load().sei_bgemm_tile_rows(int(M), int(Kpad)))


def bgemm_bf16(A, x, out, M, K, N, tile_rows, batches, b_inner, x_b, k_inner, x_k, d_b, m_inner, d_m):
    """out_b[M, N] = A[M, K] @ x_b[K, N] for every batch entry (include/sei_b200.h: sei_bgemm_bf16).
    A: zero-padded bf16 operator [rows, Kpad]; x_b / d_b: (outer, inner) batch strides; x_k / d_m: (outer, inner)
    row strides; all in elements."""
    for t, name in ((A, "A"), (x, "x"), (out, "out")):
        if not t.is_cuda or t.dtype != torch.bfloat16:
            raise SeiError(f"bgemm_bf16: {name} must be a CUDA bf16 tensor")
    _FLOPS[0] += 2.0 * int(M) * int(K) * int(N) * int(batches)
    with torch.cuda.device(x.device):
        check(_lib.load().sei_bgemm_bf16(_ptr(A), _ptr(x), _ptr(out), int(M), int(K), int(N), int(A.shape[1]),
                                         int(tile_rows), int(batches), int(b_inner), int(x_b[0]), int(x_b[1]),
                                         int(k_inner), int(x_k[0]), int(x_k[1]), int(d_b[0]), int(d_b[1]),
                                         int(m_inner), int(d_m[0]), int(d_m[1]), _stream(x)))
    return out


def ln_cl_supported(x_rows):
    """rows [T, C] of a channels-last activation that sei_ln_cl_* take: CUDA, bf16, contiguous, C % 8 == 0 and a
    channel count the parameter-gradient kernel can tile"""
    return (x_rows.is_cuda and x_rows.dtype == torch.bfloat16 and x_rows.dim() == 2 and x_rows.is_contiguous()
            and x_rows.shape[1] % 8 == 0 and _lib.load().sei_ln_cl_backward_workspace_bytes(int(x_rows.shape[1])) >= 0)


def ln_any_supported(x_rows):
    """rows the channel LayerNorm kernels take: the vectorised ones (ln_cl_supported), the one-thread-per-row ones (up to
    32 channels) or the one-warp-per-row ones (any other count)"""
    return ln_cl_supported(x_rows) or (x_rows.is_cuda and x_rows.dtype == torch.bfloat16 and x_rows.dim() == 2
                                       and x_rows.is_contiguous() and 1 <= x_rows.shape[1] <= 65536)


def padded_width(c, kind):
    """smallest channel count >= c that the vectorised kernels of `kind` ("colsum": column sums / bias pattern,
    "dwconv7") tile; zero-padded channels are exact for these per-channel operators"""
    lib = _lib.load()
    fn = lib.sei_ln_cl_backward_workspace_bytes if kind == "colsum" else lib.sei_dwconv7_workspace_bytes
    c8 = -(-int(c) // 8) * 8
    for cand in range(c8, c8 + 4096, 8):
        if fn(cand) > 0:
            return cand
    raise SeiError(f"no supported channel count at or above {c} for {kind}")


def ln_forward_raw(x, g32, b32, eps):
    """(y, mean, rstd, small) of the channel LayerNorm of rows [T, C] (no autograd bookkeeping); g32 / b32: fp32 [C]"""
    T, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(T, dtype=torch.float32, device=x.device)
    rstd = torch.empty(T, dtype=torch.float32, device=x.device)
    small = not ln_cl_supported(x)
    fwd = _lib.load().sei_ln_small_forward_bf16 if small else _lib.load().sei_ln_cl_forward_bf16
    with torch.cuda.device(x.device):
        check(fwd(_ptr(x), _ptr(g32), _ptr(b32), _ptr(y), _ptr(mean), _ptr(rstd), T, Cc, float(eps), _stream(x)))
    return y, mean, rstd, small


def ln_backward_raw(gy, x, mean, rstd, g32, small):
    """(dx, dgamma, dbeta) of the channel LayerNorm; gy, x: bf16 rows [T, C]"""
    T, Cc = x.shape
    lib = _lib.load()
    dx = torch.empty_like(x)
    dg = torch.empty(Cc, dtype=torch.float32, device=x.device)
    db = torch.empty(Cc, dtype=torch.float32, device=x.device)
    ws_bytes = lib.sei_ln_small_workspace_bytes(Cc) if small else lib.sei_ln_cl_backward_workspace_bytes(Cc)
    ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=x.device)
    bwd = lib.sei_ln_small_backward_bf16 if small else lib.sei_ln_cl_backward_bf16
    with torch.cuda.device(x.device):
        check(bwd(_ptr(gy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(g32), _ptr(dx), _ptr(dg), _ptr(db), _ptr(ws), T, Cc,
                  _stream(x)))
    return dx, dg, db


class _LayerNormCL(torch.autograd.Function):
    """y = LayerNorm_C(x) on rows [T, C] (bf16), fp32 affine parameters; backward by the hand-written kernels"""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y, mean, rstd, ctx.small = ln_forward_raw(x, g32, b32, eps)
        ctx.save_for_backward(x, mean, rstd, g32)
        ctx.param_dtypes = (gamma.dtype, beta.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, g32 = ctx.saved_tensors
        dx, dg, db = ln_backward_raw(gy.contiguous(), x, mean, rstd, g32, ctx.small)
        return dx, dg.to(ctx.param_dtypes[0]), db.to(ctx.param_dtypes[1]), None


def colsum_bf16(x_rows):
    """fp32 column sums of a bf16 [T, C] matrix (sei_colsum_bf16); rows must satisfy ln_cl_supported"""
    T, Cc = x_rows.shape
    lib = _lib.load()
    out = torch.empty(Cc, dtype=torch.float32, device=x_rows.device)
    ws = torch.empty(int(lib.sei_ln_cl_backward_workspace_bytes(Cc)), dtype=torch.uint8, device=x_rows.device)
    with torch.cuda.device(x_rows.device):
        check(lib.sei_colsum_bf16(_ptr(x_rows), _ptr(out), _ptr(ws), T, Cc, _stream(x_rows)))
    return out


def igemm_weight_chunks(w, cinp, nout):
    """w [N, C, 3, 3] -> the chunk form sei_conv3x3_igemm_bf16 reads: bf16 [NCHP][nout][8], chunk = tap * (cinp / 8) + block,
    element [chunk][n][j] = w[n][8 * block + j][ky][kx] (zero where N < nout, C < cinp or the chunk count is padded to even)"""
    N, Cc = int(w.shape[0]), int(w.shape[1])
    cch = cinp // 8
    nchp = (9 * cch + 1) & ~1
    wp = torch.zeros((nout, cinp, 3, 3), dtype=torch.float32, device=w.device)
    wp[:N, :Cc] = w.detach().float()
    t = wp.permute(2, 3, 1, 0).reshape(9, cch, 8, nout).permute(0, 1, 3, 2).reshape(9 * cch, nout, 8)
    out = torch.zeros((nchp, nout, 8), dtype=torch.bfloat16, device=w.device)
    out[: 9 * cch] = t.to(torch.bfloat16)
    return out.contiguous()


def conv3x3_igemm(x_bhwc, wg, bias, out_stride, out_valid):
    """3x3 'same' convolution of a channels-last bf16 tensor (8 or 32 channels) as an implicit GEMM on tcgen05
    (sei_conv3x3_igemm_bf16); wg from igemm_weight_chunks; returns bf16 [B, H, W, out_stride]"""
    if not x_bhwc.is_cuda or x_bhwc.dtype != torch.bfloat16 or x_bhwc.dim() != 4 or not x_bhwc.is_contiguous():
        raise SeiError("conv3x3_igemm: x must be a contiguous CUDA bf16 [B, H, W, C] tensor")
    B, H, W, Cin = x_bhwc.shape
    if bias is not None:
        bias = _t(bias, "bias")
    out = torch.empty((B, H, W, out_stride), dtype=torch.bfloat16, device=x_bhwc.device)
    nout = 32 if Cin == 8 else 16
    _FLOPS[0] += 2.0 * B * H * W * nout * 9 * Cin
    with torch.cuda.device(x_bhwc.device):
        check(_lib.load().sei_conv3x3_igemm_bf16(_ptr(x_bhwc), _ptr(wg), _ptr(bias), _ptr(out), B, H, W, Cin, out_stride,
                                                 out_valid, _stream(x_bhwc)))
    return out


def conv3x3_small_wgrad(gy4, x, cout):
    """(gw [cout, Cin, 3, 3], gb [cout]) fp32 of a 3x3 'same' convolution with <= 4 output channels: gy4 bf16 [B, H, W, 4],
    x bf16 [B, H, W, Cin] (sei_conv3x3_small_backward_bf16 without the input gradient)"""
    B, H, W, Cin = x.shape
    lib = _lib.load()
    gw = torch.empty((cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    gb = torch.empty(cout, dtype=torch.float32, device=x.device)
    ws = torch.empty(int(lib.sei_conv3x3_small_workspace_bytes(Cin, cout)), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.sei_conv3x3_small_backward_bf16(_ptr(gy4), _ptr(x), _ptr(gw), None, _ptr(gw), _ptr(gb), _ptr(ws),
                                                  B, H, W, Cin, cout, _stream(x)))
    return gw, gb


def conv3x3_small_supported(x_bhwc, out_channels):
    return (x_bhwc.is_cuda and x_bhwc.dtype == torch.bfloat16 and x_bhwc.dim() == 4 and x_bhwc.is_contiguous()
            and x_bhwc.shape[3] in (8, 16, 32, 64) and 1 <= out_channels <= 4)


class _Conv3x3Small(torch.autograd.Function):
    """x [B, H, W, Cin] bf16 (channels last) -> y [B, H, W, 4] bf16; weight [Cout, Cin, 3, 3], bias [Cout] fp32"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, H, W, Cin = x.shape
        Cout = weight.shape[0]
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty((B, H, W, 4), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            check(_lib.load().sei_conv3x3_small_forward_bf16(_ptr(x), _ptr(w32), _ptr(b32), _ptr(y), B, H, W, Cin, Cout,
                                                             _stream(x)))
        ctx.save_for_backward(x, w32)
        ctx.has_bias = bias is not None
        ctx.dtypes = (weight.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w32 = ctx.saved_tensors
        B, H, W, Cin = x.shape
        Cout = w32.shape[0]
        gy = gy.contiguous()
        lib = _lib.load()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w32)
        gb = torch.empty(Cout, dtype=torch.float32, device=x.device)
        ws = torch.empty(int(lib.sei_conv3x3_small_workspace_bytes(Cin, Cout)), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.sei_conv3x3_small_backward_bf16(_ptr(gy), _ptr(x), _ptr(w32), _ptr(gx), _ptr(gw), _ptr(gb), _ptr(ws),
                                                      B, H, W, Cin, Cout, _stream(x)))
        return gx, gw.to(ctx.dtypes[0]), (gb.to(ctx.dtypes[1]) if ctx.has_bias else None)


def conv3x3_small(x_bhwc, weight, bias):
    return _Conv3x3Small.apply(x_bhwc, weight, bias)


def dwconv7_supported(x_bhwc):
    return (x_bhwc.is_cuda and x_bhwc.dtype == torch.bfloat16 and x_bhwc.dim() == 4 and x_bhwc.is_contiguous()
            and x_bhwc.shape[3] % 8 == 0 and _lib.load().sei_dwconv7_workspace_bytes(int(x_bhwc.shape[3])) > 0)


def _dwconv7_raw(x, wt, bias, res=None, res_scale=1.0):
    """depthwise 7x7 of x [B, H, W, C] with taps wt [49, C] (+ bias) (+ res_scale * res, added in the kernel's store)"""
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        if res is None:
            check(_lib.load().sei_dwconv7_cl_bf16(_ptr(x), _ptr(wt), _ptr(bias), _ptr(y), B, H, W, Cc, _stream(x)))
        else:
            if res.shape != x.shape or res.dtype != x.dtype or not res.is_contiguous():
                raise SeiError("dwconv7: the residual must be a contiguous tensor of the input's shape and dtype")
            check(_lib.load().sei_dwconv7_cl_residual_bf16(_ptr(x), _ptr(wt), _ptr(bias), _ptr(res), float(res_scale), _ptr(y),
                                                           B, H, W, Cc, _stream(x)))
    return y


def dwconv7_wgrad_raw(gy, x):
    """(gw [C, 49], gb [C]) fp32 of the depthwise 7x7 convolution; gy, x: bf16 [B, H, W, C]"""
    B, H, W, Cc = x.shape
    lib = _lib.load()
    gw = torch.empty((Cc, 49), dtype=torch.float32, device=x.device)
    gb = torch.empty(Cc, dtype=torch.float32, device=x.device)
    ws = torch.empty(int(lib.sei_dwconv7_workspace_bytes(Cc)), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.sei_dwconv7_wgrad_cl_bf16(_ptr(gy), _ptr(x), _ptr(gw), _ptr(gb), _ptr(ws), B, H, W, Cc, _stream(x)))
    return gw, gb


def gelu_bwd_colsum(h_rows, gy_rows):
    """(gy * gelu'(h), fp32 column sums of that product) in one pass over bf16 rows [T, C] (sei_gelu_bwd_colsum_bf16)"""
    T, Cc = h_rows.shape
    lib = _lib.load()
    gx = torch.empty_like(h_rows)
    gb = torch.empty(Cc, dtype=torch.float32, device=h_rows.device)
    ws = torch.empty(int(lib.sei_ln_cl_backward_workspace_bytes(Cc)), dtype=torch.uint8, device=h_rows.device)
    with torch.cuda.device(h_rows.device):
        check(lib.sei_gelu_bwd_colsum_bf16(_ptr(h_rows), _ptr(gy_rows), _ptr(gx), _ptr(gb), _ptr(ws), T, Cc, _stream(h_rows)))
    return gx, gb


class _DwConv7(torch.autograd.Function):
    """x [B, H, W, C] bf16 (channels last), weight [C, 1, 7, 7], bias [C] (fp32 parameters)"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        Cc = x.shape[3]
        w32 = weight.detach().float().reshape(Cc, 49)
        b32 = None if bias is None else bias.detach().float().contiguous()
        ctx.save_for_backward(x, w32)
        ctx.has_bias = bias is not None
        ctx.dtypes = (weight.dtype, None if bias is None else bias.dtype)
        ctx.wshape = weight.shape
        return _dwconv7_raw(x, w32.t().contiguous(), b32)

    @staticmethod
    def backward(ctx, gy):
        x, w32 = ctx.saved_tensors
        B, H, W, Cc = x.shape
        gy = gy.contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _dwconv7_raw(gy, w32.flip(1).t().contiguous(), None)      # the same convolution with the taps flipped
        gw, gb = dwconv7_wgrad_raw(gy, x)
        return gx, gw.view(ctx.wshape).to(ctx.dtypes[0]), (gb.to(ctx.dtypes[1]) if ctx.has_bias else None)


def dwconv7(x_bhwc, weight, bias):
    return _DwConv7.apply(x_bhwc, weight, bias)


def gelu_supported(x):
    """dense (any permutation of a contiguous block) CUDA bf16 tensor with a multiple of 8 elements"""
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.numel() % 8 == 0
            and (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)))


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = torch.empty_like(x)                 # preserves x's (dense) strides: element order in memory is the same
        with torch.cuda.device(x.device):
            check(_lib.load().sei_gelu_bf16(_ptr(x), None, _ptr(y), x.numel(), _stream(x)))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        if gy.stride() != x.stride():
            gy = torch.empty_like(x).copy_(gy)
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(_lib.load().sei_gelu_bf16(_ptr(x), _ptr(gy), _ptr(gx), x.numel(), _stream(x)))
        return gx


def gelu(x):
    return _Gelu.apply(x)


class _BiasPattern(torch.autograd.Function):
    """rows [T, C] (bf16) + pat[t % period] * bias[c]; the rows are updated in place (they are a fresh GEMM output)"""

    @staticmethod
    def forward(ctx, rows, pat, bias):
        T, Cc = rows.shape
        b32 = bias.detach().float().contiguous()
        with torch.cuda.device(rows.device):
            check(_lib.load().sei_bias_pattern_add_bf16(_ptr(rows), _ptr(pat), _ptr(b32), T, Cc, pat.numel(), _stream(rows)))
        ctx.mark_dirty(rows)
        ctx.save_for_backward(pat)
        ctx.bias_dtype = bias.dtype
        return rows

    @staticmethod
    def backward(ctx, gy):
        (pat,) = ctx.saved_tensors
        gy = gy.contiguous()
        T, Cc = gy.shape
        lib = _lib.load()
        gb = torch.empty(Cc, dtype=torch.float32, device=gy.device)
        ws = torch.empty(int(lib.sei_ln_cl_backward_workspace_bytes(Cc)), dtype=torch.uint8, device=gy.device)
        with torch.cuda.device(gy.device):
            check(lib.sei_bias_pattern_grad_bf16(_ptr(gy), _ptr(pat), _ptr(gb), _ptr(ws), T, Cc, pat.numel(), _stream(gy)))
        return gy, None, gb.to(ctx.bias_dtype)


def bias_pattern_grad(gy_rows, pat):
    """fp32 [C]: sum_t pat[t % period] * gy[t, c] -- the bias gradient behind a resampler (sei_bias_pattern_grad_bf16)"""
    T, Cc = gy_rows.shape
    lib = _lib.load()
    pat = _t(pat, "pat").reshape(-1)
    gb = torch.empty(Cc, dtype=torch.float32, device=gy_rows.device)
    ws = torch.empty(int(lib.sei_ln_cl_backward_workspace_bytes(Cc)), dtype=torch.uint8, device=gy_rows.device)
    with torch.cuda.device(gy_rows.device):
        check(lib.sei_bias_pattern_grad_bf16(_ptr(gy_rows), _ptr(pat), _ptr(gb), _ptr(ws), T, Cc, pat.numel(), _stream(gy_rows)))
    return gb


def bias_pattern_add(rows, pat, bias):
    return _BiasPattern.apply(rows, pat, bias)


def layer_norm_cl(x_rows, gamma, beta, eps):
    return _LayerNormCL.apply(x_rows, gamma, beta, eps)


# ------------------------------------------------------------------------------ autograd
class _BlurCircular(torch.autograd.Function):
    """y = A x (adjoint=False) or A^T x; backward applies the other one (hand-written transpose)."""

    @staticmethod
    def forward(ctx, x, kernel_host, adjoint, path):
        ctx.kernel_host, ctx.adjoint, ctx.path = kernel_host, adjoint, path
        return blur_circular(x, kernel_host, adjoint=adjoint, path=path)

    @staticmethod
    def backward(ctx, g):
        return blur_circular(g, ctx.kernel_host, adjoint=not ctx.adjoint, path=ctx.path), None, None, None


class _BlurCircularNoise(torch.autograd.Function):
    """y = A x + sigma * noise in one kernel (physics(x) of the EI branch)."""

    @staticmethod
    def forward(ctx, x, kernel_host, noise, sigma):
        ctx.kernel_host = kernel_host
        return blur_circular(x, kernel_host, noise=noise, sigma=sigma)

    @staticmethod
    def backward(ctx, g):
        return blur_circular(g, ctx.kernel_host, adjoint=True), None, None, None


class _BlurPadded(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, filter_host, padding, transpose):
        ctx.filter_host, ctx.padding, ctx.transpose = filter_host, padding, transpose
        return blur_padded(x, filter_host, padding, transpose)

    @staticmethod
    def backward(ctx, g):
        if ctx.padding == "zero":
            raise NotImplementedError("backward through conv_transpose(padding='zero') is not supported")
        return blur_padded(g, ctx.filter_host, ctx.padding, not ctx.transpose), None, None, None


class _DownAA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rate, noise, sigma, path):
        ctx.rate, ctx.in_hw, ctx.path = rate, tuple(x.shape[-2:]), path
        return down_aa(x, rate, noise=noise, sigma=sigma, path=path)

    @staticmethod
    def backward(ctx, g):
        return down_aa_transpose(g, ctx.rate, ctx.in_hw, path=ctx.path), None, None, None, None


class _DownAATranspose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, rate, in_hw, path):
        ctx.rate, ctx.path = rate, path
        return down_aa_transpose(gy, rate, in_hw, path=path)

    @staticmethod
    def backward(ctx, g):
        return down_aa(g, ctx.rate, path=ctx.path), None, None, None


class _AddNoise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, noise, sigma):
        return add_noise(y, noise, sigma)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def scale_transform_backward(g, rate, center):
    g = _t(g, "grad")
    B, Cc, S, _ = g.shape
    rate, center = _t(rate, "downsampling_rate").reshape(-1), _t(center, "center").reshape(-1)
    gx = torch.empty_like(g)
    with torch.cuda.device(g.device):
        check(_lib.load().sei_scale_transform_backward_f32(_ptr(g), _ptr(gx), B, Cc, S, _ptr(rate), _ptr(center), _stream(g)))
    return gx


class _ScaleTransform(torch.autograd.Function):
    """x -> T(x); backward (only reached with --no-ProposedLoss__stop_gradient) is the scatter kernel."""

    @staticmethod
    def forward(ctx, x, rate, center, path):
        ctx.save_for_backward(rate, center)
        return scale_transform(x, rate, center, path=path)

    @staticmethod
    def backward(ctx, g):
        rate, center = ctx.saved_tensors
        return scale_transform_backward(g, rate, center), None, None, None


class _Mse(torch.autograd.Function):
    """nn.MSELoss(): mean((a-b)^2); backward = one elementwise kernel."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _t(a, "input"), _t(b, "target")
        if a.shape != b.shape:
            raise SeiError(f"mse: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
        out = torch.empty((), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            check(_lib.load().sei_mse_f32(_ptr(a), _ptr(b), a.numel(), _ptr(out), _ptr(_workspace(a)), _stream(a)))
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous().float()
        ga = torch.empty_like(a)
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(a.device):
            check(_lib.load().sei_mse_backward_f32(_ptr(a), _ptr(b), a.numel(), _ptr(g), _ptr(ga), _ptr(gb), _stream(a)))
        return ga, gb


class _SureLoss(torch.autograd.Function):
    """SureGaussianLoss given y1 = A(x_net), y2 = A(model(y + tau b)) -> scalar; see sei_sure_loss_f32."""

    @staticmethod
    def forward(ctx, y1, y2, y, b, margin_mse, margin_div, tau, sigma2, averaged_cst):
        y1, y2, y, b = _t(y1, "y1"), _t(y2, "y2"), _t(y, "y"), _t(b, "b")
        B, Cc, H, W = y.shape
        out = torch.empty((3,), dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            check(_lib.load().sei_sure_loss_f32(_ptr(y1), _ptr(y2), _ptr(y), _ptr(b), B, Cc, H, W, int(margin_mse),
                                                int(margin_div), float(tau), float(sigma2), int(bool(averaged_cst)),
                                                _ptr(out), _ptr(_workspace(y)), _stream(y)))
        ctx.save_for_backward(y1, y, b)
        ctx.cfg = (int(margin_mse), int(margin_div), float(tau), float(sigma2))
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g, _g_aux):
        y1, y, b = ctx.saved_tensors
        B, Cc, H, W = y.shape
        mm, md, tau, sigma2 = ctx.cfg
        g = g.contiguous().float()
        g1, g2 = torch.empty_like(y1), torch.empty_like(y1)
        with torch.cuda.device(y.device):
            check(_lib.load().sei_sure_loss_backward_f32(_ptr(y1), _ptr(y), _ptr(b), B, Cc, H, W, mm, md, tau, sigma2,
                                                         _ptr(g), _ptr(g1), _ptr(g2), _stream(y)))
        return g1, g2, None, None, None, None, None, None, None


class _McDiv(torch.autograd.Function):
    """mean over the interior of b*(y2-y1)/tau (sure.py mc_div), differentiable in y1 and y2.
    Reuses the SURE kernels: the residual term vanishes when y := y1, and sigma2 = 1/2 turns the
    2*sigma2 factor into 1."""

    @staticmethod
    def forward(ctx, y1, y2, b, margin, tau):
        y1, y2, b = _t(y1, "y1"), _t(y2, "y2"), _t(b, "b")
        B, Cc, H, W = y1.shape
        out = torch.empty((3,), dtype=torch.float32, device=y1.device)
        with torch.cuda.device(y1.device):
            check(_lib.load().sei_sure_loss_f32(_ptr(y1), _ptr(y2), _ptr(y1), _ptr(b), B, Cc, H, W, int(margin),
                                                int(margin), float(tau), 0.5, 1, _ptr(out), _ptr(_workspace(y1)),
                                                _stream(y1)))
        ctx.save_for_backward(y1, b)
        ctx.cfg = (int(margin), float(tau))
        return out[2].clone()

    @staticmethod
    def backward(ctx, g):
        y1, b = ctx.saved_tensors
        B, Cc, H, W = y1.shape
        margin, tau = ctx.cfg
        g = g.contiguous().float()
        g1, g2 = torch.empty_like(y1), torch.empty_like(y1)
        with torch.cuda.device(y1.device):
            check(_lib.load().sei_sure_loss_backward_f32(_ptr(y1), _ptr(y1), _ptr(b), B, Cc, H, W, margin, margin, tau,
                                                         0.5, _ptr(g), _ptr(g1), _ptr(g2), _stream(y1)))
        return g1, g2, None, None, None


def mc_div(y1, y2, b, margin, tau):
    return _McDiv.apply(y1, y2, b, margin, tau)


def mse(a, b):
    return _Mse.apply(a, b)


def sure_loss(y1, y2, y, b, margin_mse, margin_div, tau, sigma2, averaged_cst):
    loss, aux = _SureLoss.apply(y1, y2, y, b, margin_mse, margin_div, tau, sigma2, averaged_cst)
    return loss, aux
