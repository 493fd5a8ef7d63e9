"""The restoration CNN (reference: src/models/convolutional.py) with its dense contractions on the
tcgen05 tensor cores.

Same module tree and parameter names as the reference's ConvolutionalModel, so its state dicts load
unchanged.  What differs is where the arithmetic runs:
  * every 1x1 convolution (ConvBlock.conv2 / conv3, Downsample.conv, Upsample.seq[2]) is the tcgen05 GEMM
    (sei_gemm_bf16_tn and its fused-epilogue forms: bf16 operands, fp32 accumulation in TMEM) on channels-last
    activations: pixels x C_in  @  (C_out x C_in)^T.  The input gradient of the deep layers reads the weight matrix in
    place (sei_gemm_bf16_nn), the weight gradient reads both operands in place (sei_gemm_bf16_atb) and is added straight
    into the parameter's gradient buffer;
  * the two 3x3 edge convolutions (UNet.in_conv / out_conv) are implicit GEMMs on tcgen05 (sei_conv3x3_igemm_bf16:
    nine TMA windows with zero fill as the K-chunks of the A operand, no unfolded copy);
  * a ConvBlock is ONE autograd node (_ConvBlockFn): residual in conv3's epilogue, gelu / gelu' from conv2's epilogue at
    the deep levels, the GELU backward as a multiplier epilogue of conv3's input gradient, the incoming gradient added in
    the store of the depthwise input-gradient kernel; UNet's inner-residual and skip additions ride in GEMM epilogues too;
  * activations travel through the network in bf16, channels-last; parameters stay fp32 masters whose bf16 copies are
    rewritten by the optimizer kernel in the pass that updates them;
  * the FFT "ideal" resamplers (including the reference's quirks: fftshift applied to the half-spectrum axis,
    ifftshift results discarded) are applied as the explicit linear operators they are, by two batched
    tensor-core products per call (models/resample.py, csrc/bgemm.cu) -- no cuFFT, no layout copies;
  * depthwise 7x7 (TMA-staged halo tiles, csrc/dwconv_tile.cu), channel LayerNorm, GELU and the reductions behind the bias /
    gamma / beta gradients are hand-written kernels as well (csrc/cnn_elem.cu).

There is no library fallback: a CPU tensor, a dtype other than bf16 activations or a missing libsei_b200.so raises
SeiError.  Channel counts the vectorised kernels do not take (3-channel layers of the SR / no-in-out-conv variants) are
zero-padded to a multiple of 8 around the same kernels.  The unit tests that check the module tree against the
reference at fp32 accuracy on the CPU install their own PyTorch formulation of the five operator hooks below
(tests/torch_formulation.py); nothing in this package does.
"""
import torch
import torch.nn.functional as F
from torch.nn import Conv2d, GELU, LayerNorm as BaseLayerNorm, Module, ModuleList, Sequential

from sei_b200 import ops
from sei_b200._lib import SeiError
from sei_b200.optim import shadow_epoch
from . import resample

CL = torch.channels_last
# activation / GEMM operand dtype.  bf16 is the only dtype the tcgen05 kernel takes; the unit tests set this to
# float32 together with a patched _gemm_tn to check the network's structure against the reference at fp32 accuracy.
COMPUTE_DTYPE = torch.bfloat16
COMMUTE_DOWNSAMPLE = True      # Downsample: resample first, then the pointwise convolution on a quarter of the pixels
# Weight gradients of the pointwise convolutions are ADDED straight into an existing contiguous fp32 `param.grad` by the
# GEMM epilogue (no temporary, no separate accumulation pass over 645 M parameters); the autograd node then returns
# None for that weight.  That is what `loss.backward()` does anyway, but it is a side effect outside autograd:
# `torch.autograd.grad(loss, weights)`, tensor hooks and `backward(inputs=...)` do not see such a gradient.  Callers
# that need those semantics switch the fast path off: `set_inplace_weight_gradients(False)` (or SEI_NO_WGRAD_ACC=1).
_NO_WGRAD_ACC = __import__("os").environ.get("SEI_NO_WGRAD_ACC", "0") == "1"
# Recompute gelu / gelu' of a ConvBlock in its backward pass from the C-wide LayerNorm output instead of keeping the two
# 4C-wide tensors between the passes (SEI_RECOMPUTE_MLP=1, or set_mlp_recompute(True)): one more conv2 GEMM per block and
# pass (measured: +13 % step time at 256 x 256 batch 32 for 38 instead of 64 GB; +1 % at 512 x 512 batch 16 for 64 instead of 116 GiB) -- for the
# configurations that do not fit otherwise (SR x4 beyond batch 2, 512 x 512 at batch 16 uses 116 GiB without it).
_RECOMPUTE_MLP = __import__("os").environ.get("SEI_RECOMPUTE_MLP", "0") == "1"


def set_inplace_weight_gradients(enabled):
    """see the note above _NO_WGRAD_ACC"""
    global _NO_WGRAD_ACC
    _NO_WGRAD_ACC = not enabled


def set_mlp_recompute(enabled):
    """see the note above _RECOMPUTE_MLP"""
    global _RECOMPUTE_MLP
    _RECOMPUTE_MLP = bool(enabled)
# gelu'(h) in the epilogue of conv3's input-gradient GEMM (sei_gemm_bf16_tn_gelu_bwd).  Measured on B200: it removes the
# 9 ms GELU-backward pass but the four epilogue warps then spend longer on the erf arithmetic than the tensor cores
# on the tile (CTA-pair GEMMs 76 -> 93 ms, N = 128 GEMMs 6 -> 13 ms per step): off by default.
_GELU_FUSION = __import__("os").environ.get("SEI_GELU_FUSION", "0") == "1"
# 3x3 in / out convolutions as implicit GEMMs on tcgen05 (SEI_IGEMM_CONV=0: unfold + GEMM / direct kernels of round 1)
_IGEMM_CONV = __import__("os").environ.get("SEI_IGEMM_CONV", "1") == "1"
# ConvBlock as one autograd node with its additions fused into the neighbouring kernels (SEI_CONVBLOCK_NODE=0: the
# op-by-op path, for A/B measurements)
_CONVBLOCK_NODE = __import__("os").environ.get("SEI_CONVBLOCK_NODE", "1") == "1"
# GELU and its derivative written from conv2's GEMM epilogue, GELU backward as a multiplier in conv3's input-gradient
# epilogue (SEI_GELU_EPILOGUE=0: separate GELU kernels, for A/B measurements)
# -- from this channel count on (0 = never).  Measured per level (benchmarks/mlp_bench.py, profiles/r02_mlp_bench*.md):
# the erf arithmetic of 268 M elements costs the same at every level, but only the deep levels' tiles spend long enough
# in the tensor core for eight epilogue warps to hide it; at the shallow levels the separate, full-occupancy GELU
# kernels are faster.
_GELU_EPILOGUE_MIN_C = int(__import__("os").environ.get("SEI_GELU_EPILOGUE_MIN_C", "512"))


def _gemm_tn(a, b, bias, out_dtype):
    """D = a @ b^T (+ bias) on the tcgen05 tensor cores (include/sei_b200.h: sei_gemm_bf16_tn)"""
    return ops.gemm_bf16_tn(a, b, bias, out_dtype=out_dtype)


def _gemm_atb(a, b, out=None):
    """D (fp32) = a^T @ b with both operands read in place (sei_gemm_bf16_atb, MN-major UMMA operands); with `out`
    the product is accumulated into it"""
    return ops.gemm_bf16_atb(a, b, out=out)


# input gradients read the weight matrix in place (sei_gemm_bf16_nn) where the CTA-pair kernel takes the shape; SEI_DGRAD_NN=0:
# the transposed bf16 copies of round 1 (kept for the small layers either way)
_DGRAD_NN = __import__("os").environ.get("SEI_DGRAD_NN", "1") == "1"


def _dgrad(gy, w_bf16, wt_getter, mult=None):
    """gx = gy @ W (* mult): W = w_bf16 [C_out, C_in] read in place by the CTA-pair kernel when the shape allows, else through
    the transposed copy (wt_getter)"""
    gy = _pad_k(gy)
    if (_DGRAD_NN and COMPUTE_DTYPE == torch.bfloat16 and gy.is_cuda and w_bf16.shape[0] == gy.shape[1]
            and ops.gemm_nn_supported(gy.shape[0], w_bf16.shape[1])):
        return ops.gemm_bf16_nn(gy, w_bf16, mult)
    if mult is not None:
        return ops.gemm_bf16_tn_mul(gy, wt_getter(), mult)
    return _gemm_tn(gy, wt_getter(), None, COMPUTE_DTYPE)


def _require_device_rows(t, what):
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise SeiError(f"{what}: got a {t.dtype} tensor on {t.device}; the restoration CNN runs on CUDA (sm_100a) in bf16 "
                       "only -- there is no CPU / library fallback")


def _pad_channels(xl, to=None):
    """(..., C) -> (..., C') zero-padded to `to` channels (default: the next multiple of 8); a copy only when C' != C"""
    c = xl.shape[-1]
    to = to or -(-c // 8) * 8
    return xl if to == c else F.pad(xl, (0, to - c))


def _op_layer_norm(rows, ln):
    """channel LayerNorm of rows [T, C] (ln_fwd_kernel / ln_small kernels, csrc/cnn_elem.cu)"""
    _require_device_rows(rows, "LayerNorm")
    return ops.layer_norm_cl(rows.contiguous(), ln.weight, ln.bias, ln.eps)


def _op_dwconv7(xl, conv):
    """depthwise 7x7, padding 3, on a channels-last (B, H, W, C) tensor (dwconv7_kernel)"""
    _require_device_rows(xl, "depthwise 7x7 convolution")
    c = xl.shape[-1]
    cp = ops.padded_width(c, "dwconv7")            # channels are independent: zero-padded ones are exact and dropped
    if cp == c:
        return ops.dwconv7(xl.contiguous(), conv.weight, conv.bias)
    w = F.pad(conv.weight, (0, 0, 0, 0, 0, 0, 0, cp - c))
    b = None if conv.bias is None else F.pad(conv.bias, (0, cp - c))
    return ops.dwconv7(_pad_channels(xl, cp).contiguous(), w, b)[..., :c]


def _op_gelu(x):
    """erf GELU of a dense bf16 tensor (gelu_fwd_kernel / gelu_bwd_kernel)"""
    _require_device_rows(x, "GELU")
    if ops.gelu_supported(x):
        return ops.gelu(x)
    flat = x.contiguous(memory_format=CL).permute(0, 2, 3, 1).reshape(-1)
    n = flat.numel()
    out = ops.gelu(F.pad(flat, (0, (-n) % 8)))[:n]
    return out.view(x.shape[0], x.shape[2], x.shape[3], x.shape[1]).permute(0, 3, 1, 2)


def _op_ideal_resample(x, kind, rate):
    """IdealUpsample / IdealDownsample as batched tensor-core operator products (models/resample.py)"""
    _require_device_rows(x, f"ideal {kind}sampler")
    if resample.supported(x):
        return resample.ideal_resample(x, kind, rate)
    c = x.shape[1]
    xl = _pad_channels(x.contiguous(memory_format=CL).permute(0, 2, 3, 1)).contiguous().permute(0, 3, 1, 2)
    return resample.ideal_resample(xl, kind, rate)[:, :c]


def _op_conv3x3_small(xl, conv):
    """the output layer (hidden -> <= 4 channels, 3x3 'same'): direct kernels, or None when the shape is not theirs"""
    if ops.conv3x3_small_supported(xl, conv.out_channels):
        return ops.conv3x3_small(xl, conv.weight, conv.bias)
    return None


def _op_bias_pattern(out, rows, pat, bias):
    """out[b, c, i, j] += pat[i, j] * bias[c] on the rows view of a channels-last GEMM output (bias_pattern_add_kernel)"""
    _require_device_rows(rows, "Downsample bias")
    c = rows.shape[1]
    cp = ops.padded_width(c, "colsum")               # widths of the no-in/out-conv variant: pad the columns
    if cp != c:
        rows, bias = _pad_channels(rows, cp).contiguous(), F.pad(bias, (0, cp - c))
    rows = ops.bias_pattern_add(rows, pat.reshape(-1), bias)[:, :c]
    return rows.reshape(out.shape[0], out.shape[2], out.shape[3], c).permute(0, 3, 1, 2)


class _GemmTN(torch.autograd.Function):
    """out[T, N] = x[T, K] @ w[N, K]^T + bias[N]; all three passes on sei_gemm_bf16_tn."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_bf16, wt_getter, param=None, res=None, row_scale=None):
        ctx.save_for_backward(x, w_bf16)
        ctx.has_bias = bias is not None
        ctx.wt_getter = wt_getter
        ctx.param = param                      # nn.Parameter whose .grad may receive the weight gradient in place
        ctx.row_scale = row_scale if bias is not None else None
        if res is not None:                    # out = x w^T + bias + res, the addition in the GEMM epilogue
            return ops.gemm_bf16_tn_residual(x, w_bf16, bias, res, 1.0)
        if ctx.row_scale is not None:          # out = x w^T + bias[n] * row_scale[m % period] (bias behind a resampler)
            return ops.gemm_bf16_tn_rowscaled_bias(x, w_bf16, bias, row_scale)
        return _gemm_tn(x, w_bf16, bias, COMPUTE_DTYPE)

    @staticmethod
    def backward(ctx, gy):
        x, w_bf16 = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _dgrad(gy, w_bf16, ctx.wt_getter)                                    # dgrad: gy[T,N] @ w[N,K]
        if ctx.needs_input_grad[1]:
            if gy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0:
                g = None if (ctx.param is None or _NO_WGRAD_ACC) else ctx.param.grad
                if (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.is_cuda
                        and g.numel() == gy.shape[1] * x.shape[1] and COMPUTE_DTYPE == torch.bfloat16):
                    # accumulate straight into the parameter's gradient buffer (allocated by a previous backward /
                    # zero_grad(set_to_none=False) / the data-parallel flat buckets): no temporary, no separate
                    # accumulation pass over the 645 M parameters for each of the three network passes of a step
                    _gemm_atb(gy, x, out=g.view(gy.shape[1], x.shape[1]))
                    gw = None
                else:
                    gw = _gemm_atb(gy, x)                                            # wgrad: gy[T,N]^T @ x[T,K], in place
            elif x.shape[1] % 8 == 0:                         # 3-channel output edge: pad gy's columns, not transposes
                gw = _gemm_atb(_pad_k(gy), x)[: gy.shape[1]]
            else:
                gw = _gemm_tn(_pad_k(gy.t()), _pad_k(x.t()), None, torch.float32)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # one pass, fp32 accumulation, fixed order (csrc/cnn_elem.cu); library reduction for odd channel counts
            gb = _colsum(gy) if ctx.row_scale is None else ops.bias_pattern_grad(gy, ctx.row_scale)
        return (gx, gw, gb, None, None, None, (gy if len(ctx.needs_input_grad) > 6 and ctx.needs_input_grad[6] else None),
                None)


class _GeluGemmTN(torch.autograd.Function):
    """out[T, N] = gelu(h[T, K]) @ w[N, K]^T + bias: the GELU and the pointwise convolution after it (ConvBlock.gelu,
    ConvBlock.conv3) as one autograd node, so that the backward pass can apply gelu'(h) in the epilogue of the
    input-gradient GEMM (sei_gemm_bf16_tn_gelu_bwd) instead of a separate pass over the 4C-wide gradient."""

    @staticmethod
    def forward(ctx, h, weight, bias, w_bf16, wt_getter, param=None):
        a = ops.gelu_raw(h)
        ctx.save_for_backward(h, a, w_bf16)
        ctx.has_bias = bias is not None
        ctx.wt_getter = wt_getter
        ctx.param = param
        return _gemm_tn(a, w_bf16, bias, COMPUTE_DTYPE)

    @staticmethod
    def backward(ctx, gy):
        h, a, w_bf16 = ctx.saved_tensors
        gy = gy.contiguous()
        gh = gw = gb = None
        if ctx.needs_input_grad[0]:
            gh = ops.gemm_bf16_tn_gelu_bwd(gy, ctx.wt_getter(), h)                    # (gy @ w) * gelu'(h)
        if ctx.needs_input_grad[1]:
            g = None if (ctx.param is None or _NO_WGRAD_ACC) else ctx.param.grad
            if (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.is_cuda
                    and g.numel() == gy.shape[1] * a.shape[1]):
                _gemm_atb(gy, a, out=g.view(gy.shape[1], a.shape[1]))
            else:
                gw = _gemm_atb(gy, a)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _colsum(gy)
        return gh, gw, gb, None, None, None


def _colsum(gy):
    """fp32 column sums of a bf16 [T, N] matrix (bias gradient): colsum_kernel, one pass, fixed order; columns padded to
    a multiple of 8 for the 3-channel edge layers"""
    n = gy.shape[1]
    _require_device_rows(gy, "bias gradient")
    return ops.colsum_bf16(_pad_channels(gy, ops.padded_width(n, "colsum")).contiguous())[:n]


def _pad_k(a):
    """contiguous copy of a 2-D bf16 matrix with its last (contraction) dimension padded to a multiple of 8"""
    a = a.contiguous()
    k = a.shape[1]
    return a if k % 8 == 0 else F.pad(a, (0, 8 - k % 8))



def _wgrad_into(param, gy, x):
    """dL/dW = gy^T x of a pointwise convolution: accumulated straight into param.grad when that buffer exists (returns
    None), otherwise returned as a new fp32 matrix"""
    g = None if _NO_WGRAD_ACC else param.grad
    if (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.is_cuda
            and g.numel() == gy.shape[1] * x.shape[1]):
        _gemm_atb(gy, x, out=g.view(gy.shape[1], x.shape[1]))
        return None
    return _gemm_atb(gy, x).view(param.shape)


class _ConvBlockFn(torch.autograd.Function):
    """A whole ConvBlock (reference src/models/convolutional.py:33-51) as ONE autograd node:

        out = res_scale * x + conv3(gelu(conv2(LayerNorm(dwconv7(x)))))

    (res_scale = 1 is the block itself; 2 folds UNet's inner residual `x + xb` when the block is alone in its sequence).
    The residual is added in the epilogue of conv3's GEMM (sei_gemm_bf16_tn_residual); in the backward pass the
    incoming gradient is added in the store of the depthwise input-gradient convolution (sei_dwconv7_cl_residual_bf16)
    and conv2's bias gradient is summed inside the GELU-backward pass (sei_gelu_bwd_colsum_bf16).  As separate autograd
    nodes these were three extra passes over C-wide tensors and one over the 4C-wide gradient per block and direction
    (library additions: 6 ms of a 233 ms step; the extra column sum: 2.4 ms)."""

    @staticmethod
    def forward(ctx, xl, res_scale, block, dw_w, dw_b, ln_g, ln_b, w2, b2, w3, b3):
        B, H, W, C = xl.shape
        T = B * H * W
        dw32 = dw_w.detach().float().reshape(C, 49)
        dwb32 = None if dw_b is None else dw_b.detach().float().contiguous()
        g32, be32 = ln_g.detach().float().contiguous(), ln_b.detach().float().contiguous()
        _, w2_bf = block.conv2._weight_matrix()
        _, w3_bf = block.conv3._weight_matrix()
        t1 = ops._dwconv7_raw(xl, dw32.t().contiguous(), dwb32)
        t2, mean, rstd, small = ops.ln_forward_raw(t1.view(T, C), g32, be32, block.ln.ln.eps)
        gelu_epilogue = 0 < _GELU_EPILOGUE_MIN_C <= C
        if gelu_epilogue:
            # conv2 + GELU in one kernel: gelu(h) and gelu'(h) leave the GEMM epilogue, h itself is never stored
            a, h = ops.gemm_bf16_tn_gelu_dual(t2, w2_bf, b2)          # (`h` holds gelu'(h) on this path)
        else:
            h = _gemm_tn(t2, w2_bf, b2, COMPUTE_DTYPE)
            a = ops.gelu_raw(h)
        out = ops.gemm_bf16_tn_residual(a, w3_bf, b3, xl.view(T, C), res_scale)
        ctx.gelu_epilogue = gelu_epilogue
        ctx.recompute = _RECOMPUTE_MLP
        if ctx.recompute:                      # keep only C-wide tensors; (a, h) are rebuilt from t2 in backward()
            ctx.b2 = None if b2 is None else b2.detach()
            h = a = t2.new_empty(0)
        ctx.save_for_backward(xl, t1, mean, rstd, t2, h, a, dw32, g32)
        ctx.block, ctx.res_scale, ctx.small = block, float(res_scale), small
        ctx.dtypes = (dw_w.dtype, None if dw_b is None else dw_b.dtype, ln_g.dtype, ln_b.dtype)
        ctx.has_bias = (b2 is not None, b3 is not None)
        return out.view(B, H, W, C)

    @staticmethod
    def backward(ctx, g):
        xl, t1, mean, rstd, t2, h, a, dw32, g32 = ctx.saved_tensors
        block = ctx.block
        B, H, W, C = xl.shape
        T = B * H * W
        g = g.contiguous()
        g2 = g.view(T, C)
        need = ctx.needs_input_grad
        if ctx.recompute:
            _, w2_bf = block.conv2._weight_matrix()
            if ctx.gelu_epilogue:
                a, h = ops.gemm_bf16_tn_gelu_dual(t2, w2_bf, ctx.b2)
            else:
                h = _gemm_tn(t2, w2_bf, ctx.b2, COMPUTE_DTYPE)
                a = ops.gelu_raw(h)
        gb3 = _colsum(g2) if (ctx.has_bias[1] and need[10]) else None
        gw3 = _wgrad_into(block.conv3.weight, g2, a) if need[9] else None
        if ctx.gelu_epilogue:
            gh = _dgrad(g2, block.conv3._weight_matrix()[1], block.conv3._weight_matrix_t, mult=h)     # (g W3) * gelu'(h)
            gb2 = _colsum(gh) if (ctx.has_bias[0] and need[8]) else None
        else:
            ga = _dgrad(g2, block.conv3._weight_matrix()[1], block.conv3._weight_matrix_t)            # [T, 4C]
            gh, gb2 = ops.gelu_bwd_colsum(h, ga)
            del ga
        gw2 = _wgrad_into(block.conv2.weight, gh, t2) if need[7] else None
        gt2 = _dgrad(gh, block.conv2._weight_matrix()[1], block.conv2._weight_matrix_t)         # [T, C]
        del gh
        gt1, dgam, dbet = ops.ln_backward_raw(gt2, t1.view(T, C), mean, rstd, g32, ctx.small)
        del gt2
        gt1 = gt1.view(B, H, W, C)
        gdw, gdwb = ops.dwconv7_wgrad_raw(gt1, xl)
        gx = None
        if need[0]:
            gx = ops._dwconv7_raw(gt1, dw32.flip(1).t().contiguous(), None, res=g, res_scale=ctx.res_scale)
        dt = ctx.dtypes
        return (gx, None, None, gdw.view(C, 1, 7, 7).to(dt[0]), None if dt[1] is None else gdwb.to(dt[1]),
                dgam.to(dt[2]), dbet.to(dt[3]), gw2, gb2 if ctx.has_bias[0] else None, gw3, gb3)



class _Conv3x3Igemm(torch.autograd.Function):
    """3x3 'same' convolution of the network's edge layers as an implicit GEMM on tcgen05 (csrc/conv_igemm.cu):
    x [B, H, W, 8] (3 image channels zero-padded) -> [B, H, W, 32]   (in_conv; weight [32, <= 4, 3, 3]), or
    x [B, H, W, 32] -> [B, H, W, 4]                                  (out_conv; weight [<= 4, 32, 3, 3]).
    The input gradient is the other instantiation with the taps flipped and the channel roles exchanged; the weight
    gradient (864 numbers summed over all pixels) is the direct kernel of csrc/cnn_elem.cu, for in_conv with the roles of
    the image and the gradient exchanged."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        N = weight.shape[0]
        ctx.in_type = x.shape[-1] == 8
        if ctx.in_type:
            out = ops.conv3x3_igemm(x, ops.igemm_weight_chunks(weight, 8, 32), bias, 32, N)
        else:
            out = ops.conv3x3_igemm(x, ops.igemm_weight_chunks(weight, 32, 16), bias, 4, N)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        N, Cr = weight.shape[0], weight.shape[1]
        wt = weight.detach().transpose(0, 1).flip(2, 3)                       # [C, N, 3, 3]: the transposed convolution
        gx = gw = gb = None
        if ctx.in_type:
            if ctx.needs_input_grad[0]:
                gx4 = ops.conv3x3_igemm(g, ops.igemm_weight_chunks(wt, 32, 16), None, 4, Cr)
                gx = F.pad(gx4, (0, 4))
            gwt, _ = ops.conv3x3_small_wgrad(x[..., :4].contiguous(), g, Cr)        # roles exchanged: [Cr, 32, 3, 3]
            gw = gwt.permute(1, 0, 2, 3).flip(2, 3).contiguous()
            if ctx.has_bias:
                gb = _colsum(g.view(-1, g.shape[-1]))[:N]
        else:
            if ctx.needs_input_grad[0]:
                gx = ops.conv3x3_igemm(F.pad(g, (0, 4)), ops.igemm_weight_chunks(wt, 8, 32), None, 32, 32)
            gw, gb = ops.conv3x3_small_wgrad(g, x, N)
        return gx, gw.to(weight.dtype), (gb.to(ctx.bias_dtype) if ctx.has_bias else None)


class _GemmConv2d(Conv2d):
    """nn.Conv2d whose forward/backward contractions run on the tcgen05 GEMM.  Supports what the reference's
    network uses: 1x1 stride 1, and 3x3 stride 1 with 'same' zero padding; groups = 1."""

    def _weight_matrix(self):
        w = self.weight
        if self.kernel_size == (1, 1):
            w2 = w.reshape(w.shape[0], w.shape[1])
        else:
            w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)        # (C_out, ky, kx, C_in): matches _unfold3x3
        key = (w._version, w.data_ptr(), COMPUTE_DTYPE, shadow_epoch(w))
        if getattr(self, "_w_key", None) != key:
            with torch.no_grad():
                self._w_cache = _pad_k(w2.detach().to(COMPUTE_DTYPE))
            self._w_key = key
            self._wt_cache = None
            # Low-precision shadows for sei_b200.optim.Adam: when the cache has the parameter's own element order
            # (pointwise convolutions with C_in % 8 == 0) the optimizer kernel rewrites it in the pass that updates
            # the fp32 master, so no cast runs in the forward pass; otherwise it just invalidates this cache.
            if w.is_cuda and COMPUTE_DTYPE == torch.bfloat16:
                if self.kernel_size == (1, 1) and self._w_cache.shape == w2.shape:
                    w._sei_lowp, w._sei_lowp_t, w._sei_invalidate = self._w_cache, None, None
                else:
                    w._sei_lowp, w._sei_lowp_t, w._sei_invalidate = None, None, self._invalidate
        return w2, self._w_cache

    def _invalidate(self):
        self._w_key = None

    def _weight_matrix_t(self):
        """(K, N) copy of the low-precision weight for dgrad, made once per weight version (and then kept fresh by
        sei_b200.optim.Adam through the transposed shadow registered on the parameter)"""
        if getattr(self, "_wt_cache", None) is None:
            with torch.no_grad():
                self._wt_cache = _pad_k(self._w_cache.t())
            w = self.weight
            if getattr(w, "_sei_lowp", None) is self._w_cache and self._wt_cache.shape == (self._w_cache.shape[1], self._w_cache.shape[0]):
                w._sei_lowp_t = (self._w_cache.shape[0], self._w_cache.shape[1], self._wt_cache)
        return self._wt_cache

    def forward_nobias(self, x):
        return self.forward(x, use_bias=False)

    def gelu_fusable(self, h):
        """gelu(h) -> this convolution can run as one node with the GELU backward in the dgrad epilogue"""
        return (self.kernel_size == (1, 1) and h.is_cuda and h.dtype == torch.bfloat16 == COMPUTE_DTYPE
                and h.is_contiguous(memory_format=CL) and self.in_channels % 64 == 0 and self.out_channels % 8 == 0
                and _GELU_FUSION)

    def forward_after_gelu(self, h):
        B, C, H, W = h.shape
        h2 = h.permute(0, 2, 3, 1).reshape(B * H * W, C)                            # view of the channels-last tensor
        w2, w_bf16 = self._weight_matrix()
        out = _GeluGemmTN.apply(h2, w2, self.bias, w_bf16, self._weight_matrix_t, self.weight)
        return out.view(B, H, W, -1).permute(0, 3, 1, 2)

    def forward_rowscaled_bias(self, x, row_scale):
        """1x1 convolution whose bias is multiplied by row_scale[pixel] (Downsample behind its resampler); None when the
        shape needs the unfused path"""
        B, C, H, W = x.shape
        if not (self.kernel_size == (1, 1) and self.bias is not None and x.is_cuda and C % 8 == 0 and self.out_channels % 8 == 0
                and COMPUTE_DTYPE == torch.bfloat16 and ops.padded_width(self.out_channels, "colsum") == self.out_channels):
            return None
        xl = x.to(dtype=COMPUTE_DTYPE, memory_format=CL).permute(0, 2, 3, 1)
        w2, w_bf16 = self._weight_matrix()
        out = _GemmTN.apply(xl.reshape(B * H * W, C), w2, self.bias, w_bf16, self._weight_matrix_t,
                            self.weight if w_bf16.shape == w2.shape else None, None, row_scale.reshape(-1))
        return out.view(B, H, W, -1).permute(0, 3, 1, 2)

    def forward(self, x, use_bias=True, residual=None):
        """residual: a tensor of the output's shape added to it -- in the GEMM epilogue when the shapes allow"""
        B, C, H, W = x.shape
        xl = x.to(dtype=COMPUTE_DTYPE, memory_format=CL).permute(0, 2, 3, 1)      # (B, H, W, C) view
        if residual is not None:
            rl = residual.contiguous(memory_format=CL).permute(0, 2, 3, 1)
            if (self.kernel_size == (1, 1) and self.out_channels % 8 == 0 and self.in_channels % 8 == 0 and rl.is_cuda
                    and rl.dtype == torch.bfloat16 == COMPUTE_DTYPE and tuple(rl.shape) == (B, H, W, self.out_channels)):
                w2, w_bf16 = self._weight_matrix()
                out = _GemmTN.apply(xl.reshape(B * H * W, C), w2, self.bias if use_bias else None, w_bf16,
                                    self._weight_matrix_t, self.weight if w_bf16.shape == w2.shape else None,
                                    rl.reshape(B * H * W, self.out_channels))
                return out.view(B, H, W, -1).permute(0, 3, 1, 2)
            return self.forward(x, use_bias=use_bias) + residual
        if (self.kernel_size == (3, 3) and _IGEMM_CONV and xl.is_cuda and xl.dtype == torch.bfloat16 == COMPUTE_DTYPE
                and self.padding in ("same", (1, 1))):
            # the network's edge layers as implicit GEMMs on tcgen05 (csrc/conv_igemm.cu): no unfolded copy
            bias = self.bias if use_bias else None
            if self.out_channels == 32 and self.in_channels <= 4:
                out = _Conv3x3Igemm.apply(_pad_channels(xl, 8).contiguous(), self.weight, bias)
                return out.permute(0, 3, 1, 2)
            if self.in_channels == 32 and self.out_channels <= 4:
                out = _Conv3x3Igemm.apply(xl.contiguous(), self.weight, bias)
                return out[..., : self.out_channels].permute(0, 3, 1, 2)
        if self.kernel_size == (3, 3):
            # the network's output layer (hidden -> 3 channels): direct kernel, no unfolded copy (csrc/cnn_elem.cu)
            out = _op_conv3x3_small(xl, self)
            if out is not None:
                return out[..., : self.out_channels].permute(0, 3, 1, 2)
            xl = _unfold3x3(xl)
        x2 = xl.reshape(B * H * W, xl.shape[-1])
        w2, w_bf16 = self._weight_matrix()
        if x2.shape[1] != w_bf16.shape[1]:
            x2 = F.pad(x2, (0, w_bf16.shape[1] - x2.shape[1]))
        w2p = w2 if w2.shape[1] == w_bf16.shape[1] else F.pad(w2, (0, w_bf16.shape[1] - w2.shape[1]))
        in_place_ok = self.kernel_size == (1, 1) and w2p is w2 and w_bf16.shape == w2.shape
        out = _GemmTN.apply(x2, w2p, self.bias if use_bias else None, w_bf16, self._weight_matrix_t,
                            self.weight if in_place_ok else None)
        return out.view(B, H, W, -1).permute(0, 3, 1, 2)                        # logical NCHW, channels-last memory


def _unfold3x3(xl):
    """(B, H, W, C) -> (B, H, W, 9C): the 3x3 neighbourhood with zero padding, ordered (ky, kx, c)."""
    B, H, W, C = xl.shape
    xp = F.pad(xl, (0, 0, 1, 1, 1, 1))
    return torch.cat([xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)], dim=-1)


def _conv(in_channels, out_channels, kernel_size, **kw):
    return _GemmConv2d(in_channels, out_channels, kernel_size=kernel_size, **kw)


class LayerNorm(Module):
    """layer norm over the channels (reference :21-30: swapaxes(-3, -1) / nn.LayerNorm / swap back)"""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.ln = BaseLayerNorm(*args, **kwargs)

    def forward(self, x):
        xl = x.contiguous(memory_format=CL).permute(0, 2, 3, 1)                   # channels last: a view
        rows = xl.reshape(-1, xl.shape[-1])
        out = _op_layer_norm(rows, self.ln).view(xl.shape)                        # hand-written kernels (csrc/cnn_elem.cu)
        return out.permute(0, 3, 1, 2)


class ConvBlock(Module):
    _sei_atomic_group = True       # one autograd node: its gradients become final together (sei_b200.parallel.completion_groups)

    def __init__(self, dim):
        super().__init__()
        self.conv1 = Conv2d(in_channels=dim, out_channels=dim, kernel_size=7, padding=3, groups=dim)
        self.ln = LayerNorm(dim, eps=1e-6)
        self.conv2 = _conv(dim, 4 * dim, 1)
        self.gelu = GELU()
        self.conv3 = _conv(4 * dim, dim, 1)

    def fused_node_ok(self, xl):
        """the block runs as one autograd node (_ConvBlockFn): bf16 CUDA activations and channel counts the vectorised
        kernels tile (every width of the default network); other widths take the op-by-op path below"""
        return (_CONVBLOCK_NODE and xl.is_cuda and xl.dtype == torch.bfloat16 == COMPUTE_DTYPE and xl.shape[-1] % 8 == 0
                and ops.dwconv7_supported(xl) and ops.ln_cl_supported(xl.reshape(-1, xl.shape[-1])))

    def forward(self, x, res_scale=1.0):
        """x + block(x) for res_scale = 1 (the reference's forward); res_scale * x + block(x) in general"""
        xl = x.contiguous(memory_format=CL).permute(0, 2, 3, 1)
        if self.fused_node_ok(xl):
            out = _ConvBlockFn.apply(xl.contiguous(), res_scale, self, self.conv1.weight, self.conv1.bias, self.ln.ln.weight,
                                     self.ln.ln.bias, self.conv2.weight, self.conv2.bias, self.conv3.weight, self.conv3.bias)
            return out.permute(0, 3, 1, 2)
        x1 = _op_dwconv7(xl, self.conv1).permute(0, 3, 1, 2)      # hand-written channels-last kernels (csrc/cnn_elem.cu)
        x1 = self.ln(x1)
        x1 = self.conv2(x1)
        if self.conv3.gelu_fusable(x1):
            x1 = self.conv3.forward_after_gelu(x1)            # gelu + conv3, gelu' fused into conv3's dgrad epilogue
        else:
            x1 = _op_gelu(x1)
            x1 = self.conv3(x1)
        return (x if res_scale == 1.0 else res_scale * x) + x1


class IdealUpsample(Module):
    """zero-padding of the (fftshift-ed) half spectrum, exactly as the reference does it (:54-92)"""

    def __init__(self, rate=2):
        super().__init__()
        self.rate = rate

    def forward(self, x):
        return _op_ideal_resample(x, "up", self.rate)      # two batched tensor-core products (bgemm.cu), no cuFFT


class Upsample(Module):
    def __init__(self, in_channels, out_channels=None, rate=2):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels or in_channels // (rate ** 2)
        self.rate = rate
        self.seq = Sequential()
        self.seq.append(IdealUpsample(rate=self.rate))
        self.seq.append(LayerNorm(self.in_channels, eps=1e-6))
        self.seq.append(_conv(self.in_channels, self.out_channels, 1, stride=1))

    def forward(self, x, skip=None):
        if skip is None:
            return self.seq(x)
        x = self.seq[1](self.seq[0](x))
        return self.seq[2](x, residual=skip)


class IdealDownsample(Module):
    """mask of the (fftshift-ed) half spectrum then stride-`rate` subsampling (:113-133)"""

    def __init__(self, rate=2):
        super().__init__()
        self.rate = rate

    def forward(self, x):
        return _op_ideal_resample(x, "down", self.rate)


class Downsample(Module):
    def __init__(self, in_channels, out_channels=None, rate=2):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels or in_channels * (rate ** 2)
        self.rate = rate
        self.ln = LayerNorm(self.in_channels, eps=1e-6)
        self.conv = _conv(self.in_channels, self.out_channels, 1, stride=1)
        self.ideal_downsample = IdealDownsample(rate=self.rate)

    def forward(self, x):
        x = self.ln(x)
        if self.conv.kernel_size == (1, 1) and COMMUTE_DOWNSAMPLE:
            # The pointwise convolution (channels) and the ideal resampler (space) are linear maps on different axes,
            # so they commute: resample the C-channel tensor first, then convolve a quarter of the pixels (4x fewer
            # resampler bytes and GEMM flops than conv -> resample).  The bias is a constant image per channel; the
            # resampler maps it to bias[c] * R(1), added afterwards.
            H, W = x.shape[-2], x.shape[-1]
            xr = self.ideal_downsample(x)
            if self.conv.bias is not None:
                # the scaled bias rides in the GEMM epilogue (an in-place pass on a view of the GEMM output made autograd
                # copy the whole tensor on the way back: 2.2 ms of device-to-device copies per step)
                pat = resample.constant_response("down", H, W, self.rate, x.device)                 # (Ho, Wo) fp32
                out = self.conv.forward_rowscaled_bias(xr, pat)
                if out is not None:
                    return out
            out = self.conv.forward_nobias(xr)
            if self.conv.bias is not None:
                rows = out.permute(0, 2, 3, 1).reshape(-1, out.shape[1])                            # view of the GEMM output
                out = _op_bias_pattern(out, rows, pat, self.conv.bias)
            return out
        return self.ideal_downsample(self.conv(x))


def _run_sequence(seq, x, inner_residual):
    """`xb = x; x = seq(x); if inner_residual: x = x + xb` (reference UNet.forward).  A sequence of one ConvBlock (the
    default) returns 2 x + block(x) from the block's own node: no separate addition."""
    if inner_residual and len(seq) == 1 and isinstance(seq[0], ConvBlock):
        return seq[0](x, res_scale=2.0)
    xb = x
    x = seq(x)
    return x + xb if inner_residual else x


class UNet(Module):
    def __init__(self, in_channels, hidden_channels, inout_convs, scales, num_conv_blocks, rate, residual,
                 inner_residual):
        super().__init__()
        self.scales = scales
        self.residual = residual
        self.inner_residual = inner_residual
        self.conv_sequences = ModuleList()
        self.downsampling_layers = ModuleList()
        self.upsampling_layers = ModuleList()

        if inout_convs:
            self.in_conv = _conv(in_channels, hidden_channels, 3, padding="same")
            self.out_conv = _conv(hidden_channels, in_channels, 3, padding="same")
            width = hidden_channels
        else:
            width = in_channels

        def blocks(dim):
            return Sequential(*[ConvBlock(dim=dim) for _ in range(num_conv_blocks)])

        for _ in range(scales - 1):
            self.conv_sequences.append(blocks(width))
            self.downsampling_layers.append(Downsample(in_channels=width))
            width *= rate ** 2
        self.conv_sequences.append(blocks(width))
        for _ in range(scales - 1):
            self.upsampling_layers.append(Upsample(in_channels=width, rate=rate))
            width //= rate ** 2
            self.conv_sequences.append(blocks(width))

    def forward(self, x):
        x0 = x
        convs, downs, ups = iter(self.conv_sequences), iter(self.downsampling_layers), iter(self.upsampling_layers)
        skips = []
        if hasattr(self, "in_conv"):
            x = self.in_conv(x)
        for _ in range(self.scales - 1):
            x = _run_sequence(next(convs), x, self.inner_residual)
            skips.append(x)
            x = next(downs)(x)
        x = next(convs)(x)
        for _ in range(self.scales - 1):
            x = next(ups)(x, skip=skips.pop())       # `x = ups(x); x = x + skip` with the addition in the GEMM epilogue
            x = next(convs)(x)
        if hasattr(self, "out_conv"):
            x = self.out_conv(x)
        if self.residual:
            x = x + x0
        return x


class ConvolutionalModel(Module):
    def __init__(self, in_channels, upsampling_rate, residual, inner_residual, num_conv_blocks, hidden_channels,
                 inout_convs, scales):
        super().__init__()
        self.seq = Sequential()
        self.scales = scales
        if upsampling_rate != 1:
            self.seq.append(Upsample(in_channels=in_channels, out_channels=in_channels, rate=upsampling_rate))
        self.seq.append(UNet(in_channels=in_channels, hidden_channels=hidden_channels, inout_convs=inout_convs,
                             scales=scales, num_conv_blocks=num_conv_blocks, residual=residual,
                             inner_residual=inner_residual, rate=2))

    def forward(self, y):
        div = 2 ** (self.scales - 1)
        pad_h = (div - y.shape[-2] % div) % div
        pad_w = (div - y.shape[-1] % div) % div
        if pad_h != 0 or pad_w != 0:
            y = F.pad(y, (0, pad_w, 0, pad_h), mode="reflect")
        x_hat = self.seq(y.to(dtype=COMPUTE_DTYPE, memory_format=CL))
        x_hat = x_hat.to(dtype=y.dtype).contiguous()
        if pad_h != 0:
            x_hat = x_hat[:, :, :-pad_h, :]
        if pad_w != 0:
            x_hat = x_hat[:, :, :, :-pad_w]
        return x_hat
