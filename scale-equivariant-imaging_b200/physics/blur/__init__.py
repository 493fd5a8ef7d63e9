"""Blur physics on libsei_b200 kernels.  API of the reference's src/physics/blur/__init__.py:
conv, conv_transpose, extend_filter, Blur (v1, padded direct convolution), BlurV2 (circular
convolution; the reference evaluates it with FFTs, :205-223).  Both circular variants map onto
the same separable TMA-tiled kernel; their transpose is the hand-written adjoint kernel."""
import torch

from sei_b200 import ops
from sei_b200.linear_physics import LinearPhysics


def extend_filter(filter):
    """src/physics/blur/__init__.py:9-31: size-1 axes become 3 (centred), even axes get one
    trailing zero.  Returns a float32 tensor like the reference."""
    b, c, h, w = filter.shape
    h_new = 3 if h == 1 else h + (h % 2 == 0)
    w_new = 3 if w == 1 else w + (w % 2 == 0)
    oh, ow = int(h == 1), int(w == 1)
    out = torch.zeros((b, c, h_new, w_new), device=filter.device)
    out[:, :, oh:oh + h, ow:ow + w] = filter
    return out


def _single_channel(filter):
    if filter.dim() != 4 or filter.shape[0] != 1 or filter.shape[1] != 1:
        raise NotImplementedError("only (1,1,h,w) filters (one filter shared by all channels) are supported")
    return ops.kernel_to_host(filter)


def _is_circular_fast(k, padding):
    return padding == "circular" and k.shape[0] % 2 == 1 and k.shape[1] % 2 == 1 and k.shape[0] > 1 and k.shape[1] > 1


def conv(x, filter, padding):
    """Convolution of x (B,C,H,W) with filter (1,1,h,w); padding in valid|circular|replicate|reflect."""
    k = _single_channel(filter)
    if _is_circular_fast(k, padding):
        return ops._BlurCircular.apply(x, k, False, ops.PATH_AUTO)
    return ops._BlurPadded.apply(x, k, padding, False)


def conv_transpose(y, filter, padding):
    """Transpose of conv(., filter, padding); padding additionally accepts 'zero'."""
    k = _single_channel(filter)
    if _is_circular_fast(k, padding):
        return ops._BlurCircular.apply(y, k, True, ops.PATH_AUTO)
    return ops._BlurPadded.apply(y, k, padding, True)


class Blur(LinearPhysics):
    """y = w * x with the given padding (reference :164-194)."""

    def __init__(self, filter, padding="circular", device="cpu", **kwargs):
        super().__init__(**kwargs)
        self.padding = padding
        self.device = device
        self.filter = torch.nn.Parameter(filter, requires_grad=False).to(device)
        self._kernel_host = _single_channel(self.filter)

    def A(self, x):
        return conv(x, self.filter, self.padding)

    def A_adjoint(self, y):
        return conv_transpose(y, self.filter, self.padding)

    def measure_with_noise(self, x, noise):
        if _is_circular_fast(self._kernel_host, self.padding):
            return ops._BlurCircularNoise.apply(x, self._kernel_host, noise, self.noise_model.sigma_value())
        return super().measure_with_noise(x, noise)

    def ei_remeasure_args(self):
        if _is_circular_fast(self._kernel_host, self.padding):
            return dict(kernel_host=self._kernel_host, rate_sr=1)
        return None


class BlurV2(LinearPhysics):
    """Circular convolution with the kernel centred at k//2 (reference :197-227)."""

    def __init__(self, kernel):
        super().__init__()
        self.kernel = kernel
        self.filter = self.kernel
        self.fft_norm = "backward"   # kept for attribute compatibility; no FFT is involved here
        self._kernel_host = _single_channel(kernel)

    def A(self, x):
        return ops._BlurCircular.apply(x, self._kernel_host, False, ops.PATH_AUTO)

    def A_adjoint(self, y):
        return ops._BlurCircular.apply(y, self._kernel_host, True, ops.PATH_AUTO)

    def measure_with_noise(self, x, noise):
        return ops._BlurCircularNoise.apply(x, self._kernel_host, noise, self.noise_model.sigma_value())

    def ei_remeasure_args(self):
        return dict(kernel_host=self._kernel_host, rate_sr=1)
