"""Random scale transform (reference: src/transforms.py) on the libsei_b200 kernels.

Same public names and call signatures: sample_from, sample_downsampling_parameters,
get_downsampling_grid, padded_downsampling_transform, PaddedDownsamplingTransform,
normal_downsampling_transform, NormalDownsamplingTransform, ScalingTransform, CombinedTransform.  The sampling grid is never materialised on the hot path:
the kernel recomputes its coordinates with the reference's fp32 rounding sequence."""
import torch
from torch.nn import Module

from sei_b200 import draws, ops


def sample_from(values, shape=(1,), dtype=torch.float32, device="cpu"):
    """values[floor(len(values) * U)], U ~ rand(shape)   (reference :5-11)."""
    u = draws.rand(tuple(shape), device, dtype)
    idx = torch.floor(len(values) * u).to(torch.int)
    return torch.tensor(values, device=device, dtype=dtype)[idx]


def sample_downsampling_parameters(image_count, device, dtype, rates):
    """Draw order and shapes of the reference (:14-24): rand((B,)) for the rates, then
    rand((B, 2)) for the centres.  Returns rate (B,) and center (B,1,1,2) in [-1, 1]."""
    u_rate = draws.rand((image_count,), device, dtype)
    u_center = draws.rand((image_count, 2), device, dtype)
    if u_rate.is_cuda and dtype == torch.float32:
        return ops.scale_params(u_rate, u_center, rates)
    values = torch.tensor(rates, device=device, dtype=dtype)
    idx = torch.floor(len(rates) * u_rate).to(torch.int)
    return values[idx], 2 * u_center.view(image_count, 1, 1, 2) - 1


def get_downsampling_grid(shape, downsampling_rate, center, dtype, device):
    """The sampling grid of the reference (:27-43), for inspection only: the kernels evaluate it
    analytically.  Square images only, like the reference (its .view scrambles h != w)."""
    b, _, h, w = shape
    u = 2 / w * torch.arange(w, dtype=dtype, device=device) - 1
    v = 2 / h * torch.arange(h, dtype=dtype, device=device) - 1
    U, V = torch.meshgrid(u, v, indexing="ij")
    grid = torch.stack([V, U], dim=-1).view(1, h, w, 2).repeat(b, 1, 1, 1)
    inv = 1 / downsampling_rate.view(b, 1, 1, 1).expand_as(grid)
    return inv * (grid - center) + center


def padded_downsampling_transform(x, downsampling_rate, center, mode, padding_mode, antialiased):
    """Zoom-out of each image by its rate about its centre; bicubic taps, reflection padding,
    same output size (reference :60-83).  downsampling_rate: (B,), center: (B,1,1,2)."""
    if mode != "bicubic" or padding_mode != "reflection":
        raise NotImplementedError("only mode='bicubic', padding_mode='reflection' (the only combination the "
                                  "reference ever passes, src/transforms.py:105-106) is built")
    if antialiased:
        return _antialiased_padded_transform(x, downsampling_rate, center)
    return ops._ScaleTransform.apply(x, downsampling_rate, center, ops.PATH_AUTO)


def _antialiased_padded_transform(x, downsampling_rate, center):
    """alias_free_interpolate (reference :44-57: per-image F.interpolate(scale_factor=rate_i.item(), antialias=True), then
    torch.stack) followed by grid_sample of the SMALLER image with the grid of the original shape (:63-82).  Like the
    reference it reads the rates on the host, and like the reference it only works when every image of the batch drew
    the same rate: torch.stack of differently sized images raises a RuntimeError there, and so does this."""
    rates = [float(r) for r in downsampling_rate.reshape(-1).tolist()]
    if len(set(rates)) > 1:
        raise RuntimeError("stack expects each tensor to be equal size: the anti-aliasing pre-filter resizes every image by "
                           f"its own rate {sorted(set(rates))} (reference src/transforms.py:44-57 fails the same way)")
    small = ops.resize_bicubic(x, rates[0], True)
    return ops.scale_transform_from(small, x.shape[-1], downsampling_rate, center)


class PaddedDownsamplingTransform(Module):
    def __init__(self, antialias, downsampling_rates):
        super().__init__()
        self.antialias = antialias
        self.downsampling_rates = downsampling_rates

    def _sample(self, x):
        return sample_downsampling_parameters(image_count=x.shape[0], device=x.device, dtype=x.dtype,
                                              rates=self.downsampling_rates)

    def forward(self, x):
        rate, center = self._sample(x)
        return padded_downsampling_transform(x, downsampling_rate=rate, center=center, antialiased=self.antialias,
                                             mode="bicubic", padding_mode="reflection")

    def fused_remeasure(self, x_net, physics, apply_noise=True):
        """x2 = T(x_net) and y = A(x2) + sigma * n in one kernel (sei_ei_remeasure_f32).  The
        random tensors are drawn in the reference's order: rates, centres, then the noise."""
        args = physics.ei_remeasure_args()
        if self.antialias or args is None:
            return None
        rate, center = self._sample(x_net)
        B, C, S, _ = x_net.shape
        r = args["rate_sr"]
        So = S if r == 1 else ops.down_out_size(S, r)
        noise = draws.randn((B, C, So, So), x_net.device, x_net.dtype) if apply_noise else None
        sigma = physics.noise_model.sigma_value() if apply_noise else 0.0
        return ops.ei_remeasure(x_net, rate, center, args["kernel_host"], r, noise, sigma)


def normal_downsampling_transform(x, downsampling_rate, mode, antialiased):
    """Every image resized by the same factor (reference :112-124: a per-image F.interpolate loop); one kernel here."""
    if mode != "bicubic":
        raise NotImplementedError("only mode='bicubic' (the only one the reference passes, src/transforms.py:139) is built")
    return ops.resize_bicubic(x, float(downsampling_rate), antialiased)


class NormalDownsamplingTransform(Module):
    """reference :127-145: ONE rate for the whole batch, drawn with sample_from(shape=()) and read on the host"""

    def __init__(self, antialias, downsampling_rates):
        super().__init__()
        self.antialias = antialias
        self.downsampling_rates = downsampling_rates

    def forward(self, x):
        rate = sample_from(self.downsampling_rates, shape=(), dtype=x.dtype, device=x.device).item()
        return normal_downsampling_transform(x, downsampling_rate=rate, mode="bicubic", antialiased=self.antialias)


class ScalingTransform(Module):
    def __init__(self, kind, antialias):
        super().__init__()
        downsampling_rates = [0.75, 0.5]
        if kind == "padded":
            self.transform = PaddedDownsamplingTransform(antialias=antialias, downsampling_rates=downsampling_rates)
        elif kind == "normal":
            self.transform = NormalDownsamplingTransform(antialias=antialias, downsampling_rates=downsampling_rates)
        else:
            raise ValueError(f"Unknown kind: {kind}")

    def forward(self, x):
        return self.transform(x)

    def fused_remeasure(self, x_net, physics, apply_noise=True):
        fused = getattr(self.transform, "fused_remeasure", None)          # only the padded transform has a fused kernel
        out = fused(x_net, physics, apply_noise=apply_noise) if fused is not None else None
        if out is None:
            x2 = self.transform(x_net)
            return x2, (physics(x2) if apply_noise else physics.A(x2))
        return out


class CombinedTransform(Module):
    def __init__(self, transforms):
        super().__init__()
        self.transforms = transforms

    def forward(self, x):
        for transform in self.transforms:
            x = transform(x)
        return x
