"""The part of deepinv v0.2.0 (the reference's un-vendored dependency, requirements.txt:1) that
the reference's physics and losses build on, restated for this package: LinearPhysics,
GaussianNoise, adjoint_function, EILoss, SupLoss, mse.  Definitions follow
tests/golden/deepinv_shim (the written specification used to generate the golden vectors); the
arithmetic runs in libsei_b200 kernels."""
import torch
from torch import nn

from . import draws, ops


class GaussianNoise(nn.Module):
    """y = x + randn_like(x) * sigma   (deepinv physics/noise.py; attached at reference
    src/physics/__init__.py:53).  `sigma` is a float32 Parameter like upstream."""

    def __init__(self, sigma=0.1):
        super().__init__()
        self.sigma = nn.Parameter(torch.tensor(sigma), requires_grad=False)

    def sigma_value(self):
        if not hasattr(self, "_sigma_host"):
            self._sigma_host = float(self.sigma.detach().cpu())
        return self._sigma_host

    def forward(self, x):
        noise = draws.randn_like(x)
        return ops._AddNoise.apply(x, noise, self.sigma_value())


class LinearPhysics(nn.Module):
    """forward(x) = sensor(noise(A(x))) with identity sensor; A_dagger by conjugate gradient."""

    def __init__(self, max_iter=50, tol=1e-3, **kwargs):
        super().__init__()
        self.noise_model = lambda x: x
        self.sensor_model = lambda x: x
        self.max_iter = max_iter
        self.tol = tol

    def A(self, x):
        raise NotImplementedError

    def A_adjoint(self, y):
        raise NotImplementedError

    def noise(self, x):
        return self.noise_model(x)

    def sensor(self, x):
        return self.sensor_model(x)

    def forward(self, x):
        return self.sensor(self.noise(self.A(x)))

    def measure_with_noise(self, x, noise):
        """A(x) + sigma * noise with the noise add fused into the operator's epilogue; `noise` is
        the standard-normal draw.  Subclasses override; this is the unfused composition."""
        return ops._AddNoise.apply(self.A(x), noise, self.noise_model.sigma_value())

    def A_dagger(self, y):
        """Least-squares pseudo-inverse by conjugate gradient, following deepinv v0.2.0 (restated in
        tests/golden/deepinv_shim/deepinv/physics/forward.py; upstream is not vendored): when A^T y has fewer
        elements than y solve A^T A x = A^T y, otherwise (deblurring, super-resolution) solve A A^T z = y and return
        A^T z.  With the reference's default non-adjoint `A_adjoint` of the SR operator (plain bicubic upsample) the
        two are different systems, so the branch matters.  The operators run in libsei_b200; the CG vector updates are
        torch (test-time path only: demo/test.py:122, src/models/__init__.py:28)."""
        Aty = self.A_adjoint(y)
        overcomplete = Aty.numel() < y.numel()
        if not overcomplete:
            op, b = (lambda v: self.A(self.A_adjoint(v))), y
        else:
            op, b = (lambda v: self.A_adjoint(self.A(v))), Aty
        x = conjugate_gradient(op, b, max_iter=self.max_iter, tol=self.tol)
        if not overcomplete:
            x = self.A_adjoint(x)
        return x


def conjugate_gradient(A, b, max_iter=1e2, tol=1e-5):
    """deepinv.optim.utils.conjugate_gradient (v0.2.0): CG from x0 = 0, stops once |r| < tol"""
    x = torch.zeros_like(b)
    r = b
    p = r
    rsold = (r * r).flatten().sum()
    for _ in range(int(max_iter)):
        Ap = A(p)
        alpha = rsold / (p * Ap).flatten().sum()
        x = x + p * alpha
        r = r + Ap * (-alpha)
        rsnew = (r * r).flatten().sum()
        if rsnew.sqrt() < tol:
            break
        p = r + p * (rsnew / rsold)
        rsold = rsnew
    return x


def adjoint_function(A, input_size, device="cpu", dtype=torch.float):
    """deepinv.physics.adjoint_function: y -> vjp of A at ones(input_size).  For the linear
    operators of this package the vjp is the hand-written transpose kernel (autograd backward)."""
    x = torch.ones(input_size, device=device, dtype=dtype, requires_grad=True)
    out = A(x)
    batches = input_size[0]

    def adjoint(y):
        if y.size(0) > batches:
            raise ValueError("Batch size of A_adjoint input is larger than expected")
        if y.size(0) < batches:
            y2 = torch.zeros((batches,) + tuple(y.shape[1:]), device=y.device, dtype=y.dtype)
            y2[: y.size(0)] = y
            return torch.autograd.grad(out, x, y2, retain_graph=True)[0][: y.size(0)]
        return torch.autograd.grad(out, x, y, retain_graph=True)[0]

    return adjoint


class MSE(nn.Module):
    """nn.MSELoss() computed by the reduction kernel (forward) and one elementwise kernel (backward)."""

    def forward(self, input, target):
        return ops.mse(input, target)


def mse():
    return MSE()


class SupLoss(nn.Module):
    """deepinv.loss.SupLoss: metric(x_net, x)."""

    def __init__(self, metric=None):
        super().__init__()
        self.name = "sup"
        self.metric = metric if metric is not None else MSE()

    def forward(self, x_net, x, **kwargs):
        return self.metric(x_net, x)


class EILoss(nn.Module):
    """deepinv.loss.EILoss:  x2 = T(x_net) (no grad if no_grad); y = physics(x2) (noise if
    apply_noise); x3 = model(y, physics); weight * metric(x3, x2).

    When T is this package's padded ScalingTransform and the physics is one of this package's
    operators, T, A and the noise add run as ONE fused kernel (sei_ei_remeasure_f32)."""

    def __init__(self, transform, metric=None, apply_noise=True, weight=1.0, no_grad=True):
        super().__init__()
        self.name = "ei"
        self.metric = metric if metric is not None else MSE()
        self.weight = weight
        self.T = transform
        self.noise = apply_noise
        self.no_grad = no_grad

    def _remeasure(self, x_net, physics):
        fused = getattr(self.T, "fused_remeasure", None)
        if fused is not None and self.no_grad and hasattr(physics, "ei_remeasure_args"):
            with torch.no_grad():
                return fused(x_net, physics, apply_noise=self.noise)
        if self.no_grad:
            with torch.no_grad():
                x2 = self.T(x_net)
        else:
            x2 = self.T(x_net)
        y = physics(x2) if self.noise else physics.A(x2)
        return x2, y

    def forward(self, x_net, physics, model, **kwargs):
        x2, y = self._remeasure(x_net, physics)
        x3 = model(y, physics)
        return self.weight * self.metric(x3, x2)


class Shift(nn.Module):
    """deepinv.transform.Shift (v0.2.0, n_trans = 1): one random circular shift of the whole batch in both axes.
    Draws follow tests/golden/deepinv_shim (two CPU randperm draws); the roll itself is sei_roll_f32."""

    def __init__(self, n_trans=1, shift_max=1.0):
        super().__init__()
        if n_trans != 1:
            raise NotImplementedError("Shift(n_trans != 1) is not used by the reference's losses")
        self.n_trans = n_trans
        self.shift_max = shift_max

    def forward(self, x):
        H, W = x.shape[-2:]
        assert self.n_trans <= H - 1 and self.n_trans <= W - 1
        H_max, W_max = int(self.shift_max * H), int(self.shift_max * W)
        sh = int(torch.arange(-H_max, H_max)[draws.randperm(2 * H_max)][0])
        sw = int(torch.arange(-W_max, W_max)[draws.randperm(2 * W_max)][0])
        return ops._Roll.apply(x, sh, sw)


class Rotate(nn.Module):
    """deepinv.transform.Rotate (v0.2.0, n_trans = 1): one random rotation of the whole batch by an integer angle in
    1..359 degrees, nearest-neighbour, same size, zero fill (torchvision's rotate defaults).  The draw follows
    tests/golden/deepinv_shim (one CPU randperm(359)); the resampling is sei_rotate_nearest_f32."""

    def __init__(self, n_trans=1, degrees=360):
        super().__init__()
        if n_trans != 1:
            raise NotImplementedError("Rotate(n_trans != 1) is not used by the reference's losses")
        self.n_trans, self.group_size = n_trans, degrees

    def forward(self, x):
        theta = torch.arange(0, 360)[1:][draws.randperm(359)][: self.n_trans]
        return ops.rotate_nearest(x, float(theta[0]))
