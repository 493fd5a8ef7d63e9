"""Recorrupted-to-Recorrupted loss and its equivariant companion (reference: src/losses/r2r.py), selected by
--ProposedLoss__sure_alternative r2r.  Same class names, constructor arguments and attributes; every recorruption
y + level * n is one streaming kernel (sei_add_noise_f32) on a standard-normal draw, A and the MSE reductions are the
libsei_b200 operators."""
import torch
from torch import nn

from sei_b200 import draws, ops
from sei_b200.linear_physics import mse


def _recorrupt(y, noise, level):
    """y + level * noise (noise ~ N(0, 1) drawn by the caller, so that one draw can be used twice)"""
    return ops._AddNoise.apply(y, noise, level)


class R2RLoss(nn.Module):
    """metric(A(model(y + eta alpha n)), y - (eta / alpha) n) for ONE draw n (reference :8-24)"""

    def __init__(self, metric=None, eta=0.1, alpha=0.5):
        super().__init__()
        self.name = "r2r"
        self.metric = mse() if metric is None else metric
        self.eta, self.alpha = eta, alpha

    def recorrupted_pair(self, y):
        n = draws.randn_like(y)
        return _recorrupt(y, n, self.eta * self.alpha), _recorrupt(y, n, -self.eta / self.alpha)

    def forward(self, y, physics, model, **kwargs):
        network_input, target = self.recorrupted_pair(y)
        return self.metric(physics.A(model(network_input, physics)), target)


class R2REILoss(nn.Module):
    """R2R data term + an EI term whose two network inputs carry noise of the same total level 1.5 sigma: the measurement
    (noise sigma) is recorrupted with 0.5 sigma, the clean re-measurement A(T(x1)) with 1.5 sigma (reference :27-57)"""

    FIRST_PASS_LEVEL, SECOND_PASS_LEVEL = 0.5, 1.5

    def __init__(self, transform, sigma, no_grad=True, metric=None):
        super().__init__()
        self.T, self.sigma, self.no_grad = transform, sigma, no_grad
        self.metric = mse() if metric is None else metric
        self.r2r_loss = R2RLoss(eta=self.sigma, alpha=0.5)

    def _transformed(self, x):
        # the stop-gradient of the EI target: T runs without a graph when no_grad is set
        with torch.set_grad_enabled(torch.is_grad_enabled() and not self.no_grad):
            return self.T(x)

    def ei_loss(self, y, physics, model, **kwargs):
        x1 = model(_recorrupt(y, draws.randn_like(y), self.FIRST_PASS_LEVEL * self.sigma), physics)
        x2 = self._transformed(x1)
        y2 = physics.A(x2)
        x3 = model(_recorrupt(y2, draws.randn_like(y2), self.SECOND_PASS_LEVEL * self.sigma), physics)
        return self.metric(x3, x2)

    def forward(self, *kargs, **kwargs):
        # draw order of the reference: the R2R pair first, then the two EI recorruptions
        return self.r2r_loss(*kargs, **kwargs) + self.ei_loss(*kargs, **kwargs)
