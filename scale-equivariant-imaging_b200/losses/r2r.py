"""Recorrupted-to-Recorrupted loss and its equivariant companion (reference: src/losses/r2r.py), selected by
--ProposedLoss__sure_alternative r2r.  Same classes and signatures; the scaled noise additions, A and the
MSE reductions run in libsei_b200 kernels."""
import torch
import torch.nn as nn
from torch.nn import Module

from sei_b200 import draws, ops
from sei_b200.linear_physics import mse


class R2RLoss(nn.Module):
    def __init__(self, metric=None, eta=0.1, alpha=0.5):
        super().__init__()
        self.name = "r2r"
        self.metric = metric if metric is not None else mse()
        self.eta = eta
        self.alpha = alpha

    def forward(self, y, physics, model, **kwargs):
        n = draws.randn_like(y)                                     # pert = n * eta
        y_plus = ops._AddNoise.apply(y, n, self.eta * self.alpha)   # y + pert * alpha
        y_minus = ops._AddNoise.apply(y, n, -self.eta / self.alpha)  # y - pert / alpha
        output = model(y_plus, physics)
        return self.metric(physics.A(output), y_minus)


class R2REILoss(Module):
    def __init__(self, transform, sigma, no_grad=True, metric=None):
        super().__init__()
        self.T = transform
        self.sigma = sigma
        self.no_grad = no_grad
        self.metric = metric if metric is not None else mse()
        self.r2r_loss = R2RLoss(eta=self.sigma, alpha=0.5)

    def forward(self, *kargs, **kwargs):
        return self.r2r_loss(*kargs, **kwargs) + self.ei_loss(*kargs, **kwargs)

    def ei_loss(self, y, physics, model, **kwargs):
        """EI with consistent input noise (reference :37-57): both network inputs carry noise of level 1.5 sigma"""
        x1 = model(ops._AddNoise.apply(y, draws.randn_like(y), 0.5 * self.sigma), physics)
        if self.no_grad:
            with torch.no_grad():
                x2 = self.T(x1)
        else:
            x2 = self.T(x1)
        y2 = physics.A(x2)
        x3 = model(ops._AddNoise.apply(y2, draws.randn_like(y2), 1.5 * self.sigma), physics)
        return self.metric(x3, x2)
