"""Gaussian SURE with a Monte-Carlo divergence (reference: src/losses/sure.py, itself taken from
deepinv).  Same signatures: mc_div(y1, y, model, physics, tau, margin) and
SureGaussianLoss(sigma, tau, margin, cropped_div, averaged_cst).forward(y, x_net, physics, model).

The probe y + tau*b (b ~ N(0,1) inside the margin, 0 on the border) is one kernel; the two
reductions (interior mean of (y1-y)^2 and of b*(y2-y1)/tau) are one deterministic reduction
kernel with a hand-written backward."""
import torch
import torch.nn as nn

from sei_b200 import draws, ops


def _probe(y, tau, margin):
    B, C, H, W = y.shape
    if margin == 0:
        draw = draws.randn_like(y)
    else:
        draw = draws.randn((B, C, H - 2 * margin, W - 2 * margin), y.device, y.dtype)
    return ops.sure_perturb(y, draw, margin, tau)   # (y + tau*b, b)


def mc_div(y1, y, model, physics, tau, margin=0):
    """Monte-Carlo divergence estimate: mean over the interior of b * (A(model(y + tau b)) - y1) / tau."""
    assert margin is not None
    y_pert, b = _probe(y, tau, margin)
    y2 = physics.A(model(y_pert))
    return ops.mc_div(y1, y2, b, margin, tau)


class SureGaussianLoss(nn.Module):
    def __init__(self, sigma, tau=1e-2, margin=0, cropped_div=False, averaged_cst=False):
        super().__init__()
        self.name = "SureGaussian"
        self.sigma2 = sigma ** 2
        self.tau = tau
        assert margin is not None
        self.margin = margin
        self.cropped_div = cropped_div
        self.averaged_cst = averaged_cst

    def forward(self, y, x_net, physics, model, **kwargs):
        y1 = physics.A(x_net)
        margin_div = self.margin if self.cropped_div else 0
        y_pert, b = _probe(y, self.tau, margin_div)
        y2 = physics.A(model(y_pert))
        loss, _ = ops.sure_loss(y1, y2, y, b, self.margin, margin_div, self.tau, self.sigma2, self.averaged_cst)

        from os import environ
        if "_TEMPORARY_HOTFIX" in environ:
            assert physics.rate is not None
            return physics.rate ** 2 * loss
        return loss
