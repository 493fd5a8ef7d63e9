"""A tiny differentiable stand-in for the restoration network, shared by the golden
generator (tests/golden/make_golden.py, run against the reference) and the parity tests.

x_hat = w0*u + w1*roll(u, (1, 2)) + w2*u^2 + c   with u = nearest-upsample(y, rate)

It is deliberately non-linear and spatially mixing so that the loss gradients w.r.t. its
three outputs per step (x_net, x_net of the perturbed input, x3) are all exercised.
Accepts and ignores extra positional arguments like the reference's Model.forward
(src/models/__init__.py:148-149).
"""
import torch
from torch import nn
import torch.nn.functional as F


class ToyModel(nn.Module):
    def __init__(self, rate=1):
        super().__init__()
        self.rate = rate
        self.w = nn.Parameter(torch.tensor([0.8, 0.15, 0.05]))
        self.c = nn.Parameter(torch.tensor(0.01))

    def forward(self, y, *args):
        u = y
        if self.rate != 1:
            u = F.interpolate(u, scale_factor=self.rate, mode="nearest")
        return self.w[0] * u + self.w[1] * torch.roll(u, (1, 2), (-2, -1)) + self.w[2] * u * u + self.c
