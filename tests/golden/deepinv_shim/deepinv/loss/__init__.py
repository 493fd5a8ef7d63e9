import torch
from . import metric  # noqa: F401


class SupLoss(torch.nn.Module):
    """deepinv v0.2.0 loss/sup.py: metric(x_net, x)."""

    def __init__(self, metric=torch.nn.MSELoss()):
        super().__init__()
        self.name = "sup"
        self.metric = metric

    def forward(self, x_net, x, **kwargs):
        return self.metric(x_net, x)


class EILoss(torch.nn.Module):
    """deepinv v0.2.0 loss/ei.py."""

    def __init__(self, transform, metric=torch.nn.MSELoss(), apply_noise=True, weight=1.0, no_grad=True):
        super().__init__()
        self.name = "ei"
        self.metric = metric
        self.weight = weight
        self.T = transform
        self.noise = apply_noise
        self.no_grad = no_grad

    def forward(self, x_net, physics, model, **kwargs):
        if self.no_grad:
            with torch.no_grad():
                x2 = self.T(x_net)
        else:
            x2 = self.T(x_net)
        if self.noise:
            y = physics(x2)
        else:
            y = physics.A(x2)
        x3 = model(y, physics)
        return self.weight * self.metric(x3, x2)
