import torch
from .forward import LinearPhysics, Physics  # noqa: F401


class GaussianNoise(torch.nn.Module):
    """deepinv v0.2.0 physics/noise.py: y = x + randn_like(x) * sigma."""

    def __init__(self, sigma=0.1):
        super().__init__()
        self.sigma = torch.nn.Parameter(torch.tensor(sigma), requires_grad=False)

    def forward(self, x):
        return x + torch.randn_like(x) * self.sigma


def adjoint_function(A, input_size, device="cpu", dtype=torch.float):
    """deepinv v0.2.0 physics/forward.py: vjp of A taken at ones(input_size)."""
    x = torch.ones(input_size, device=device, dtype=dtype)
    (_, vjpfunc) = torch.func.vjp(A, x)
    batches = x.size()[0]

    def adjoint(y):
        if y.size()[0] < batches:
            y2 = torch.zeros((batches,) + y.size()[1:], device=y.device, dtype=y.dtype)
            y2[: y.size()[0], ...] = y
            return vjpfunc(y2)[0][: y.size()[0], ...]
        elif y.size()[0] > batches:
            raise ValueError("Batch size of A_adjoint input is larger than expected")
        return vjpfunc(y)[0]

    return adjoint
