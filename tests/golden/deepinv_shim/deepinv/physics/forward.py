import torch


class Physics(torch.nn.Module):
    """deepinv v0.2.0: forward(x) = sensor(noise(A(x))), identity sensor by default."""

    def __init__(self, A=lambda x: x, noise_model=lambda x: x, sensor_model=lambda x: x, **kwargs):
        super().__init__()
        self.noise_model = noise_model
        self.sensor_model = sensor_model
        self.forw = A

    def forward(self, x):
        return self.sensor(self.noise(self.A(x)))

    def A(self, x):
        return self.forw(x)

    def sensor(self, x):
        return self.sensor_model(x)

    def noise(self, x):
        return self.noise_model(x)


class LinearPhysics(Physics):
    def __init__(self, A=lambda x: x, A_adjoint=lambda x: x, noise_model=lambda x: x,
                 sensor_model=lambda x: x, max_iter=50, tol=1e-3, **kwargs):
        super().__init__(A=A, noise_model=noise_model, sensor_model=sensor_model)
        self.max_iter = max_iter
        self.tol = tol
        self.adjoint = A_adjoint

    def A_adjoint(self, y):
        return self.adjoint(y)

    def A_dagger(self, y):
        """Least-squares pseudo-inverse by conjugate gradient, as deepinv v0.2.0 states it (RECALLED, not vendored):
        when A^T y is smaller than y (an overcomplete system) solve A^T A x = A^T y; otherwise -- deblurring (equal
        sizes) and super-resolution -- solve A A^T z = y and return A^T z.  max_iter / tol as attributes."""
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
    """deepinv.optim.utils.conjugate_gradient (v0.2.0, RECALLED): plain CG from x0 = 0; stops once |r| < tol"""
    def dot(s1, s2):
        return (s1.conj() * s2).flatten().sum()

    x = torch.zeros_like(b)
    r = b
    p = r
    rsold = dot(r, r)
    for _ in range(int(max_iter)):
        Ap = A(p)
        alpha = rsold / dot(p, Ap)
        x = x + p * alpha
        r = r + Ap * (-alpha)
        rsnew = torch.real(dot(r, r))
        if rsnew.sqrt() < tol:
            break
        p = r + p * (rsnew / rsold)
        rsold = rsnew
    return x
