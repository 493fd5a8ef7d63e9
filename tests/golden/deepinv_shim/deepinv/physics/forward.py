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
        """Least-squares pseudo-inverse by conjugate gradient on A^T A x = A^T y
        (deepinv v0.2.0 uses its conjugate_gradient helper with max_iter / tol)."""
        b = self.A_adjoint(y)
        x = torch.zeros_like(b)
        r = b.clone()
        p = r
        rs = (r * r).flatten().sum()
        tol2 = self.tol ** 2
        for _ in range(int(self.max_iter)):
            Ap = self.A_adjoint(self.A(p))
            alpha = rs / (p * Ap).flatten().sum()
            x = x + p * alpha
            r = r + Ap * (-alpha)
            rs_new = (r * r).flatten().sum()
            if rs_new < tol2:
                break
            p = r + p * (rs_new / rs)
            rs = rs_new
        return x
