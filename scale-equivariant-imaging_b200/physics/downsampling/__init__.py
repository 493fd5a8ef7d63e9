"""SR physics (reference: src/physics/downsampling/__init__.py): antialiased bicubic decimation by
`rate`, its true transpose (autograd backward / true_adjoint=True) and the deprecated plain
bicubic upsample the reference returns from A_adjoint by default."""
from sei_b200 import ops
from sei_b200.linear_physics import LinearPhysics


class Downsampling(LinearPhysics):
    def __init__(self, rate, antialias, true_adjoint=False):
        super().__init__()
        self.rate = rate
        self.antialias = antialias
        self.true_adjoint = true_adjoint
        if not antialias:
            raise NotImplementedError("Downsampling(antialias=False) is not built; the reference's factory "
                                      "always passes antialias=True (src/physics/__init__.py:48)")

    def A(self, x):
        return ops._DownAA.apply(x, self.rate, None, 0.0, ops.PATH_AUTO)

    def A_adjoint(self, y):
        if self.true_adjoint:
            in_hw = (y.shape[2] * self.rate, y.shape[3] * self.rate)
            return ops._DownAATranspose.apply(y, self.rate, in_hw, ops.PATH_AUTO)
        return ops.up_bicubic(y, self.rate)

    def measure_with_noise(self, x, noise):
        return ops._DownAA.apply(x, self.rate, noise, self.noise_model.sigma_value(), ops.PATH_AUTO)

    def ei_remeasure_args(self):
        return dict(kernel_host=None, rate_sr=self.rate)
