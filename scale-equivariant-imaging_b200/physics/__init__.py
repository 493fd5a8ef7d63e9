"""Degradation physics factory (reference: src/physics/__init__.py).  Same entry points and
attributes -- get_physics(args, device), PhysicsManager, BlurKernel, physics.task,
getattr(physics, "__manager").randomly_degrade -- on the libsei_b200 operators."""
from os.path import exists

import torch

from rng import fork_rng
from sei_b200 import draws, ops
from sei_b200.linear_physics import GaussianNoise
from .blur import Blur, BlurV2
from .downsampling import Downsampling
from .kernels import get_kernel


class BlurKernel:
    """A kernel given either by a .pt file or by one of the names of physics.kernels."""

    def __init__(self, kernel_path):
        self.kernel_path = kernel_path

    def to_tensor(self, device):
        kernel = torch.load(self.kernel_path) if exists(self.kernel_path) else get_kernel(name=self.kernel_path)
        return kernel[None, None].to(device)


class CTLikeFilter:
    def __init__(self, *a, **k):
        raise NotImplementedError("the CT-like filter task is outside the scope of this package (SURVEY.md section 2, row 5)")


class PhysicsManager:
    def __init__(self, blueprint, task, device, noise_level, v2):
        if task == "deblurring":
            kernel = BlurKernel(**blueprint[BlurKernel.__name__]).to_tensor(device)
            physics = BlurV2(kernel=kernel) if v2 else Blur(filter=kernel, padding="circular", device=device)
        elif task == "sr":
            physics = Downsampling(antialias=True, **blueprint[Downsampling.__name__])
        elif task == "invert_a_tomography_like_filter":
            physics = CTLikeFilter()
        else:
            raise ValueError(f"Unknown task: {task}")

        physics.noise_model = GaussianNoise(sigma=noise_level / 255).to(device)
        self.task = task
        physics.task = task
        setattr(physics, "__manager", self)
        self.physics = physics

    def get_physics(self):
        return self.physics

    def randomly_degrade(self, x, seed):
        """A then noise; with a seed the global RNG is forked, seeded and restored (reference :65-74)."""
        with fork_rng(enabled=seed is not None):
            if seed is not None:
                torch.manual_seed(seed)
            return self.physics.noise_model(self.physics.A(x))

    def randomly_degrade_batch(self, x, seeds):
        """The measurements the reference's datasets produce one image at a time (SyntheticDataset.__getitem__,
        src/datasets/synthetic_dataset.py:26-40: x.unsqueeze(0) -> randomly_degrade(x, seed)), for a whole batch of
        equally sized images with ONE operator launch: image i gets the standard-normal draw of
        `torch.manual_seed(seeds[i]); randn((1, C, h, w))` (or the next draw of the global stream where seeds[i] is
        None / seeds is None, like the CSS re-degradation, src/datasets/__init__.py:70-75), and the noise add is the
        operator's epilogue."""
        B = x.shape[0]
        seeds = [None] * B if seeds is None else list(seeds)
        if len(seeds) != B:
            raise ValueError(f"{B} images but {len(seeds)} seeds")
        y_shape = tuple(x.shape)
        if self.task == "sr":
            r = self.physics.rate
            y_shape = y_shape[:2] + (ops.down_out_size(x.shape[2], r), ops.down_out_size(x.shape[3], r))
        noise = torch.empty(y_shape, device=x.device, dtype=x.dtype)
        for i, seed in enumerate(seeds):
            with fork_rng(enabled=seed is not None):
                if seed is not None:
                    torch.manual_seed(seed)
                noise[i:i + 1] = draws.randn((1,) + tuple(y_shape[1:]), x.device, x.dtype)
        return self.physics.measure_with_noise(x, noise)


def get_physics(args, device):
    blueprint = {
        PhysicsManager.__name__: dict(task=args.task, noise_level=args.noise_level, v2=args.physics_v2),
        BlurKernel.__name__: dict(kernel_path=args.kernel),
        Downsampling.__name__: dict(rate=args.sr_factor, true_adjoint=args.physics_true_adjoint),
    }
    manager = PhysicsManager(blueprint=blueprint, device=device, **blueprint[PhysicsManager.__name__])
    return manager.get_physics()
