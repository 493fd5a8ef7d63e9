"""Named blur kernels (reference: src/physics/kernels.py:3-28), float64 like the reference."""
import torch

# name -> (family, level)
KERNEL_FAMILIES = {f"Gaussian_R{r}": ("gaussian", r) for r in (1, 2, 3)}
KERNEL_FAMILIES.update({f"Box_R{r}": ("box", r) for r in (2, 3, 4)})


def gaussian_taps(level, dtype=torch.float64):
    """1-D factor of the Gaussian family: size 6*level+1, std = level, sum 1."""
    size = 6 * level + 1
    u = torch.arange(size, dtype=dtype) - (size - 1) / 2
    g = torch.exp(-(u ** 2) / (2 * level ** 2))
    return g / g.sum()


def get_kernel(name, dtype=torch.float64):
    assert name in KERNEL_FAMILIES, f"Unsupported kernel: {name}"
    family, level = KERNEL_FAMILIES[name]
    if family == "gaussian":
        size = 6 * level + 1
        u = torch.arange(size, dtype=dtype) - (size - 1) / 2
        sq = u[:, None] ** 2 + u[None, :] ** 2
        kernel = torch.exp(-sq / (2 * level ** 2))
    else:
        size = 2 * level + 1
        kernel = torch.ones(size, size, dtype=dtype)
    return kernel / kernel.sum()
