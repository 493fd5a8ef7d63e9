"""sei_b200 -- Python binding of libsei_b200.so, the B200 (sm_100a) implementation of the
Scale-Equivariant-Imaging per-step hot path.  The reference-facing modules (physics, transforms,
losses, crop, rng, training) live next to this package and import it."""
from . import _lib, draws, ops  # noqa: F401
from ._lib import SeiError, launch_count, last_kernel  # noqa: F401
