"""Random draws of the hot path, with an injection hook for parity runs.

The reference draws with torch.rand / torch.randn / torch.randn_like on the tensors' device
inside the operators (SURVEY.md section 3.1 lists the order: SURE probe, transform rates,
transform centres, EI measurement noise).  By default the same torch calls are made here, in
the same order and shapes, so a seeded run consumes the generator identically.  Parity tests
replace the draws by the tensors the reference actually drew (recorded in tests/golden)."""
from contextlib import contextmanager

import torch

_queue = None


@contextmanager
def inject(tensors):
    """Within the context every draw pops the next tensor of `tensors` instead of sampling."""
    global _queue
    prev, _queue = _queue, list(tensors)
    try:
        yield
        if _queue:
            raise RuntimeError(f"{len(_queue)} injected draws were not consumed")
    finally:
        _queue = prev


def _next(shape, device, dtype):
    t = _queue.pop(0)
    t = torch.as_tensor(t).to(device=device, dtype=dtype)
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"injected draw has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def rand(shape, device, dtype):
    if _queue is not None:
        return _next(shape, device, dtype)
    return torch.rand(shape, device=device, dtype=dtype)


def randn(shape, device, dtype):
    if _queue is not None:
        return _next(shape, device, dtype)
    return torch.randn(shape, device=device, dtype=dtype)


def randn_like(x):
    return randn(tuple(x.shape), x.device, x.dtype)


def randperm(n):
    """torch.randperm(n) on the CPU generator (deepinv's Shift / Rotate draw their group element this way)"""
    if _queue is not None:
        return _next((n,), "cpu", torch.int64)
    return torch.randperm(n)


def randint(low, high):
    """torch.randint(low, high, size=(1,)).item() on the CPU generator (CropPair's offsets, reference src/crop.py:26-27)"""
    if _queue is not None:
        return int(_next((1,), "cpu", torch.int64).item())
    return int(torch.randint(low, high, size=(1,)).item())
