"""torch.optim.Adam's update rule (reference: demo/train.py:11,167-186 `Adam(params, lr=lr, betas=(0.9, beta2))`,
`optimizer.step()` :266) as one streaming kernel per parameter tensor (csrc/optim.cu).

Same constructor arguments and state-dict layout as torch.optim.Adam (`step`, `exp_avg`, `exp_avg_sq` per parameter),
no weight decay / amsgrad / maximize (the reference uses none of them).  The step count lives in device memory, so a
captured CUDA graph of `step()` replays correctly; the learning rate is a launch argument (re-capture after changing it).

Parameters may carry low-precision shadows registered by the modules that consume them
(`p._sei_lowp`: bf16 copy with the parameter's element order; `p._sei_lowp_t`: (rows, cols, tensor) of its transposed
copy): they are refreshed inside the same pass / by a tiled transpose, so the forward pass never re-casts weights.
Modules whose cached copy is NOT maintained here register `p._sei_invalidate` (called after every step).
"""
import torch
from torch.optim.optimizer import register_optimizer_step_post_hook

from . import _lib
from .ops import SeiError, _ptr, check


# Every optimizer step (of any torch optimizer) advances this counter; modules that cache low-precision copies of
# their parameters include it in the cache key unless the copy is maintained by the Adam below.  (A parameter's
# `_version` alone is not enough: torch's fused Adam updates parameters without bumping it.)
OPT_EPOCH = [0]


def _post_step(optimizer, args, kwargs):
    OPT_EPOCH[0] += 1
    if not isinstance(optimizer, Adam):
        for group in optimizer.param_groups:
            for p in group["params"]:
                if getattr(p, "_sei_maintained", False):
                    p._sei_maintained = False


register_optimizer_step_post_hook(_post_step)


def shadow_epoch(p):
    """cache-key component for a low-precision copy of parameter p (None: kept fresh by sei_b200.optim.Adam)"""
    return None if getattr(p, "_sei_maintained", False) else OPT_EPOCH[0]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unused):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("sei_b200.optim.Adam: weight_decay / amsgrad are not used by the reference and not built")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _init_state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            return st
        # State that came through load_state_dict: torch leaves `step` on the CPU (or as a Python number in older
        # checkpoints) unless the group is capturable / fused, and a map_location="cpu" checkpoint leaves everything
        # there.  The kernel dereferences these pointers on the device, so they are coerced / checked here.
        step = st.get("step", 0.0)
        if not (torch.is_tensor(step) and step.device == p.device and step.dtype == torch.float32 and step.dim() == 0):
            st["step"] = torch.as_tensor(float(step), dtype=torch.float32).to(p.device)
        for name in ("exp_avg", "exp_avg_sq"):
            t = st.get(name)
            if t is None:
                raise SeiError(f"sei_b200.optim.Adam: optimizer state lacks '{name}'")
            if t.device != p.device or t.dtype != torch.float32 or not t.is_contiguous():
                st[name] = t.to(device=p.device, dtype=torch.float32).contiguous()
            if st[name].shape != p.shape:
                raise SeiError(f"sei_b200.optim.Adam: state '{name}' has shape {tuple(st[name].shape)}, parameter {tuple(p.shape)}")
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
            todo = [p for p in group["params"] if p.grad is not None]
            if not todo:
                continue
            steps = [self._init_state(p)["step"] for p in todo]
            torch._foreach_add_(steps, 1.0)
            for p in todo:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise SeiError("sei_b200.optim.Adam needs contiguous float32 CUDA parameters (no CPU fallback)")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                lowp = getattr(p, "_sei_lowp", None)
                stream = torch.cuda.current_stream(p.device).cuda_stream
                with torch.cuda.device(p.device):
                    check(lib.sei_adam_step_f32(_ptr(p), _ptr(g), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]), _ptr(lowp),
                                                _ptr(st["step"]), p.numel(), lr, float(b1), float(b2), eps, stream))
                    lowp_t = getattr(p, "_sei_lowp_t", None)
                    if lowp is not None and lowp_t is not None:
                        rows, cols, dst = lowp_t
                        check(lib.sei_transpose_bf16(_ptr(lowp), _ptr(dst), rows, cols, stream))
                if lowp is not None:
                    p._sei_maintained = True
                inval = getattr(p, "_sei_invalidate", None)
                if inval is not None:
                    inval()
        return loss
