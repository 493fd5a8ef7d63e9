"""The host logic of the reference-facing mirrors (physics, transforms, losses, sei_b200.linear_physics and the autograd
Functions of sei_b200.ops) run end to end on the CPU, with the raw CUDA operators replaced by oracle-backed stand-ins
(tests/fake_ops.py): draw order, fused / unfused dispatch, loss assembly and the forward / backward pairing of every
Function are checked against the fixtures produced by the reference for EVERY loss configuration -- the same assertions
as tests/test_gpu_parity.py::test_loss_and_gradients_match_reference, which runs the real kernels."""
import numpy as np
import pytest
import torch

import fake_ops
from test_gpu_parity import LOSS_ARGS, LOSS_CASES, Tap, base_args
from util import rel_err

TOL = 1e-5


@pytest.fixture
def cpu_ops(monkeypatch):
    fake_ops.install(monkeypatch)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_host_path_of_every_loss_configuration(golden, cpu_ops, name):
    import losses
    import physics
    from sei_b200 import draws
    from toy_model import ToyModel
    g = golden(f"loss_{name}_f32")
    args = base_args(**LOSS_ARGS[name])
    phys = physics.get_physics(args, device="cpu")
    loss_fn = losses.get_loss(args=args, physics=phys)
    model = Tap(ToyModel(rate=int(g["rate"])))
    injected = [g[k] for k in sorted((k for k in g if k.startswith("draw")), key=lambda s: int(s[4:].split("_")[0]))]
    with draws.inject(injected):                      # raises if a draw is missing, left over or of the wrong shape
        loss = loss_fn(x=torch.from_numpy(g["x"]), y=torch.from_numpy(g["y"]), model=model)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    for i, o in enumerate(model.outs):
        assert rel_err(o.detach().numpy(), g[f"model_out{i}"]) < TOL, f"model_out{i}"
        if f"model_out{i}_grad" in g:
            assert rel_err(o.grad.numpy(), g[f"model_out{i}_grad"]) < 2 * TOL, f"model_out{i}_grad"
    for pname, p in model.model.named_parameters():
        ref = g[f"grad_{pname}"]
        assert np.allclose(p.grad.numpy(), ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max() + 1e-9), pname


def test_host_path_physics_and_transforms(golden, cpu_ops):
    """factories -> operators: A, A_adjoint, physics(x) with its noise draw, the batched dataset degradation and the module
    form of the scale transform, on the CPU stand-ins"""
    import physics
    import transforms
    from sei_b200 import draws
    g = golden("transform")
    with draws.inject([g["module_draw1_rand"], g["module_draw2_rand"]]):      # draw0 is the image itself
        y = transforms.ScalingTransform(kind="padded", antialias=False)(torch.from_numpy(g["module_x"]))
    assert rel_err(y.numpy(), g["module_T"]) < TOL
    for kw in (dict(), dict(physics_v2=False), dict(task="sr", kernel=None, sr_factor=2)):
        phys = physics.get_physics(base_args(**kw), device="cpu")
        mgr = getattr(phys, "__manager")
        r = getattr(phys, "rate", 1)
        x = torch.rand(3, 3, 16 * r, 16 * r)
        v = torch.randn(3, 3, 16, 16)
        xg = x.clone().requires_grad_(True)
        (phys.A(xg) * v).sum().backward()                         # autograd of A = the transpose operator
        lhs, rhs = float((phys.A(x) * v).sum()), float((x * xg.grad).sum())
        assert abs(lhs - rhs) < 1e-4 * abs(lhs)
        yb = mgr.randomly_degrade_batch(x, [5, 6, 5])
        ys = torch.cat([mgr.randomly_degrade(x[i:i + 1], seed=s) for i, s in enumerate([5, 6, 5])])
        assert float((yb - ys).abs().max()) < 5e-7
        n0, n2 = yb[0] - phys.A(x[:1])[0], yb[2] - phys.A(x[2:3])[0]
        assert float((n0 - n2).abs().max()) < 5e-7                # same seed, same noise
