"""sei_b200.optim.Adam (csrc/optim.cu) against torch.optim.Adam on identical parameters and gradients: five steps,
fp32 parameters within 2e-6 relative; the bf16 shadows equal the rounded parameters bit for bit; the transposed
shadow equals their transpose; a captured CUDA graph of step() keeps advancing the device-side step count."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_adam_matches_torch(dev):
    from sei_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(513,), (128, 64, 1, 1), (7, 3, 3, 3), (1,), (64, 256)]
    ours = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    lowp = ours[1].detach().reshape(128, 64).to(torch.bfloat16).clone()
    lowp_t = lowp.t().contiguous()
    ours[1]._sei_lowp, ours[1]._sei_lowp_t = lowp, (128, 64, lowp_t)
    calls = []
    ours[2]._sei_invalidate = lambda: calls.append(1)
    a, b = Adam(ours, lr=3e-3, betas=(0.9, 0.99)), torch.optim.Adam(ref, lr=3e-3, betas=(0.9, 0.99))
    for step in range(5):
        for p, q in zip(ours, ref):
            g = torch.randn_like(p) * (1 + step)
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
        for p, q in zip(ours, ref):
            assert float((p - q).abs().max() / q.abs().max()) < 2e-6
        assert torch.equal(lowp, ours[1].detach().reshape(128, 64).to(torch.bfloat16))
        assert torch.equal(lowp_t, lowp.t())
    assert len(calls) == 5
    sd = a.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0


def test_adam_step_in_cuda_graph(dev):
    from sei_b200.optim import Adam
    torch.manual_seed(1)
    p = torch.nn.Parameter(torch.randn(1000, device=dev))
    q = torch.nn.Parameter(p.detach().clone())
    p.grad, q.grad = torch.randn_like(p), None
    q.grad = p.grad.clone()
    a, b = Adam([p], lr=1e-2), torch.optim.Adam([q], lr=1e-2)
    a.step(); b.step()                                   # state initialised eagerly
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph):
            a.step()
    torch.cuda.current_stream().wait_stream(side)       # (capture records the step, it does not run it)
    for _ in range(3):
        graph.replay()
        b.step()
    torch.cuda.synchronize()
    assert float((p - q).abs().max() / q.abs().max()) < 2e-6
    assert float(a.state[p]["step"]) == 4.0


def test_weight_shadows_follow_any_optimizer(dev):
    """the bf16 weight copies the GEMMs read must track the fp32 masters whichever optimizer updates them: torch's
    fused Adam does not bump Tensor._version, sei_b200.optim.Adam rewrites the copy itself"""
    import models.convolutional as mc
    from sei_b200.optim import Adam as SeiAdam
    for make in (lambda ps: torch.optim.Adam(ps, lr=1e-2, fused=True), lambda ps: torch.optim.Adam(ps, lr=1e-2),
                 lambda ps: SeiAdam(ps, lr=1e-2)):
        torch.manual_seed(0)
        conv = mc._GemmConv2d(16, 32, kernel_size=1).to(dev)
        opt = make(conv.parameters())
        x = torch.randn(2, 16, 8, 8, device=dev)
        for _ in range(3):
            opt.zero_grad()
            y = conv(x)
            ref = torch.nn.functional.conv2d(x.bfloat16().float(), conv.weight.detach().bfloat16().float(), conv.bias.detach())
            assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1e-2      # current weights, not stale ones
            y.float().square().mean().backward()
            opt.step()
        _, cache = conv._weight_matrix()
        assert torch.equal(cache, conv.weight.detach().reshape(32, 16).bfloat16())
        if conv._wt_cache is not None and isinstance(opt, SeiAdam):
            assert torch.equal(conv._wt_cache, cache.t())


def test_adam_loads_torch_and_cpu_state_dicts(dev, tmp_path):
    """Optimizer.load_state_dict leaves `step` on the CPU for non-capturable groups, a map_location="cpu"
    checkpoint leaves the whole state there, and old checkpoints store `step` as a Python number.  The kernel reads
    these through device pointers, so step() must coerce them (not dereference a host pointer)."""
    from sei_b200.optim import Adam
    torch.manual_seed(2)
    p = torch.nn.Parameter(torch.randn(256, 8, device=dev))
    q = torch.nn.Parameter(p.detach().clone())
    grads = [torch.randn_like(p) for _ in range(4)]
    ref = torch.optim.Adam([q], lr=1e-2)
    for g in grads[:2]:
        q.grad = g.clone()
        ref.step()
    sd = ref.state_dict()
    path = tmp_path / "opt.pt"
    torch.save(sd, path)
    for variant in ("torch_state", "cpu_checkpoint", "python_step"):
        p2 = torch.nn.Parameter(q.detach().clone())
        ours = Adam([p2], lr=1e-2)
        # (a live state_dict() shares its tensors with the optimizer it came from: load a deep copy, as a resume would)
        state = torch.load(path, map_location="cpu") if variant != "torch_state" else copy.deepcopy(ref.state_dict())
        if variant == "python_step":
            state["state"][0]["step"] = float(state["state"][0]["step"])
        ours.load_state_dict(state)
        q3 = torch.nn.Parameter(q.detach().clone())
        ref3 = torch.optim.Adam([q3], lr=1e-2)
        ref3.load_state_dict(copy.deepcopy(ref.state_dict()))
        for g in grads[2:]:
            p2.grad, q3.grad = g.clone(), g.clone()
            ours.step()
            ref3.step()
        torch.cuda.synchronize()
        assert float((p2 - q3).abs().max() / q3.abs().max()) < 2e-6, variant
        st = ours.state[p2]
        assert st["step"].device == p2.device and float(st["step"]) == 4.0
        assert st["exp_avg"].is_cuda and st["exp_avg_sq"].is_cuda
