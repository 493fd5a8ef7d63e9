#!/usr/bin/env python
"""Which Python lines issue large tensor copies during one training step (contiguous / clone / copy_ / to / pad that
really move > 4 MB)?  torch.profiler's stacks are empty in this build, so the tensor methods are wrapped instead."""
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

HITS = collections.Counter()
BYTES = collections.Counter()


def site():
    for fr in reversed(traceback.extract_stack()[:-2]):
        if "site-packages" not in fr.filename and "find_copies" not in fr.filename:
            return f"{os.path.relpath(fr.filename, ROOT)}:{fr.lineno}"
    return "?"


def wrap(owner, name, moved):
    orig = getattr(owner, name)

    def f(*a, **k):
        out = orig(*a, **k)
        try:
            t = a[0]
            if torch.is_tensor(t) and torch.is_tensor(out) and t.is_cuda and out.numel() * out.element_size() > (4 << 20) and moved(t, out):
                s = f"{name} @ {site()}"
                HITS[s] += 1
                BYTES[s] += out.numel() * out.element_size()
        except Exception:  # noqa: BLE001
            pass
        return out
    setattr(owner, name, f)


def main():
    from argparse import Namespace
    import losses, models, physics  # noqa: E401
    from sei_b200.optim import Adam as SeiAdam
    dev = torch.device("cuda:0")
    largs = Namespace(task="deblurring", noise_level=5, physics_v2=True, kernel="Gaussian_R2", sr_factor=None,
                      physics_true_adjoint=False, partial_sure=True, sure_margin=None, partial_sure_sr=False,
                      Loss__crop_training_pairs=False, Loss__crop_size=48, ProposedLoss__stop_gradient=True,
                      ProposedLoss__sure_alternative=None, ProposedLoss__alpha_tradeoff=1.0,
                      ProposedLoss__transforms="Scaling_Transforms", ScalingTransform__kind="padded",
                      ScalingTransform__antialias=False, method="proposed", sure_cropped_div=True, sure_averaged_cst=None)
    phys = physics.get_physics(largs, device=dev)
    loss_fn = losses.get_loss(largs, phys)
    margs = Namespace(task="deblurring", sr_factor=None, noise_level=5, model_kind="Proposed",
                      ProposedModel__architecture="Convolutional", ConvolutionalModel__residual=True,
                      ConvolutionalModel__inner_residual=True, ConvolutionalModel__inout_convs=True,
                      ConvolutionalModel__hidden_channels=32, ConvolutionalModel__scales=5,
                      ConvolutionalModel__num_conv_blocks=1, data_parallel_devices=None)
    model = models.get_model(margs, physics=phys, device=dev).to(dev)
    opt = SeiAdam(model.parameters(), lr=1e-4)
    x = torch.rand(8, 3, 256, 256, device=dev)
    y = phys(x)

    def step():
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(x=x, y=y, model=model)
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    diff = lambda t, o: o.data_ptr() != t.data_ptr()
    wrap(torch.Tensor, "contiguous", diff)
    wrap(torch.Tensor, "clone", lambda t, o: True)
    wrap(torch.Tensor, "copy_", lambda t, o: True)
    wrap(torch.Tensor, "to", diff)
    wrap(torch.Tensor, "float", diff)
    wrap(F, "pad", lambda t, o: True)
    step()
    torch.cuda.synchronize()
    print("| copies | MB | call site |")
    print("|---|---|---|")
    for s, n in sorted(HITS.items(), key=lambda kv: -BYTES[kv[0]]):
        print(f"| {n} | {BYTES[s] / 1e6:.0f} | {s} |")


if __name__ == "__main__":
    main()
