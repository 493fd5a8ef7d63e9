#!/usr/bin/env python
"""Kernel-time breakdown of one eager `proposed` training step with the CNN (torch.profiler, CUDA activity):
which kernels the step's GPU time goes to.  Shares only -- absolute numbers come from bench.py."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--cnn-hidden", type=int, default=32)
    ap.add_argument("--cnn-scales", type=int, default=5)
    ap.add_argument("--rows", type=int, default=45)
    ap.add_argument("--torch-adam", action="store_true")
    ap.add_argument("--sr", type=int, default=0, help="SR factor (0: deblurring)")
    ap.add_argument("--copies", action="store_true", help="attribute copy / cat / add kernels to source lines")
    args = ap.parse_args()
    import losses
    import models
    import physics
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    largs, margs = bench.loss_args(), bench.model_args(args.cnn_hidden, args.cnn_scales)
    if args.sr:                                     # BASELINE configs[2]: SR x2 / x4 instead of deblurring
        for a in (largs, margs):
            a.task, a.sr_factor = "sr", args.sr
        largs.kernel = None
    phys = physics.get_physics(largs, device=dev)
    loss_fn = losses.get_loss(largs, phys)
    model = models.get_model(margs, physics=phys, device=dev).to(dev)
    from sei_b200.optim import Adam as SeiAdam
    opt = SeiAdam(model.parameters(), lr=1e-4) if not args.torch_adam else torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = torch.rand(args.batch, 3, 256 * (args.sr or 1), 256 * (args.sr or 1), device=dev)
    y = phys(x)

    def step():
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(x=x, y=y, model=model)
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if args.copies:
        # attribute the layout / dtype copies to the Python lines that issue them
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
            step()
            torch.cuda.synchronize()
        evs = [e for e in prof.key_averages(group_by_stack_n=8) if e.key in ("aten::copy_", "aten::cat", "aten::add", "aten::add_", "aten::fill_", "aten::zero_")]
        for e in sorted(evs, key=lambda e: -e.self_device_time_total)[: args.rows]:
            frames = [f for f in e.stack if "site-packages" not in f][:4]
            print(f"{e.self_device_time_total / 1e3:8.2f} ms  x{e.count:<4d} {e.key:12s} " + " <- ".join(f.strip()[-70:] for f in frames))
        return
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = prof.key_averages()
    total = sum(e.device_time_total for e in evs)
    print(f"# batch {args.batch}, total CUDA time of one step: {total / 1e3:.1f} ms")
    print("| share | calls | total ms | kernel |")
    print("|---|---|---|---|")
    for e in sorted(evs, key=lambda e: -e.device_time_total)[: args.rows]:
        print(f"| {100 * e.device_time_total / total:.1f}% | {e.count} | {e.device_time_total / 1e3:.2f} | {e.key[:110]} |")


if __name__ == "__main__":
    main()
