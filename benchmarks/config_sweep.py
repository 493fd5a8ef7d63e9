#!/usr/bin/env python
"""Step throughput over the BASELINE.json configurations other than bench.py's headline one: SR x2 / x4,
Box_R3 deblurring, and the method ablations (supervised, css, sure, proposed + Shifts) at 512 x 512 -- each a
full training step (loss forward, backward, Adam) through the reference-facing API, CUDA-graph captured,
timed with CUDA events.  One markdown row per configuration.

    python benchmarks/config_sweep.py [--network cnn|standin] [--only sr2]
    torchrun --nproc-per-node N benchmarks/config_sweep.py ...      (data parallel, aggregate imgs/s)
"""
import argparse
import gc
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

CONFIGS = [
    # name, task, kernel, sr, method, transforms, measurement size, batch per GPU
    ("cfg2 deblur Gaussian_R2 proposed 256 b32", "deblurring", "Gaussian_R2", None, "proposed", "Scaling_Transforms", 256, 32),
    ("cfg3 SR x2 proposed 256 b8/GPU", "sr", None, 2, "proposed", "Scaling_Transforms", 256, 8),
    ("cfg4 SR x4 proposed 256 b2", "sr", None, 4, "proposed", "Scaling_Transforms", 256, 2),
    ("cfg4 SR x4 proposed 256 b8", "sr", None, 4, "proposed", "Scaling_Transforms", 256, 8),
    ("cfg4 deblur Box_R3 proposed 256 b32", "deblurring", "Box_R3", None, "proposed", "Scaling_Transforms", 256, 32),
    ("cfg5 supervised 512 b16", "deblurring", "Gaussian_R2", None, "supervised", "Scaling_Transforms", 512, 16),
    ("cfg5 css 512 b16", "deblurring", "Gaussian_R2", None, "css", "Scaling_Transforms", 512, 16),
    ("cfg5 sure 512 b16", "deblurring", "Gaussian_R2", None, "sure", "Scaling_Transforms", 512, 16),
    ("cfg5 proposed+Shifts (ei-shift) 512 b16", "deblurring", "Gaussian_R2", None, "proposed", "Shifts", 512, 16),
    ("cfg5 proposed (scale) 512 b16", "deblurring", "Gaussian_R2", None, "proposed", "Scaling_Transforms", 512, 16),
]


def run(cfg, args, dev, rank, world):
    import torch.distributed as dist
    import losses
    import models
    import physics
    from sei_b200 import parallel
    from toy_model import ToyModel
    name, task, kernel, sr, method, transforms, size, batch = cfg
    largs = Namespace(task=task, noise_level=5, physics_v2=True, kernel=kernel, sr_factor=sr, physics_true_adjoint=False,
                      partial_sure=True, sure_margin=None, partial_sure_sr=False, Loss__crop_training_pairs=False,
                      Loss__crop_size=48, ProposedLoss__stop_gradient=True, ProposedLoss__sure_alternative=None,
                      ProposedLoss__alpha_tradeoff=1.0, ProposedLoss__transforms=transforms,
                      ScalingTransform__kind="padded", ScalingTransform__antialias=False, method=method,
                      sure_cropped_div=True, sure_averaged_cst=None)
    torch.manual_seed(0)
    phys = physics.get_physics(largs, device=dev)
    loss_fn = losses.get_loss(largs, phys)
    rate = sr or 1
    if args.network == "cnn":
        margs = Namespace(task=task, sr_factor=sr, noise_level=5, model_kind="Proposed",
                          ProposedModel__architecture="Convolutional", ConvolutionalModel__residual=True,
                          ConvolutionalModel__inner_residual=True, ConvolutionalModel__inout_convs=True,
                          ConvolutionalModel__hidden_channels=args.cnn_hidden, ConvolutionalModel__scales=args.cnn_scales,
                          ConvolutionalModel__num_conv_blocks=1, data_parallel_devices=None)
        model = models.get_model(margs, physics=phys, device=dev).to(dev)
    else:
        model = ToyModel(rate=rate).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True, fused=True)
    reducer = parallel.GradAllReducer(model.parameters())
    torch.manual_seed(1 + rank)
    x = torch.rand(batch, 3, size * rate, size * rate, device=dev)
    y = phys(x)

    def fwd_bwd():
        opt.zero_grad(set_to_none=False)
        loss_fn(x=x, y=y, model=model).backward()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fwd_bwd(); reducer(); opt.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gc.collect(); torch.cuda.empty_cache()
    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1, stream=side):
        fwd_bwd()
    with torch.cuda.graph(g2, stream=side):
        opt.step()

    def step():
        g1.replay(); reducer(); g2.replay()

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del g1, g2, model, opt, x, y
    gc.collect(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats(dev)
    return float(ms), batch * world / (float(ms) * 1e-3), peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--network", default="cnn", choices=["cnn", "standin"])
    ap.add_argument("--cnn-hidden", type=int, default=32)
    ap.add_argument("--cnn-scales", type=int, default=5)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from sei_b200 import parallel
    rank, world, local = parallel.init_distributed(backend="nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rows = []
    for cfg in CONFIGS:
        if args.only and args.only not in cfg[0]:
            continue
        try:
            ms, ips, peak = run(cfg, args, dev, rank, world)
            rows.append((cfg[0], f"{ms:.2f}", f"{ips:.1f}", f"{peak:.1f}"))
        except Exception as e:  # noqa: BLE001
            rows.append((cfg[0], "failed", f"{type(e).__name__}: {str(e)[:90]}", ""))
            gc.collect(); torch.cuda.empty_cache()
    if rank == 0:
        print(f"# network={args.network} (hidden {args.cnn_hidden}, scales {args.cnn_scales}), {world} GPU(s), {args.steps} timed steps")
        print("| config | ms/step | imgs/s (all GPUs) | peak GiB |")
        print("|---|---|---|---|")
        for r in rows:
            print("| " + " | ".join(r) + " |")
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
