#!/usr/bin/env python
"""Operator micro-benchmark (BASELINE.json configs[3]): physics / transform / loss kernels of
libsei_b200 vs the HBM roofline, next to the stock-PyTorch formulation the reference uses on the
same GPU (FFT blur, F.interpolate(antialias), F.grid_sample) -- "the existing Blackwell kernel to
beat".  Inputs are larger than L2 (rotating buffers), timing by CUDA events after warm-up.

    python benchmarks/op_sweep.py [--batch 128] [--reps 20] [--only blur] [--json out.jsonl]
    torchrun --nproc-per-node N benchmarks/op_sweep.py      (every rank runs the sweep on its GPU;
                                                              rank 0 prints aggregate GB/s = sum over ranks / max time)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def time_us(fn, n_rot, reps, warmup=3, graph=True):
    """device time per call: `reps` back-to-back calls over rotating inputs, captured in a CUDA graph so that the host's
    per-call cost (ctypes + torch.empty, 15 - 30 us: more than several of these kernels take) is not what is measured"""
    for i in range(warmup):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                with torch.cuda.graph(g):
                    for i in range(reps):
                        fn(i % n_rot)
            torch.cuda.current_stream().wait_stream(side)
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / reps
        except Exception:  # noqa: BLE001  (an op that cannot be captured, e.g. torch.fft plans on first use)
            torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--json", default="")
    args = ap.parse_args()

    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from physics.kernels import get_kernel
    from sei_b200 import ops
    peak, peak_src = peak_gbs()
    B, S, C, NR = args.batch, args.size, 3, 3
    P = B * C * S * S
    torch.manual_seed(rank)
    rows = []

    def add(name, alg_bytes, fn, torch_fn=None):
        if args.only and args.only not in name:
            return
        us = time_us(fn, NR, args.reps)
        t_us = time_us(torch_fn, NR, max(3, args.reps // 4)) if (torch_fn and not args.no_torch) else None
        if world > 1:
            t = torch.tensor([us], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t)
        gbs = alg_bytes * world / us / 1e3
        rows.append(dict(op=name, us=round(us, 2), alg_MB=round(alg_bytes / 1e6, 1), GBs=round(gbs, 1),
                         frac_of_peak=round(gbs / (peak * world), 3), torch_us=None if t_us is None else round(t_us, 2),
                         speedup_vs_torch=None if t_us is None else round(t_us / us, 2), n_gpus=world))

    ys = [torch.rand(B, C, S, S, device=dev) for _ in range(NR)]
    ns = [torch.randn(B, C, S, S, device=dev) for _ in range(NR)]
    sigma = 5 / 255

    def fft_blur(x, k):  # the reference's BlurV2.A formulation (src/physics/blur/__init__.py:205-223)
        psf = torch.zeros(x.shape[-2:], device=x.device, dtype=x.dtype)
        psf[:k.shape[-2], :k.shape[-1]] = k
        psf = psf.roll((-(k.shape[-2] // 2), -(k.shape[-1] // 2)), dims=(-2, -1))
        return torch.fft.irfft2(torch.fft.rfft2(psf) * torch.fft.rfft2(x), s=x.shape[-2:])

    for kname in ("Gaussian_R2", "Box_R3", "Gaussian_R3"):
        k64 = get_kernel(kname)
        kh = ops.kernel_to_host(k64)
        kd = k64.to(dev, torch.float32)
        add(f"blur {kname} A", 8 * P, lambda i: ops.blur_circular(ys[i], kh), lambda i: fft_blur(ys[i], kd))
        add(f"blur {kname} A^T", 8 * P, lambda i: ops.blur_circular(ys[i], kh, adjoint=True))
        add(f"blur {kname} A+noise", 12 * P, lambda i: ops.blur_circular(ys[i], kh, noise=ns[i], sigma=sigma),
            lambda i: fft_blur(ys[i], kd) + ns[i] * sigma)

    for r in (2, 4):
        b = max(1, B // (r * r))
        xs = [torch.rand(b, C, S * r, S * r, device=dev) for _ in range(NR)]
        gs = [torch.randn(b, C, S, S, device=dev) for _ in range(NR)]
        Pr = b * C * S * S
        add(f"SR x{r} A", 4 * (r * r + 1) * Pr, lambda i: ops.down_aa(xs[i], r),
            lambda i: F.interpolate(xs[i], scale_factor=1 / r, mode="bicubic", antialias=True))
        add(f"SR x{r} A^T", 4 * (r * r + 1) * Pr, lambda i: ops.down_aa_transpose(gs[i], r, (S * r, S * r)))
        del xs, gs

    rate = torch.tensor([0.75, 0.5] * (B // 2), device=dev)
    center = 2 * torch.rand(B, 1, 1, 2, device=dev) - 1

    def torch_T(x):  # the reference's grid build + grid_sample (src/transforms.py:27-83)
        b, _, h, w = x.shape
        u = 2 / w * torch.arange(w, device=x.device, dtype=x.dtype) - 1
        U, V = torch.meshgrid(u, u, indexing="ij")
        grid = torch.stack([V, U], dim=-1).view(1, h, w, 2).repeat(b, 1, 1, 1)
        grid = 1 / rate.view(b, 1, 1, 1).expand_as(grid) * (grid - center) + center
        return F.grid_sample(x, grid, mode="bicubic", padding_mode="reflection", align_corners=True)

    add("scale transform T", 8 * P, lambda i: ops.scale_transform(ys[i], rate, center), lambda i: torch_T(ys[i]))
    kh = ops.kernel_to_host(get_kernel("Gaussian_R2"))
    kd = get_kernel("Gaussian_R2").to(dev, torch.float32)
    add("fused EI re-measure T->A->+noise (Gaussian_R2)", 16 * P,
        lambda i: ops.ei_remeasure(ys[i], rate, center, kh, 1, ns[i], sigma),
        lambda i: fft_blur(torch_T(ys[i]), kd) + ns[i] * sigma)
    add("MSE(x3,x2) reduction", 8 * P, lambda i: ops.mse(ys[i], ns[i]), lambda i: F.mse_loss(ys[i], ns[i]))
    add("SURE reductions (4 inputs)", 16 * P,
        lambda i: ops.sure_loss(ys[i], ns[i], ys[(i + 1) % NR], ns[(i + 1) % NR], 6, 6, 1e-2, sigma ** 2, None))
    add("add noise", 12 * P, lambda i: ops.add_noise(ys[i], ns[i], sigma), lambda i: ys[i] + ns[i] * sigma)
    draw = torch.randn(B, C, S - 12, S - 12, device=dev)
    add("SURE probe y + tau*b", 16 * P, lambda i: ops.sure_perturb(ys[i], draw, 6, 1e-2))

    if rank == 0:
        print(f"# op sweep: batch {B} x {C} x {S} x {S} fp32 per GPU, {world} GPU(s); HBM peak {peak} GB/s ({peak_src}) per GPU")
        print("| op | us | algorithmic MB | GB/s | frac of peak | stock torch us | speed-up |")
        print("|---|---|---|---|---|---|---|")
        for r_ in rows:
            print(f"| {r_['op']} | {r_['us']} | {r_['alg_MB']} | {r_['GBs']} | {r_['frac_of_peak']} | {r_['torch_us']} | {r_['speedup_vs_torch']} |")
        if args.json:
            with open(args.json, "a") as f:
                for r_ in rows:
                    r_["env"] = {k: v for k, v in os.environ.items() if k.startswith("SEI_")}
                    f.write(json.dumps(r_) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
