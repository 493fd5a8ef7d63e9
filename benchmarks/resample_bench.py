#!/usr/bin/env python
"""The ideal resamplers of the CNN (models/resample.py -> sei_bgemm_bf16) per level and per pass at the training shapes
(batch 32, 256x256, hidden 32): microseconds and algorithmic GB/s (operand read + result written) next to the copy peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sei_b200 import ops, last_kernel  # noqa: E402
from models import resample  # noqa: E402


def bench(fn, reps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    peak = 6556.2
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["hbm_gbs"])
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print(f"| op | level | x (B,H,W,C) | pass | M | K | N | items | us | GB/s | frac of {peak:.0f} | kernel |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    total = 0.0
    for kind in ("down", "up"):
        for lvl in range(4):
            # Downsample resamples BEFORE its pointwise convolution (C_l channels); Upsample is IdealUpsample -> LayerNorm ->
            # conv (models/convolutional.py, reference :136-150), so it resamples at the WIDE channel count C_{l+1}
            C = 32 * 4 ** lvl if kind == "down" else 32 * 4 ** (lvl + 1)
            S = 256 >> lvl if kind == "down" else 256 >> (lvl + 1)
            B, H, W = batch, S, S
            pk = resample._packed(kind, H, W, 2, dev)
            Ho, Wo = pk["Ho"], pk["Wo"]
            x = torch.randn(B, H, W, C, device=dev).bfloat16()
            out = torch.empty((B, Ho, Wo, C), dtype=x.dtype, device=dev)
            gx = torch.empty_like(x)
            y = torch.empty((B, H, 2, Wo, C), dtype=x.dtype, device=dev)
            calls = [("fwd width", pk["A1"], C, B * H, x, y), ("fwd height", pk["A2"], Wo * C, B, y, out),
                     ("bwd height", pk["A2T"], Wo * C, B, out, y), ("bwd width", pk["A1T"], C, B * H, y, gx)]
            calls = [(n_, a_, N_, it_, s_, d_, (lambda a_=a_, s_=s_, d_=d_, N_=N_, it_=it_: a_(s_, d_, N_, it_))) for n_, a_, N_, it_, s_, d_ in calls]
            for name, a, N, items, src, dst, fn in calls:
                us = bench(fn)
                total += us
                gb = (src.numel() + dst.numel()) * 2 / 1e9
                Pk = a._pack_factor(items)
                print(f"| {kind} | {lvl} | {tuple(x.shape)} | {name} | {a.A.shape[0]} x{Pk} | {a.A.shape[1]} x{Pk} | {N} | {items // Pk} | {us:.1f} | {gb / us * 1e6:.0f} | "
                      f"{gb / us * 1e6 / peak:.2f} | {last_kernel()} |")
            del x, y, out, gx
    print(f"\nsum of the 32 calls: {total / 1e3:.2f} ms (a proposed step runs each forward pass 3 times and each backward pass about 3 times)")


if __name__ == "__main__":
    main()
