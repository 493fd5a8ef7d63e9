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
            y = torch.empty((B, 2, H, Wo, C), dtype=x.dtype, device=dev)
            out = torch.empty((B, Ho, Wo, C), dtype=x.dtype, device=dev)
            gx = torch.empty_like(x)
            a1, a2, a2t, a1t = pk["A1"], pk["A2"], pk["A2T"], pk["A1T"]
            calls = [
                ("fwd width", a1, C, B * H, x, y, lambda: ops.bgemm_bf16(a1.data, x, y, a1.M, a1.K, C, a1.tile, B * H, H, (H * W * C, W * C), W, (0, C),
                                                                      (2 * H * Wo * C, Wo * C), Wo, (H * Wo * C, C))),
                ("fwd height", a2, Wo * C, B, y, out, lambda: ops.bgemm_bf16(a2.data, y, out, a2.M, a2.K, Wo * C, a2.tile, B, 1, (2 * H * Wo * C, 0), 2 * H,
                                                                          (0, Wo * C), (Ho * Wo * C, 0), Ho, (0, Wo * C))),
                ("bwd height", a2t, Wo * C, B, out, y, lambda: ops.bgemm_bf16(a2t.data, out, y, a2t.M, a2t.K, Wo * C, a2t.tile, B, 1, (Ho * Wo * C, 0), Ho,
                                                                           (0, Wo * C), (2 * H * Wo * C, 0), 2 * H, (0, Wo * C))),
                ("bwd width", a1t, C, B * H, y, gx, lambda: ops.bgemm_bf16(a1t.data, y, gx, a1t.M, a1t.K, C, a1t.tile, B * H, H, (2 * H * Wo * C, Wo * C), Wo,
                                                                        (H * Wo * C, C), (H * W * C, W * C), W, (0, C))),
            ]
            for name, a, N, items, src, dst, fn in calls:
                us = bench(fn)
                total += us
                gb = (src.numel() + dst.numel()) * 2 / 1e9
                print(f"| {kind} | {lvl} | {tuple(x.shape)} | {name} | {a.M} | {a.K} | {N} | {items} | {us:.1f} | {gb / us * 1e6:.0f} | "
                      f"{gb / us * 1e6 / peak:.2f} | {last_kernel()} |")
            del x, y, out, gx
    print(f"\nsum of the 32 calls: {total / 1e3:.2f} ms (a proposed step runs each forward pass 3 times and each backward pass about 3 times)")


if __name__ == "__main__":
    main()
