#!/usr/bin/env python
"""Channel LayerNorm forward / backward per level of the default network (batch 32, 256x256): us and fraction of the HBM copy peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sei_b200 import ops  # noqa: E402
from gemm_bench import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6556.2
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["hbm_gbs"])
    print("| level | T x C | forward us | frac | backward (dx + dgamma, dbeta) us | frac |")
    print("|---|---|---|---|---|---|")
    for s in range(5):
        C, T = 32 * 4 ** s, 32 * (256 >> s) ** 2
        x = torch.randn(T, C, device=dev).bfloat16()
        gy = torch.randn(T, C, device=dev).bfloat16()
        g32, b32 = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        y, mean, rstd, small = ops.ln_forward_raw(x, g32, b32, 1e-6)
        f = 1e3 * bench(lambda: ops.ln_forward_raw(x, g32, b32, 1e-6))
        b = 1e3 * bench(lambda: ops.ln_backward_raw(gy, x, mean, rstd, g32, small))
        mb = x.numel() * 2 / 1e6
        print(f"| s{s} | {T} x {C} | {f:.1f} | {2 * mb / f * 1e3 / peak:.2f} | {b:.1f} | {3 * mb / b * 1e3 / peak:.2f} |", flush=True)


if __name__ == "__main__":
    main()
