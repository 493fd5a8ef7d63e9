#!/usr/bin/env python
"""Depthwise 7x7 (ConvBlock.conv1) forward and weight gradient per level of the default network (batch 32, 256x256)."""
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
    print("| level | B x H x W x C | forward us | forward + residual us | weight gradient us | FMA floor us (1.9 GHz) |")
    print("|---|---|---|---|---|---|")
    for s in range(5):
        C, S = 32 * 4 ** s, 256 >> s
        x = torch.randn(32, S, S, C, device=dev).bfloat16()
        g = torch.randn(32, S, S, C, device=dev).bfloat16()
        wt = (torch.randn(49, C, device=dev) / 7).contiguous()
        bias = torch.randn(C, device=dev)
        f = 1e3 * bench(lambda: ops._dwconv7_raw(x, wt, bias))
        fr = 1e3 * bench(lambda: ops._dwconv7_raw(x, wt, None, res=g, res_scale=1.0))
        wg = 1e3 * bench(lambda: ops.dwconv7_wgrad_raw(g, x))
        floor = x.numel() * 49 / (148 * 128 * 1.9e9) * 1e6
        print(f"| s{s} | 32 x {S} x {S} x {C} | {f:.1f} | {fr:.1f} | {wg:.1f} | {floor:.1f} |", flush=True)


if __name__ == "__main__":
    main()
