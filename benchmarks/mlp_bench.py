#!/usr/bin/env python
"""The pointwise MLP of a ConvBlock (conv2 -> GELU -> conv3, reference src/models/convolutional.py:40-42) per level of
the default network (hidden 32, 5 scales, 256x256 input, batch 32): the separate kernels next to the fused epilogues.

  forward : gemm + gelu kernel            vs  gemm with gelu / gelu' written from the epilogue
  backward: gemm + gelu-backward(+colsum) vs  gemm with the multiplier epilogue + colsum
"""
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
    batch, size = 32, 256
    print("| level | T | C | gemm us | gelu us | gemm+gelu-dual us | dgrad us | gelu-bwd+colsum us | dgrad*mul us | colsum us | conv3+res us |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for s in range(5):
        C, T = 32 * 4 ** s, batch * (size >> s) ** 2
        torch.manual_seed(s)
        t2 = torch.randn(T, C, device=dev).bfloat16()
        w2 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).bfloat16()
        w3t = (torch.randn(4 * C, C, device=dev) / C ** 0.5).bfloat16()
        w3 = (torch.randn(C, 4 * C, device=dev) / C ** 0.5).bfloat16()
        b2 = torch.randn(4 * C, device=dev)
        g = torch.randn(T, C, device=dev).bfloat16()
        h = ops.gemm_bf16_tn(t2, w2, b2)
        a, d = ops.gemm_bf16_tn_gelu_dual(t2, w2, b2)
        ga = ops.gemm_bf16_tn(g, w3t, None)
        us = lambda f: 1e3 * bench(f)
        r = [us(lambda: ops.gemm_bf16_tn(t2, w2, b2)), us(lambda: ops.gelu_raw(h)),
             us(lambda: ops.gemm_bf16_tn_gelu_dual(t2, w2, b2)), us(lambda: ops.gemm_bf16_tn(g, w3t, None)),
             us(lambda: ops.gelu_bwd_colsum(h, ga)), us(lambda: ops.gemm_bf16_tn_mul(g, w3t, d)),
             us(lambda: ops.colsum_bf16(ga)), us(lambda: ops.gemm_bf16_tn_residual(a, w3, None, g, 1.0))]
        print(f"| s{s} | {T} | {C} | " + " | ".join(f"{v:.1f}" for v in r) + " |", flush=True)
        del t2, h, a, d, ga, g


if __name__ == "__main__":
    main()
