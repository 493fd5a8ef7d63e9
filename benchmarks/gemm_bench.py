#!/usr/bin/env python
"""tcgen05 GEMM (sei_gemm_bf16_tn) throughput on the pointwise-convolution shapes of the reference's
ConvolutionalModel (hidden 32, 5 scales; 256x256 input, batch 32), next to cuBLAS (torch.matmul)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sei_b200 import ops  # noqa: E402


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
    return e0.elapsed_time(e1) / reps


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda:0")
    peak = 1685.6
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = float(json.load(open(path))["bf16_tflops"])
    batch, size = 32, 256
    shapes = []
    for s in range(5):
        dim, hw = 32 * 4 ** s, (size >> s) ** 2
        shapes.append((f"s{s} ConvBlock {dim}->{4 * dim}", batch * hw, 4 * dim, dim))
        shapes.append((f"s{s} ConvBlock {4 * dim}->{dim}", batch * hw, dim, 4 * dim))
    shapes.append(("square 8192^3", 8192, 8192, 8192))
    if only == "--wgrad":
        # weight-gradient products D[Cout, Cin] += gy[pixels, Cout]^T x[pixels, Cin] (sei_gemm_bf16_atb_accumulate), in place
        print(f"| layer | pixels | Cout | Cin | sei us | sei TFLOP/s | frac of {peak} | cuBLAS (gy.t() @ x) us | kernel |")
        print("|---|---|---|---|---|---|---|---|---|")
        from sei_b200 import last_kernel
        for name, M, N, K in shapes[:-1]:
            if M * K * 2 > 8e9 or M * N * 2 > 8e9:
                M = M // 4
                name += " (M/4)"
            gy = torch.randn(M, N, device=dev).bfloat16()
            x = torch.randn(M, K, device=dev).bfloat16()
            out = torch.zeros(N, K, device=dev)
            flops = 2.0 * M * N * K
            ms = bench(lambda: ops.gemm_bf16_atb(gy, x, out=out))
            kern = last_kernel()
            ms_ref = bench(lambda: gy.t() @ x)
            print(f"| {name} | {M} | {N} | {K} | {ms * 1e3:.1f} | {flops / ms / 1e9:.1f} | {flops / ms / 1e9 / peak:.3f} | "
                  f"{ms_ref * 1e3:.1f} | {kern} |")
            del gy, x, out
        return
    print(f"| layer | M | N | K | sei us | sei TFLOP/s | frac of {peak} | cuBLAS us | cuBLAS TFLOP/s |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, M, N, K in shapes:
        if only and only not in name:
            continue
        if M * K * 2 > 8e9 or M * N * 2 > 8e9:
            M = M // 4
            name += " (M/4)"
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        flops = 2.0 * M * N * K
        ms = bench(lambda: ops.gemm_bf16_tn(a, b))
        ms_ref = bench(lambda: a @ b.t())
        print(f"| {name} | {M} | {N} | {K} | {ms * 1e3:.1f} | {flops / ms / 1e9:.1f} | {flops / ms / 1e9 / peak:.3f} | "
              f"{ms_ref * 1e3:.1f} | {flops / ms_ref / 1e9:.1f} |")
        del a, b


if __name__ == "__main__":
    main()
