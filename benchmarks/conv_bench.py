#!/usr/bin/env python
"""The 3x3 edge convolutions of the restoration CNN (UNet.in_conv 3 -> 32, UNet.out_conv 32 -> 3; batch 32, 256x256):
implicit GEMM on tcgen05 (sei_conv3x3_igemm_bf16) next to round 1's formulation (pad + 9-slice cat + GEMM / direct
kernel) and to cuDNN (torch conv2d, bf16 channels-last)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scale-equivariant-imaging_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from sei_b200 import ops  # noqa: E402
from gemm_bench import bench  # noqa: E402
import models.convolutional as mc  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, S = 32, 256
    print("| layer | implicit GEMM us | algorithmic MB | GB/s | round-1 path us | cuDNN us |")
    print("|---|---|---|---|---|---|")
    for name, cin, cout in (("in_conv 3->32", 3, 32), ("out_conv 32->3", 32, 3)):
        conv = mc._conv(cin, cout, 3, padding="same").to(dev)
        x = torch.randn(B, cin, S, S, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        xl = x.permute(0, 2, 3, 1)
        with torch.no_grad():
            if cin == 3:
                x8 = F.pad(xl, (0, 5)).contiguous()
                wg = ops.igemm_weight_chunks(conv.weight, 8, 32)
                t_new = 1e3 * bench(lambda: ops.conv3x3_igemm(x8, wg, conv.bias, 32, 32))
                mb = (x8.numel() + B * S * S * 32) * 2 / 1e6
            else:
                xc = xl.contiguous()
                wg = ops.igemm_weight_chunks(conv.weight, 32, 16)
                t_new = 1e3 * bench(lambda: ops.conv3x3_igemm(xc, wg, conv.bias, 4, 3))
                mb = (xc.numel() + B * S * S * 4) * 2 / 1e6
            mc._IGEMM_CONV = False
            t_old = 1e3 * bench(lambda: conv(x))
            mc._IGEMM_CONV = True
            wb = conv.weight.detach().bfloat16().contiguous(memory_format=torch.channels_last)
            bb = conv.bias.detach().bfloat16()
            t_dnn = 1e3 * bench(lambda: F.conv2d(x, wb, bb, padding=1))
        print(f"| {name} | {t_new:.1f} | {mb:.1f} | {mb / t_new * 1e3:.0f} | {t_old:.1f} | {t_dnn:.1f} |", flush=True)


if __name__ == "__main__":
    main()
