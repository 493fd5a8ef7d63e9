#!/bin/bash
# ncu --set full of the operator kernels that sit below 0.6 of the copy peak (scale transform, SR A x2 / x4)
set -x
mkdir -p gpurun_out
bash benchmarks/ncu_ops.sh r02_scale_rows "scale transform" "scale_rows_kernel"
bash benchmarks/ncu_ops.sh r02_sr2_A "SR x2 A" "down_(stream|rows)_kernel"
bash benchmarks/ncu_ops.sh r02_sr4_A "SR x4 A" "down_(stream|rows)_kernel"
bash benchmarks/ncu_ops.sh r02_blur_noise "blur Gaussian_R2 A+noise" "blur_band_kernel"
ls -la gpurun_out | tail
