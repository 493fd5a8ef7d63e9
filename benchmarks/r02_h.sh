#!/bin/bash
# full GPU suite of the tree, the edge-convolution table, copies attribution, ncu of the new kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02h_pytest_gpu.log
timeout 300 python benchmarks/conv_bench.py > gpurun_out/r02h_conv_bench.md 2>&1; cat gpurun_out/r02h_conv_bench.md
timeout 400 python benchmarks/profile_step.py --batch 32 --copies > gpurun_out/r02h_profile_copies.md 2>&1; grep -n -i "memcpy\|copy\|cat\|add" gpurun_out/r02h_profile_copies.md | head -40
bash benchmarks/ncu_one.sh r02_dwconv_tile "dwconv7_tile_kernel" 4 1 -- python benchmarks/dw_bench.py
bash benchmarks/ncu_one.sh r02_dwconv_wgrad_tile "dwconv7_wgrad_tile_kernel" 4 1 -- python benchmarks/dw_bench.py
bash benchmarks/ncu_one.sh r02_conv_igemm "conv3x3_igemm_kernel" 2 2 -- python benchmarks/conv_bench.py
ls gpurun_out | tail -5
