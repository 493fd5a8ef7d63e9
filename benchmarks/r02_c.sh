#!/bin/bash
# GEMM epilogue with 8 warps / staged bias / residual, ConvBlock as one node: tests, GEMM table (8 vs 4 epilogue warps), bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_cnn_kernels.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; tail -15 gpurun_out/r02c_pytest.log
timeout 300 python benchmarks/gemm_bench.py > gpurun_out/r02c_gemm_bench_ew8.md 2>&1; tail -13 gpurun_out/r02c_gemm_bench_ew8.md
SEI_GEMM_EW=4 timeout 300 python benchmarks/gemm_bench.py > gpurun_out/r02c_gemm_bench_ew4.md 2>&1; tail -13 gpurun_out/r02c_gemm_bench_ew4.md
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -3 gpurun_out/r02c_bench.err; cut -c1-300 gpurun_out/r02c_bench.json
SEI_CONVBLOCK_NODE=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench_nonode.json 2> gpurun_out/r02c_bench_nonode.err; cut -c1-300 gpurun_out/r02c_bench_nonode.json
timeout 400 python benchmarks/profile_step.py --batch 32 > gpurun_out/r02c_profile_step_b32.md 2>&1; head -40 gpurun_out/r02c_profile_step_b32.md
