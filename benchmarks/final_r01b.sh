#!/bin/bash
# resampler products after the block-diagonal packing: parity, then the per-shape table
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_kernels.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python benchmarks/resample_bench.py > gpurun_out/resample_bench_packed.md 2>&1; cat gpurun_out/resample_bench_packed.md
