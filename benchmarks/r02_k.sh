#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_kernels.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -2 gpurun_out/r02k_bench.err; cut -c1-220 gpurun_out/r02k_bench.json
timeout 400 python benchmarks/profile_step.py --batch 32 > gpurun_out/r02k_profile_step_b32.md 2>&1; grep -E "total CUDA|ln_|colsum" gpurun_out/r02k_profile_step_b32.md | cut -c1-150
