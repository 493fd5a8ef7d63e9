#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_kernels.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log
timeout 300 python benchmarks/dw_bench.py > gpurun_out/r02e_dw_bench_tile.md 2>&1; cat gpurun_out/r02e_dw_bench_tile.md
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err; cut -c1-300 gpurun_out/r02e_bench.json
SEI_DWCONV_TILE=0 timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02e_bench_regwin.json 2> gpurun_out/r02e_bench_regwin.err; cut -c1-300 gpurun_out/r02e_bench_regwin.json
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02e_bench2.json 2> gpurun_out/r02e_bench2.err; cut -c1-300 gpurun_out/r02e_bench2.json
timeout 400 python benchmarks/profile_step.py --batch 32 > gpurun_out/r02e_profile_step_b32.md 2>&1; head -44 gpurun_out/r02e_profile_step_b32.md
