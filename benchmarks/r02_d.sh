#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_cnn_kernels.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; tail -5 gpurun_out/r02d_pytest.log
timeout 300 python benchmarks/mlp_bench.py > gpurun_out/r02d_mlp_bench.md 2>&1; cat gpurun_out/r02d_mlp_bench.md
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -3 gpurun_out/r02d_bench.err; cut -c1-300 gpurun_out/r02d_bench.json
SEI_GELU_EPILOGUE_MIN_C=8192 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench_s4only.json 2> gpurun_out/r02d_bench_s4only.err; cut -c1-300 gpurun_out/r02d_bench_s4only.json
SEI_GELU_EPILOGUE_MIN_C=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02d_bench_nofuse.json 2> gpurun_out/r02d_bench_nofuse.err; cut -c1-300 gpurun_out/r02d_bench_nofuse.json
