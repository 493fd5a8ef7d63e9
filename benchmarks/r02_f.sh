#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_kernels.py -m gpu -x -q -k "implicit" > gpurun_out/r02f_pytest_igemm.log 2>&1; tail -25 gpurun_out/r02f_pytest_igemm.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_cnn_kernels.py -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; tail -8 gpurun_out/r02f_pytest.log
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err; cut -c1-300 gpurun_out/r02f_bench.json
SEI_IGEMM_CONV=0 timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02f_bench_noigemm.json 2> gpurun_out/r02f_bench_noigemm.err; cut -c1-300 gpurun_out/r02f_bench_noigemm.json
