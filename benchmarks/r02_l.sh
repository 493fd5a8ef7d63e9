#!/bin/bash
# grouped tile order of the CTA-pair GEMMs: parity, GEMM tables (forward + weight gradient), DRAM traffic of the deepest
# layer under ncu, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; tail -3 gpurun_out/r02l_pytest.log
timeout 300 python benchmarks/gemm_bench.py > gpurun_out/r02l_gemm_bench.md 2>&1; tail -12 gpurun_out/r02l_gemm_bench.md
timeout 300 python benchmarks/gemm_bench.py --wgrad > gpurun_out/r02l_gemm_bench_wgrad.md 2>&1; tail -8 gpurun_out/r02l_gemm_bench_wgrad.md
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -2 gpurun_out/r02l_bench.err; cut -c1-220 gpurun_out/r02l_bench.json
bash benchmarks/ncu_one.sh r02_gemm_2cta_s4 "gemm_bf16_tn_2cta_kernel" 4 2 -- python benchmarks/gemm_bench.py "s4 ConvBlock 8192->32768"
grep -A4 "DRAM traffic" gpurun_out/ncu_r02_gemm_2cta_s4.txt
bash benchmarks/ncu_one.sh r02_gemm_atb_2cta_s4 "gemm_bf16_atb_2cta_kernel" 2 1 -- python benchmarks/gemm_bench.py --wgrad
grep -A4 "DRAM traffic" gpurun_out/ncu_r02_gemm_atb_2cta_s4.txt
