#!/bin/bash
# correctness of the CTA-pair weight-gradient kernel, then its throughput next to the single-CTA kernel, then the step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python benchmarks/gemm_bench.py --wgrad > gpurun_out/wgrad_2cta.md 2>&1
cat gpurun_out/wgrad_2cta.md
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_2cta_wgrad.json 2> gpurun_out/bench_2cta_wgrad.err
tail -3 gpurun_out/bench_2cta_wgrad.err; cat gpurun_out/bench_2cta_wgrad.json
