#!/bin/bash
# deeper ring for the narrow resampler tiles (A/B), then the last validation pass of the round-1 tree
set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/resample_bench.py 2>&1 | grep -E "down \| 0|sum of" 
SEI_BGEMM_NO_DEEP_RING=1 timeout 300 python benchmarks/resample_bench.py 2>&1 | grep -E "down \| 0|sum of"
SEI_BGEMM_NO_DEEP_RING=1 timeout 600 python -m pytest tests/test_cnn_kernels.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final3.log 2>&1; tail -3 gpurun_out/pytest_gpu_final3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; tail -2 gpurun_out/bench_final3.err; cut -c1-400 gpurun_out/bench_final3.json
timeout 300 python benchmarks/resample_bench.py > gpurun_out/resample_bench_final.md 2>&1
