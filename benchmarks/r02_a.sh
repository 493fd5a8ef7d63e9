#!/bin/bash
# round-2 baseline pass of the restored tree: GPU tests, smoke, bench, operator sweep, per-kernel step profile, GEMM table
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -3 gpurun_out/r02a_bench.err; cut -c1-1500 gpurun_out/r02a_bench.json
timeout 400 python benchmarks/op_sweep.py --no-torch > gpurun_out/r02a_op_sweep.md 2>&1; tail -40 gpurun_out/r02a_op_sweep.md
timeout 400 python benchmarks/profile_step.py > gpurun_out/r02a_profile_step.md 2>&1; head -50 gpurun_out/r02a_profile_step.md
timeout 300 python benchmarks/gemm_bench.py > gpurun_out/r02a_gemm_bench.md 2>&1; tail -30 gpurun_out/r02a_gemm_bench.md
