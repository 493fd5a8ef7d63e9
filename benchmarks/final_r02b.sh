#!/bin/bash
# last validation of the round-2 tree on one B200: full GPU suite, smoke, bench (default arguments), step profile, GEMM table
set -x
mkdir -p gpurun_out
T=r02y
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_bench_n1.json
timeout 400 python benchmarks/profile_step.py --batch 32 --rows 70 > gpurun_out/${T}_profile_step_b32.md 2>&1; head -30 gpurun_out/${T}_profile_step_b32.md | cut -c1-140
timeout 300 python benchmarks/gemm_bench.py > gpurun_out/${T}_gemm_bench.md 2>&1; tail -12 gpurun_out/${T}_gemm_bench.md
timeout 300 python benchmarks/ln_bench.py > gpurun_out/${T}_ln_bench.md 2>&1
timeout 300 python benchmarks/mlp_bench.py > gpurun_out/${T}_mlp_bench.md 2>&1
