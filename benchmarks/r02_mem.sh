#!/bin/bash
# memory option (gelu / gelu' rebuilt in the backward pass): the two configurations that do not fit or barely fit without it;
# then the last validation of the tree (GPU suite, smoke, bench)
set -x
mkdir -p gpurun_out
SEI_RECOMPUTE_MLP=1 timeout 300 python benchmarks/config_sweep.py --network cnn --only "cfg5 proposed (scale)" --steps 3 --warmup 1 > gpurun_out/r02_mem_cfg5_recompute.md 2>&1; tail -2 gpurun_out/r02_mem_cfg5_recompute.md
SEI_RECOMPUTE_MLP=1 timeout 300 python benchmarks/config_sweep.py --network cnn --only "SR x4 proposed 256 b8" --steps 3 --warmup 1 > gpurun_out/r02_mem_sr4_b8_recompute.md 2>&1; tail -2 gpurun_out/r02_mem_sr4_b8_recompute.md
timeout 300 python benchmarks/config_sweep.py --network cnn --only "SR x4 proposed 256 b8" --steps 3 --warmup 1 > gpurun_out/r02_mem_sr4_b8_default.md 2>&1; tail -2 gpurun_out/r02_mem_sr4_b8_default.md
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02x_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02x_smoke.log 2>&1; tail -3 gpurun_out/r02x_smoke.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02x_bench_n1.json 2>/dev/null; cut -c1-200 gpurun_out/r02x_bench_n1.json
