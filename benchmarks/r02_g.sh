#!/bin/bash
# two GPUs: NCCL data-parallel test, bench with the serial and the overlapped gradient all-reduce
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r02g_pytest_dist.log 2>&1; tail -5 gpurun_out/r02g_pytest_dist.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; cut -c1-200 gpurun_out/r02g_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --dp-mode serial > gpurun_out/r02g_bench_n2_serial.json 2> gpurun_out/r02g_bench_n2_serial.err; tail -2 gpurun_out/r02g_bench_n2_serial.err; cut -c1-200 gpurun_out/r02g_bench_n2_serial.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --dp-mode overlap > gpurun_out/r02g_bench_n2_overlap.json 2> gpurun_out/r02g_bench_n2_overlap.err; tail -2 gpurun_out/r02g_bench_n2_overlap.err; cut -c1-200 gpurun_out/r02g_bench_n2_overlap.json
