#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -2
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --dp-mode serial > gpurun_out/r02n_bench_n2_serial.json 2> gpurun_out/r02n_bench_n2_serial.err; tail -1 gpurun_out/r02n_bench_n2_serial.err; cut -c1-230 gpurun_out/r02n_bench_n2_serial.json
timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --dp-mode pipelined > gpurun_out/r02n_bench_n2_pipelined.json 2> gpurun_out/r02n_bench_n2_pipelined.err; tail -3 gpurun_out/r02n_bench_n2_pipelined.err; cut -c1-230 gpurun_out/r02n_bench_n2_pipelined.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_n1.json 2>/dev/null; cut -c1-230 gpurun_out/r02n_bench_n1.json
timeout 600 $TR --master-port 29533 benchmarks/op_sweep.py --no-torch > gpurun_out/r02m_op_sweep_n2.md 2>&1; tail -3 gpurun_out/r02m_op_sweep_n2.md
