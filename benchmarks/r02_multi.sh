#!/bin/bash
# usage: r02_multi.sh N   -- data-parallel runs on N GPUs of one box: bench, SR x2 / 512^2 configurations, operator sweep
N=$1
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_n$N.json 2> gpurun_out/r02m_bench_n$N.err; tail -1 gpurun_out/r02m_bench_n$N.err; cut -c1-230 gpurun_out/r02m_bench_n$N.json
timeout 600 $TR --master-port 29522 benchmarks/config_sweep.py --network cnn --only "cfg3" > gpurun_out/r02m_config_cfg3_n$N.md 2>&1; tail -3 gpurun_out/r02m_config_cfg3_n$N.md
timeout 600 $TR --master-port 29523 benchmarks/config_sweep.py --network cnn --only "cfg5 proposed (scale)" > gpurun_out/r02m_config_cfg5_n$N.md 2>&1; tail -3 gpurun_out/r02m_config_cfg5_n$N.md
timeout 600 $TR --master-port 29524 benchmarks/op_sweep.py --no-torch > gpurun_out/r02m_op_sweep_n$N.md 2>&1; tail -22 gpurun_out/r02m_op_sweep_n$N.md
