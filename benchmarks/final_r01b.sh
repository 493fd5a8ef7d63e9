#!/bin/bash
# resampler products at the real training shapes + a last test / smoke / bench pass
set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/resample_bench.py > gpurun_out/resample_bench.md 2>&1; cat gpurun_out/resample_bench.md
