#!/bin/bash
# usage: ncu_ops.sh TAG "ONLY-PATTERN" KERNEL-REGEX   (under gpurun; op_sweep runs the op 3 + 20 times)
tag=$1; only=$2; regex=$3
python benchmarks/op_sweep.py --no-torch --only "$only" --reps 5 > gpurun_out/ncu_${tag}_plain.log 2>&1 &&
bash benchmarks/ncu_one.sh $tag "$regex" 4 2 -- python benchmarks/op_sweep.py --no-torch --only "$only" --reps 5
