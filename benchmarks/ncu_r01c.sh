#!/bin/bash
# ncu --set full of the kernels changed in this round; only text summaries travel back (reports are deleted)
set -x
cap() {  # name, regex, count, skip, command...
  name=$1; regex=$2; count=$3; skip=$4; shift 4
  ncu --set full --clock-control none -k regex:"$regex" --launch-skip $skip --launch-count $count -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python benchmarks/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/ncu_$name.txt 25
  rm -f /tmp/$name.ncu-rep
}
cap dwconv "dwconv7_kernel" 2 0 python benchmarks/profile_step.py --batch 8
cap dwwgrad "dwconv7_wgrad" 1 0 python benchmarks/profile_step.py --batch 8
cap scale "scale_band" 1 1 python benchmarks/op_sweep.py --no-torch --only "transform T" --reps 2
cap blur "blur_band" 1 3 python benchmarks/op_sweep.py --no-torch --only "Gaussian_R2 A" --reps 2
ls -la gpurun_out/ncu_*.txt
