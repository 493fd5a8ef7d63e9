#!/bin/bash
out=gpurun_out/variants_scale.md
: > $out
run() { echo "## $1" >> $out; env $1 python benchmarks/op_sweep.py --no-torch --only "$2" --reps 30 2>&1 | grep "|" | grep -v "^| op\|^|---" >> $out; }
run "SEI_SCALE_TH=16" "transform T"
run "SEI_SCALE_TH=8" "transform T"
run "SEI_SCALE_TH=24" "transform T"
run "SEI_SCALE_TH=16" "fused EI"
cat $out
