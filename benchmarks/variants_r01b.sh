#!/bin/bash
# A/B sweep of the blur CTA shapes and scale-transform band heights (tuning run, results -> gpurun_out/variants.md)
out=gpurun_out/variants.md
: > $out
run() { echo "## $1" >> $out; env $1 python benchmarks/op_sweep.py --no-torch --only "$2" --reps 30 2>&1 | grep -v "^#" | grep "|" | grep -v "^| op\|^|---" >> $out; }
run "SEI_BLUR_THREADS=128 SEI_BLUR_STAGES=1" blur
run "SEI_BLUR_THREADS=128 SEI_BLUR_STAGES=2" blur
run "SEI_BLUR_THREADS=256" blur
run "SEI_BLUR_H16=0" blur
run "SEI_BLUR_THREADS=128 SEI_BLUR_STAGES=1 SEI_BLUR_CTAS=3" "Gaussian_R2 A"
run "SEI_BLUR_THREADS=128 SEI_BLUR_STAGES=1 SEI_BLUR_TH=8" "Gaussian_R2 A"
run "SEI_BLUR_THREADS=128 SEI_BLUR_STAGES=1 SEI_BLUR_TH=24" "Gaussian_R2 A"
run "SEI_SCALE_TH=8" scale
run "SEI_SCALE_TH=16" scale
run "SEI_SCALE_TH=24" scale
run "SEI_SCALE_TH=32" scale
cat $out
