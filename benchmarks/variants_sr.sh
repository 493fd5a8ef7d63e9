#!/bin/bash
out=gpurun_out/variants_sr.md
: > $out
run() { echo "## $1" >> $out; env $1 python benchmarks/op_sweep.py --no-torch --only "$2" --reps 30 2>&1 | grep "A |" >> $out; }
for ch in 16 32 48; do for th in 8 16; do
run "SEI_DOWN_CH_KB=$ch SEI_DOWN_TH=$th SEI_DOWN_SMEM_KB=110" "SR x4"
done; done
for ch in 16 32 48; do run "SEI_DOWN_CH_KB=$ch SEI_DOWN_TH=8 SEI_DOWN_SMEM_KB=75" "SR x4"; done
for ch in 16 32; do for th in 8 16; do for kb in 110 56; do
run "SEI_DOWN_CH_KB=$ch SEI_DOWN_TH=$th SEI_DOWN_SMEM_KB=$kb" "SR x2"
done; done; done
cat $out
