#!/bin/bash
# usage: ncu_one.sh NAME REGEX SKIP COUNT -- command...   -> gpurun_out/ncu_NAME.txt (text summary only)
name=$1; regex=$2; skip=$3; count=$4; shift 5
ncu --set full --clock-control none -k regex:"$regex" --launch-skip $skip --launch-count $count -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
python benchmarks/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/ncu_$name.txt 30
rm -f /tmp/$name.ncu-rep
