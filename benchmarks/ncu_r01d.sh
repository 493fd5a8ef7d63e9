#!/bin/bash
# Round-1 evidence for the final build: GEMM table, ncu --set full of the CTA-pair GEMM and the tcgen05 resampler product
# (text summaries only), and the ncu launch list of one eager CNN step (shares per kernel).
set -x
python benchmarks/gemm_bench.py > gpurun_out/gemm_bench_2cta.md 2>&1
cap() {  # name, regex, count, skip, command...
  name=$1; regex=$2; count=$3; skip=$4; shift 4
  ncu --set full --clock-control none -k regex:"$regex" --launch-skip $skip --launch-count $count -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python benchmarks/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/ncu_$name.txt 20
  rm -f /tmp/$name.ncu-rep
}
cap gemm2cta "gemm_bf16_tn_2cta" 1 4 python benchmarks/gemm_bench.py "s3 ConvBlock 2048->8192"
cap bgemmtc "bgemm_tc_kernel" 2 0 python benchmarks/profile_step.py --batch 8
cap dwconv2 "dwconv7_kernel" 1 0 python benchmarks/profile_step.py --batch 8
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_cnn_b8_final.csv python bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_cnn_b8_final.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
t, c = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > mv and r[mv].replace(".", "").replace(",", "").isdigit():
        t[r[kn]] += float(r[mv].replace(",", "")); c[r[kn]] += 1
tot = sum(t.values())
with open("gpurun_out/launches_cnn_b8_final.md", "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, cold-cache, serialised), first 6000 launches of `bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline`; total {tot / 1e6:.1f} ms; shares only\n\n| share | launches | mean us | kernel |\n|---|---|---|---|\n")
    sei = 0.0
    for k, v in t.most_common(45):
        f.write(f"| {100 * v / tot:.1f}% | {c[k]} | {v / c[k] / 1e3:.1f} | {k[:120]} |\n")
    sei = sum(v for k, v in t.items() if "sei::" in k or k.startswith("sei"))
    f.write(f"\nlibsei_b200 kernels: {100 * sei / tot:.1f}% of the captured GPU time.\n")
PY
rm -f gpurun_out/launches_cnn_b8_final.csv
ls -la gpurun_out/ncu_*.txt gpurun_out/launches_cnn_b8_final.md gpurun_out/gemm_bench_2cta.md
