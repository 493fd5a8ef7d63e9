#!/bin/bash
# ncu launch list (shares only) of the final tree: first 4000 launches of an eager CNN step at batch 8
mkdir -p gpurun_out
T=r02_last
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --network cnn --batch 8 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
T = "r02_last"
rows = list(csv.reader(open(f"gpurun_out/{T}_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
t, c = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > mv and r[mv].replace(".", "").replace(",", "").isdigit():
        t[r[kn]] += float(r[mv].replace(",", "")); c[r[kn]] += 1
tot = sum(t.values())
with open(f"gpurun_out/{T}_launches_cnn_step_b8.md", "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, cold-cache, serialised), first 4000 launches of `bench.py --network cnn --batch 8 --steps 1 --warmup 1 --no-graph --no-cpu-baseline` (final round-2 tree); total {tot / 1e6:.1f} ms; shares only\n\n| share | launches | mean us | kernel |\n|---|---|---|---|\n")
    for k, v in t.most_common(50):
        f.write(f"| {100 * v / tot:.1f}% | {c[k]} | {v / c[k] / 1e3:.1f} | {k[:120]} |\n")
    sei = sum(v for k, v in t.items() if "sei::" in k or k.startswith("sei"))
    f.write(f"\nlibsei_b200 kernels: {100 * sei / tot:.1f}% of the captured GPU time.\n")
PY
rm -f gpurun_out/${T}_launches.csv
head -16 gpurun_out/${T}_launches_cnn_step_b8.md | cut -c1-150; tail -2 gpurun_out/${T}_launches_cnn_step_b8.md
