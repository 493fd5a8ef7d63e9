#!/bin/bash
# Final validation of the round-1 tree on one B200: full GPU suite, smoke, bench line, ncu evidence (text summaries only).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; tail -3 gpurun_out/pytest_gpu_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -2 gpurun_out/bench_final2.err; cat gpurun_out/bench_final2.json
ncu --set full --clock-control none -k regex:"gemm_bf16_atb_2cta" --launch-skip 30 --launch-count 1 -o /tmp/wgrad2cta -f python benchmarks/gemm_bench.py --wgrad > gpurun_out/ncu_wgrad2cta.log 2>&1
python benchmarks/ncu_summary.py /tmp/wgrad2cta.ncu-rep gpurun_out/ncu_wgrad2cta.txt 20; rm -f /tmp/wgrad2cta.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_cnn_b8_final2.csv python bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_cnn_b8_final2.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
t, c = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > mv and r[mv].replace(".", "").replace(",", "").isdigit():
        t[r[kn]] += float(r[mv].replace(",", "")); c[r[kn]] += 1
tot = sum(t.values())
with open("gpurun_out/launches_cnn_b8_final2.md", "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, cold-cache, serialised), first 6000 launches of `bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline`; total {tot / 1e6:.1f} ms; shares only\n\n| share | launches | mean us | kernel |\n|---|---|---|---|\n")
    for k, v in t.most_common(45):
        f.write(f"| {100 * v / tot:.1f}% | {c[k]} | {v / c[k] / 1e3:.1f} | {k[:120]} |\n")
    sei = sum(v for k, v in t.items() if "sei::" in k or k.startswith("sei"))
    f.write(f"\nlibsei_b200 kernels: {100 * sei / tot:.1f}% of the captured GPU time.\n")
PY
rm -f gpurun_out/launches_cnn_b8_final2.csv
head -12 gpurun_out/launches_cnn_b8_final2.md | cut -c1-150
