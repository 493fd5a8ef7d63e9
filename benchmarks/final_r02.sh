#!/bin/bash
# Validation + evidence pass of the round-2 tree on one B200: full GPU suite, smoke, both bench arms, the operator / GEMM /
# MLP / depthwise / edge-convolution / resampler tables, the per-kernel step profile, the other BASELINE configurations,
# the ncu launch list of one step and ncu --set full captures (text summaries + DRAM traffic per launch).
set -x
mkdir -p gpurun_out
T=r02f
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err; cat gpurun_out/${T}_bench_n1.json
timeout 600 python benchmarks/op_sweep.py > gpurun_out/${T}_op_sweep.md 2>&1; tail -24 gpurun_out/${T}_op_sweep.md
timeout 300 python benchmarks/gemm_bench.py > gpurun_out/${T}_gemm_bench.md 2>&1
timeout 300 python benchmarks/mlp_bench.py > gpurun_out/${T}_mlp_bench.md 2>&1
timeout 300 python benchmarks/dw_bench.py > gpurun_out/${T}_dw_bench.md 2>&1
timeout 300 python benchmarks/conv_bench.py > gpurun_out/${T}_conv_bench.md 2>&1
timeout 300 python benchmarks/resample_bench.py > gpurun_out/${T}_resample_bench.md 2>&1
timeout 400 python benchmarks/profile_step.py --batch 32 --rows 70 > gpurun_out/${T}_profile_step_b32.md 2>&1
timeout 900 python benchmarks/config_sweep.py --network cnn > gpurun_out/${T}_config_sweep_n1.md 2>&1; tail -12 gpurun_out/${T}_config_sweep_n1.md
# ncu: launch list of one eager step (shares), then --set full of the dominant GEMM and of the blur (summaries + traffic)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
T = "r02f"
rows = list(csv.reader(open(f"gpurun_out/{T}_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
t, c = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) > mv and r[mv].replace(".", "").replace(",", "").isdigit():
        t[r[kn]] += float(r[mv].replace(",", "")); c[r[kn]] += 1
tot = sum(t.values())
with open(f"gpurun_out/{T}_launches_cnn_step_b8.md", "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, cold-cache, serialised), first 8000 launches of `bench.py --network cnn --batch 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline`; total {tot / 1e6:.1f} ms; shares only\n\n| share | launches | mean us | kernel |\n|---|---|---|---|\n")
    for k, v in t.most_common(50):
        f.write(f"| {100 * v / tot:.1f}% | {c[k]} | {v / c[k] / 1e3:.1f} | {k[:120]} |\n")
    sei = sum(v for k, v in t.items() if "sei::" in k or k.startswith("sei"))
    f.write(f"\nlibsei_b200 kernels: {100 * sei / tot:.1f}% of the captured GPU time.\n")
PY
rm -f gpurun_out/${T}_launches.csv
head -14 gpurun_out/${T}_launches_cnn_step_b8.md | cut -c1-150
bash benchmarks/ncu_one.sh r02_gemm_2cta_s4 "gemm_bf16_tn_2cta_kernel" 4 2 -- python benchmarks/gemm_bench.py "s4 ConvBlock 8192->32768"
bash benchmarks/ncu_one.sh r02_blur_band "blur_band_kernel" 4 2 -- python benchmarks/op_sweep.py --no-torch --only "blur Gaussian_R2 A" --reps 5
bash benchmarks/ncu_one.sh r02_gemm_128_s0 "gemm_bf16_tn_kernel" 4 1 -- python benchmarks/gemm_bench.py "s0 ConvBlock 32->128"
ls gpurun_out | grep -c .
