#!/usr/bin/env python
"""bench.py -- SEI training imgs/sec @256x256 on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one batch: the reference's `proposed` training step
(demo/train.py:258-268: zero_grad, loss(x, y, model) with SURE + equivariant terms, backward,
optimizer step) on the workload BASELINE.json's configs[1] names: deblurring Gaussian_R2,
synthetic 256x256 RGB crops, batch 32 per GPU.  Physics, scale transform and loss reductions run
in the libsei_b200 kernels behind the reference's own Python API (physics.get_physics,
losses.get_loss, models.get_model).  --network cnn (default) drives the reference's restoration CNN
(--ProposedModel__architecture Convolutional at its default flags: hidden 32, 5 scales, 645 M
parameters) with every dense contraction on the tcgen05 GEMM; --network standin swaps in a 4-parameter
pointwise network so that the line isolates the operator + loss-assembly path.

Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the same step
(oracle/, the reference is pure Python and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "scale-equivariant-imaging_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH, SIZE, CH = 32, 256, 3
KERNEL, NOISE_LEVEL, MARGIN = "Gaussian_R2", 5, 6
WORKLOAD = f"deblurring {KERNEL} method=proposed, synthetic {SIZE}x{SIZE} RGB crops, batch {BATCH} per GPU (BASELINE configs[1])"


def loss_args():
    from argparse import Namespace
    return Namespace(task="deblurring", noise_level=NOISE_LEVEL, physics_v2=True, kernel=KERNEL, sr_factor=None,
                     physics_true_adjoint=False, partial_sure=True, sure_margin=None, partial_sure_sr=False,
                     Loss__crop_training_pairs=False, Loss__crop_size=48, ProposedLoss__stop_gradient=True,
                     ProposedLoss__sure_alternative=None, ProposedLoss__alpha_tradeoff=1.0,
                     ProposedLoss__transforms="Scaling_Transforms", ScalingTransform__kind="padded",
                     ScalingTransform__antialias=False, method="proposed", sure_cropped_div=True,
                     sure_averaged_cst=None)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


def model_args(hidden, scales):
    from argparse import Namespace
    return Namespace(task="deblurring", sr_factor=None, noise_level=NOISE_LEVEL, model_kind="Proposed",
                     ProposedModel__architecture="Convolutional", ConvolutionalModel__residual=True,
                     ConvolutionalModel__inner_residual=True, ConvolutionalModel__inout_convs=True,
                     ConvolutionalModel__hidden_channels=hidden, ConvolutionalModel__scales=scales,
                     ConvolutionalModel__num_conv_blocks=1, data_parallel_devices=None)


def network_description(args):
    if args.network == "cnn":
        return (f"reference ConvolutionalModel (hidden {args.cnn_hidden}, {args.cnn_scales} scales, 1 block per scale; "
                "random init), bf16 channels-last activations; pointwise / input 3x3 convolutions on the tcgen05 GEMM "
                "(fp32 accumulation), ideal resamplers as batched tcgen05 operator products, depthwise 7x7, channel "
                "LayerNorm, GELU, output 3x3 convolution and bias / gamma / beta reductions as hand-written kernels; "
                "Adam as one streaming kernel per tensor (also refreshes the bf16 weight copies); residual adds are PyTorch library ops")
    return ("4-parameter pointwise stand-in (tests/toy_model.py): isolates the operator + loss-assembly path; "
            "the restoration CNN is not in this line")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    import losses
    import physics
    import sei_b200
    from sei_b200 import ops
    from toy_model import ToyModel

    from sei_b200 import parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sei_b200 hot path has no CPU fallback "
                         "(use --impl reference for the CPU restatement)")
    rank, world, local_rank = parallel.init_distributed(backend="nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    global BATCH
    BATCH = args.batch
    torch.manual_seed(0)                       # identical initial weights on every rank
    largs = loss_args()
    phys = physics.get_physics(largs, device=dev)
    loss_fn = losses.get_loss(largs, phys)
    if args.network == "cnn":
        import models
        model = models.get_model(model_args(args.cnn_hidden, args.cnn_scales), physics=phys, device=dev).to(dev)
    else:
        model = ToyModel(rate=1).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    from sei_b200.optim import Adam as SeiAdam        # torch.optim.Adam's rule as one streaming kernel per tensor
    opt = SeiAdam(model.parameters(), lr=1e-4, betas=(0.9, 0.999))
    params = [p for p in model.parameters()]
    torch.manual_seed(1 + rank)                # per-rank data and draws

    # synthetic data: NBUF resident batches (x ~ U[0,1), y = A x + sigma n) -> 2*NBUF*25 MB > L2 (126 MB)
    NBUF = 8
    xs = [torch.rand(BATCH, CH, SIZE, SIZE, device=dev) for _ in range(NBUF)]
    ys = [phys(x) for x in xs]
    host_x = [x.cpu().pin_memory() for x in xs[:2]]
    host_y = [y.cpu().pin_memory() for y in ys[:2]]
    x_static, y_static = torch.empty_like(xs[0]), torch.empty_like(ys[0])
    loss_static = torch.zeros((), device=dev)
    allreduce_grads = parallel.GradAllReducer(params)      # one bucketed NCCL all-reduce (average) per step

    def fwd_bwd():
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(x=x_static, y=y_static, model=model)
        loss.backward()
        loss_static.copy_(loss.detach())

    # capture forward+backward and the optimizer step as CUDA graphs (the step is ~40 small launches)
    use_graph = not args.no_graph
    g_fb = g_opt = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    launches_per_step = 0
    with torch.cuda.stream(side):
        x_static.copy_(xs[0]); y_static.copy_(ys[0])
        for _ in range(3):
            n_before = sei_b200.launch_count()
            fwd_bwd(); allreduce_grads(); opt.step()
            launches_per_step = sei_b200.launch_count() - n_before
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    import gc
    gc.collect()
    torch.cuda.empty_cache()     # the eager warm-up's cached blocks would otherwise sit beside the graph's private pool
    print(f"[bench] rank {rank}: after eager warm-up {torch.cuda.memory_allocated(dev) / 2 ** 30:.1f} GiB allocated, "
          f"peak {torch.cuda.max_memory_allocated(dev) / 2 ** 30:.1f} GiB", file=sys.stderr)
    if use_graph:
        try:
            g_fb, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb, stream=side):
                fwd_bwd()
            with torch.cuda.graph(g_opt, stream=side):
                opt.step()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {str(e)[:200]}); running eagerly", file=sys.stderr)
            g_fb = g_opt = None
            torch.cuda.synchronize()
            gc.collect()
            torch.cuda.empty_cache()

    torch.cuda.synchronize()

    def step_resident(i):
        x_static.copy_(xs[i % NBUF]); y_static.copy_(ys[i % NBUF])
        if g_fb is not None:
            g_fb.replay(); allreduce_grads(); g_opt.replay()
        else:
            fwd_bwd(); allreduce_grads(); opt.step()

    def step_e2e(i):
        x_static.copy_(host_x[i % 2], non_blocking=True); y_static.copy_(host_y[i % 2], non_blocking=True)
        if g_fb is not None:
            g_fb.replay(); allreduce_grads(); g_opt.replay()
        else:
            fwd_bwd(); allreduce_grads(); opt.step()
        return float(loss_static.item())       # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_step = timed(step_resident, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    clock_summary = clocks.stop() if rank == 0 else None
    final_loss = float(loss_static.item())
    assert np.isfinite(final_loss), "loss diverged"

    # rooflines, measured live: (1) the operator kernel that recurs most in the step (circular blur A / A^T), back to
    # back over rotating inputs larger than L2; (2) for the CNN step, the dominant kernel: the tcgen05 GEMM
    roofline = roofline_ops = None
    if rank == 0:
        peaks = measured_peaks()
        khost = phys._kernel_host
        big = [torch.rand(128, CH, SIZE, SIZE, device=dev) for _ in range(3)]   # 3 x 100 MB in, +100 MB out
        for b_ in big:
            ops.blur_circular(b_, khost)
        torch.cuda.synchronize()
        reps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            ops.blur_circular(big[i % 3], khost)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        alg_bytes = 8.0 * big[0].numel()
        achieved = alg_bytes / (us * 1e-6) / 1e9
        roofline_ops = {"kernel": "blur_band_kernel<13> (circular Gaussian_R2 blur, A / A^T)", "bound": "hbm",
                        "achieved": round(achieved, 1), "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": round(achieved / peaks["hbm"], 4), "traffic": 151.0e6,
                        "peak_source": peaks["src"], "us_per_launch": round(us, 2),
                        "algorithmic_bytes_per_launch": alg_bytes,
                        "how": f"{reps} back-to-back launches on 128x{CH}x{SIZE}x{SIZE} fp32, 3 rotating inputs (400 MB > L2); "
                               "traffic = dram bytes of one launch from ncu --set full (profiles/)"}
        del big
        roofline = roofline_ops
        if args.network == "cnn":
            dim = args.cnn_hidden * 4 ** (args.cnn_scales - 1)
            T = BATCH * (SIZE >> (args.cnn_scales - 1)) ** 2
            a_ = torch.randn(T, dim, device=dev).bfloat16()
            w_ = torch.randn(4 * dim, dim, device=dev).bfloat16()
            for _ in range(3):
                ops.gemm_bf16_tn(a_, w_)
            torch.cuda.synchronize()
            reps = 10
            e0.record()
            for _ in range(reps):
                ops.gemm_bf16_tn(a_, w_)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            flops = 2.0 * T * dim * 4 * dim
            tf = flops / (us * 1e-6) / 1e12
            roofline = {"kernel": f"{sei_b200.last_kernel()} (tcgen05 cta_group::2; deepest ConvBlock conv2: {T}x{dim} @ ({4 * dim}x{dim})^T)",
                        "bound": "tensor", "achieved": round(tf, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "frac": round(tf / peaks["tf_sustained"], 4), "frac_of_burst_peak": round(tf / peaks["tf_burst"], 4),
                        "traffic": None, "peak_source": peaks["src"] + ", sustained cuBLAS bf16 figure",
                        "us_per_launch": round(us, 1), "algorithmic_flops_per_launch": flops,
                        "how": f"{reps} back-to-back launches, CUDA events"}
            del a_, w_

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_reference_sample(args, seconds_budget=25.0)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    imgs = BATCH * world
    out = {
        "metric": "SEI training imgs/sec @256x256 (proposed step)", "value": round(imgs / (ms_step * 1e-3), 2),
        "unit": "imgs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 network (fp32 accumulate), f32 operators" if args.network == "cnn" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.replace(f"batch 32", f"batch {BATCH}"),
                   "network": network_description(args), "n_params": n_params,
                   "global_batch": imgs, "parallelism": f"dp{world}", "cuda_graph": g_fb is not None,
                   "l2": f"inputs rotate over {NBUF} resident batches ({2 * NBUF * 25} MB > 126 MB L2)",
                   "final_loss": final_loss},
        "e2e": {"value": round(imgs / (ms_e2e * 1e-3), 2), "unit": "imgs/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": int(2 * xs[0].numel() * 4), "d2h_bytes_per_step": 4,
                "how": "losses.get_loss(...)(x, y, model) + backward + Adam from pinned host x,y; loss.item() each step"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clock_summary, "roofline": roofline, "roofline_operators": roofline_ops, "cpu_baseline": cpu_baseline,
        "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2),
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------- CPU restatement arm
def cpu_network(args):
    """The network of the step on the CPU in fp32: the same module tree with its contractions as plain torch
    matmuls (what the reference's nn.Conv2d does on the host), or the stand-in network."""
    if args.network != "cnn":
        from toy_model import ToyModel
        return ToyModel(rate=1)
    import models
    import models.convolutional as mc
    mc.COMPUTE_DTYPE = torch.float32
    mc._gemm_tn = lambda a, b, bias, out_dtype: (a @ b.t() + (bias if bias is not None else 0)).to(out_dtype)
    mc._gemm_atb = lambda a, b: (a.t() @ b).float()
    torch.manual_seed(0)
    return models.get_model(model_args(args.cnn_hidden, args.cnn_scales), physics=None, device="cpu")


def reference_step_factory(args, batch):
    """One `proposed` training step on the host: physics, transform and loss reductions by the C oracle (OpenMP),
    the network and its backward by PyTorch on the CPU (as in the reference), Adam on the parameters."""
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    kern = orc.named_kernel(KERNEL)
    phys = orc.OraclePhysics("deblurring", kernel=kern, sigma=float(np.float32(NOISE_LEVEL / 255)))
    x = rng.random((batch, CH, SIZE, SIZE), dtype=np.float32)
    y = orc.add_noise(phys.A(x), rng.standard_normal(x.shape).astype(np.float32), phys.sigma)
    net = cpu_network(args)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    tau, sigma2, alpha = np.float32(1e-2), (NOISE_LEVEL / 255) ** 2, 1.0

    def step():
        b = np.zeros_like(y)
        b[:, :, MARGIN:-MARGIN, MARGIN:-MARGIN] = rng.standard_normal((batch, CH, SIZE - 2 * MARGIN, SIZE - 2 * MARGIN)).astype(np.float32)
        u_rate, u_center = rng.random(batch, dtype=np.float32), rng.random((batch, 2), dtype=np.float32)
        noise = rng.standard_normal(y.shape).astype(np.float32)
        opt.zero_grad(set_to_none=True)
        x_net = net(torch.from_numpy(y))
        x_net2 = net(torch.from_numpy(y + b * tau))
        xn, xn2 = x_net.detach().numpy(), x_net2.detach().numpy()
        y1, y2 = phys.A(xn), phys.A(xn2)
        l_sure, _, _ = orc.sure_loss(y1, y2, y, b, MARGIN, MARGIN, float(tau), sigma2, None)
        rate, center = orc.sample_params_from_uniforms(u_rate, u_center)
        x2 = orc.scale_transform(xn, rate, center)
        y_ei = orc.add_noise(phys.A(x2), noise, phys.sigma)
        x3 = net(torch.from_numpy(y_ei))
        x3n = x3.detach().numpy()
        loss = l_sure + alpha * orc.mse(x3n, x2)
        # backward: seeds through A^T (oracle), then the network's autograd
        n_int = batch * CH * (SIZE - 2 * MARGIN) ** 2
        mask = np.zeros_like(y)
        mask[:, :, MARGIN:-MARGIN, MARGIN:-MARGIN] = 1
        g_y2 = np.float32(2.0 * sigma2 / (float(tau) * n_int)) * b * mask
        g_y1 = np.float32(2.0 / n_int) * (y1 - y) * mask - g_y2
        g1, g2 = phys.A_vjp(g_y1), phys.A_vjp(g_y2)
        g3 = np.float32(2.0 * alpha / x3n.size) * (x3n - x2)
        torch.autograd.backward([x_net, x_net2, x3], [torch.from_numpy(g1), torch.from_numpy(g2), torch.from_numpy(g3)])
        opt.step()
        return loss

    return step


def run_reference_sample(args, seconds_budget=25.0, batch=None, steps=None, warmup=1):
    """Time the CPU restatement of the step on a bounded sample, using every host core."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_threads(cores)               # torchrun exports OMP_NUM_THREADS=1; the baseline uses every host core
    torch.set_num_threads(cores)
    batch = batch or (1 if args.network == "cnn" else 8)
    step = reference_step_factory(args, batch)
    t0 = time.perf_counter()
    for _ in range(warmup):
        step()
    one = (time.perf_counter() - t0) / max(warmup, 1)
    n = steps or max(1, min(10, int(seconds_budget / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    return {"value": round(batch / dt, 4), "unit": "imgs/s", "cores": cores, "kind": "port",
            "sample": f"{n} step(s) of the same proposed step on a batch of {batch} {SIZE}x{SIZE} RGB crop(s): operators and "
                      f"loss by the oracle/ C restatement (OpenMP {cores} threads), network fwd/bwd + Adam by PyTorch CPU fp32 "
                      f"({cores} threads), {dt * 1e3:.1f} ms/step",
            "ms_per_step": round(dt * 1e3, 2), "batch": batch}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base = run_reference_sample(args, steps=min(args.steps, 3 if args.network == "cnn" else 10), warmup=1)
    out = {"impl": "reference", "metric": "SEI training imgs/sec @256x256 (proposed step)", "value": base["value"],
           "unit": "imgs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "network": network_description(args).split(";")[0] + " (fp32 on the host)",
                      "global_batch": base["batch"], "parallelism": "cpu"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--network", default="cnn", choices=["cnn", "standin"])
    ap.add_argument("--cnn-hidden", type=int, default=32)
    ap.add_argument("--cnn-scales", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
