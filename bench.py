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

Prints ONE JSON line (rank 0).  --impl reference times the same step on the host cores: the reference's OWN
modules (physics, losses, transforms, models/convolutional.py, staged unmodified under the git-ignored
baseline/_ref/ by __graft_entry__.build(); deepinv is replaced by tests/golden/deepinv_shim) when that copy is
present (cpu_baseline.kind = "reference"), else the oracle/ C restatement (kind = "port").  Each of its K + W steps
is a bounded sample of the workload (a smaller batch of the same 256x256 crops) so that the run ends within minutes.
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
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")
SHIM = os.path.join(ROOT, "tests", "golden", "deepinv_shim")


def _path(*dirs):
    for p in dirs:
        if p not in sys.path:
            sys.path.insert(0, p)


# the B200 arm imports the package's mirrors (physics, losses, ...); the reference arm imports the reference's own
# modules of the same names from baseline/_ref/src, so the two never share a process (see run_reference_subprocess)
_path(ROOT, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH, SIZE, CH = 32, 256, 3
KERNEL, NOISE_LEVEL, MARGIN = "Gaussian_R2", 5, 6
METRIC = "SEI training imgs/sec @256x256 (proposed step)"
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
                "random init), bf16 channels-last activations; pointwise convolutions on the tcgen05 GEMM (fp32 accumulation; "
                "bias, residual / skip additions, GELU and its backward in the epilogues), 3x3 in / out convolutions as "
                "implicit GEMMs on tcgen05, ideal resamplers as batched tcgen05 operator products, depthwise 7x7 on TMA-staged "
                "halo tiles, channel LayerNorm, GELU and bias / gamma / beta reductions as hand-written kernels; every "
                "ConvBlock one autograd node; Adam as one streaming kernel per tensor (also refreshes the bf16 weight copies)")
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


def bench_config(args, world):
    """`config` of the JSON line; the reference arm prints the B200 arm's config verbatim (same workload, same network)"""
    return {"workload": WORKLOAD.replace("batch 32", f"batch {args.batch}"),
            "network": network_description(args), "global_batch": args.batch * world, "parallelism": f"dp{world}",
            "dp_mode": args.dp_mode if world > 1 else "none",
            "l2": "inputs rotate over 8 resident batches (2 x 8 x 25 MB > 126 MB L2)"}


# ----------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    _path(PKG)
    import torch.distributed as dist
    import losses
    import physics
    import sei_b200
    from sei_b200 import ops
    from toy_model import ToyModel

    from sei_b200 import parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sei_b200 hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
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
    # gradient buckets follow the module tree (sei_b200.parallel.completion_groups) so that each can be all-reduced as
    # soon as the backward pass has produced it
    pipelined = args.dp_mode == "pipelined" and world > 1
    reducer = parallel.GradAllReducer(params, module=model if args.dp_mode == "overlap" else None,
                                      **({"max_elems": 64 * 2 ** 20} if pipelined else {}))
    #   pipelined: graph(forward + backward), then per bucket (three for the default network: the two 268 M-parameter
    #            convolutions of the deepest block end a bucket each) an asynchronous all-reduce followed by that
    #            bucket's own Adam graph -- the optimizer pass of bucket i runs while bucket i + 1 is on the wire.
    opt_parts = None
    if pipelined:
        opt_parts = [SeiAdam(b, lr=1e-4, betas=(0.9, 0.999)) for b in reducer.buckets if b]

        class _Parts:
            def zero_grad(self, set_to_none=False):
                for o in opt_parts:
                    o.zero_grad(set_to_none=set_to_none)

            def step(self):
                for o in opt_parts:
                    o.step()
        opt = _Parts()

    # Data-parallel modes (--dp-mode, only matters for N > 1):
    #   overlap: forward + backward + the per-group gradient all-reduces (launched from inside backward() on NCCL's
    #            stream, overlapping the rest of the backward pass) + Adam captured as ONE CUDA graph;
    #   serial:  round 1's scheme -- graph(forward + backward), one eager all-reduce of every bucket, graph(Adam).
    overlap = args.dp_mode == "overlap" and world > 1

    def fwd_bwd(reduce_inside):
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(x=x_static, y=y_static, model=model)
        if reduce_inside:
            reducer.arm()
        loss.backward()
        if reduce_inside:
            reducer.finish()
        loss_static.copy_(loss.detach())

    # Capturing torch's NCCL collectives in a CUDA graph hangs on this stack (torch 2.11 / NCCL 2.28.9, both from the
    # autograd thread and deferred to the capturing thread behind events; measured twice on 2 GPUs), so the overlapped
    # mode runs forward + backward eagerly and only the optimizer step as a graph.
    use_graph = not args.no_graph
    g_step = g_opt = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    launches_per_step, flops_per_step = 0, 0.0
    with torch.cuda.stream(side):
        x_static.copy_(xs[0]); y_static.copy_(ys[0])
        for _ in range(3):
            n_before, f_before = sei_b200.launch_count(), ops.flop_count()
            fwd_bwd(overlap)
            if not overlap:
                reducer()
            opt.step()
            launches_per_step = sei_b200.launch_count() - n_before
            flops_per_step = ops.flop_count() - f_before
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    import gc
    gc.collect()
    torch.cuda.empty_cache()     # the eager warm-up's cached blocks would otherwise sit beside the graph's private pool
    print(f"[bench] rank {rank}: after eager warm-up {torch.cuda.memory_allocated(dev) / 2 ** 30:.1f} GiB allocated, "
          f"peak {torch.cuda.max_memory_allocated(dev) / 2 ** 30:.1f} GiB", file=sys.stderr)
    if use_graph:
        try:
            g_step = torch.cuda.CUDAGraph()
            # thread_local: NCCL's watchdog thread keeps making CUDA calls while this thread captures
            mode = dict(capture_error_mode="thread_local") if world > 1 else {}
            if not overlap:
                with torch.cuda.graph(g_step, stream=side, **mode):
                    fwd_bwd(False)
                    if world == 1:
                        opt.step()
            else:
                g_step = None
            if pipelined:
                g_opt = []
                for o in opt_parts:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side, **mode):
                        o.step()
                    g_opt.append(g)
            elif world > 1:
                g_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_opt, stream=side, **mode):
                    opt.step()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {str(e)[:200]}); running eagerly", file=sys.stderr)
            g_step = g_opt = None
            torch.cuda.synchronize()
            gc.collect()
            torch.cuda.empty_cache()

    torch.cuda.synchronize()

    def one_step():
        if g_step is not None:
            g_step.replay()
        else:
            fwd_bwd(overlap)
        if pipelined and g_opt is not None:
            works = [dist.all_reduce(reducer.flat[i], op=dist.ReduceOp.AVG, async_op=True)
                     for i, b in enumerate(reducer.buckets) if b]
            for w, g in zip(works, g_opt):
                w.wait()                      # the compute stream waits for THIS bucket only; the next one is on the wire
                g.replay()
            return
        if world > 1 and not overlap:
            reducer()
        if g_opt is not None:
            g_opt.replay()
        elif g_step is None or world > 1:
            opt.step()

    def step_resident(i):
        x_static.copy_(xs[i % NBUF]); y_static.copy_(ys[i % NBUF])
        one_step()

    def step_e2e(i):
        x_static.copy_(host_x[i % 2], non_blocking=True); y_static.copy_(host_y[i % 2], non_blocking=True)
        one_step()
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
        # the launches are captured in a CUDA graph: the host's per-call cost (ctypes + torch.empty) is longer than this
        # kernel, and an eager loop would time the host (65 us per call against 43 us of kernel)
        g_blur, s_blur = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        s_blur.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_blur):
            with torch.cuda.graph(g_blur, stream=s_blur):
                for i in range(reps):
                    ops.blur_circular(big[i % 3], khost)
        torch.cuda.current_stream().wait_stream(s_blur)
        g_blur.replay()
        torch.cuda.synchronize()
        e0.record()
        g_blur.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        del g_blur
        alg_bytes = 8.0 * big[0].numel()
        achieved = alg_bytes / (us * 1e-6) / 1e9
        roofline_ops = {"kernel": "blur_band_kernel<13> (circular Gaussian_R2 blur, A / A^T)", "bound": "hbm",
                        "achieved": round(achieved, 1), "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": round(achieved / peaks["hbm"], 4), "traffic": None,
                        "peak_source": peaks["src"], "us_per_launch": round(us, 2),
                        "algorithmic_bytes_per_launch": alg_bytes,
                        "how": f"{reps} back-to-back launches (one CUDA-graph replay) on 128x{CH}x{SIZE}x{SIZE} fp32, 3 rotating inputs (400 MB > L2)"}
        roofline_ops["traffic"], roofline_ops["traffic_source"] = ncu_traffic("blur_band_kernel", "r02_ncu_blur_band_traffic.json")
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
            step_tf = flops_per_step / (ms_step * 1e-3) / 1e12
            # the kernel is timed alone (10 launches, ~30 ms): the burst cuBLAS figure is its denominator; the whole
            # step (seconds under the power cap) is compared with the sustained figure
            roofline = {"kernel": f"{sei_b200.last_kernel()} (tcgen05 cta_group::2; deepest ConvBlock conv2: {T}x{dim} @ ({4 * dim}x{dim})^T)",
                        "bound": "tensor", "achieved": round(tf, 1), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                        "frac": round(tf / peaks["tf_burst"], 4), "frac_of_sustained_peak": round(tf / peaks["tf_sustained"], 4),
                        "traffic": None, "peak_source": peaks["src"] + ", burst cuBLAS bf16 figure (kernel timed alone)",
                        "us_per_launch": round(us, 1), "algorithmic_flops_per_launch": flops,
                        "step_tensor_flops": flops_per_step, "step_tflops": round(step_tf, 1),
                        "step_frac_of_sustained_peak": round(step_tf / peaks["tf_sustained"], 4),
                        "how": f"{reps} back-to-back launches, CUDA events; step_*: tensor-core flops of every GEMM launched in one "
                               "step (counted by sei_b200.ops) / the timed step"}
            roofline["traffic"], roofline["traffic_source"] = ncu_traffic("gemm_bf16_tn_2cta_kernel", "r02_ncu_gemm_2cta_s4_traffic.json")
            del a_, w_

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_reference_subprocess(args, seconds_budget=25.0)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    imgs = BATCH * world
    out = {
        "metric": METRIC, "value": round(imgs / (ms_step * 1e-3), 2),
        "unit": "imgs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 network (fp32 accumulate), f32 operators" if args.network == "cnn" else "f32", "data": "synthetic",
        "config": bench_config(args, world),
        "n_params": n_params, "cuda_graph": g_step is not None or g_opt is not None, "final_loss": final_loss,
        "e2e": {"value": round(imgs / (ms_e2e * 1e-3), 2), "unit": "imgs/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": int(2 * xs[0].numel() * 4), "d2h_bytes_per_step": 4,
                "how": "losses.get_loss(...)(x, y, model) + backward + Adam from pinned host x,y; loss.item() each step"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clock_summary, "roofline": roofline, "roofline_operators": roofline_ops, "cpu_baseline": cpu_baseline,
        "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2),
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------- CPU arm
def reference_step_own(args, batch):
    """One `proposed` training step by the REFERENCE'S OWN CODE on the host (fp32, PyTorch CPU): get_physics,
    get_loss and ConvolutionalModel are the unmodified modules staged under baseline/_ref/src (demo/train.py:79-186,
    258-270: zero_grad, loss(x=, y=, model=), backward, Adam step, loss.item())."""
    _path(SHIM, REF_SRC)
    import importlib.util
    import physics as ref_physics
    import losses as ref_losses
    assert os.path.abspath(ref_physics.__file__).startswith(REF_SRC), ref_physics.__file__
    largs = loss_args()
    dev = "cpu"
    phys = ref_physics.get_physics(largs, device=dev)
    loss_fn = ref_losses.get_loss(args=largs, physics=phys)
    if args.network == "cnn":
        # src/models/__init__.py imports deepinv.models / bm3d (not installed): the CNN's own file is loaded by path
        spec = importlib.util.spec_from_file_location("ref_convolutional", os.path.join(REF_SRC, "models", "convolutional.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)
        # the constructor arguments src/models/__init__.py:76-86 passes for task=deblurring
        net = mod.ConvolutionalModel(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1,
                                     hidden_channels=args.cnn_hidden, inout_convs=True, scales=args.cnn_scales)
    else:
        from toy_model import ToyModel
        net = ToyModel(rate=1)

    class Wrapped(torch.nn.Module):          # Model.forward(x, *args) of src/models/__init__.py:148-149
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, *a):
            return self.m(x)

    model = Wrapped(net)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.999))
    torch.manual_seed(1)
    x = torch.rand(batch, CH, SIZE, SIZE)
    with torch.no_grad():
        y = phys(x)

    def step():
        opt.zero_grad()
        loss = loss_fn(x=x, y=y, model=model)
        loss.backward()
        opt.step()
        return float(loss.item())

    return step, "reference"


def cpu_network(args):
    """The network of the port's step on the CPU in fp32: this package's module tree with its contractions as plain
    torch matmuls, or the stand-in network.  Only used when baseline/_ref/ is absent."""
    if args.network != "cnn":
        from toy_model import ToyModel
        return ToyModel(rate=1)
    _path(PKG)
    import models
    import models.convolutional as mc
    import torch_formulation                 # tests/torch_formulation.py: fp32 library formulation of the operator hooks
    torch_formulation.install(mc)
    torch.manual_seed(0)
    return models.get_model(model_args(args.cnn_hidden, args.cnn_scales), physics=None, device="cpu")


def reference_step_port(args, batch):
    """One `proposed` training step on the host: physics, transform and loss reductions by the C oracle (OpenMP),
    the network and its backward by PyTorch on the CPU (as in the reference), Adam on the parameters."""
    from oracle import oracle as orc
    orc.set_threads(os.cpu_count() or 1)
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

    return step, "port"


def run_reference_sample(args, seconds_budget, steps, warmup):
    """Time `warmup` + `steps` steps of the CPU arm, each on a bounded sample of the workload (a batch of `b` of the
    same 256x256 crops, b chosen from a one-crop probe so that the whole run fits `seconds_budget`), every host core."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)         # torchrun exports OMP_NUM_THREADS=1; the baseline uses every host core
    own = os.path.isdir(REF_SRC) and not args.force_port
    factory = reference_step_own if own else reference_step_port
    n_total = steps + warmup
    batch = args.ref_batch
    probe_s = None
    if not batch:
        step, kind = factory(args, 1)
        step()                           # first call pays allocator / thread-pool start-up
        t0 = time.perf_counter()
        step()
        probe_s = time.perf_counter() - t0
        batch = int(max(1, min(args.batch, seconds_budget / max(n_total, 1) / max(probe_s, 1e-3))))
        del step
    step, kind = factory(args, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    what = ("the reference's own physics / losses / transforms / ConvolutionalModel (baseline/_ref/src, deepinv names from "
            "tests/golden/deepinv_shim), PyTorch CPU fp32, torch.optim.Adam" if kind == "reference" else
            "operators and loss by the oracle/ C restatement (OpenMP), network fwd/bwd + Adam by PyTorch CPU fp32")
    return {"value": round(batch / dt, 4), "unit": "imgs/s", "cores": cores, "kind": kind,
            "sample": f"{steps} timed + {warmup} warm-up step(s) of the same proposed step, each on a batch of {batch} of the "
                      f"workload's {SIZE}x{SIZE} RGB crops (of {args.batch} per GPU step): {what}, {cores} threads, "
                      f"{dt * 1e3:.1f} ms/step" + (f"; one-crop probe {probe_s * 1e3:.0f} ms" if probe_s else ""),
            "ms_per_step": round(dt * 1e3, 2), "batch": batch, "steps": steps, "warmup": warmup}


def run_reference_subprocess(args, seconds_budget):
    """cpu_baseline leg of the B200 arm: the CPU arm in a process of its own (it imports the reference's modules, whose
    names -- physics, losses, transforms -- are those of this package's mirrors already imported here)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--network", args.network, "--cnn-hidden", str(args.cnn_hidden), "--cnn-scales", str(args.cnn_scales),
           "--batch", str(args.batch), "--ref-budget-s", str(seconds_budget)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "imgs/s", "cores": os.cpu_count(), "kind": "unavailable",
                "sample": f"CPU arm failed: {type(e).__name__}: {str(e)[:200]}"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base = run_reference_sample(args, seconds_budget=args.ref_budget_s, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": base["value"],
           "unit": "imgs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": bench_config(args, args.gpus),
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))



def ncu_traffic(kernel_substr, fname):
    """DRAM bytes per launch of a kernel from a committed `ncu --set full` capture of the same launch shape
    (profiles/<fname>, written on the GPU box by benchmarks/ncu_summary.py: dram__bytes_read.sum + dram__bytes_write.sum).
    Returns (bytes, provenance) or (None, why)."""
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(path):
        return None, f"no capture at profiles/{fname}"
    try:
        d = json.load(open(path))
        ls = [l for l in d["launches"] if kernel_substr in l["kernel"]]
        if not ls:
            return None, f"profiles/{fname} has no launch of {kernel_substr}"
        b = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in ls) / len(ls)
        return b, f"profiles/{fname}: mean of {len(ls)} launch(es), {d['source']}"
    except Exception as e:  # noqa: BLE001
        return None, f"profiles/{fname} unreadable ({type(e).__name__})"


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
    ap.add_argument("--dp-mode", default="serial", choices=["overlap", "serial", "pipelined"],
                    help="N > 1: all-reduce the gradient buckets from inside backward() (one captured graph) or after it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=200.0,
                    help="--impl reference: wall-clock budget of the whole K + W step run (sizes the per-step sample)")
    ap.add_argument("--ref-batch", type=int, default=0, help="--impl reference: crops per step (0 = from the budget)")
    ap.add_argument("--force-port", action="store_true", help="--impl reference: use the oracle port even if baseline/_ref exists")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
