#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):

    python tests/golden/make_golden.py

It puts the reference's own, unmodified `src/` on sys.path together with the small
`deepinv` stand-in next to this file (deepinv v0.2.0 is the reference's un-vendored
dependency; see deepinv_shim/deepinv/__init__.py), executes the reference operators and
losses on seeded inputs on the CPU, records every random tensor the reference draws, and
stores inputs, draws and outputs as .npz.  The fixtures pin the oracle (oracle/) and the
CUDA path; nothing here is imported by the product.

Reference entry points exercised (file:line in /root/reference):
  src/physics/kernels.py:13-28            get_kernel
  src/physics/blur/__init__.py:34-161     conv / conv_transpose (v1, all paddings)
  src/physics/blur/__init__.py:164-227    Blur, BlurV2 (A, A_adjoint)
  src/physics/downsampling/__init__.py    Downsampling (A, autograd vjp, A_adjoint)
  src/physics/__init__.py:29-102          PhysicsManager / get_physics / randomly_degrade
  src/transforms.py:5-109                 sample_downsampling_parameters, padded transform
  src/losses/sure.py:7-76                 mc_div, SureGaussianLoss
  src/losses/__init__.py:13-266           Loss / ProposedLoss / SURELoss / SupervisedLoss / get_loss
  src/crop.py:8-57                        CropPair (incl. the batched-input quirk)
"""
import os
import sys
import zlib
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
sys.path.insert(0, os.path.join(HERE, "deepinv_shim"))
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.join(HERE, ".."))

import physics as ref_physics  # noqa: E402  (the reference's package)
from physics import get_physics, Blur, BlurV2, Downsampling  # noqa: E402
from physics.kernels import get_kernel  # noqa: E402
from physics.blur import conv, conv_transpose  # noqa: E402
import transforms as ref_transforms  # noqa: E402
import losses as ref_losses  # noqa: E402
from losses.sure import SureGaussianLoss  # noqa: E402
from crop import CropPair  # noqa: E402

from toy_model import ToyModel  # noqa: E402  (tests/toy_model.py, shared with the tests)

assert ref_physics.__file__.startswith(REF_SRC), ref_physics.__file__
assert ref_transforms.__file__.startswith(REF_SRC)
assert ref_losses.__file__.startswith(REF_SRC)

torch.set_num_threads(4)


class DrawRecorder:
    """Wrap torch.rand / randn / randn_like / randint so every draw is logged in order."""

    NAMES = ("rand", "randn", "randn_like", "randint", "randperm")

    def __init__(self):
        self.draws = []
        self._orig = {}

    def __enter__(self):
        for n in self.NAMES:
            self._orig[n] = getattr(torch, n)

            def wrapped(*a, __n=n, **k):
                out = self._orig[__n](*a, **k)
                self.draws.append((__n, out.detach().clone()))
                return out

            setattr(torch, n, wrapped)
        return self

    def __exit__(self, *exc):
        for n in self.NAMES:
            setattr(torch, n, self._orig[n])

    def as_dict(self, prefix="draw"):
        return {f"{prefix}{i}_{n}": t.numpy() for i, (n, t) in enumerate(self.draws)}


def np_(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


def base_args(**kw):
    a = dict(task="deblurring", noise_level=5, physics_v2=True, kernel="Gaussian_R2", sr_factor=None,
             physics_true_adjoint=False, partial_sure=True, sure_margin=None, partial_sure_sr=False,
             Loss__crop_training_pairs=False, Loss__crop_size=48, ProposedLoss__stop_gradient=True,
             ProposedLoss__sure_alternative=None, ProposedLoss__alpha_tradeoff=1.0,
             ProposedLoss__transforms="Scaling_Transforms", ScalingTransform__kind="padded",
             ScalingTransform__antialias=False, method="proposed", sure_cropped_div=True,
             sure_averaged_cst=None)
    a.update(kw)
    return Namespace(**a)


# --------------------------------------------------------------------------------------
def gen_kernels():
    out = {}
    for name in ["Gaussian_R1", "Gaussian_R2", "Gaussian_R3", "Box_R2", "Box_R3", "Box_R4"]:
        out[name] = np_(get_kernel(name))
    save("kernels", **out)


def gen_blur():
    out = {}
    g = torch.Generator().manual_seed(1234)
    cases = [("Gaussian_R2", (2, 3, 20, 28)), ("Box_R3", (2, 3, 20, 28)), ("Gaussian_R1", (1, 3, 17, 15)),
             ("Gaussian_R3", (1, 2, 19, 40)), ("Box_R2", (3, 1, 16, 16)), ("Box_R4", (1, 3, 9, 12)),
             ("Gaussian_R2", (3, 3, 48, 48))]
    for ci, (kname, shape) in enumerate(cases):
        kernel = get_kernel(kname).unsqueeze(0).unsqueeze(0)
        x64 = torch.rand(shape, dtype=torch.float64, generator=g)
        gy64 = torch.randn(shape, dtype=torch.float64, generator=g)
        out[f"c{ci}_kernel_name"] = np.array(kname)
        out[f"c{ci}_x"] = np_(x64)
        out[f"c{ci}_gy"] = np_(gy64)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            x, gy = x64.to(dt), gy64.to(dt)
            v2 = BlurV2(kernel=kernel)
            out[f"c{ci}_v2_A_{tag}"] = np_(v2.A(x))
            out[f"c{ci}_v2_At_{tag}"] = np_(v2.A_adjoint(gy))
            # autograd backward of A (what a training step uses)
            xr = x.clone().requires_grad_(True)
            (v2.A(xr) * gy).sum().backward()
            out[f"c{ci}_v2_vjp_{tag}"] = np_(xr.grad)
        # v1 only runs in fp32 (extend_filter builds a float32 filter)
        v1 = Blur(filter=kernel, padding="circular", device="cpu")
        x, gy = x64.float(), gy64.float()
        out[f"c{ci}_v1_A_f32"] = np_(v1.A(x))
        out[f"c{ci}_v1_At_f32"] = np_(v1.A_adjoint(gy))
    save("blur", **out)

    # non-default paddings of the v1 operator, incl. even-sized and non-separable filters
    out = {}
    x = torch.rand((2, 3, 18, 22), generator=g)
    out["x"] = np_(x)
    filt = {"g5": get_kernel("Gaussian_R1")[1:6, 1:6].clone(), "box7": get_kernel("Box_R3"),
            "rand4x5": torch.rand((4, 5), dtype=torch.float64, generator=g),
            "rand3x3": torch.rand((3, 3), dtype=torch.float64, generator=g),
            "row1x5": torch.rand((1, 5), dtype=torch.float64, generator=g)}
    for fname, f in filt.items():
        f = (f / f.sum()).float().unsqueeze(0).unsqueeze(0)
        out[f"{fname}_filter"] = np_(f)
        for padding in ["valid", "circular", "replicate", "reflect"]:
            y = conv(x, f, padding)
            out[f"{fname}_{padding}_A"] = np_(y)
            gy = torch.randn(y.shape, generator=g)
            out[f"{fname}_{padding}_gy"] = np_(gy)
            out[f"{fname}_{padding}_At"] = np_(conv_transpose(gy, f, padding))
        gy = torch.randn(x.shape, generator=g)
        out[f"{fname}_zero_gy"] = np_(gy)
        out[f"{fname}_zero_At"] = np_(conv_transpose(gy, f, "zero"))
    save("blur_paddings", **out)


def gen_downsampling():
    out = {}
    g = torch.Generator().manual_seed(4321)
    cases = [(2, (2, 3, 32, 24)), (4, (2, 3, 32, 24)), (2, (1, 3, 31, 27)), (4, (1, 2, 30, 45)),
             (3, (1, 3, 30, 27)), (2, (2, 3, 48, 48)), (4, (1, 3, 64, 48)), (2, (1, 1, 6, 4))]
    for ci, (rate, shape) in enumerate(cases):
        x64 = torch.rand(shape, dtype=torch.float64, generator=g)
        out[f"c{ci}_rate"] = np.array(rate)
        out[f"c{ci}_x"] = np_(x64)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            phys = Downsampling(rate=rate, antialias=True)
            x = x64.to(dt)
            y = phys.A(x)
            out[f"c{ci}_A_{tag}"] = np_(y)
            if tag == "f64":
                gy64 = torch.randn(y.shape, dtype=torch.float64, generator=g)
                out[f"c{ci}_gy"] = np_(gy64)
            gy = gy64.to(dt)
            xr = x.clone().requires_grad_(True)
            (phys.A(xr) * gy).sum().backward()
            out[f"c{ci}_vjp_{tag}"] = np_(xr.grad)
            # the deprecated non-adjoint "adjoint" (plain bicubic upsample)
            out[f"c{ci}_At_plain_{tag}"] = np_(phys.A_adjoint(gy))
            if shape[-2] % rate == 0 and shape[-1] % rate == 0:
                phys_t = Downsampling(rate=rate, antialias=True, true_adjoint=True)
                out[f"c{ci}_At_true_{tag}"] = np_(phys_t.A_adjoint(gy))
    save("downsampling", **out)


def gen_transform():
    out = {}
    g = torch.Generator().manual_seed(777)
    cases = [(4, 3, 24), (3, 1, 48), (2, 3, 33), (8, 3, 48)]
    for ci, (B, C, S) in enumerate(cases):
        x64 = torch.rand((B, C, S, S), dtype=torch.float64, generator=g)
        rate64 = torch.tensor([0.75, 0.5], dtype=torch.float64)[torch.randint(0, 2, (B,), generator=g)]
        rate64[0], rate64[-1] = 0.75, 0.5
        center64 = 2 * torch.rand((B, 1, 1, 2), dtype=torch.float64, generator=g) - 1
        if ci == 0:
            center64[0] = 0.0
            center64[1, 0, 0, 0], center64[1, 0, 0, 1] = 1.0, -1.0
        out[f"c{ci}_x"] = np_(x64)
        out[f"c{ci}_rate"] = np_(rate64)
        out[f"c{ci}_center"] = np_(center64)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            y = ref_transforms.padded_downsampling_transform(
                x64.to(dt), downsampling_rate=rate64.to(dt), center=center64.to(dt),
                mode="bicubic", padding_mode="reflection", antialiased=False)
            out[f"c{ci}_T_{tag}"] = np_(y)
            grid = ref_transforms.get_downsampling_grid(
                shape=x64.shape, downsampling_rate=rate64.to(dt), center=center64.to(dt), dtype=dt, device="cpu")
            out[f"c{ci}_grid_{tag}"] = np_(grid)
    # parameter sampling: draw order and mapping from uniforms to (rate, centre)
    with DrawRecorder() as rec:
        torch.manual_seed(0)
        rate, center = ref_transforms.sample_downsampling_parameters(
            image_count=16, device="cpu", dtype=torch.float32, rates=[0.75, 0.5])
    out.update(rec.as_dict("params_draw"))
    out["params_rate"] = np_(rate)
    out["params_center"] = np_(center)
    # module form (ScalingTransform): draws + output
    with DrawRecorder() as rec:
        torch.manual_seed(5)
        x = torch.rand((4, 3, 24, 24), generator=g)
        T = ref_transforms.ScalingTransform(kind="padded", antialias=False)
        y = T(x)
    out.update(rec.as_dict("module_draw"))
    out["module_x"] = np_(x)
    out["module_T"] = np_(y)
    save("transform", **out)


class Tap(torch.nn.Module):
    """Wrap a model; keep every output (with retain_grad) so that dL/d(output) can be saved."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.outs = []

    def forward(self, y, *args):
        out = self.model(y)
        if out.requires_grad:
            out.retain_grad()
        self.outs.append(out)
        return out


def gen_losses():
    g = torch.Generator().manual_seed(99)
    cases = [
        ("deblur_gauss2_proposed", dict(task="deblurring", kernel="Gaussian_R2", method="proposed"), (4, 3, 32, 32)),
        ("deblur_box3_proposed", dict(task="deblurring", kernel="Box_R3", method="proposed"), (2, 3, 24, 24)),
        ("deblur_gauss2_v1_proposed", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", physics_v2=False), (2, 3, 32, 32)),
        ("sr2_proposed", dict(task="sr", kernel=None, sr_factor=2, method="proposed"), (4, 3, 16, 16)),
        ("sr4_proposed", dict(task="sr", kernel=None, sr_factor=4, method="proposed"), (2, 3, 12, 12)),
        ("sr2_partial_proposed", dict(task="sr", kernel=None, sr_factor=2, method="proposed", partial_sure_sr=True), (2, 3, 16, 16)),
        ("deblur_gauss2_sure", dict(task="deblurring", kernel="Gaussian_R2", method="sure"), (2, 3, 32, 32)),
        ("deblur_gauss2_sure_avgcst", dict(task="deblurring", kernel="Gaussian_R2", method="sure", sure_averaged_cst=True), (2, 3, 32, 32)),
        ("deblur_gauss2_sure_nocrop", dict(task="deblurring", kernel="Gaussian_R2", method="sure", sure_cropped_div=False), (2, 3, 32, 32)),
        ("deblur_gauss2_supervised", dict(task="deblurring", kernel="Gaussian_R2", method="supervised"), (2, 3, 32, 32)),
        ("sr2_css", dict(task="sr", kernel=None, sr_factor=2, method="css"), (2, 3, 16, 16)),
        ("deblur_gauss2_proposed_alpha", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__alpha_tradeoff=0.3), (2, 3, 32, 32)),
        ("cfg1_deblur_gauss2_proposed", dict(task="deblurring", kernel="Gaussian_R2", method="proposed"), (8, 3, 48, 48)),
        # section 8(f) N3 variants
        ("deblur_gauss2_r2r", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__sure_alternative="r2r"), (2, 3, 32, 32)),
        ("sr2_r2r", dict(task="sr", kernel=None, sr_factor=2, method="proposed", ProposedLoss__sure_alternative="r2r"), (2, 3, 16, 16)),
        ("deblur_gauss2_nostopgrad", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__stop_gradient=False), (2, 3, 32, 32)),
        ("deblur_gauss2_shifts", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__transforms="Shifts"), (2, 3, 32, 32)),
        ("sr2_nostopgrad", dict(task="sr", kernel=None, sr_factor=2, method="proposed", ProposedLoss__stop_gradient=False), (2, 3, 16, 16)),
        ("deblur_gauss2_normalT", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ScalingTransform__kind="normal"), (2, 3, 32, 32)),
        ("deblur_gauss2_rotations", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__transforms="Rotations"), (2, 3, 32, 32)),
        ("deblur_gauss2_rotshift", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ProposedLoss__transforms="Rotations+Shifts"), (2, 3, 32, 32)),
        ("deblur_gauss2_normalT_aa", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", ScalingTransform__kind="normal", ScalingTransform__antialias=True), (2, 3, 32, 32)),
        # the reference's DEFAULT Loss path (demo/train.py:39,53): one random 48x48 crop of the batch (two CPU randint
        # draws, src/crop.py:26-27) -- with the 4-D MinSizePadding quirk -- before the method loss
        ("deblur_gauss2_proposed_crop", dict(task="deblurring", kernel="Gaussian_R2", method="proposed", Loss__crop_training_pairs=True), (2, 3, 64, 64)),
        ("sr2_proposed_crop", dict(task="sr", kernel=None, sr_factor=2, method="proposed", Loss__crop_training_pairs=True), (2, 3, 56, 56)),
        ("deblur_gauss2_supervised_crop32", dict(task="deblurring", kernel="Gaussian_R2", method="supervised", Loss__crop_training_pairs=True, Loss__crop_size=32), (2, 3, 40, 48)),
    ]
    only = os.environ.get("GOLDEN_ONLY", "")
    for name, kw, yshape in cases:
        if only and only not in name:
            continue
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            if tag == "f64" and kw.get("physics_v2") is False:
                continue  # v1 Blur is fp32-only
            args = base_args(**kw)
            physics = get_physics(args, device="cpu")
            loss_fn = ref_losses.get_loss(args=args, physics=physics)
            rate = args.sr_factor if args.task == "sr" else 1
            B, C, H, W = yshape
            gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000 + 17)
            x = torch.rand((B, C, H * rate, W * rate), dtype=torch.float64, generator=gen).to(dt)
            with torch.no_grad():
                y = physics.A(x) + (5 / 255) * torch.randn((B, C, H, W), dtype=torch.float64, generator=gen).to(dt)
            model = Tap(ToyModel(rate=rate).to(dt))
            with DrawRecorder() as rec:
                torch.manual_seed(3)
                loss = loss_fn(x=x, y=y, model=model)
            loss.backward()
            out = {"x": np_(x), "y": np_(y), "loss": np_(loss), "rate": np.array(rate)}
            out.update(rec.as_dict())
            for pname, p in model.model.named_parameters():
                out[f"param_{pname}"] = np_(p)
                out[f"grad_{pname}"] = np_(p.grad)
            for i, o in enumerate(model.outs):
                out[f"model_out{i}"] = np_(o)
                if o.grad is not None:
                    out[f"model_out{i}_grad"] = np_(o.grad)
            save(f"loss_{name}_{tag}", **out)


def gen_crop():
    out = {}
    g = torch.Generator().manual_seed(31)
    # dataset-style 3-D use (src/datasets/__init__.py:84): (C,H,W) inputs
    x3 = torch.rand((3, 64, 80), generator=g)
    y3 = torch.rand((3, 32, 40), generator=g)
    with DrawRecorder() as rec:
        torch.manual_seed(11)
        xc, yc = CropPair(location="random", size=24)(x3, y3, xy_size_ratio=2)
    out.update(rec.as_dict("d3_draw"))
    out.update(d3_x=np_(x3), d3_y=np_(y3), d3_xc=np_(xc), d3_yc=np_(yc))
    xc, yc = CropPair(location="center", size=24)(x3, y3)
    out.update(d3c_xc=np_(xc), d3c_yc=np_(yc))
    # Loss.forward-style 4-D use (src/losses/__init__.py:205): the size test reads C and H
    x4 = torch.rand((2, 3, 60, 60), generator=g)
    y4 = torch.rand((2, 3, 60, 60), generator=g)
    for seed in (0, 1, 2, 3):
        with DrawRecorder() as rec:
            torch.manual_seed(seed)
            xc, yc = CropPair(location="random", size=48)(x4, y4, xy_size_ratio=1)
        out.update(rec.as_dict(f"d4_s{seed}_draw"))
        out[f"d4_s{seed}_xc"] = np_(xc)
        out[f"d4_s{seed}_yc"] = np_(yc)
    out.update(d4_x=np_(x4), d4_y=np_(y4))
    save("crop", **out)


def gen_degrade():
    """PhysicsManager.randomly_degrade (src/physics/__init__.py:65-74): seeded A + noise, RNG state preserved."""
    out = {}
    g = torch.Generator().manual_seed(55)
    for name, kw, shape in [("deblur", dict(task="deblurring", kernel="Gaussian_R2"), (1, 3, 40, 56)),
                            ("sr2", dict(task="sr", kernel=None, sr_factor=2), (1, 3, 40, 56))]:
        physics = get_physics(base_args(**kw), device="cpu")
        mgr = getattr(physics, "__manager")
        x = torch.rand(shape, generator=g)
        torch.manual_seed(123)
        before = torch.get_rng_state().clone()
        with DrawRecorder() as rec:
            y = mgr.randomly_degrade(x, seed=42)
        after = torch.get_rng_state()
        assert torch.equal(before, after)
        out.update(rec.as_dict(f"{name}_draw"))
        out[f"{name}_x"] = np_(x)
        out[f"{name}_y"] = np_(y)
        out[f"{name}_sigma"] = np_(physics.noise_model.sigma)
    save("degrade", **out)


def gen_model():
    """The reference's ConvolutionalModel (src/models/convolutional.py) at a small configuration: state dict,
    forward output and parameter gradients in fp32 on the CPU.  Loaded by file path because
    src/models/__init__.py imports deepinv.models / bm3d, which are not installed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_convolutional", os.path.join(REF_SRC, "models", "convolutional.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, kw, shape in [
        ("deblur", dict(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1,
                        hidden_channels=8, inout_convs=True, scales=3), (2, 3, 32, 32)),
        ("sr2", dict(in_channels=3, upsampling_rate=2, residual=True, inner_residual=True, num_conv_blocks=2,
                     hidden_channels=8, inout_convs=True, scales=2), (2, 3, 16, 16)),
        ("pad", dict(in_channels=3, upsampling_rate=1, residual=False, inner_residual=False, num_conv_blocks=1,
                     hidden_channels=8, inout_convs=False, scales=3), (1, 3, 30, 27)),
    ]:
        torch.manual_seed(7)
        model = mod.ConvolutionalModel(**kw)
        # non-trivial biases / norm parameters so that every term is exercised
        with torch.no_grad():
            for p_ in model.parameters():
                if p_.dim() == 1:
                    p_.add_(0.1 * torch.randn_like(p_))
        y = torch.rand(shape)
        out = model(y)
        gout = torch.randn_like(out)
        (out * gout).sum().backward()
        arrays = {"y": np_(y), "out": np_(out), "gout": np_(gout)}
        for k_, v_ in model.state_dict().items():
            arrays[f"sd::{k_}"] = np_(v_)
        for k_, p_ in model.named_parameters():
            arrays[f"grad::{k_}"] = np_(p_.grad)
        arrays["kwargs"] = np.array(repr(kw))
        save(f"model_{name}", **arrays)
    # default flags (hidden 32, 5 scales): names and shapes only (645 M parameters)
    with torch.device("meta"):
        big = mod.ConvolutionalModel(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True,
                                     num_conv_blocks=1, hidden_channels=32, inout_convs=True, scales=5)
    names = sorted(big.state_dict().keys())
    shapes = [list(big.state_dict()[n].shape) for n in names]
    save("model_default_layout", names=np.array(names), shapes=np.array([repr(s_) for s_ in shapes]),
         n_params=np.array(sum(p_.numel() for p_ in big.parameters())))


def gen_normal_transform():
    """normal_downsampling_transform (src/transforms.py:112-124) for both rates and both antialias settings, fp64 + fp32"""
    g = torch.Generator().manual_seed(4242)
    out = {}
    shapes = [(2, 3, 32, 32), (1, 2, 48, 40), (1, 1, 21, 37)]
    for i, shape in enumerate(shapes):
        x = torch.rand(shape, generator=g, dtype=torch.float64)
        out[f"x{i}"] = np_(x)
        for rate in (0.75, 0.5):
            for aa in (False, True):
                tag = f"{i}_r{int(rate * 100)}_aa{int(aa)}"
                out[f"y64_{tag}"] = np_(ref_transforms.normal_downsampling_transform(x, rate, "bicubic", aa))
                out[f"y32_{tag}"] = np_(ref_transforms.normal_downsampling_transform(x.float(), rate, "bicubic", aa))
    save("normal_transform", **out)


def gen_transform_aa():
    """padded_downsampling_transform(..., antialiased=True) (src/transforms.py:44-83): equal rates per batch (the only case
    the reference's torch.stack accepts), fp64 + fp32; and the error it raises for mixed rates"""
    g = torch.Generator().manual_seed(9090)
    out = {}
    for ci, (B, C, S, rate) in enumerate([(3, 3, 32, 0.75), (2, 1, 48, 0.5), (1, 3, 33, 0.75), (4, 2, 24, 0.5)]):
        x64 = torch.rand((B, C, S, S), dtype=torch.float64, generator=g)
        rate64 = torch.full((B,), rate, dtype=torch.float64)
        center64 = 2 * torch.rand((B, 1, 1, 2), dtype=torch.float64, generator=g) - 1
        out[f"c{ci}_x"], out[f"c{ci}_rate"], out[f"c{ci}_center"] = np_(x64), np_(rate64), np_(center64)
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            y = ref_transforms.padded_downsampling_transform(
                x64.to(dt), downsampling_rate=rate64.to(dt), center=center64.to(dt),
                mode="bicubic", padding_mode="reflection", antialiased=True)
            assert y.shape == x64.shape
            out[f"c{ci}_T_{tag}"] = np_(y)
    try:
        ref_transforms.padded_downsampling_transform(
            torch.rand(2, 1, 16, 16), downsampling_rate=torch.tensor([0.75, 0.5]), center=torch.zeros(2, 1, 1, 2),
            mode="bicubic", padding_mode="reflection", antialiased=True)
        msg = "no error"
    except RuntimeError as e:
        msg = str(e).splitlines()[0]
    out["mixed_rates_error"] = np.array(msg)
    save("transform_aa", **out)


def gen_rotate():
    """deepinv Rotate = torchvision.transforms.functional.rotate(x, angle), defaults (nearest, same size, zero fill): the real
    torchvision function on CPU for a set of angles and shapes, plus the module form of the shim (draw + output)"""
    from torchvision.transforms.functional import rotate
    from deepinv.transform import Rotate
    g = torch.Generator().manual_seed(606)
    out = {}
    shapes = [(2, 3, 32, 32), (1, 2, 48, 40), (1, 1, 21, 37), (1, 1, 96, 96)]
    angles = [1, 17, 36, 45, 50, 90, 123, 180, 200, 270, 301, 359]
    out["angles"] = np.array(angles)
    for i, shape in enumerate(shapes):
        x = torch.rand(shape, generator=g)
        out[f"x{i}"] = np_(x)
        for a in angles:
            out[f"y{i}_a{a}"] = np_(rotate(x, float(a)))
    x = torch.rand((2, 3, 24, 24), generator=g)
    with DrawRecorder() as rec:
        torch.manual_seed(8)
        y = Rotate()(x)
    out.update(rec.as_dict("module_draw"))
    out["module_x"], out["module_y"] = np_(x), np_(y)
    save("rotate", **out)


def gen_resample():
    """IdealUpsample / IdealDownsample of the reference's CNN (src/models/convolutional.py:48-133), float64, with the
    vector-Jacobian product autograd gives for a random output gradient (pins the transposed operators)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_convolutional", os.path.join(REF_SRC, "models", "convolutional.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(2024)
    out = {}
    shapes = [(2, 3, 16, 16), (1, 2, 32, 32), (1, 2, 48, 48), (1, 1, 24, 40), (1, 1, 6, 12), (1, 2, 64, 64)]
    for i, shape in enumerate(shapes):
        for kind, layer in (("down", mod.IdealDownsample(rate=2)), ("up", mod.IdealUpsample(rate=2))):
            x = torch.randn(shape, generator=g, dtype=torch.float64, requires_grad=True)
            y = layer(x)
            gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
            (gx,) = torch.autograd.grad(y, x, gy)
            out[f"{kind}{i}_x"], out[f"{kind}{i}_y"] = np_(x), np_(y)
            out[f"{kind}{i}_gy"], out[f"{kind}{i}_gx"] = np_(gy), np_(gx)
    save("resample", **out)


def gen_step():
    """BASELINE configs[0]: deblurring Gaussian_R2, method=proposed, one training step's loss + gradients on
    synthetic 48x48 RGB crops, batch 8, on the CPU in fp32, with the reference's own ConvolutionalModel (small
    flags: hidden 8, 3 scales) and the reference's own losses/physics."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_convolutional", os.path.join(REF_SRC, "models", "convolutional.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    kw = dict(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1,
              hidden_channels=8, inout_convs=True, scales=3)
    torch.manual_seed(11)
    net = mod.ConvolutionalModel(**kw)

    class Wrapped(torch.nn.Module):     # Model.forward(x, *args) of src/models/__init__.py:148-149
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, *args):
            return self.m(x)

    model = Wrapped(net)
    args = base_args()
    physics = get_physics(args, device="cpu")
    loss_fn = ref_losses.get_loss(args=args, physics=physics)
    gen = torch.Generator().manual_seed(2024)
    x = torch.rand((8, 3, 48, 48), generator=gen)
    with torch.no_grad():
        y = physics.A(x) + (5 / 255) * torch.randn((8, 3, 48, 48), generator=gen)
    with DrawRecorder() as rec:
        torch.manual_seed(3)
        loss = loss_fn(x=x, y=y, model=model)
    loss.backward()
    out = {"x": np_(x), "y": np_(y), "loss": np_(loss), "kwargs": np.array(repr(kw))}
    out.update(rec.as_dict())
    for k_, v_ in net.state_dict().items():
        out[f"sd::{k_}"] = np_(v_)
    for k_, p_ in net.named_parameters():
        out[f"grad::{k_}"] = np_(p_.grad)
    save("step_cfg1_cnn", **out)


def gen_step_default():
    """BASELINE configs[0] with the reference's DEFAULT network flags (--ProposedModel__architecture Convolutional:
    hidden 32, 5 scales, 645 M parameters; src/settings.py:58-60): one proposed step's loss and parameter gradients on
    synthetic 48x48 crops, batch 8, CPU fp32.  The weights are the seeded initialisation (torch.manual_seed(11); the
    mirror's constructor consumes the generator identically, checked at small flags), so only digests of the 645 M
    gradients are stored: every tensor's norm, the full gradient of tensors up to 64 Ki elements, and 16 Ki evenly
    strided samples of the larger ones."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_convolutional", os.path.join(REF_SRC, "models", "convolutional.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    kw = dict(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1,
              hidden_channels=32, inout_convs=True, scales=5)
    torch.set_num_threads(os.cpu_count() or 4)
    torch.manual_seed(11)
    net = mod.ConvolutionalModel(**kw)

    class Wrapped(torch.nn.Module):     # Model.forward(x, *args) of src/models/__init__.py:148-149
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, *args):
            return self.m(x)

    model = Wrapped(net)
    args = base_args()
    physics = get_physics(args, device="cpu")
    loss_fn = ref_losses.get_loss(args=args, physics=physics)
    gen = torch.Generator().manual_seed(2025)
    x = torch.rand((8, 3, 48, 48), generator=gen)
    with torch.no_grad():
        y = physics.A(x) + (5 / 255) * torch.randn((8, 3, 48, 48), generator=gen)
    with DrawRecorder() as rec:
        torch.manual_seed(3)
        loss = loss_fn(x=x, y=y, model=model)
    loss.backward()
    out = {"x": np_(x), "y": np_(y), "loss": np_(loss), "kwargs": np.array(repr(kw)), "init_seed": np.array(11)}
    out.update(rec.as_dict())
    with torch.no_grad():
        out["net_out"] = np_(net(y))
    for k_, p_ in net.named_parameters():
        g_ = p_.grad.detach().flatten()
        out[f"gnorm::{k_}"] = np.array(float(g_.double().norm()))
        out[f"wsum::{k_}"] = np.array(float(p_.detach().double().sum()))       # pins the seeded initialisation
        if g_.numel() <= 65536:
            out[f"grad::{k_}"] = np_(g_)
        else:
            idx = torch.linspace(0, g_.numel() - 1, 16384, dtype=torch.float64).long()
            out[f"gsample::{k_}"] = np_(g_[idx])
    save("step_cfg1_cnn_default", **out)


def gen_dagger():
    """A_dagger of the reference's physics objects (deepinv LinearPhysics.A_dagger through the shim's statement of
    upstream's conjugate gradient) for deblurring and SR x2 with both adjoint kinds, a fixed small iteration count."""
    out = {}
    g = torch.Generator().manual_seed(99)
    cases = [("deblur_g1", base_args(kernel="Gaussian_R1"), (2, 3, 24, 24)),
             ("deblur_box2_v1", base_args(kernel="Box_R2", physics_v2=False), (1, 3, 20, 28)),
             ("sr2_plain", base_args(task="sr", kernel=None, sr_factor=2), (2, 3, 32, 32)),
             ("sr2_true", base_args(task="sr", kernel=None, sr_factor=2, physics_true_adjoint=True), (2, 3, 32, 32))]
    for name, args, shape in cases:
        for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            if dt == torch.float64 and not getattr(args, "physics_v2", True):
                continue                      # v1 Blur extends its filter in fp32: float64 inputs raise in the reference
            phys = get_physics(args, device="cpu")
            phys.max_iter, phys.tol = 6, 1e-12
            x = torch.rand(shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64).to(dt)
            with torch.no_grad():
                y = phys.A(x)
                rec = phys.A_dagger(y)
            out[f"{name}_y_{tag}"], out[f"{name}_dagger_{tag}"] = np_(y), np_(rec)
        out[f"{name}_x"] = np_(x)
    save("dagger", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["step", "step_default", "dagger", "kernels", "blur", "downsampling", "transform", "losses", "crop", "degrade", "model"]
    for w in which:
        globals()[f"gen_{w}"]()
