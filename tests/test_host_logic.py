"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, argument
validation fails loudly without touching a device, the host-side mirrors of the reference's
glue (kernels, crop, factories, draw order) behave like the reference (golden vectors)."""
import os
import re
from argparse import Namespace

import numpy as np
import pytest
import torch

import sei_b200
from sei_b200 import _lib, draws, ops
from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sei_b200.h")).read()
    declared = set(re.findall(r"\b(sei_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"libsei_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sei_abi_version() == 1
    assert lib.sei_reduce_workspace_bytes() > 0
    integration = open(os.path.join(ROOT, "INTEGRATION.md")).read()       # every entry point is mapped to the code it replaces
    assert not [name for name in declared if name not in integration]


def test_sass_is_blackwell_native():
    """the tiled kernels stage their tiles with the TMA engine (bulk async copies -> UBLKCP)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert out.count("UBLKCP") >= 8
    assert "SYNCS.ARRIVE.TRANS64" in out
    # tcgen05 / TMEM / TMA in the GEMMs: single-CTA UMMA, the CTA-pair kernel (cta_group::2 MMA, multicast commit, 2-CTA
    # TMA loads, cluster barrier) and the batched operator product (5-D tensor maps in both directions)
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG.2D", "UTMASTG.2D", "UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST",
                     "UTMALDG.2D.2CTA", "UCGABAR_ARV", "UTMALDG.5D", "UTMASTG.5D"):
        assert mnemonic in out, mnemonic


def test_argument_validation_needs_no_device():
    lib = _lib.load()
    k = np.ones((3, 3))
    rc = lib.sei_blur_circular_f32(None, None, 1, 8, 8, k.ctypes.data, 3, 3, 0, None, 0.0, 0, None)
    assert rc == -22 and b"null" in lib.sei_last_error()
    rc = lib.sei_down_aa_f32(1, 1, 1, 8, 8, 7, None, 0.0, 0, None)
    assert rc == -22 and b"rate" in lib.sei_last_error()
    with pytest.raises(_lib.SeiError):
        _lib.check(rc)
    # the CNN-side entry points validate before touching a device as well
    assert lib.sei_ln_cl_forward_bf16(None, None, None, None, None, None, 4, 32, 1e-6, None) == -22
    assert lib.sei_ln_cl_forward_bf16(1, 1, 1, 1, 1, 1, 4, 12, 1e-6, None) == -22 and b"multiple of 8" in lib.sei_last_error()
    assert lib.sei_ln_small_forward_bf16(1, 1, 1, 1, 1, 1, 4, 0, 1e-6, None) == -22          # (any C >= 1 is taken: 1..32 one thread per row, above one warp per row)
    assert lib.sei_dwconv7_cl_bf16(None, None, None, None, 1, 8, 8, 8, None) == -22
    assert lib.sei_dwconv7_cl_bf16(16, 16, None, 16, 1, 8, 8, 12, None) == -22
    assert lib.sei_conv3x3_small_forward_bf16(16, 16, None, 16, 1, 8, 8, 32, 5, None) == -22 and b"Cout" in lib.sei_last_error()
    assert lib.sei_gelu_bf16(16, None, 16, 12, None) == -22 and b"multiple of 8" in lib.sei_last_error()
    assert lib.sei_bgemm_bf16(16, 16, 16, 8, 8, 12, 64, 16, 1, 1, 0, 0, 8, 0, 16, 0, 0, 8, 0, 16, None) == -22
    assert lib.sei_adam_step_f32(None, None, None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, None) == -22
    assert lib.sei_gemm_bf16_tn_gelu_bwd(16, 16, 16, None, 128, 64, 64, 64, 64, 64, 64, None) == -22


def test_cpu_tensors_fail_loudly():
    x = torch.rand(1, 3, 16, 16)
    k = ops.kernel_to_host(torch.ones(1, 1, 3, 3) / 9)
    with pytest.raises(sei_b200.SeiError, match="no CPU fallback"):
        ops.blur_circular(x, k)
    with pytest.raises(sei_b200.SeiError):
        ops.mse(x, x)
    import physics
    phys = physics.get_physics(base_args(), device="cpu")
    with pytest.raises(sei_b200.SeiError):
        phys.A(x)
    if not torch.cuda.is_available():
        lib = _lib.load()
        assert lib.sei_device_info(None, None, None, None) != 0   # no device -> error code, not a fallback


def test_float64_is_rejected():
    if not torch.cuda.is_available():
        pytest.skip("needs a device to get past the device check")
    with pytest.raises(sei_b200.SeiError, match="float32"):
        ops.blur_circular(torch.rand(1, 1, 16, 16, dtype=torch.float64, device="cuda"), np.ones((3, 3)) / 9)


def test_named_kernels(golden):
    from physics.kernels import get_kernel
    g = golden("kernels")
    for name, ref in g.items():
        k = get_kernel(name)
        assert k.dtype == torch.float64
        assert rel_err(k.numpy(), ref) < 1e-15
    with pytest.raises(AssertionError):
        get_kernel("Gaussian_R7")


def test_factories_build_the_reference_objects():
    import losses
    import physics
    for kw, cls, margin in [(dict(), physics.BlurV2, 6), (dict(physics_v2=False), physics.Blur, 6),
                            (dict(kernel="Box_R3"), physics.BlurV2, 3),
                            (dict(task="sr", kernel=None, sr_factor=2), physics.Downsampling, 0),
                            (dict(task="sr", kernel=None, sr_factor=4, partial_sure_sr=True), physics.Downsampling, 2)]:
        args = base_args(**kw)
        phys = physics.get_physics(args, device="cpu")
        assert isinstance(phys, cls)
        assert phys.task == args.task
        mgr = getattr(phys, "__manager")
        assert mgr.physics is phys and mgr.task == args.task
        assert abs(float(phys.noise_model.sigma) - 5 / 255) < 1e-8
        if args.task == "deblurring":
            assert phys.filter.shape[:2] == (1, 1) and phys.filter.dtype == torch.float64
            assert not hasattr(phys, "rate")
        else:
            assert phys.rate == args.sr_factor
        loss = losses.get_loss(args, phys)
        assert isinstance(loss.loss, losses.ProposedLoss)
        sure, ei = loss.loss.loss_fns
        assert sure.margin == margin and sure.cropped_div and not sure.averaged_cst
        assert abs(sure.sigma2 - (5 / 255) ** 2) < 1e-12 and sure.tau == 1e-2
        assert ei.no_grad and ei.weight == 1.0 and ei.noise
        assert loss.crop_fn is None
    with pytest.raises(ValueError):
        physics.get_physics(base_args(task="nope"), device="cpu")
    phys = physics.get_physics(base_args(), device="cpu")
    with pytest.raises(ValueError):
        losses.get_loss(base_args(method="ei-shift"), phys)   # README names the code does not accept
    for m, cls in [("supervised", losses.SupervisedLoss), ("css", losses.CSSLoss), ("sure", losses.SURELoss),
                   ("noise2inverse", losses.Noise2InverseLoss)]:
        assert isinstance(losses.get_loss(base_args(method=m), phys).loss, cls)
    crop_loss = losses.get_loss(base_args(Loss__crop_training_pairs=True), phys)
    assert crop_loss.crop_fn is not None and crop_loss.xy_size_ratio == 1


def test_batched_degradation_checks_its_seeds():
    import physics
    mgr = getattr(physics.get_physics(base_args(), device="cpu"), "__manager")
    with pytest.raises(ValueError, match="3 images but 2 seeds"):
        mgr.randomly_degrade_batch(torch.zeros(3, 3, 8, 8), [1, 2])
    with pytest.raises(sei_b200.SeiError):                         # CPU tensors: no fallback, like every operator
        mgr.randomly_degrade_batch(torch.zeros(2, 3, 16, 16), [1, 2])


def test_crop_pair_matches_reference(golden):
    from crop import CropPair
    g = golden("crop")
    x3, y3 = torch.from_numpy(g["d3_x"]), torch.from_numpy(g["d3_y"])
    torch.manual_seed(11)
    xc, yc = CropPair(location="random", size=24)(x3, y3, xy_size_ratio=2)
    assert np.array_equal(xc.numpy(), g["d3_xc"]) and np.array_equal(yc.numpy(), g["d3_yc"])
    xc, yc = CropPair(location="center", size=24)(x3, y3)
    assert np.array_equal(xc.numpy(), g["d3c_xc"]) and np.array_equal(yc.numpy(), g["d3c_yc"])
    # batched inputs: the reference pads size - C zero rows before cropping (quirk kept)
    x4, y4 = torch.from_numpy(g["d4_x"]), torch.from_numpy(g["d4_y"])
    for seed in (0, 1, 2, 3):
        torch.manual_seed(seed)
        xc, yc = CropPair(location="random", size=48)(x4, y4, xy_size_ratio=1)
        assert np.array_equal(xc.numpy(), g[f"d4_s{seed}_xc"]) and np.array_equal(yc.numpy(), g[f"d4_s{seed}_yc"])


def test_transform_parameter_draw_order(golden):
    """rates first, then centres; same mapping from uniforms as the reference (CPU branch)"""
    import transforms
    g = golden("transform")
    torch.manual_seed(0)
    rate, center = transforms.sample_downsampling_parameters(16, "cpu", torch.float32, [0.75, 0.5])
    assert np.array_equal(rate.numpy(), g["params_rate"])
    assert np.array_equal(center.numpy(), g["params_center"])
    with draws.inject([g["params_draw0_rand"], g["params_draw1_rand"]]):
        rate2, center2 = transforms.sample_downsampling_parameters(16, "cpu", torch.float32, [0.75, 0.5])
    assert torch.equal(rate, rate2) and torch.equal(center, center2)
    grid = transforms.get_downsampling_grid((4, 3, 24, 24), torch.from_numpy(g["c0_rate"]).float(),
                                            torch.from_numpy(g["c0_center"]).float(), torch.float32, "cpu")
    assert np.array_equal(grid.numpy(), g["c0_grid_f32"])


def test_draw_injection_is_strict():
    with pytest.raises(RuntimeError, match="shape"):
        with draws.inject([np.zeros((2, 2), np.float32)]):
            draws.rand((3,), "cpu", torch.float32)
    with pytest.raises(RuntimeError, match="not consumed"):
        with draws.inject([np.zeros((2,), np.float32)]):
            pass


def test_unbuilt_variants_say_so():
    import transforms
    t = transforms.ScalingTransform(kind="normal", antialias=False)          # built: resize kernel (GPU tests)
    assert isinstance(t.transform, transforms.NormalDownsamplingTransform)
    # antialias=True is built too; mixed rates fail like the reference's torch.stack does (src/transforms.py:44-57)
    with pytest.raises(RuntimeError, match="equal size"):
        transforms.padded_downsampling_transform(torch.zeros(2, 1, 8, 8), torch.tensor([0.75, 0.5]), torch.zeros(2, 1, 1, 2),
                                                 "bicubic", "reflection", True)
    with pytest.raises(NotImplementedError):
        transforms.padded_downsampling_transform(torch.zeros(1, 1, 8, 8), torch.ones(1), torch.zeros(1, 1, 1, 2),
                                                 "nearest", "reflection", False)
    with pytest.raises(ValueError):
        transforms.ScalingTransform(kind="other", antialias=False)


def test_checkpoint_round_trip(tmp_path):
    """training.save_training_state / get_weights keep the reference's file format (src/training.py:6-45)"""
    import training
    net = torch.nn.Linear(3, 2)
    net.get_weights = net.state_dict
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, 10)
    path = tmp_path / "run" / "checkpoints" / "ckp_007.pt"
    training.save_training_state(7, net, opt, sched, str(path))
    state = torch.load(path)
    assert tuple(state) == ("epoch", "params", "optimizer", "scheduler") and state["epoch"] == 7
    assert state["optimizer"]["param_groups"][0]["lr"] == 1e-3 and "last_epoch" in state["scheduler"]
    w = training.get_weights(str(path), "cpu")                       # a training state is reduced to its params
    assert set(w) == {"weight", "bias"} and torch.equal(w["weight"], net.weight)
    bare = tmp_path / "weights.pt"
    torch.save(net.state_dict(), bare)
    assert torch.equal(training.get_weights(str(bare), "cpu")["bias"], net.bias)
    training.save_training_state(0, net, opt, sched, "ckp_in_cwd_test.pt")      # no directory part
    os.remove("ckp_in_cwd_test.pt")


def test_adam_host_logic():
    """sei_b200.optim.Adam: torch's constructor arguments and state-dict layout, no CPU fallback, unsupported options raise,
    and every optimizer step of any optimizer advances the epoch that keys the low-precision weight copies"""
    from sei_b200 import optim
    w = torch.nn.Parameter(torch.zeros(4, 3))
    with pytest.raises(NotImplementedError):
        optim.Adam([w], weight_decay=0.1)
    with pytest.raises(NotImplementedError):
        optim.Adam([w], amsgrad=True)
    opt = optim.Adam([w], lr=5e-4, betas=(0.9, 0.99))
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(4, 3))], lr=5e-4, betas=(0.9, 0.99))
    g, gr = opt.state_dict()["param_groups"][0], ref.state_dict()["param_groups"][0]
    assert (g["lr"], tuple(g["betas"]), g["eps"], g["params"]) == (gr["lr"], tuple(gr["betas"]), gr["eps"], gr["params"])
    opt.step()                                        # nothing has a gradient: a no-op like torch's
    assert not opt.state_dict()["state"]
    w.grad = torch.ones_like(w)
    with pytest.raises(sei_b200.SeiError, match="no CPU fallback"):
        opt.step()
    assert set(opt.state[w]) == {"step", "exp_avg", "exp_avg_sq"}         # torch.optim.Adam's per-parameter state
    # the cache key of bf16 weight copies: advanced by ANY optimizer's step; None only while this Adam maintains the copy
    before = optim.shadow_epoch(w)
    ref.param_groups[0]["params"][0].grad = torch.ones(4, 3)
    ref.step()
    assert optim.shadow_epoch(w) == before + 1
    w._sei_maintained = True
    assert optim.shadow_epoch(w) is None
    other = torch.optim.SGD([w], lr=0.1)
    other.step()                                      # a foreign optimizer touched w: the maintained flag is dropped
    assert optim.shadow_epoch(w) is not None


def test_igemm_weight_chunks_reproduce_the_convolution():
    """the chunk form the implicit-GEMM kernel reads ([tap * blocks + block][n][8], zero padded): contracting it with the
    shifted, zero-padded windows (what the nine TMA copies deliver) is F.conv2d(padding=1)"""
    import torch.nn.functional as F
    from sei_b200 import ops
    torch.manual_seed(0)
    for cin, cinp, cout, nout in ((3, 8, 32, 32), (32, 32, 3, 16)):
        w = torch.randn(cout, cin, 3, 3)
        x = torch.randn(2, cin, 9, 11)
        wg = ops.igemm_weight_chunks(w, cinp, nout).float()                  # bf16-rounded weights
        cch = cinp // 8
        assert wg.shape == (((9 * cch + 1) // 2) * 2, nout, 8)
        xp = F.pad(F.pad(x, (0, 0, 0, 0, 0, cinp - cin)), (1, 1, 1, 1))      # channel padding, then the 'same' halo
        out = torch.zeros(2, nout, 9, 11)
        for tap in range(9):
            ky, kx = divmod(tap, 3)
            win = xp[:, :, ky:ky + 9, kx:kx + 11]                            # the window of this tap
            for blk in range(cch):
                out += torch.einsum("bchw,nc->bnhw", win[:, 8 * blk:8 * blk + 8], wg[tap * cch + blk])
        ref = F.conv2d(x, w.bfloat16().float(), padding=1)
        assert torch.allclose(out[:, :cout], ref, atol=1e-4) and float(out[:, cout:].abs().max() if nout > cout else 0) == 0.0
        assert float(wg[9 * cch:].abs().max() if wg.shape[0] > 9 * cch else 0) == 0.0


def test_convblock_is_one_completion_group():
    """a ConvBlock runs as one autograd node, so the data-parallel reducer must treat it as one bucket however large it is
    (its children are never called as modules: hooks on them would never fire)"""
    import models.convolutional as mc
    from sei_b200 import parallel
    net = mc.ConvolutionalModel(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1,
                                hidden_channels=8, inout_convs=True, scales=3)
    groups = parallel.completion_groups(net, max_elems=64)                   # smaller than any block
    blocks = [m for m in net.modules() if isinstance(m, mc.ConvBlock)]
    assert all(any(g is b for g in groups) for b in blocks)
    inner = {id(m) for b in blocks for m in b.modules() if m is not b}
    assert not any(id(g) in inner for g in groups)
    seen = [id(p) for g in groups for p in g.parameters()]
    assert len(seen) == len(set(seen)) == len(list(net.parameters()))         # a partition of the parameters
