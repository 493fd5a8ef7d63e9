"""The CNN mirror (scale-equivariant-imaging_b200/models) against the reference's ConvolutionalModel
(golden fixtures from tests/golden/make_golden.py gen_model): identical parameter tree, and -- with the
tensor-core GEMM replaced by an fp32 matmul for this test only -- identical outputs and gradients at fp32
accuracy on the CPU.  (The bf16 tcgen05 path itself is checked on the GPU in test_gpu_model.py.)"""
import numpy as np
import pytest
import torch

from util import rel_err


def _load(golden, name):
    import models.convolutional as mc
    g = golden(f"model_{name}")
    model = mc.ConvolutionalModel(**eval(str(g["kwargs"])))
    model.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")})
    return g, model


def test_default_parameter_tree_matches_reference(golden):
    import models.convolutional as mc
    g = golden("model_default_layout")
    with torch.device("meta"):
        m = mc.ConvolutionalModel(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True,
                                  num_conv_blocks=1, hidden_channels=32, inout_convs=True, scales=5)
    names = sorted(m.state_dict().keys())
    assert names == list(g["names"])
    assert [repr(list(m.state_dict()[n].shape)) for n in names] == list(g["shapes"])
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 645063043


@pytest.mark.parametrize("name", ["deblur", "sr2", "pad"])
def test_network_structure_fp32(golden, name, monkeypatch):
    import models.convolutional as mc
    g, model = _load(golden, name)
    import torch_formulation
    torch_formulation.install(mc, monkeypatch.setattr)      # fp32 library formulation of the operator hooks: tests only
    y = torch.from_numpy(g["y"])
    out = model(y)
    assert rel_err(out.detach().numpy(), g["out"]) < 2e-5
    (out * torch.from_numpy(g["gout"])).sum().backward()
    for k, p in model.named_parameters():
        ref = g[f"grad::{k}"]
        assert np.allclose(p.grad.numpy(), ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max() + 1e-7), k


def test_cpu_forward_fails_loudly(golden):
    import sei_b200
    _, model = _load(golden, "deblur")
    with pytest.raises(sei_b200.SeiError):
        model(torch.rand(1, 3, 32, 32))


def test_model_factory():
    from argparse import Namespace
    import models
    args = Namespace(task="deblurring", sr_factor=None, noise_level=5, model_kind="Proposed",
                     ProposedModel__architecture="Convolutional", ConvolutionalModel__residual=True,
                     ConvolutionalModel__inner_residual=True, ConvolutionalModel__inout_convs=True,
                     ConvolutionalModel__hidden_channels=4, ConvolutionalModel__scales=2,
                     ConvolutionalModel__num_conv_blocks=1, data_parallel_devices=None)
    m = models.get_model(args, physics=None, device="cpu")
    sd = m.get_weights()
    assert "seq.0.in_conv.weight" in sd
    m.load_weights(sd)
    assert m.get_backbone() is m.model.model
    args.ProposedModel__architecture = "Transformer"
    with pytest.raises(NotImplementedError):
        models.get_model(args, physics=None, device="cpu")
