"""Model factory (reference: src/models/__init__.py): get_model(args, physics, device) -> Model whose
forward(x, *args) ignores extra arguments, with get_weights / load_weights on the backbone.

Built here: kind "Proposed" with architecture "Convolutional" (the restoration CNN, convolutions on the
tcgen05 GEMM) and kind "Identity".  The transformer backbone (deepinv SwinIR) and the evaluation baselines
(PnP, DIP, BM3D, DiffPIR, DPS, TV, Upsample) live in third-party packages and are outside the hot path
(SURVEY.md section 2, rows 14-15)."""
from torch.nn import Module

from .convolutional import ConvolutionalModel


class Identity(Module):
    def forward(self, y):
        return y


class ProposedModel(Module):
    def __init__(self, blueprint, architecture, sampling_rate):
        super().__init__()
        if architecture == "Convolutional":
            self.model = ConvolutionalModel(in_channels=3, upsampling_rate=sampling_rate,
                                            **blueprint[ConvolutionalModel.__name__])
        elif architecture == "Transformer":
            raise NotImplementedError("ProposedModel__architecture=Transformer (deepinv SwinIR) is out of scope; "
                                      "pass --ProposedModel__architecture Convolutional")
        else:
            raise ValueError(f"Unknown model kind: {architecture}")

    def forward(self, y):
        return self.model(y)

    def get_backbone(self):
        return self.model


class Model(Module):
    def __init__(self, blueprint, kind, physics, task, sr_factor, device, noise_level, data_parallel_devices):
        super().__init__()
        sampling_rate = sr_factor if task == "sr" else 1
        if kind == "Proposed":
            self.model = ProposedModel(blueprint=blueprint, sampling_rate=sampling_rate,
                                       **blueprint[ProposedModel.__name__])
        elif kind == "Identity":
            self.model = Identity()
        else:
            raise NotImplementedError(f"model kind {kind} is an evaluation baseline outside the scope of this package")
        if data_parallel_devices is not None:
            raise NotImplementedError("--data_parallel_devices (torch DataParallel) is replaced by one process per GPU "
                                      "with an NCCL gradient all-reduce; launch bench.py / the trainer under torchrun")

    def forward(self, x, *args):
        return self.model(x)

    def get_backbone(self):
        return self.model.get_backbone() if isinstance(self.model, ProposedModel) else self.model

    def get_weights(self):
        return self.get_backbone().state_dict()

    def load_weights(self, state_dict):
        self.get_backbone().load_state_dict(state_dict)


def get_model(args, physics, device):
    blueprint = {
        ConvolutionalModel.__name__: dict(residual=args.ConvolutionalModel__residual,
                                          inner_residual=args.ConvolutionalModel__inner_residual,
                                          num_conv_blocks=args.ConvolutionalModel__num_conv_blocks,
                                          inout_convs=args.ConvolutionalModel__inout_convs,
                                          hidden_channels=args.ConvolutionalModel__hidden_channels,
                                          scales=args.ConvolutionalModel__scales),
        Model.__name__: dict(task=args.task, sr_factor=args.sr_factor, noise_level=args.noise_level,
                             kind=args.model_kind),
        ProposedModel.__name__: dict(architecture=args.ProposedModel__architecture),
    }
    dp = args.data_parallel_devices.split(",") if getattr(args, "data_parallel_devices", None) is not None else None
    return Model(blueprint=blueprint, physics=physics, device=device, data_parallel_devices=dp,
                 **blueprint[Model.__name__])
