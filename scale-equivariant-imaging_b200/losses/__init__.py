"""Training losses (reference: src/losses/__init__.py): SupervisedLoss, CSSLoss, Noise2InverseLoss,
SURELoss, ProposedLoss, Loss and the get_loss(args, physics) factory, with the same constructor
arguments, flags and forward(x, y, model) -> scalar protocol, assembled from libsei_b200 kernels."""
from os import environ

from torch.nn import Module
from torch.nn.functional import l1_loss

from crop import CropPair
from sei_b200.linear_physics import SupLoss, EILoss, Rotate, Shift, mse
from transforms import ScalingTransform, CombinedTransform  # noqa: F401
from .r2r import R2REILoss
from .sure import SureGaussianLoss


class _ModelThenLoss(Module):
    """x_net = model(y); loss(x=, x_net=, y=, physics=, model=)."""

    def __init__(self, physics, loss):
        super().__init__()
        self.physics = physics
        self.loss = loss

    def forward(self, x, y, model):
        x_net = model(y)
        return self.loss(x=x, x_net=x_net, y=y, physics=self.physics, model=model)


class SupervisedLoss(_ModelThenLoss):
    def __init__(self, physics):
        metric = mse()
        if "SUPERVISED_L1" in environ:
            print("SUPERVISED_L1")
            metric = l1_loss
        super().__init__(physics, SupLoss(metric=metric))


class CSSLoss(_ModelThenLoss):
    def __init__(self, physics):
        super().__init__(physics, SupLoss(metric=mse()))


class Noise2InverseLoss(_ModelThenLoss):
    def __init__(self, physics):
        super().__init__(physics, SupLoss(metric=mse()))


class SURELoss(_ModelThenLoss):
    def __init__(self, noise_level, cropped_div, averaged_cst, margin, physics):
        super().__init__(physics, SureGaussianLoss(sigma=noise_level / 255, cropped_div=cropped_div,
                                                   averaged_cst=averaged_cst, margin=margin))


_EI_TRANSFORMS = {
    # ProposedLoss__transforms -> factory(blueprint)   (reference :84-96)
    "Scaling_Transforms": lambda blueprint: ScalingTransform(**blueprint[ScalingTransform.__name__]),
    "Shifts": lambda blueprint: Shift(),
    "Rotations": lambda blueprint: Rotate(),
    "Rotations+Shifts": lambda blueprint: CombinedTransform([Rotate(), Shift()]),
}


def _ei_transform(transforms, blueprint):
    if transforms not in _EI_TRANSFORMS:
        raise ValueError(f"Unknown transforms: {transforms}")
    return _EI_TRANSFORMS[transforms](blueprint)


class ProposedLoss(Module):
    """SURE + equivariant-imaging loss (reference :67-142)."""

    def __init__(self, blueprint, sure_alternative, noise_level, stop_gradient, sure_cropped_div,
                 sure_averaged_cst, sure_margin, alpha_tradeoff, transforms, physics):
        super().__init__()
        self.physics = physics
        ei_transform = _ei_transform(transforms, blueprint)
        assert sure_alternative in [None, "r2r"]
        if sure_alternative == "r2r":
            self.loss_fns = [R2REILoss(transform=ei_transform, sigma=noise_level / 255, no_grad=stop_gradient,
                                       metric=mse())]
        else:
            self.loss_fns = [
                SureGaussianLoss(sigma=noise_level / 255, cropped_div=sure_cropped_div,
                                 averaged_cst=sure_averaged_cst, margin=sure_margin),
                EILoss(metric=mse(), transform=ei_transform, no_grad=stop_gradient, weight=alpha_tradeoff),
            ]
        self.compute_x_net = sure_alternative != "r2r"

    def forward(self, x, y, model):
        x_net = model(y) if self.compute_x_net else None
        loss = 0
        for loss_fn in self.loss_fns:
            loss = loss + loss_fn(x=x, x_net=x_net, y=y, physics=self.physics, model=model)
        return loss


class Loss(Module):
    def __init__(self, physics, blueprint, noise_level, sure_cropped_div, sure_averaged_cst, sure_margin,
                 method, crop_training_pairs, crop_size):
        super().__init__()
        sure = dict(noise_level=noise_level, sure_cropped_div=sure_cropped_div, sure_averaged_cst=sure_averaged_cst,
                    sure_margin=sure_margin)
        methods = {
            "supervised": lambda: SupervisedLoss(physics=physics),
            "css": lambda: CSSLoss(physics=physics),
            "noise2inverse": lambda: Noise2InverseLoss(physics=physics),
            "sure": lambda: SURELoss(physics=physics, noise_level=noise_level, cropped_div=sure_cropped_div,
                                     averaged_cst=sure_averaged_cst, margin=sure_margin),
            "proposed": lambda: ProposedLoss(physics=physics, blueprint=blueprint, **sure, **blueprint[ProposedLoss.__name__]),
        }
        if method not in methods:
            raise ValueError(f"Unknwon method: {method}")          # (the reference's spelling)
        self.loss = methods[method]()

        self.crop_fn = None
        if crop_training_pairs:
            self.xy_size_ratio = physics.rate if hasattr(physics, "rate") else 1
            self.crop_fn = CropPair(location="random", size=crop_size)
        if "HOMOGENEOUS_SWINIR" in environ:
            self.crop_fn = None

    def forward(self, x, y, model):
        if self.crop_fn is not None:
            x, y = self.crop_fn(x, y, xy_size_ratio=self.xy_size_ratio)
        return self.loss(x=x, y=y, model=model)


def _sure_margin(args, physics):
    if not args.partial_sure:
        assert args.sure_margin is None
        return 0
    if args.sure_margin is not None:
        return args.sure_margin
    if args.task == "deblurring":
        assert physics.task == "deblurring"
        kernel = physics.filter
        return (max(kernel.shape[-2], kernel.shape[-1]) - 1) // 2
    if args.task == "sr":
        return 2 if args.partial_sure_sr else 0
    raise ValueError(f"no SURE margin rule for task {args.task}")


def get_loss(args, physics):
    blueprint = {
        Loss.__name__: dict(crop_training_pairs=args.Loss__crop_training_pairs, crop_size=args.Loss__crop_size),
        ProposedLoss.__name__: dict(stop_gradient=args.ProposedLoss__stop_gradient,
                                    sure_alternative=args.ProposedLoss__sure_alternative,
                                    alpha_tradeoff=args.ProposedLoss__alpha_tradeoff,
                                    transforms=args.ProposedLoss__transforms),
        ScalingTransform.__name__: dict(kind=args.ScalingTransform__kind, antialias=args.ScalingTransform__antialias),
    }
    return Loss(physics=physics, blueprint=blueprint, method=args.method, noise_level=args.noise_level,
                sure_cropped_div=args.sure_cropped_div, sure_averaged_cst=args.sure_averaged_cst,
                sure_margin=_sure_margin(args, physics), **blueprint[Loss.__name__])
