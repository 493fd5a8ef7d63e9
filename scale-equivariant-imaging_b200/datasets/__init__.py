"""Training / test dataset wrappers (reference: src/datasets/__init__.py) on the libsei_b200 operators.

Same classes and `__getitem__` semantics as the reference (PrepareTrainingPairs, TrainingDataset with the CSS
re-degradation and the SR `_HOTFIX` 48-pixel crop, TestDataset, Dataset, get_dataset), plus the batched device path the
reference lacks: `TrainingDataset.get_batch(indices)` produces the stacked items of a whole batch with one launch per
stage -- measurement, CSS re-measurement (src/datasets/__init__.py:70-76), paired random crop (sei_crop_batch_f32) --
consuming the random generators exactly like the per-item loop of a DataLoader with num_workers=0 (demo/train.py:127):
per item, the CSS noise from the device generator and two crop offsets from the CPU generator, in item order.
Image-file ground-truth datasets (download + PNG decoding) are out of scope: pass `ground_truth_dataset`."""
from os import environ

import torch
from torch.nn import Module
from torch.utils.data import Dataset as BaseDataset

from crop import CropPair, MinSizePadding
from sei_b200 import draws, ops
from .synthetic_dataset import SyntheticDataset


class PrepareTrainingPairs(Module):
    def __init__(self, physics, crop_size, crop_location):
        super().__init__()
        self.physics = physics
        self.crop_size = 48 if "HOMOGENEOUS_SWINIR" in environ else crop_size
        self.crop_location = crop_location

    def xy_size_ratio(self):
        if self.physics.task == "sr" and "HOMOGENEOUS_SWINIR" not in environ:
            return self.physics.rate
        return 1

    def forward(self, x, y):
        return CropPair(location=self.crop_location, size=self.crop_size)(x, y, xy_size_ratio=self.xy_size_ratio())


class TrainingDataset(BaseDataset):
    def __init__(self, synthetic_dataset, physics, css, noise2inverse, prepare_training_pairs, _HOTFIX):
        super().__init__()
        self.synthetic_dataset = synthetic_dataset
        self.physics = physics
        self.css = css
        self.noise2inverse = noise2inverse
        self.prepare_training_pairs = prepare_training_pairs
        self.important_unnamed_flag = _HOTFIX

    def _crop_plan(self):
        """(location, size, ratio) of the crop the reference applies to an item (src/datasets/__init__.py:78-90)"""
        if self.important_unnamed_flag and "HOMOGENEOUS_SWINIR" not in environ:
            return "random", 48, self.physics.rate
        ptp = self.prepare_training_pairs
        return ptp.crop_location, ptp.crop_size, ptp.xy_size_ratio()

    def __getitem__(self, index):
        x, y = self.synthetic_dataset[index]
        if self.css:
            z = getattr(self.physics, "__manager").randomly_degrade(y.unsqueeze(0), seed=None).squeeze(0)
            x, y = y, z
        location, size, ratio = self._crop_plan()
        return CropPair(location=location, size=size)(x, y, xy_size_ratio=ratio)

    def get_batch(self, indices):
        """The items `indices` stacked into (x, y) batches -- values identical to [self[i] for i in indices] -- with one
        launch per stage on the device."""
        x, y = self.synthetic_dataset.get_batch(indices)
        if self.css:
            # CSS: the measurement becomes the target and is measured again with unseeded noise (:70-76).  Per item
            # the reference draws randn((1, C, h, w)) from the device generator; so does randomly_degrade_batch(None)
            z = getattr(self.physics, "__manager").randomly_degrade_batch(y, None)
            x, y = y, z
        location, size, ratio = self._crop_plan()
        r = int(ratio)
        # MinSizePadding on a (C, H, W) item pads rows / columns up to the crop size (src/crop.py:42-57)
        x = _pad_min(x, size * r)
        y = _pad_min(y, size)
        h, w = y.shape[-2:]
        tops, lefts = [], []
        for _ in indices:                                   # two CPU-generator draws per item, rows first (:26-27)
            if location == "random":
                tops.append(draws.randint(0, h - size + 1))
                lefts.append(draws.randint(0, w - size + 1))
            else:
                tops.append((h - size) // 2)
                lefts.append((w - size) // 2)
        ty, ly = torch.tensor(tops, dtype=torch.int32), torch.tensor(lefts, dtype=torch.int32)
        xc = ops.crop_batch(x, ty * r, ly * r, size * r, size * r)
        yc = ops.crop_batch(y, ty, ly, size, size)
        return xc, yc

    def __len__(self):
        return len(self.synthetic_dataset)


def _pad_min(t, size):
    """batched MinSizePadding of (B, C, H, W): zero rows / columns so that H, W >= size (per-item semantics)"""
    ph, pw = max(0, size - t.shape[-2]), max(0, size - t.shape[-1])
    return t if ph == 0 and pw == 0 else torch.nn.functional.pad(t, (0, pw, 0, ph))


class TestDataset(BaseDataset):
    def __init__(self, synthetic_dataset, noise2inverse, physics):
        super().__init__()
        self.synthetic_dataset = synthetic_dataset
        self.noise2inverse = noise2inverse
        self.physics = physics

    def __getitem__(self, index):
        x, y = self.synthetic_dataset[index]
        if self.noise2inverse and self.physics.task == "deblurring":
            w, h = 2 * (y.shape[1] // 2), 2 * (y.shape[2] // 2)      # (the reference's names; rows first)
            y = y[:, :w, :h]
        if x.shape != y.shape:
            h, w = y.shape[1], y.shape[2]
            f = self.physics.rate if self.physics.task == "sr" else 1
            x = x[:, : h * f, : w * f]
        return x, y

    def __len__(self):
        return len(self.synthetic_dataset)


class Dataset(BaseDataset):
    def __init__(self, blueprint, purpose, physics, css, noise2inverse, device, _HOTFIX):
        super().__init__()
        synthetic_dataset = SyntheticDataset(blueprint=blueprint, device=device, physics=physics,
                                             **blueprint[SyntheticDataset.__name__])
        if purpose == "train":
            ptp = PrepareTrainingPairs(physics=physics, **blueprint[PrepareTrainingPairs.__name__])
            self.dataset = TrainingDataset(synthetic_dataset=synthetic_dataset, physics=physics, css=css,
                                           noise2inverse=noise2inverse, prepare_training_pairs=ptp, _HOTFIX=_HOTFIX)
        elif purpose == "test":
            self.dataset = TestDataset(synthetic_dataset=synthetic_dataset, noise2inverse=noise2inverse, physics=physics)
        else:
            raise ValueError(f"Unknown purpose: {purpose}")

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        return self.dataset[index]

    def get_batch(self, indices):
        return self.dataset.get_batch(indices)


def get_dataset(args, purpose, physics, device, _HOTFIX):
    if purpose == "test":
        noise2inverse, css = args.noise2inverse, False
    elif purpose == "train":
        noise2inverse, css = args.method == "noise2inverse", args.method == "css"
    else:
        raise ValueError(f"Unknown purpose: {purpose}")
    blueprint = {
        "ground_truth_dataset": getattr(args, "ground_truth_dataset", None),
        PrepareTrainingPairs.__name__: dict(crop_size=args.PrepareTrainingPairs__crop_size,
                                            crop_location=args.PrepareTrainingPairs__crop_location),
        SyntheticDataset.__name__: dict(unique_seeds=args.SyntheticDataset__unique_seeds,
                                        deterministic_measurements=args.SyntheticDataset__deterministic_measurements),
    }
    return Dataset(blueprint=blueprint, device=device, physics=physics, purpose=purpose, css=css,
                   noise2inverse=noise2inverse, _HOTFIX=_HOTFIX)
