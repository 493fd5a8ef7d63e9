"""The dataset wrappers (scale-equivariant-imaging_b200/datasets, reference src/datasets/__init__.py and
synthetic_dataset.py): per-item semantics against the reference's own classes run on the CPU (structure of the returned
pairs), and the batched device path against the per-item path (identical values: same seeds, same draws)."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from test_gpu_parity import base_args


class GT:
    """a ground-truth dataset: indexable of (C, H, W) tensors with unique ids (reference GroundTruthDataset protocol)"""

    def __init__(self, n, shape, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.items = [torch.rand(shape, generator=g) for _ in range(n)]

    def __getitem__(self, i):
        return self.items[i]

    def __len__(self):
        return len(self.items)

    def get_unique_id(self, i):
        return 1000 + 7 * i


def _dataset_args(method, gt, crop_size=32, location="random"):
    return Namespace(method=method, noise2inverse=False, ground_truth_dataset=gt,
                     PrepareTrainingPairs__crop_size=crop_size, PrepareTrainingPairs__crop_location=location,
                     SyntheticDataset__unique_seeds=True, SyntheticDataset__deterministic_measurements=True)


def _check_batch_equals_items(device, task_kw, method, hotfix, shape):
    import datasets
    import physics
    phys = physics.get_physics(base_args(**task_kw), device=device)
    gt = GT(5, shape)
    ds = datasets.get_dataset(_dataset_args(method, gt), "train", phys, device, hotfix)
    idx = [3, 0, 4, 1]
    torch.manual_seed(21)
    items = [ds[i] for i in idx]
    torch.manual_seed(21)
    xb, yb = ds.get_batch(idx)
    xs, ys = torch.stack([a for a, _ in items]), torch.stack([b for _, b in items])
    assert xb.shape == xs.shape and yb.shape == ys.shape
    return float((xb - xs).abs().max()), float((yb - ys).abs().max()), xb, yb


def test_batched_path_equals_items_cpu(monkeypatch):
    """host logic on the CPU stand-ins: seeds, draw order of the crops, MinSizePadding, SR crop ratio (no CSS here: on the
    CPU the CSS noise and the crop offsets share ONE generator, so the batched order differs by construction)"""
    import fake_ops
    fake_ops.install(monkeypatch)
    for task_kw, hotfix, shape in ((dict(), False, (3, 40, 56)), (dict(task="sr", kernel=None, sr_factor=2), True, (3, 112, 128))):
        ex, ey, xb, yb = _check_batch_equals_items("cpu", task_kw, "proposed", hotfix, shape)
        assert ex == 0.0 and ey < 1e-6
    assert tuple(xb.shape[-2:]) == (96, 96) and tuple(yb.shape[-2:]) == (48, 48)      # _HOTFIX: 48-pixel SR crops


def test_test_dataset_and_factory(monkeypatch):
    import datasets
    import fake_ops
    import physics
    fake_ops.install(monkeypatch)
    phys = physics.get_physics(base_args(task="sr", kernel=None, sr_factor=2), device="cpu")
    ds = datasets.get_dataset(_dataset_args("proposed", GT(2, (3, 33, 41))), "test", phys, "cpu", False)
    x, y = ds[1]
    assert tuple(y.shape) == (3, 16, 20) and tuple(x.shape) == (3, 32, 40)            # x cropped to a multiple of y
    with pytest.raises(ValueError):
        datasets.get_dataset(_dataset_args("proposed", GT(1, (3, 8, 8))), "validate", phys, "cpu", False)
    with pytest.raises(NotImplementedError):
        datasets.get_dataset(_dataset_args("proposed", None), "train", phys, "cpu", False)


@pytest.mark.gpu
@pytest.mark.parametrize("task_kw,hotfix,shape", [(dict(), False, (3, 64, 80)), (dict(physics_v2=False), False, (3, 40, 40)),
                                                  (dict(task="sr", kernel=None, sr_factor=2), True, (3, 128, 160)),
                                                  (dict(task="sr", kernel=None, sr_factor=4), True, (3, 256, 224))])
@pytest.mark.parametrize("method", ["proposed", "css"])
def test_batched_device_path_equals_items(task_kw, hotfix, shape, method):
    """measurement, CSS re-measurement and paired random crops of a batch, one launch each, equal the per-item pipeline of
    the reference's DataLoader loop (device and CPU generators consumed in the same order)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sei_b200 import launch_count
    n0 = launch_count()
    ex, ey, xb, yb = _check_batch_equals_items(torch.device("cuda:0"), task_kw, method, hotfix, shape)
    assert ex < 1e-6 and ey < 1e-6, (ex, ey)
    assert launch_count() - n0 >= 4 and xb.is_cuda
