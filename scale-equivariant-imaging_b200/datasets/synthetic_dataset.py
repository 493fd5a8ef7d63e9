"""Seeded synthetic measurements of a ground-truth dataset (reference: src/datasets/synthetic_dataset.py).

`SyntheticDataset[i]` is the reference's per-item path (one batch-of-1 physics call).  `get_batch(indices)` is the
B200 path: the same items -- same seeds, same draws, same values -- with ONE operator launch for the whole batch
(PhysicsManager.randomly_degrade_batch)."""
import torch
from torch.utils.data import Dataset


class SyntheticDataset(Dataset):
    def __init__(self, blueprint, device, deterministic_measurements, unique_seeds, physics, ground_truth_dataset=None):
        super().__init__()
        self.device = device
        self.deterministic_measurements = deterministic_measurements
        self.unique_seeds = unique_seeds
        self.physics_manager = getattr(physics, "__manager")
        if ground_truth_dataset is None:
            ground_truth_dataset = (blueprint or {}).get("ground_truth_dataset")
        if ground_truth_dataset is None:
            # DIV2K / Urban100 / FMD download + PNG decoding (reference src/datasets/ground_truth.py) is outside the
            # hot path: pass any indexable of (C, H, W) tensors
            raise NotImplementedError("pass ground_truth_dataset=<indexable of (C, H, W) float tensors>; the image-file "
                                      "datasets of the reference are out of scope")
        self.ground_truth_dataset = ground_truth_dataset

    def _seed(self, index):
        if not self.deterministic_measurements:
            return None
        if not self.unique_seeds:
            return 0
        gt = self.ground_truth_dataset
        return gt.get_unique_id(index) if hasattr(gt, "get_unique_id") else index

    def __getitem__(self, index):
        x = self.ground_truth_dataset[index].to(self.device)
        y = self.physics_manager.randomly_degrade(x.unsqueeze(0), seed=self._seed(index)).squeeze(0)
        return x, y

    def get_batch(self, indices):
        """(x, y) of `indices` stacked, one operator launch; identical to stacking the per-item results (the images
        must have one size)"""
        x = torch.stack([self.ground_truth_dataset[i].to(self.device) for i in indices])
        y = self.physics_manager.randomly_degrade_batch(x, [self._seed(i) for i in indices])
        return x, y

    def __len__(self):
        return len(self.ground_truth_dataset)
