"""Per-epoch seed shuffling (fast_trainer/shufflers.py:6-45,92-100): which seeds each rank
samples.  Host logic, identical arithmetic to the reference."""
from __future__ import annotations

import torch


class Shuffler:
    DEFAULT_INITIAL_SEED = 2147483647

    def __init__(self, idx: torch.Tensor, initial_seed: int = DEFAULT_INITIAL_SEED):
        assert idx.dim() == 1
        self.initial_idx = idx
        self.initial_seed = initial_seed
        self.generator = torch.Generator(device="cpu")
        self.set_epoch(0)

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def get_idx(self):
        self.generator.manual_seed(self.initial_seed + self.epoch)
        perm = torch.randperm(self.initial_idx.numel(), generator=self.generator)
        return self.initial_idx[perm.to(self.initial_idx.device)]


class DistributedShuffler(Shuffler):
    """Rank r takes ``[n*r/W, n*(r+1)/W)`` of a permutation common to all ranks (:32-45)."""

    def __init__(self, idx, world_size, initial_seed=Shuffler.DEFAULT_INITIAL_SEED):
        super().__init__(idx, initial_seed)
        self.world_size = world_size

    def get_idx(self, rank):
        shuffled = super().get_idx()
        n = shuffled.numel()
        return shuffled[(n * rank) // self.world_size:(n * (rank + 1)) // self.world_size]


class FederatedDistributedShuffler(Shuffler):
    """Each rank shuffles its own partition's seeds (:92-100)."""
    pass
