"""Per-epoch seed shuffling (fast_trainer/shufflers.py:6-45,92-100): which seeds each rank
samples.

Default (``device=None``): host logic with the reference's exact arithmetic -- a CPU generator
seeded with ``initial_seed + epoch`` and ``torch.randperm`` -- so an epoch visits the seeds in
the same order as the reference.

``device="cuda"``: the epoch's permutation is drawn ON the GPU (a device generator with the same
seed rule) and the shuffled seed list stays in HBM; a Session given a device-resident ``idx``
reads each batch's seeds in place (no per-batch H2D copy, ``fast_sampler.Session``), so nothing
of the epoch set-up touches the host.  The permutation then differs from the reference's (a
sequential Fisher-Yates over mt19937 cannot be reproduced in parallel); it is still a function of
``(initial_seed, epoch)`` only, identical on every rank, which is what the slicing below needs.
"""
from __future__ import annotations

from typing import Optional

import torch


class Shuffler:
    DEFAULT_INITIAL_SEED = 2147483647

    def __init__(self, idx: torch.Tensor, initial_seed: int = DEFAULT_INITIAL_SEED, device: Optional[str] = None):
        assert idx.dim() == 1
        self.device = torch.device(device) if device is not None else None
        on_gpu = self.device is not None and self.device.type == "cuda"
        self.initial_idx = idx.to(self.device) if on_gpu else idx
        self.initial_seed = initial_seed
        self.generator = torch.Generator(device=self.device if on_gpu else "cpu")
        self._on_gpu = on_gpu
        self.set_epoch(0)

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def get_idx(self):
        self.generator.manual_seed(self.initial_seed + self.epoch)
        if self._on_gpu:
            perm = torch.randperm(self.initial_idx.numel(), generator=self.generator, device=self.device)
            return self.initial_idx[perm]
        perm = torch.randperm(self.initial_idx.numel(), generator=self.generator)
        return self.initial_idx[perm.to(self.initial_idx.device)]


class DistributedShuffler(Shuffler):
    """Rank r takes ``[n*r/W, n*(r+1)/W)`` of a permutation common to all ranks (:32-45)."""

    def __init__(self, idx, world_size, initial_seed=Shuffler.DEFAULT_INITIAL_SEED, device: Optional[str] = None):
        super().__init__(idx, initial_seed, device)
        self.world_size = world_size

    def get_idx(self, rank):
        shuffled = super().get_idx()
        n = shuffled.numel()
        return shuffled[(n * rank) // self.world_size:(n * (rank + 1)) // self.world_size]


class FederatedDistributedShuffler(Shuffler):
    """Each rank shuffles its own partition's seeds (:92-100)."""
    pass
