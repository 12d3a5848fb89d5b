"""Host-side mirror of ``fast_trainer/transferers.py``: the ``DeviceIterator`` protocol
(:20-29) and its four implementations, yielding ``[PreparedBatch]`` on the GPU.

``DeviceDistributedPrefetcher`` replaces the reference's 10-stage pipeline with three NCCL
all_to_alls per batch (:33-887): the Session already produced ``x`` in MFG order with the fused
partition-book + cache + peer-to-peer gather kernel, so the iterator only hands batches over
with the reference's ``record_stream`` / one-batch-ahead semantics.  ``NcclAllToAllPrefetcher``
keeps the reference's request/response exchange (counts -> ids -> rows, then cat + permute) as
the comparison path.
"""
from __future__ import annotations

import ctypes
from typing import Iterator, List

import torch

from . import _lib, fast_sampler
from ._lib import check
from .samplers import OwnedPreparedBatch, PreparedBatch, ProtoBatch, ProtoDistributedBatch


def _current_stream(device) -> torch.cuda.Stream:
    d = device if isinstance(device, torch.device) else torch.device(device)
    idx = d.index if d.index is not None else torch.cuda.current_device()
    return fast_sampler.current_stream_cached(idx)


class DeviceIterator(Iterator[List[PreparedBatch]]):
    """Abstract class that returns PreparedBatch on devices (fast_trainer/transferers.py:20-29)."""

    def __init__(self, devices):
        assert len(devices) > 0
        self.devices = devices
        self.device = self.devices[0]

    def print_stats(self):
        return


class DevicePrefetcher(DeviceIterator):
    """fast_trainer/transferers.py:890-970.  Batches are born in HBM, so "prefetch" reduces to
    holding the next batch (already enqueued by the Session on its side streams) and making the
    consumer stream the owner of its memory via ``record_stream``."""

    def __init__(self, devices, it: Iterator[PreparedBatch], pipeline_on=True):
        super().__init__(devices)
        self.it = it
        self.next: List[PreparedBatch] = []
        self.preload()

    def preload(self, timing=True):
        self.next = []
        for device in self.devices:
            batch = next(self.it, None)
            if batch is None:
                break
            self.next.append(batch.to(device, non_blocking=True))

    def __next__(self):
        ret = self.next
        if not ret:
            raise StopIteration
        for device, batch in zip(self.devices, ret):
            batch.record_stream(_current_stream(device))
        self.preload()
        return ret


class DeviceTransferer(DeviceIterator):
    """fast_trainer/transferers.py:973-985"""

    def __init__(self, devices, it: Iterator[PreparedBatch], pipeline_on=True):
        super().__init__(devices)
        self.it = it

    def __next__(self):
        ret = [batch.to(device, non_blocking=True) for device, batch in zip(self.devices, self.it)]
        if len(ret) == 0:
            raise StopIteration
        return ret


class DeviceSlicerTransferer(DeviceIterator):
    """fast_trainer/transferers.py:988-1009: slice x / y by a ProtoBatch's n_id on the device."""

    def __init__(self, devices, x: torch.Tensor, y: torch.Tensor, it: Iterator[ProtoBatch]):
        super().__init__(devices)
        self.x, self.y, self.it = x, y, it

    def __next__(self):
        ret = [PreparedBatch.from_proto_batch(self.x, self.y, pb).to(device, non_blocking=True)
               for device, pb in zip(self.devices, self.it)]
        if len(ret) == 0:
            raise StopIteration
        return ret


def _feature_bytes(batch: ProtoDistributedBatch, rank: int, row_bytes: int) -> int:
    return sum(p.numel() for r, p in enumerate(batch.partition_nids) if r != rank) * row_bytes


class DeviceDistributedPrefetcher(DeviceIterator):
    """Drop-in for fast_trainer/transferers.py:33-887 (constructor ``(devices, it, pipeline_on)``,
    attributes ``it``, ``NUMBER_OF_SENT_BYTES``, ``print_stats``)."""

    def __init__(self, devices, it: Iterator[ProtoDistributedBatch], pipeline_on=True):
        super().__init__(devices)
        self.it = it
        self.pipeline_on = pipeline_on
        cfg = it.session.config
        self.partition_book = cfg.partition_book
        self.cache = cfg.cache
        self.use_cache = cfg.use_cache
        self.rank = int(cfg.partition_book.rank)
        self.NUMBER_OF_SENT_BYTES = 0  # bytes this rank pulled from peers over NVLink
        self.next: List[PreparedBatch] = []
        self.preload()

    def _prepare(self, batch: ProtoDistributedBatch) -> PreparedBatch:
        if batch.x is None:
            raise _lib.SalientB200Error(
                "peer feature tables are not reachable from this process (no process group matching the "
                "partition count and no Config.partition_tables); use NcclAllToAllPrefetcher for the "
                "all_to_all comparison path")
        self.NUMBER_OF_SENT_BYTES += _feature_bytes(batch, self.rank, batch.x.size(1) * batch.x.element_size())
        y = batch.y_flat
        if y is None and batch.sliced_cpu_labels is not None:
            y = batch.sliced_cpu_labels.squeeze()
        if batch.owners:
            b = OwnedPreparedBatch(batch.x, y, batch.adjs, batch.idx_range)
            b.owners = batch.owners
            return b
        return PreparedBatch(batch.x, y, batch.adjs, batch.idx_range)

    def preload(self, timing=True):
        batch = next(self.it, None)
        self.next = [] if batch is None else [self._prepare(batch)]

    def __next__(self):
        ret = self.next
        if not ret:
            raise StopIteration
        ret[0].record_stream(_current_stream(self.device))
        self.preload()
        return ret


class NcclAllToAllPrefetcher(DeviceIterator):
    """Comparison path C1: the reference's request/response protocol
    (fast_trainer/transferers.py:507-766) with everything resident in HBM:
      1. all_to_all of the per-owner request counts        (:751-766, then the D2H at :725-748)
      2. all_to_all of the requested global ids            (:690-722)
      3. owners gather the requested rows                  (:633-687)  -- spp_gather_rows
      4. all_to_all of the feature rows                    (:507-531)
      5. cat(partitions..., cached)[perm]                  (:462-505)  -- spp_gather_rows by perm
    """

    def __init__(self, devices, it: Iterator[ProtoDistributedBatch], pipeline_on=True, group=None):
        import torch.distributed as dist
        super().__init__(devices)
        self.it = it
        self.group = group
        self.dist = dist
        cfg = it.session.config
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.session = it.session
        self.off = [int(v) for v in cfg.partition_book.partition_offsets.tolist()]
        self.x_local = it.session._x_local
        self.cache_feats = it.session._cache_feats if cfg.use_cache else None
        self.NUMBER_OF_SENT_BYTES = 0
        self.next: List[PreparedBatch] = []
        self.preload()

    def _exchange(self, b: ProtoDistributedBatch) -> PreparedBatch:
        dist, W, r = self.dist, self.world, self.rank
        dev = self.device
        F = self.x_local.size(1)
        send_counts = torch.tensor([0 if p == r else b.partition_nids[p].numel() for p in range(W)],
                                   dtype=torch.int64, device=dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()          # host sync, like the reference
        send_ids = torch.cat([b.partition_nids[p] for p in range(W) if p != r]) if W > 1 else \
            torch.empty(0, dtype=torch.int64, device=dev)
        recv_ids = torch.empty(sum(rc), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv_ids, send_ids, rc, sc, group=self.group)
        rows_out = fast_sampler.serial_index(self.x_local, recv_ids - self.off[r])
        rows_in = torch.empty((sum(sc), F), dtype=self.x_local.dtype, device=dev)
        dist.all_to_all_single(rows_in, rows_out, sc, rc, group=self.group)
        own = fast_sampler.serial_index(self.x_local, b.partition_nids[r] - self.off[r])
        parts, pos = [], 0
        for p in range(W):
            if p == r:
                parts.append(own)
            else:
                parts.append(rows_in[pos:pos + sc[p]])
                pos += sc[p]
        if self.cache_feats is not None:
            parts.append(fast_sampler.serial_index(self.cache_feats, b.cached_nids))
        x = fast_sampler.serial_index(torch.cat(parts, dim=0), b.perm_partition_to_mfg)
        self.NUMBER_OF_SENT_BYTES += sum(sc) * (8 + F * x.element_size())
        y = b.sliced_cpu_labels
        return PreparedBatch(x, y.squeeze() if y is not None else None, b.adjs, b.idx_range)

    def preload(self, timing=True):
        batch = next(self.it, None)
        self.next = [] if batch is None else [self._exchange(batch)]

    def __next__(self):
        ret = self.next
        if not ret:
            raise StopIteration
        ret[0].record_stream(_current_stream(self.device))
        self.preload()
        return ret
