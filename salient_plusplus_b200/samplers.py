"""Host-side mirror of ``fast_trainer/samplers.py`` on top of the GPU ``fast_sampler`` module:
``FastSamplerConfig`` (:271-305), ``FastSampler`` (:372-399), ``FastSamplerIter`` (:331-357),
``ProtoDistributedBatch`` (:32-164), ``PreparedBatch`` (:213-268), ``ProtoBatch`` (:197-210),
``FastSamplerStats`` (:308-328).  Same names, fields and iterator protocol, so the reference's
``driver/`` and ``fast_trainer/train.py`` run on it unchanged."""
from __future__ import annotations

import datetime
from dataclasses import dataclass, fields
from typing import Iterable, Iterator, List, NamedTuple, Optional, Sized

import torch

from . import fast_sampler
from .adj import Adj, Adj__from_fast_sampler
from .fast_sampler import Cache, RangePartitionBook


class ProtoDistributedBatch(NamedTuple):
    partition_nids: List[torch.Tensor]
    sliced_cpu_features: torch.Tensor
    sliced_cpu_labels: torch.Tensor
    cached_nids: torch.Tensor
    perm_partition_to_mfg: torch.Tensor
    adjs: List[Adj]
    idx_range: slice
    # extensions: MFG node list and features already gathered in MFG order by the fused kernel
    n_id: Optional[torch.Tensor] = None
    x: Optional[torch.Tensor] = None
    # the distinct allocations every tensor above is a view of (record_stream on those is enough)
    owners: tuple = ()

    @classmethod
    def from_fast_sampler(cls, batch):
        assert batch.sliced_cpu_features is not None
        (start, stop) = batch.idx_range
        return cls(partition_nids=batch.partition_nids, sliced_cpu_features=batch.sliced_cpu_features,
                   sliced_cpu_labels=batch.sliced_cpu_labels, cached_nids=batch.cached_nids,
                   perm_partition_to_mfg=batch.perm_partition_to_mfg,
                   adjs=[Adj__from_fast_sampler(adj) for adj in batch.adjs], idx_range=slice(start, stop),
                   n_id=getattr(batch, "n_id", None), x=getattr(batch, "x", None),
                   owners=getattr(batch, "owners", ()))

    def record_stream(self, stream):
        if self.owners:
            for t in self.owners:
                t.record_stream(stream)
            return
        for part in self.partition_nids:
            if part.is_cuda:
                part.record_stream(stream)
        if self.perm_partition_to_mfg.is_cuda:
            self.perm_partition_to_mfg.record_stream(stream)
        for adj in self.adjs:
            adj.record_stream(stream)
        for t in (self.n_id, self.x):
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def to(self, device, stream=None, non_blocking=False, streams_to_sync=None, delay_feature_transfer=True):
        with torch.cuda.stream(stream):
            mv = lambda t: t.to(device, non_blocking=non_blocking)
            return self._replace(adjs=[adj.to(device, non_blocking=non_blocking) for adj in self.adjs],
                                 partition_nids=[mv(p) for p in self.partition_nids],
                                 perm_partition_to_mfg=mv(self.perm_partition_to_mfg),
                                 sliced_cpu_features=(self.sliced_cpu_features if delay_feature_transfer
                                                      else mv(self.sliced_cpu_features)))

    @property
    def num_total_nodes(self):
        return self.perm_partition_to_mfg.size(0)

    @property
    def num_cached_nodes(self):
        return self.cached_nids.size(0)


class ProtoBatch(NamedTuple):
    n_id: torch.Tensor
    adjs: List[Adj]
    idx_range: slice

    @classmethod
    def from_fast_sampler(cls, proto_sample):
        n_id, adjs, (start, stop) = proto_sample
        return cls(n_id=n_id, adjs=[Adj__from_fast_sampler(adj) for adj in adjs], idx_range=slice(start, stop))

    @property
    def batch_size(self):
        return self.idx_range.stop - self.idx_range.start


class PreparedBatch(NamedTuple):
    x: torch.Tensor
    y: Optional[torch.Tensor]
    adjs: List[Adj]
    idx_range: slice
    # Not a field (the reference unpacks four, driver/models.py:464): the distinct allocations
    # x / y / adjs are views of.  A Session cuts every structure tensor of a batch out of one arena,
    # so record_stream on the owners covers all of them (set on OwnedPreparedBatch instances).
    owners = ()

    @classmethod
    def from_proto_batch(cls, x: torch.Tensor, y: Optional[torch.Tensor], proto_batch: ProtoBatch):
        n_id = proto_batch.n_id
        return cls(x=fast_sampler.serial_index(x, n_id),
                   y=fast_sampler.serial_index(y.view(y.size(0), -1), n_id[:proto_batch.batch_size]).view(
                       (-1,) + tuple(y.shape[1:])) if y is not None else None,
                   adjs=proto_batch.adjs, idx_range=proto_batch.idx_range)

    @classmethod
    def from_fast_sampler(cls, prepared_sample):
        x, y, adjs, (start, stop) = prepared_sample
        owners = getattr(prepared_sample, "owners", None)
        if owners and cls is PreparedBatch:
            b = OwnedPreparedBatch(x, y.squeeze() if y is not None else None,
                                   [Adj__from_fast_sampler(adj) for adj in adjs], slice(start, stop))
            b.owners = owners
            return b
        return cls(x=x, y=y.squeeze() if y is not None else None,
                   adjs=[Adj__from_fast_sampler(adj) for adj in adjs], idx_range=slice(start, stop))

    def record_stream(self, stream):
        if self.owners:
            for t in self.owners:
                t.record_stream(stream)
            return
        if self.x is not None and self.x.is_cuda:
            self.x.record_stream(stream)
        if self.y is not None and self.y.is_cuda:
            self.y.record_stream(stream)
        for adj in self.adjs:
            adj.record_stream(stream)

    def to(self, device, non_blocking=False):
        # batches are born on the GPU: moving to the device they already live on is the identity
        dev = torch.device(device)
        if self.x is not None and self.x.is_cuda and dev.type == "cuda" and \
                (dev.index is None or dev.index == self.x.device.index):
            return self
        return PreparedBatch(
            x=self.x.to(device=device, non_blocking=non_blocking) if self.x is not None else None,
            y=self.y.to(device=device, non_blocking=non_blocking) if self.y is not None else None,
            adjs=[adj.to(device=device, non_blocking=non_blocking) for adj in self.adjs], idx_range=self.idx_range)

    @property
    def num_total_nodes(self):
        return self.x.size(0)

    @property
    def batch_size(self):
        return self.idx_range.stop - self.idx_range.start


class OwnedPreparedBatch(PreparedBatch):
    """A PreparedBatch (same four fields) that also carries ``owners`` as an instance attribute."""


@dataclass
class FastSamplerConfig:
    x_cpu: torch.Tensor
    x_gpu: torch.Tensor
    y: torch.Tensor
    rowptr: torch.Tensor
    col: torch.Tensor
    idx: torch.Tensor
    batch_size: int
    sizes: List[int]
    skip_nonfull_batch: bool
    pin_memory: bool
    distributed: bool
    partition_book: Optional[RangePartitionBook] = None
    cache: Optional[Cache] = None
    force_exact_num_batches: bool = False
    exact_num_batches: int = 0
    count_remote_frequency: bool = False
    use_cache: bool = False
    # extensions (see fast_sampler.Config)
    partition_tables: Optional[list] = None
    peer_table_ptrs: Optional[list] = None
    peer_table_pitch: int = 0
    fused_gather: bool = True

    def to_fast_sampler(self) -> fast_sampler.Config:
        c = fast_sampler.Config()
        for field in fields(self):
            if not self.distributed and field.name == "partition_book":
                continue
            v = getattr(self, field.name)
            if field.name == "cache" and v is None:
                v = Cache()
            setattr(c, field.name, v)
        return c

    def get_num_batches(self) -> int:
        if self.force_exact_num_batches:
            return self.exact_num_batches
        num_batches, r = divmod(self.idx.numel(), self.batch_size)
        if not self.skip_nonfull_batch and r > 0:
            num_batches += 1
        return num_batches


class FastSamplerStats(NamedTuple):
    total_blocked_dur: datetime.timedelta
    total_blocked_occasions: int

    @classmethod
    def from_session(cls, session):
        return cls(total_blocked_dur=session.total_blocked_dur,
                   total_blocked_occasions=session.total_blocked_occasions)


class FastSamplerDistributedStats(NamedTuple):
    remote_frequency_tensor: torch.Tensor
    remote_vertices_ordered_by_freq: torch.Tensor

    @classmethod
    def from_session(cls, session):
        assert session.num_consumed_batches == session.num_total_batches
        session.reduce_multithreaded_frequency_counts()
        return cls(remote_frequency_tensor=session.remote_frequency_tensor,
                   remote_vertices_ordered_by_freq=session.remote_vertices_ordered_by_freq)


class FastSamplerIter(Iterator[PreparedBatch]):
    session: fast_sampler.Session

    def __init__(self, num_threads: int, max_items_in_queue: int, cfg: FastSamplerConfig):
        ncfg = cfg.to_fast_sampler()
        self.session = fast_sampler.Session(num_threads, max_items_in_queue, ncfg)
        assert self.session.num_total_batches == cfg.get_num_batches()

    def __next__(self):
        if not self.session.config.distributed:
            sample = self.session.blocking_get_batch()
            if sample is None:
                raise StopIteration
            return PreparedBatch.from_fast_sampler(sample)
        sample = self.session.blocking_get_batch_distributed()
        if sample is None:
            raise StopIteration
        return ProtoDistributedBatch.from_fast_sampler(sample)

    def get_stats(self) -> FastSamplerStats:
        return FastSamplerStats.from_session(self.session)

    def get_distributed_stats(self) -> FastSamplerDistributedStats:
        return FastSamplerDistributedStats.from_session(self.session)


class ABCNeighborSampler(Iterable[PreparedBatch], Sized):
    pass


@dataclass
class FastSampler(ABCNeighborSampler):
    num_threads: int
    max_items_in_queue: int
    cfg: FastSamplerConfig

    @property
    def idx(self):
        return self.cfg.idx

    @idx.setter
    def idx(self, idx: torch.Tensor) -> None:
        self.cfg.idx = idx

    @property
    def cache(self):
        return self.cfg.cache

    @cache.setter
    def cache(self, cache: Cache) -> None:
        self.cfg.cache = cache

    def __iter__(self):
        return FastSamplerIter(self.num_threads, self.max_items_in_queue, self.cfg)

    def __len__(self):
        return self.cfg.get_num_batches()
