"""Batch containers and the sampler front end the reference's training code talks to, on top of
the GPU ``fast_sampler`` module.  The public names, fields and iterator protocol are those of
``fast_trainer/samplers.py`` -- ``FastSamplerConfig`` (:271-305), ``FastSampler`` (:372-399),
``FastSamplerIter`` (:331-357), ``ProtoDistributedBatch`` (:32-164), ``PreparedBatch`` (:213-268),
``ProtoBatch`` (:197-210), ``FastSamplerStats`` (:308-328) -- so ``driver/`` and
``fast_trainer/train.py`` run on it unchanged; the bodies are written for batches
that are born in HBM (one arena per batch, ``owners``) rather than moved there."""
from __future__ import annotations

import dataclasses
import datetime
from typing import Iterable, Iterator, List, NamedTuple, Optional, Sized

import torch

from . import fast_sampler
from .adj import Adj, Adj__from_fast_sampler
from .fast_sampler import Cache, RangePartitionBook


# ---- helpers shared by the containers -------------------------------------------------------------
def _wrap(raw_adjs) -> List[Adj]:
    """(rowptr, col, e_id, sizes) tuples of the sampler -> Adj records the models consume."""
    return list(map(Adj__from_fast_sampler, raw_adjs))


def _as_slice(rng) -> slice:
    lo, hi = rng
    return slice(lo, hi)


def _mark(stream, owners, tensors=(), adjs=()):
    """record_stream on the owning allocations when they are known, else on every device tensor."""
    if owners:
        for block in owners:
            block.record_stream(stream)
        return
    for t in tensors:
        if t is not None and t.is_cuda:
            t.record_stream(stream)
    for a in adjs:
        a.record_stream(stream)


def _flat_labels(y: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if y is None else y.squeeze()


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
    # sliced_cpu_labels already squeezed (None: not precomputed)
    y_flat: Optional[torch.Tensor] = None

    @classmethod
    def from_fast_sampler(cls, batch):
        if batch.sliced_cpu_features is None:
            raise AssertionError("distributed batch without sliced_cpu_features")
        return cls(batch.partition_nids, batch.sliced_cpu_features, batch.sliced_cpu_labels, batch.cached_nids,
                   batch.perm_partition_to_mfg, _wrap(batch.adjs), _as_slice(batch.idx_range),
                   getattr(batch, "n_id", None), getattr(batch, "x", None), getattr(batch, "owners", ()),
                   getattr(batch, "y_flat", None))

    def record_stream(self, stream):
        _mark(stream, self.owners, (*self.partition_nids, self.perm_partition_to_mfg, self.n_id, self.x), self.adjs)

    def to(self, device, stream=None, non_blocking=False, streams_to_sync=None, delay_feature_transfer=True):
        def move(t):
            return t.to(device, non_blocking=non_blocking)

        with torch.cuda.stream(stream):
            moved = dict(partition_nids=list(map(move, self.partition_nids)),
                         perm_partition_to_mfg=move(self.perm_partition_to_mfg),
                         adjs=[a.to(device, non_blocking=non_blocking) for a in self.adjs])
            if not delay_feature_transfer:
                moved["sliced_cpu_features"] = move(self.sliced_cpu_features)
            return self._replace(**moved)

    num_total_nodes = property(lambda self: self.perm_partition_to_mfg.size(0))
    num_cached_nodes = property(lambda self: self.cached_nids.size(0))


class ProtoBatch(NamedTuple):
    n_id: torch.Tensor
    adjs: List[Adj]
    idx_range: slice

    @classmethod
    def from_fast_sampler(cls, proto_sample):
        nodes, raw, rng = proto_sample
        return cls(nodes, _wrap(raw), _as_slice(rng))

    batch_size = property(lambda self: self.idx_range.stop - self.idx_range.start)


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
        nodes = proto_batch.n_id
        labels = None
        if y is not None:  # labels of the seeds = the first batch_size entries of n_id
            rows = fast_sampler.serial_index(y.view(y.size(0), -1), nodes[:proto_batch.batch_size])
            labels = rows.view((-1,) + tuple(y.shape[1:]))
        return cls(fast_sampler.serial_index(x, nodes), labels, proto_batch.adjs, proto_batch.idx_range)

    @classmethod
    def from_fast_sampler(cls, prepared_sample):
        feats, labels, raw, rng = prepared_sample
        owners = getattr(prepared_sample, "owners", None)
        kind = OwnedPreparedBatch if (owners and cls is PreparedBatch) else cls
        flat = getattr(prepared_sample, "y_flat", None)
        made = kind(feats, flat if flat is not None else _flat_labels(labels), _wrap(raw), _as_slice(rng))
        if kind is OwnedPreparedBatch:
            made.owners = owners
        return made

    def record_stream(self, stream):
        _mark(stream, self.owners, (self.x, self.y), self.adjs)

    def to(self, device, non_blocking=False):
        # batches are born on the GPU: moving to the device they already live on is the identity
        target = torch.device(device)
        here = self.x.device if self.x is not None else None
        if here is not None and here.type == "cuda" and target.type == "cuda" and target.index in (None, here.index):
            return self

        def move(t):
            return None if t is None else t.to(device=device, non_blocking=non_blocking)

        return PreparedBatch(move(self.x), move(self.y),
                             [a.to(device=device, non_blocking=non_blocking) for a in self.adjs], self.idx_range)

    num_total_nodes = property(lambda self: self.x.size(0))
    batch_size = property(lambda self: self.idx_range.stop - self.idx_range.start)


class OwnedPreparedBatch(PreparedBatch):
    """A PreparedBatch (same four fields) that also carries ``owners`` as an instance attribute."""


@dataclasses.dataclass
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
    local_parts: Optional[list] = None
    fused_gather: bool = True

    def to_fast_sampler(self) -> fast_sampler.Config:
        """The module-level Config with the same field values (a partition book only travels for
        distributed runs; a missing cache becomes the empty Cache the module expects)."""
        native = fast_sampler.Config()
        values = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}
        if not self.distributed:
            values.pop("partition_book")
        if values["cache"] is None:
            values["cache"] = Cache()
        for name, value in values.items():
            setattr(native, name, value)
        return native

    def get_num_batches(self) -> int:
        if self.force_exact_num_batches:
            return self.exact_num_batches
        seeds, per = self.idx.numel(), self.batch_size
        return seeds // per if self.skip_nonfull_batch else -(-seeds // per)


class FastSamplerStats(NamedTuple):
    total_blocked_dur: datetime.timedelta
    total_blocked_occasions: int

    @classmethod
    def from_session(cls, session):
        return cls(session.total_blocked_dur, session.total_blocked_occasions)


class FastSamplerDistributedStats(NamedTuple):
    remote_frequency_tensor: torch.Tensor
    remote_vertices_ordered_by_freq: torch.Tensor

    @classmethod
    def from_session(cls, session):
        if session.num_consumed_batches != session.num_total_batches:
            raise AssertionError("remote-frequency statistics need a fully consumed Session")
        session.reduce_multithreaded_frequency_counts()
        return cls(session.remote_frequency_tensor, session.remote_vertices_ordered_by_freq)


class FastSamplerIter(Iterator[PreparedBatch]):
    session: fast_sampler.Session

    def __init__(self, num_threads: int, max_items_in_queue: int, cfg: FastSamplerConfig):
        self.session = fast_sampler.Session(num_threads, max_items_in_queue, cfg.to_fast_sampler())
        if self.session.num_total_batches != cfg.get_num_batches():
            raise AssertionError("Session and FastSamplerConfig disagree on the number of batches")
        # one (fetch, wrap) pair per mode, chosen once
        if self.session.config.distributed:
            self._fetch, self._wrap = self.session.blocking_get_batch_distributed, ProtoDistributedBatch.from_fast_sampler
        else:
            self._fetch, self._wrap = self.session.blocking_get_batch, PreparedBatch.from_fast_sampler

    def __next__(self):
        got = self._fetch()
        if got is None:
            raise StopIteration
        return self._wrap(got)

    def get_stats(self) -> FastSamplerStats:
        return FastSamplerStats.from_session(self.session)

    def get_distributed_stats(self) -> FastSamplerDistributedStats:
        return FastSamplerDistributedStats.from_session(self.session)


class ABCNeighborSampler(Iterable[PreparedBatch], Sized):
    pass


def _cfg_attr(name: str) -> property:
    """Read/write pass-through to ``self.cfg.<name>`` (the driver swaps idx / cache between epochs)."""
    return property(lambda self: getattr(self.cfg, name), lambda self, value: setattr(self.cfg, name, value))


@dataclasses.dataclass
class FastSampler(ABCNeighborSampler):
    num_threads: int
    max_items_in_queue: int
    cfg: FastSamplerConfig

    idx = _cfg_attr("idx")
    cache = _cfg_attr("cache")

    def __iter__(self):
        return FastSamplerIter(self.num_threads, self.max_items_in_queue, self.cfg)

    def __len__(self):
        return self.cfg.get_num_batches()
