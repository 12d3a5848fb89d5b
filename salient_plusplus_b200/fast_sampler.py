"""GPU-native stand-in for the reference's pybind11 module ``fast_sampler``
(fast_sampler/fast_sampler.cpp:1280-1396): same names, argument meaning and error behaviour --
``Config``, ``Session``, ``ProtoDistributedBatch``, ``RangePartitionBook``, ``Cache``,
``sample_adj``, ``multilayer_sample``, ``full_sample``, ``serial_index``, ``to_row_major`` -- with
every operation executed by the sm_100a kernels of ``libsalient_b200.so`` through its C ABI
(``include/salient_b200.h``).  PyTorch is used for device memory, streams and events only.

Differences a caller can observe (all documented in INTEGRATION.md):
  * returned tensors live in HBM (the reference returns pinned host tensors that the caller then
    copies to the GPU; ``PreparedBatch.to(device)`` on our tensors is a no-op);
  * stochastic sampling uses a counter-based generator and the *correct* Floyd step, so sampled
    neighbourhoods differ from the reference's mt19937 stream (which is provably non-uniform,
    fast_sampler/sample_cpu.hpp:99); deterministic paths are bit-exact;
  * ``num_threads`` is accepted and ignored: parallelism comes from the GPU, batches are kept in
    flight on ``min(max_items_in_queue, SPP_SESSION_DEPTH=6)`` CUDA streams.
There is no CPU fallback: without a CUDA device or without the built library every entry point
raises.
"""
from __future__ import annotations

import copy
import ctypes
import datetime
import os
import time
from collections import OrderedDict, deque
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (META_EDGES0, META_OVERFLOW, SPP_MAX_HOPS, SPP_MAX_PARTS, SPP_META_WORDS,
                   BatchJob, FeatureMap, Graph, SalientB200Error, SamplerSizes, SamplerWs, check)

__all__ = ["Config", "Session", "ProtoDistributedBatch", "RangePartitionBook", "Cache", "sample_adj",
           "multilayer_sample", "full_sample", "serial_index", "to_row_major"]

c_vp = ctypes.c_void_p


# ------------------------------------------------------------------------------------------------
# device plumbing
# ------------------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise SalientB200Error("no CUDA device visible: salient_plusplus_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr(stream: Optional[torch.cuda.Stream] = None) -> int:
    return (stream or torch.cuda.current_stream()).cuda_stream


# Per-batch host work runs on the consumer's thread, so its cost bounds the public-API rate once
# the GPU needs < 100 us per batch.  ``torch.cuda.stream(...)`` costs ~15 us per use (two Stream
# objects + device-index normalisation); the raw (stream_id, device_index, device_type) triples
# torch itself passes to its C layer are switched directly instead.
_get_stream_raw = getattr(torch._C, "_cuda_getCurrentStream", None)
_set_stream_raw = getattr(torch._C, "_cuda_setStream", None)
_FAST_STREAMS = _get_stream_raw is not None and _set_stream_raw is not None
_STREAM_OBJS: dict = {}


def current_stream_cached(device_index: int) -> torch.cuda.Stream:
    """``torch.cuda.current_stream(device)`` without building a new Stream object every call."""
    if not _FAST_STREAMS:
        return torch.cuda.current_stream(device_index)
    raw = _get_stream_raw(device_index)
    st = _STREAM_OBJS.get(raw)
    if st is None:
        st = _STREAM_OBJS[raw] = torch.cuda.Stream(stream_id=raw[0], device_index=raw[1], device_type=raw[2])
    return st


class OwnedSample(tuple):
    """The tuple a non-distributed Session returns, plus ``owners``: the distinct allocations its
    tensors are views of (``record_stream`` on those covers every view), and ``y_flat``: the labels
    already squeezed the way the training loop takes them (None: squeeze ``[1]`` yourself)."""
    owners: tuple = ()
    y_flat = None


_RESIDENT: "OrderedDict[tuple, tuple]" = OrderedDict()
_RESIDENT_MAX = 32


def _resident(t: torch.Tensor, dtype: Optional[torch.dtype] = None, tag: str = "") -> torch.Tensor:
    """Device-resident, contiguous copy of ``t`` (uploaded once and cached while the source
    tensor is alive -- graph structure and feature tables are set-up state, not per-batch input)."""
    dev = _device()
    if t.is_cuda and t.is_contiguous() and (dtype is None or t.dtype == dtype):
        return t
    key = (t.data_ptr(), tuple(t.shape), t.dtype, t._version, dtype, tag, dev.index, t.device.type)
    hit = _RESIDENT.get(key)
    if hit is not None:
        _RESIDENT.move_to_end(key)
        return hit[1]
    d = t.to(device=dev, non_blocking=False)
    if dtype is not None and d.dtype != dtype:
        d = d.to(dtype)
    d = d.contiguous()
    _RESIDENT[key] = (t, d)  # keep the source alive so its data_ptr cannot be recycled
    while len(_RESIDENT) > _RESIDENT_MAX:
        _RESIDENT.popitem(last=False)
    return d


def clear_resident_cache() -> None:
    _RESIDENT.clear()


class _FeatureTable:
    """A feature matrix resident in HBM.  Rows whose byte size is not a multiple of 128 would
    straddle up to three 128-byte DRAM lines when gathered at random (ogbn-products: 200-byte
    rows -> 2.5 lines = 320 bytes fetched per row, measured with ncu), so the resident copy is laid
    out with a row pitch rounded up to 128 bytes when that costs at most 1/3 extra memory
    (``SPP_PAD_FEATURE_ROWS=0`` disables).  The gather output stays dense."""
    __slots__ = ("storage", "ptr", "pitch", "row_bytes", "dim", "dtype", "rows", "_src")

    def __init__(self, x: torch.Tensor):
        dev = _device()
        self.dim, self.dtype, self.rows = x.size(-1), x.dtype, x.size(0)
        es = x.element_size()
        self.row_bytes = self.dim * es
        pitch = self.row_bytes
        if (os.environ.get("SPP_PAD_FEATURE_ROWS", "1") != "0" and self.row_bytes % 128 != 0
                and self.row_bytes >= 96):
            cand = (self.row_bytes + 127) // 128 * 128
            if cand * 3 <= self.row_bytes * 4 and cand % es == 0:
                pitch = cand
        self.pitch = pitch
        self._src = x
        if pitch == self.row_bytes:
            self.storage = _resident(x, None, "features")
        else:
            st = torch.empty((self.rows, pitch // es), dtype=x.dtype, device=dev)
            chunk = max(1, (256 << 20) // max(self.row_bytes, 1))
            for r0 in range(0, self.rows, chunk):
                r1 = min(self.rows, r0 + chunk)
                st[r0:r1, :self.dim].copy_(x[r0:r1], non_blocking=False)
            self.storage = st
        self.ptr = self.storage.data_ptr()


_FEATURE_TABLES: "OrderedDict[tuple, _FeatureTable]" = OrderedDict()


def feature_table(x: torch.Tensor) -> _FeatureTable:
    key = (x.data_ptr(), tuple(x.shape), x.dtype, x._version, x.device.type,
           torch.cuda.current_device() if torch.cuda.is_available() else -1,
           os.environ.get("SPP_PAD_FEATURE_ROWS", "1"))
    t = _FEATURE_TABLES.get(key)
    if t is None:
        t = _FEATURE_TABLES[key] = _FeatureTable(x)
        while len(_FEATURE_TABLES) > 8:
            _FEATURE_TABLES.popitem(last=False)
    else:
        _FEATURE_TABLES.move_to_end(key)
    return t


class _DeviceGraph:
    """CSR in HBM: int64 rowptr, int32 col (ids < 2^31, the reference narrows too,
    fast_sampler/fast_sampler.cpp:196-199)."""

    _cache: "OrderedDict[tuple, _DeviceGraph]" = OrderedDict()

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor):
        if rowptr.dim() != 1 or col.dim() != 1 or rowptr.numel() < 1:
            raise RuntimeError("rowptr/col must be 1-D CSR arrays")
        self.num_nodes = rowptr.numel() - 1
        if self.num_nodes >= 2 ** 31:
            raise RuntimeError("graphs with >= 2^31 nodes are not supported")
        self.rowptr = _resident(rowptr, torch.int64, "rowptr")
        if col.dtype == torch.int32:
            self.col = _resident(col, torch.int32, "col")
        else:
            self.col = _resident(col, torch.int32, "col32")
        self.nnz = col.numel()
        if self.num_nodes > 0:
            deg = self.rowptr[1:] - self.rowptr[:-1]
            self.max_degree = int(deg.max().item())
        else:
            self.max_degree = 0
        self.c = Graph(self.rowptr.data_ptr(), self.col.data_ptr(), 0, 0, self.num_nodes)
        self._deg = None

    def degrees(self) -> torch.Tensor:
        """int64 out-degrees (built on first use: only full-neighbourhood Sessions need them)."""
        if self._deg is None:
            self._deg = self.rowptr[1:] - self.rowptr[:-1]
        return self._deg

    @classmethod
    def get(cls, rowptr: torch.Tensor, col: torch.Tensor) -> "_DeviceGraph":
        key = (rowptr.data_ptr(), rowptr.numel(), rowptr._version, col.data_ptr(), col.numel(), col._version,
               torch.cuda.current_device() if torch.cuda.is_available() else -1)
        g = cls._cache.get(key)
        if g is None:
            g = cls(rowptr, col)
            g._keep = (rowptr, col)
            cls._cache[key] = g
            while len(cls._cache) > 8:
                cls._cache.popitem(last=False)
        return g


def _sampler_sizes(bs: int, sizes: Sequence[int], g: _DeviceGraph, replace: bool = False) -> SamplerSizes:
    L = len(sizes)
    arr = (ctypes.c_int32 * max(L, 1))(*[int(s) for s in sizes])
    out = SamplerSizes()
    check(_lib.load().spp_sampler_sizes(int(bs), arr, L, int(bool(replace)), g.num_nodes, g.max_degree,
                                        ctypes.byref(out)), "spp_sampler_sizes")
    return out


class _Workspace:
    """Per-stream scratch of the sampler (hash table, discovery list, scan state, meta block)."""

    def __init__(self, sz: SamplerSizes, device: torch.device):
        self.max_nodes = int(sz.max_nodes)
        self.max_targets = int(sz.max_targets)
        self.table = torch.empty(int(sz.table_slots), dtype=torch.int64, device=device)
        self.table_direct = int(sz.table_direct)
        self.n_ids = torch.empty(self.max_nodes, dtype=torch.int32, device=device)
        self.tgt_start = torch.empty(self.max_targets, dtype=torch.int64, device=device)
        self.tgt_deg = torch.empty(self.max_targets, dtype=torch.int32, device=device)
        self.tile_state = torch.zeros(int(sz.tile_words), dtype=torch.int64, device=device)
        self.meta = torch.zeros(SPP_META_WORDS, dtype=torch.int64, device=device)
        self.cand = torch.empty(int(sz.cand_words), dtype=torch.int32, device=device)
        self._refresh()

    def _refresh(self):
        self.c = SamplerWs(self.table.data_ptr(), self.table.numel(), self.n_ids.data_ptr(), self.max_nodes,
                           self.tgt_start.data_ptr(), self.tgt_deg.data_ptr(), self.max_targets,
                           self.tile_state.data_ptr(), self.tile_state.numel(), self.meta.data_ptr(),
                           self.cand.data_ptr(), self.cand.numel(), self.table_direct, 0)

    def ensure_tiles(self, items: int):
        words = 2 + (max(items, 1) + 1023) // 1024 + 30
        if words > self.tile_state.numel():
            self.tile_state = torch.zeros(words, dtype=torch.int64, device=self.table.device)
            self._refresh()

    def meta_ptr(self, word: int) -> int:
        return self.meta.data_ptr() + 8 * word


Adj = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Tuple[int, int]]


def _sample_fused(g: _DeviceGraph, ws: _Workspace, sz: SamplerSizes, seeds_ptr: int, bs: int,
                  sizes: Sequence[int], replace: bool, rng_seed: int, stream_ptr: int, device,
                  want_nid: bool):
    """All hops sampled (fanout >= 0): one C call, no host synchronisation.  Output tensors are
    allocated at their upper bounds; the exact sizes arrive later through the meta block."""
    L = len(sizes)
    rowptrs = [torch.empty(int(sz.hop_targets[h]) + 1, dtype=torch.int64, device=device) for h in range(L)]
    cols = [torch.empty(max(int(sz.hop_edges[h]), 1), dtype=torch.int64, device=device) for h in range(L)]
    n_id = torch.empty(ws.max_nodes, dtype=torch.int64, device=device) if want_nid else None
    rp = (c_vp * max(L, 1))(*[t.data_ptr() for t in rowptrs])
    cp = (c_vp * max(L, 1))(*[t.data_ptr() for t in cols])
    caps = (ctypes.c_int64 * max(L, 1))(*[int(sz.hop_edges[h]) for h in range(L)])
    sz_arr = (ctypes.c_int32 * max(L, 1))(*[int(s) for s in sizes])
    check(_lib.load().spp_sample_minibatch(ctypes.byref(g.c), seeds_ptr, bs, sz_arr, L, int(bool(replace)),
                                           ctypes.c_uint64(rng_seed & 0xFFFFFFFFFFFFFFFF), ctypes.byref(ws.c),
                                           rp, cp, caps, n_id.data_ptr() if want_nid else None, stream_ptr),
          "spp_sample_minibatch")
    return rowptrs, cols, n_id


def _finish_fused(meta_host: torch.Tensor, rowptrs, cols, L: int) -> Tuple[int, List[Adj]]:
    """Cut the upper-bound buffers down to the sizes in the (host copy of the) meta block; the
    adjacency list is reversed like fast_sampler.cpp:224."""
    m = meta_host.tolist()
    if m[META_OVERFLOW]:
        raise SalientB200Error("sampler buffer bound exceeded on the device (SPP_META_OVERFLOW)")
    adjs: List[Adj] = []
    for h in range(L):
        T, E, S = m[h], m[META_EDGES0 + h], m[h + 1]
        e_id = torch.empty(0, dtype=torch.int64, device=rowptrs[h].device)
        adjs.append((rowptrs[h][:T + 1], cols[h][:E], e_id, (T, S)))
    adjs.reverse()
    return m[L], adjs


def _sample_stepwise(g: _DeviceGraph, ws: _Workspace, seeds: torch.Tensor, sizes: Sequence[int], replace: bool,
                     rng_seed: int, device) -> Tuple[int, List[Adj]]:
    """Hops whose edge count is data dependent (full neighbourhood): count -> read E -> allocate
    -> fill, one host synchronisation per phase (the layer-wise inference path)."""
    L = _lib.load()
    sp = _stream_ptr()
    bs = seeds.numel()
    check(L.spp_sample_begin(ctypes.byref(g.c), seeds.data_ptr(), bs, ctypes.byref(ws.c), sp), "spp_sample_begin")
    T = bs
    adjs: List[Adj] = []
    for h, k in enumerate(sizes):
        k = int(k)
        ws.ensure_tiles(T)
        rowptr = torch.empty(T + 1, dtype=torch.int64, device=device)
        check(L.spp_sample_hop_count(ctypes.byref(g.c), h, k, int(bool(replace)), T, ctypes.byref(ws.c),
                                     rowptr.data_ptr(), sp), "spp_sample_hop_count")
        E = int(ws.meta[META_EDGES0 + h].item())
        if T + E >= 2 ** 32 - 16:
            raise SalientB200Error(f"hop {h}: {T}+{E} candidates exceed the 32-bit position range")
        ws.ensure_tiles(E)
        col = torch.empty(max(E, 1), dtype=torch.int64, device=device)
        check(L.spp_sample_hop_fill(ctypes.byref(g.c), h, k, int(bool(replace)),
                                    ctypes.c_uint64(rng_seed & 0xFFFFFFFFFFFFFFFF), T, E, ctypes.byref(ws.c),
                                    rowptr.data_ptr(), col.data_ptr(), sp), "spp_sample_hop_fill")
        meta = ws.meta.tolist()
        if meta[META_OVERFLOW]:
            raise SalientB200Error("sampler buffer bound exceeded on the device (SPP_META_OVERFLOW)")
        S = meta[h + 1]
        adjs.append((rowptr, col[:E], torch.empty(0, dtype=torch.int64, device=device), (T, S)))
        T = S
    adjs.reverse()
    return T, adjs


_free_call_counter = [5489]


def _next_free_seed() -> int:
    _free_call_counter[0] += 1
    return _free_call_counter[0]


def _sample(rowptr, col, idx, sizes: Sequence[int], replace: bool, seed: Optional[int]):
    device = _device()
    if len(sizes) > SPP_MAX_HOPS:
        raise RuntimeError(f"at most {SPP_MAX_HOPS} hops are supported")
    g = _DeviceGraph.get(rowptr, col)
    seeds = idx.to(device=device, dtype=torch.int64).contiguous()
    sz = _sampler_sizes(seeds.numel(), sizes, g, replace)
    ws = _Workspace(sz, device)
    rng_seed = _next_free_seed() if seed is None else int(seed)
    if any(int(s) < 0 for s in sizes):
        nb, adjs = _sample_stepwise(g, ws, seeds, sizes, replace, rng_seed, device)
    else:
        rowptrs, cols, _ = _sample_fused(g, ws, sz, seeds.data_ptr(), seeds.numel(), sizes, replace, rng_seed,
                                         _stream_ptr(), device, False)
        nb, adjs = _finish_fused(ws.meta.cpu(), rowptrs, cols, len(sizes))
    return g, ws, nb, adjs


def sample_adj(rowptr: torch.Tensor, col: torch.Tensor, idx: torch.Tensor, num_neighbors: int, replace: bool,
               pin_memory: bool = False, *, seed: Optional[int] = None):
    """``fast_sampler.sample_adj`` (fast_sampler/sample_cpu.hpp:154-165, bound at
    fast_sampler.cpp:1339-1344): one hop; returns ``(rowptr, col, n_id int32, e_id)``."""
    _, ws, nb, adjs = _sample(rowptr, col, idx, [int(num_neighbors)], bool(replace), seed)
    rp, cl, e_id, _ = adjs[0]
    return rp, cl, ws.n_ids[:nb].clone(), e_id


def multilayer_sample(idx: torch.Tensor, sizes: Sequence[int], rowptr: torch.Tensor, col: torch.Tensor,
                      pin_memory: bool = False, *, seed: Optional[int] = None):
    """``fast_sampler.multilayer_sample`` (fast_sampler/fast_sampler.cpp:229-236, bound at
    :1345-1350): returns ``(n_id int64, [(rowptr, col, e_id, (T, S)) ...])``, outermost hop first."""
    _, ws, nb, adjs = _sample(rowptr, col, idx, list(sizes), False, seed)
    return ws.n_ids[:nb].to(torch.int64), adjs


def serial_index(input: torch.Tensor, idx: torch.Tensor, n: Optional[int] = None, pin_memory: bool = False):
    """``fast_sampler.serial_index`` (fast_sampler/fast_sampler.cpp:238-279): ``out[i] = in[idx[i]]``
    for ``i < min(len(idx), n)``; rows past ``len(idx)`` are left uninitialised like the reference."""
    if isinstance(n, bool):  # serial_index(in, idx, pin_memory) overload
        pin_memory, n = n, None
    if not ((input.dim() == 2 and input.stride(-1) == 1) or input.size(-1) == 1):
        raise RuntimeError("input must be 2D row-major tensor")
    device = _device()
    table = _resident(input, None, "table")
    ids = idx.to(device=device)
    if ids.dtype not in (torch.int64, torch.int32):
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    n = ids.numel() if n is None else int(n)
    f = input.size(-1)
    out = torch.empty((n, f), dtype=input.dtype, device=device)
    row_bytes = f * input.element_size()
    if n > 0 and ids.numel() > 0 and row_bytes > 0:
        check(_lib.load().spp_gather_rows(table.data_ptr(), row_bytes, ids.data_ptr(), int(ids.dtype == torch.int64),
                                          ids.numel(), None, out.data_ptr(), n, _stream_ptr()), "spp_gather_rows")
    return out


def to_row_major(t: torch.Tensor) -> torch.Tensor:
    """``fast_sampler.to_row_major`` (fast_sampler/fast_sampler.cpp:281-308).  Set-up-time layout
    conversion, done by the framework's strided copy."""
    if t.dim() != 2:
        raise RuntimeError("only support 2D tensors")
    tr, tc = t.shape
    if t.stride(0) == tc and t.stride(1) == 1:
        return t
    if not (t.stride(0) == 1 and t.stride(1) == tr):
        raise RuntimeError("input has unrecognizable stides")
    return t.contiguous()


def full_sample(*args, **kwargs):
    """``fast_sampler.full_sample`` (fast_sampler/fast_sampler.cpp:310-366).  Its only caller in the
    reference (FastPreSampler, fast_trainer/samplers.py:402-423) is broken; kept as a stub."""
    raise SalientB200Error("full_sample is not part of the mini-batch generation path (SURVEY.md section 2, #9)")


# ------------------------------------------------------------------------------------------------
# RangePartitionBook / Cache
# ------------------------------------------------------------------------------------------------
def _offsets_host(partition_offsets: torch.Tensor):
    off = [int(v) for v in partition_offsets.tolist()]
    if len(off) < 2 or len(off) - 1 > SPP_MAX_PARTS:
        raise RuntimeError(f"partition_offsets must hold between 2 and {SPP_MAX_PARTS + 1} entries")
    return off, (ctypes.c_int64 * (SPP_MAX_PARTS + 1))(*(off + [off[-1]] * (SPP_MAX_PARTS + 1 - len(off))))


class RangePartitionBook:
    """``fast_sampler.RangePartitionBook`` (fast_sampler/range_partition_book.hpp:31-57,
    .cpp:85-112; bound at fast_sampler.cpp:1368-1382)."""

    def __init__(self, rank: int, world_size: int, partition_offsets: torch.Tensor):
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.partition_offsets = partition_offsets

    def _off(self):
        return _offsets_host(self.partition_offsets)

    def _apply(self, nids: torch.Tensor, fn, out_dtype):
        device = _device()
        ids = nids.to(device=device, dtype=torch.int64).contiguous()
        out = torch.empty(ids.shape, dtype=out_dtype, device=device)
        fn(ids, out)
        return out if nids.is_cuda else out.cpu()

    def nid2localnid(self, nids: torch.Tensor, partition_idx: int) -> torch.Tensor:
        off, arr = self._off()
        P = len(off) - 1
        return self._apply(nids, lambda i, o: check(_lib.load().spp_nid2localnid(
            arr, P, int(partition_idx), i.data_ptr(), i.numel(), o.data_ptr(), _stream_ptr()), "spp_nid2localnid"),
            torch.int64)

    def nid2partid(self, nids: torch.Tensor) -> torch.Tensor:
        off, arr = self._off()
        P = len(off) - 1
        return self._apply(nids, lambda i, o: check(_lib.load().spp_nid2partid(
            arr, P, i.data_ptr(), i.numel(), o.data_ptr(), _stream_ptr()), "spp_nid2partid"), torch.int64)

    def nid_is_local(self, nids: torch.Tensor) -> torch.Tensor:
        off, arr = self._off()
        P = len(off) - 1
        return self._apply(nids, lambda i, o: check(_lib.load().spp_nid_is_local(
            arr, P, self.rank, i.data_ptr(), i.numel(), o.data_ptr(), _stream_ptr()), "spp_nid_is_local"),
            torch.bool)

    def partid2nids(self, partition_idx: int) -> torch.Tensor:
        off, _ = self._off()
        return torch.arange(off[partition_idx], off[partition_idx + 1], dtype=torch.int64)


class Cache:
    """``fast_sampler.Cache`` (fast_sampler/range_partition_book.hpp:60-91, .cpp:116-195; bound at
    fast_sampler.cpp:1383-1394).  The dense id -> cache-row map lives in HBM (int32[num_nodes],
    -1 = not cached); it is sized on demand instead of the reference's fixed 200 M entries."""

    def __init__(self, rank: int = 0, world_size: int = 0, cached_vertices: Optional[torch.Tensor] = None,
                 cached_features: Optional[torch.Tensor] = None):
        self._rank = int(rank)
        self._world_size = int(world_size)
        self._cached_vertices = (cached_vertices if cached_vertices is not None
                                 else torch.empty(0, dtype=torch.int64))
        self._cached_features = (cached_features if cached_features is not None
                                 else torch.empty((0, 0), dtype=torch.float16))
        self._map: Optional[torch.Tensor] = None
        self._map_nodes = 0

    rank = property(lambda self: self._rank)
    world_size = property(lambda self: self._world_size)
    cached_vertices = property(lambda self: self._cached_vertices)
    cached_features = property(lambda self: self._cached_features)

    def device_index(self, num_nodes: int):
        """``(index, nodes)``: the L2-resident lookup structure over node ids ``[0, nodes)``
        (``spp_cache_build_index``: membership bits + rank -> cache row), built on first use."""
        if self._map is None or self._map_nodes < num_nodes:
            device = _device()
            cv = self._cached_vertices.to(device=device, dtype=torch.int64).contiguous()
            if cv.numel() > 0:
                num_nodes = max(num_nodes, int(cv.max().item()) + 1)
            num_nodes = max(int(num_nodes), 1)
            L = _lib.load()
            m = torch.empty(int(L.spp_cache_index_bytes(num_nodes, cv.numel())), dtype=torch.uint8, device=device)
            check(L.spp_cache_build_index(cv.data_ptr(), cv.numel(), num_nodes, m.data_ptr(), _stream_ptr()),
                  "spp_cache_build_index")
            self._map, self._map_nodes = m, num_nodes
        return self._map, self._map_nodes

    def device_features(self) -> torch.Tensor:
        return _resident(self._cached_features, None, "cache_features")

    def device_table(self) -> Optional["_FeatureTable"]:
        if self._cached_features.dim() != 2 or self._cached_features.numel() == 0:
            return None
        return feature_table(self._cached_features)

    def _lookup(self, nids: torch.Tensor, fn, dtype):
        device = _device()
        ids = nids.to(device=device, dtype=torch.int64).contiguous()
        m, nodes = self.device_index(1)  # ids beyond the index are "not cached"
        out = torch.empty(ids.shape, dtype=dtype, device=device)
        check(fn(m.data_ptr(), nodes, ids.data_ptr(), ids.numel(), out.data_ptr(), _stream_ptr()), "cache lookup")
        return out if nids.is_cuda else out.cpu()

    def nid_is_cached(self, nids: torch.Tensor) -> torch.Tensor:
        return self._lookup(nids, _lib.load().spp_nid_is_cached, torch.bool)

    def nid2cachenid(self, nids: torch.Tensor) -> torch.Tensor:
        return self._lookup(nids, _lib.load().spp_nid2cachenid, torch.int64)


def make_feature_map(offsets: Sequence[int], rank: int, tables: Sequence[Optional[torch.Tensor]],
                     cache_table: Optional[torch.Tensor] = None, cache_index=None,
                     table_ptrs: Optional[Sequence[int]] = None, table_pitch: int = 0,
                     cache_pitch: int = 0, local_parts: Sequence[int] = ()) -> FeatureMap:
    """Fill the C ``spp_feature_map``: ``tables[p]`` are device tensors (local partitions) and/or
    ``table_ptrs[p]`` raw device pointers of IPC-mapped peer partitions; ``cache_index`` is
    ``Cache.device_index(num_nodes)``; ``local_parts`` lists further partitions resident on this
    GPU (fewer GPUs than partitions)."""
    P = len(offsets) - 1
    fm = FeatureMap()
    fm.num_parts = P
    fm.rank = int(rank)
    for p in range(SPP_MAX_PARTS + 1):
        fm.offsets[p] = int(offsets[min(p, P)])
    for p in range(P):
        ptr = None
        if table_ptrs is not None and table_ptrs[p]:
            ptr = int(table_ptrs[p])
        elif tables is not None and tables[p] is not None and tables[p].numel() > 0:
            ptr = tables[p].data_ptr()
        fm.tables[p] = ptr
    fm.cache_table = cache_table.data_ptr() if cache_table is not None and cache_table.numel() > 0 else None
    if cache_index is not None and fm.cache_table:
        fm.cache_index, fm.cache_index_nodes = cache_index[0].data_ptr(), int(cache_index[1])
    else:
        fm.cache_index, fm.cache_index_nodes = None, 0
    fm.table_pitch = int(table_pitch)
    fm.cache_pitch = int(cache_pitch)
    mask = 0
    for p in local_parts:
        mask |= 1 << int(p)
    fm.local_parts = mask
    return fm


# ------------------------------------------------------------------------------------------------
# Config / ProtoDistributedBatch / Session
# ------------------------------------------------------------------------------------------------
class Config:
    """``fast_sampler.Config`` (fast_sampler/fast_sampler.cpp:515-531,1290-1309): default
    constructible, read-write fields.  ``partition_tables`` / ``peer_table_ptrs`` are optional
    extensions: per-partition feature tables reachable from this GPU (local tensors or
    IPC-mapped peer pointers); when absent in distributed mode they are exchanged over
    ``torch.distributed`` by :mod:`salient_plusplus_b200.peer`."""

    def __init__(self):
        self.x_cpu = torch.empty((0, 0))
        self.x_gpu = torch.empty((0, 0))
        self.y = None
        self.rowptr = torch.zeros(1, dtype=torch.int64)
        self.col = torch.empty(0, dtype=torch.int64)
        self.idx = torch.empty(0, dtype=torch.int64)
        self.batch_size = 0
        self.sizes: List[int] = []
        self.skip_nonfull_batch = False
        self.pin_memory = False
        self.distributed = False
        self.partition_book: Optional[RangePartitionBook] = None
        self.cache: Cache = Cache()
        self.force_exact_num_batches = False
        self.exact_num_batches = 0
        self.count_remote_frequency = False
        self.use_cache = False
        # extensions (not in the reference)
        self.partition_tables = None
        self.peer_table_ptrs = None
        self.peer_table_pitch = 0
        self.local_parts = None  # further partitions resident on this GPU (fewer GPUs than partitions)
        self.fused_gather = True


class ProtoDistributedBatch:
    """``fast_sampler.ProtoDistributedBatch`` (fast_sampler/fast_sampler.cpp:180-188,1281-1289).
    ``n_id`` and ``x`` are extensions: the MFG node list and, when the partition tables are
    reachable, the features already gathered in MFG order by the fused P2P kernel."""

    def __init__(self):
        self.partition_nids: List[torch.Tensor] = []
        self.sliced_cpu_features = None
        self.sliced_cpu_labels = None
        self.cached_nids = None
        self.perm_partition_to_mfg = None
        self.adjs: List[Adj] = []
        self.idx_range: Tuple[int, int] = (0, 0)
        self.n_id = None
        self.x = None
        self.owners: tuple = ()  # distinct allocations the tensors above are views of
        self.y_flat = None       # sliced_cpu_labels.squeeze(), when the native host path made it


def _batch_ranges(n: int, cfg: Config) -> List[Tuple[int, int]]:
    """Batch ranges exactly as the reference enqueues them (fast_sampler/fast_sampler.cpp:587-627)."""
    out: List[Tuple[int, int]] = []
    if cfg.force_exact_num_batches:
        B = int(cfg.exact_num_batches)
        if B <= 0:
            raise RuntimeError("exact_num_batches must be positive")
        avg = n // B - 1
        if avg < 0:
            raise RuntimeError("force_exact_num_batches: fewer seeds than batches")
        sizes = [avg] * B
        rem = n - avg * B
        i = 0
        while rem > 0:
            sizes[i % B] += 1
            rem -= 1
            i += 1
        s = 0
        for b in sizes:
            out.append((s, s + b))
            s += b
    else:
        bs = int(cfg.batch_size)
        if bs <= 0:
            raise RuntimeError("batch_size must be positive")
        for i in range(0, n, bs):
            this = min(n, i + bs) - i
            if cfg.skip_nonfull_batch and this < bs:
                continue
            out.append((i, i + this))
    return out


class _Slot:
    """One in-flight mini-batch: sampler workspace, CUDA stream, completion event, pinned meta
    block.  Slots are recycled across Sessions (a new Session is created every epoch,
    fast_trainer/samplers.py:394-396) through ``_SLOT_POOL`` so that epoch start-up does not
    re-allocate hash tables and streams."""

    def __init__(self, sz: SamplerSizes, device, max_bs: int, split_words: int):
        self.ws = _Workspace(sz, device)
        self.stream = torch.cuda.Stream(device)
        self.stream_raw = (self.stream.stream_id, self.stream.device_index, self.stream.device_type)
        self.event = torch.cuda.Event()
        self.meta_host = torch.empty(SPP_META_WORDS + SPP_MAX_PARTS + 2, dtype=torch.int64).pin_memory()
        self.seeds = torch.empty(max(max_bs, 1), dtype=torch.int64, device=device)
        self.seeds_host = torch.empty(max(max_bs, 1), dtype=torch.int64).pin_memory()  # H2D staging
        if split_words:
            self.split_scratch = torch.empty(split_words, dtype=torch.int32, device=device)
            self.counts = torch.zeros(SPP_MAX_PARTS + 2, dtype=torch.int64, device=device)
        self.job = None
        self.ticket = None
        self.cjob = BatchJob()
        # graph replay: device-resident per-batch block + its pinned staging copy (spp_device_job)
        self.job_dev = torch.zeros(ctypes.sizeof(_lib.DeviceJob) // 8, dtype=torch.int64, device=device)
        self.job_host = torch.zeros(ctypes.sizeof(_lib.DeviceJob) // 8, dtype=torch.int64).pin_memory()


_EXECUTORS: dict = {}


def _executor(device_index: int):
    """The native enqueue executor of a device (one worker thread inside libsalient_b200)."""
    ex = _EXECUTORS.get(device_index)
    if ex is None:
        ex = _lib.load().spp_executor_create(device_index)
        if not ex:
            raise SalientB200Error("spp_executor_create failed: " + _lib.load().spp_last_error().decode())
        _EXECUTORS[device_index] = ex
    return ex


def _arena_cuts(lay, m: Sequence[int], L: int, P: Optional[int]) -> List[int]:
    """Piece sizes that cut one batch's arena into its exact-size tensors with a single
    ``split_with_sizes``.  ``lay`` = Session._layout(...), ``m`` = host copy of the size block
    (meta words, then the owner-split counts), ``P`` = partitions (None: not distributed).
    Pieces: per hop ``[rowptr (T+1) | slack | col (E) | slack]`` -> indices 4h and 4h+2; then,
    distributed, ``[n_id | slack | P partition buckets | cached | slack | perm | slack | tail]``
    starting at index 4L, else one tail piece."""
    a_off, a_nid, _, a_words, mx = lay
    cuts: List[int] = []
    for h in range(L):
        ro, co = a_off[h]
        T, E = m[h], m[META_EDGES0 + h]
        end = a_off[h + 1][0] if h + 1 < L else a_nid
        cuts += [T + 1, co - ro - T - 1, E, end - co - E]
    if P is None:
        cuts.append(a_words - a_nid)
    else:
        counts = list(m[SPP_META_WORDS:SPP_META_WORDS + P + 1])
        nb = m[L]
        cuts += [nb, mx - nb] + counts + [mx - sum(counts), nb, mx - nb, a_words - a_nid - 3 * mx]
    return cuts


_SLOT_POOL: dict = {}
_EMPTY_EID: dict = {}


def _empty_eid(device) -> torch.Tensor:
    """e_id is always empty (fast_sampler/sample_cpu.hpp:120); one shared tensor per device."""
    t = _EMPTY_EID.get(device)
    if t is None:
        t = _EMPTY_EID[device] = torch.empty(0, dtype=torch.int64, device=device)
    return t


def clear_slot_pool() -> None:
    _SLOT_POOL.clear()


def host_session_spec(*, device, n_hops: int, num_parts: int, layout, has_x: bool, feat_dim: int, feat_dtype,
                      y: Optional[torch.Tensor], y_in_arena: bool, ranges, batch_edges, idx_host_ptr: int,
                      idx_dev_ptr: int, executor, entry_points, slots, e_id: torch.Tensor) -> dict:
    """The construction record of ``_spp_host.HostSession`` (csrc/host_session.cpp).  ``layout`` =
    ``Session._layout(None)``; ``num_parts`` < 0: not distributed; ``entry_points`` = the ctypes
    functions (submit, poll, wait, last_error) -- of libsalient_b200.so in a Session, of a
    recording stand-in in the CPU tests; ``slots``: objects with ``cjob`` / ``meta_host`` /
    ``seeds`` / ``stream_raw`` (a ``_Slot``)."""
    a_off, a_nid, a_y, a_words, node_bound = layout
    fn_addr = [ctypes.cast(f, ctypes.c_void_p).value for f in entry_points]
    return {
        "device": device, "n_hops": int(n_hops), "num_parts": int(num_parts),
        "hop_offsets": [(int(r), int(c)) for r, c in a_off], "nid_offset": int(a_nid), "y_offset": int(a_y),
        "arena_words": int(a_words), "node_bound": int(node_bound),
        "has_x": bool(has_x), "feat_dim": int(feat_dim), "feat_dtype": feat_dtype,
        "has_y": y is not None, "y_in_arena": bool(y_in_arena),
        "y_cols": int(y.size(-1)) if y is not None else 0, "y_dtype": y.dtype if y is not None else torch.int64,
        "ranges": [(int(a), int(b)) for a, b in ranges],
        "batch_edges": [int(v) for v in batch_edges] if batch_edges is not None else [],
        "idx_host_ptr": int(idx_host_ptr or 0), "idx_dev_ptr": int(idx_dev_ptr or 0),
        "executor": int(executor), "submit_fn": fn_addr[0], "poll_fn": fn_addr[1], "wait_fn": fn_addr[2],
        "last_error_fn": fn_addr[3],
        "slots": [(ctypes.addressof(s.cjob), s.meta_host.data_ptr(), s.seeds.data_ptr(), int(s.stream_raw[0]),
                   int(s.stream_raw[1]), int(s.stream_raw[2])) for s in slots],
        "e_id": e_id, "error_cls": SalientB200Error,
    }


class Session:
    """``fast_sampler.Session`` (fast_sampler/fast_sampler.cpp:533-936, bound at :1310-1338)."""

    def __init__(self, num_threads: int, max_items_in_queue: int, config: Config):
        if max_items_in_queue <= 0:
            raise RuntimeError(f"max_items_in_queue ({max_items_in_queue}) must be positive")
        _t = [time.perf_counter()] if os.environ.get("SPP_DEBUG_TIMING") else None
        self._config = copy.copy(config)  # copied by value like fast_sampler.cpp:541-542
        cfg = self._config
        self._device = _device()
        self._lib = _lib.load()
        self._sizes = [int(s) for s in cfg.sizes]
        if len(self._sizes) > SPP_MAX_HOPS:
            raise RuntimeError(f"at most {SPP_MAX_HOPS} hops are supported")
        self._g = _DeviceGraph.get(cfg.rowptr, cfg.col)
        # Seeds: a device-resident idx is used in place; a host idx (what the reference's driver
        # passes) is pinned once and each batch's slice is copied H2D on that batch's stream.
        # (each batch's slice is staged through the slot's pinned buffer: the reference's driver
        # passes pageable tensors, and pinning a fresh idx every epoch would cost milliseconds)
        idx = cfg.idx.to(torch.int64).contiguous()
        if idx.is_cuda:
            self._idx, self._idx_host = idx, None
        else:
            self._idx, self._idx_host = None, idx
            self._idx_host_ptr = idx.data_ptr()
        self._ranges = _batch_ranges(idx.numel(), cfg)
        self._num_total = len(self._ranges)
        self._num_consumed = 0
        self._next = 0
        self._blocked_dur = datetime.timedelta(0)
        self._blocked_occasions = 0
        self._native = None
        self._full = any(s < 0 for s in self._sizes)
        # Layer-wise inference (driver/models.py:455-480) samples ONE full-neighbourhood hop per
        # batch: its edge count is the degree sum of the seeds, known before anything is launched,
        # so those batches take the asynchronous, pipelined path like sampled ones.  Other
        # full-neighbourhood configurations need the device's counts between hops (stepwise path).
        self._edge_bound = None
        self._batch_edges = None
        if self._full and len(self._sizes) == 1 and self._ranges and os.environ.get("SPP_ASYNC_FULL", "1") != "0":
            self._batch_edges = self._leading_hop_edges(idx)
            # the workspace bound is rounded up so Sessions over different seed sets (one per layer
            # and epoch) find their slots in the pool again; outputs are sized per batch
            self._edge_bound = -(-max(max(self._batch_edges), 1) // 65536) * 65536
            self._full = False
        self._y = _resident(cfg.y, None, "y") if cfg.y is not None else None
        if self._y is not None and self._y.dim() == 1:
            self._y = self._y.view(-1, 1)
        if _t: _t.append(time.perf_counter())
        self._setup_features()
        if _t: _t.append(time.perf_counter())
        self._row_bytes = self._feat_shape[0] * torch.empty(0, dtype=self._feat_shape[1]).element_size()
        max_bs = max((e - s for s, e in self._ranges), default=0)
        if self._edge_bound is not None:
            # spp_sampler_sizes bounds a full-neighbourhood hop by targets * max_degree: hand it the
            # smallest per-target figure that covers the largest batch's exact edge count
            per_target = -(-self._edge_bound // max(max_bs, 1))
            sz = SamplerSizes()
            arr = (ctypes.c_int32 * 1)(self._sizes[0])
            check(self._lib.spp_sampler_sizes(int(max_bs), arr, 1, 0, self._g.num_nodes, int(per_target),
                                              ctypes.byref(sz)), "spp_sampler_sizes")
            self._sz = sz
        else:
            self._sz = _sampler_sizes(max_bs, self._sizes, self._g)
        depth = int(os.environ.get("SPP_SESSION_DEPTH", "6"))
        depth = max(1, min(depth, int(max_items_in_queue), max(self._num_total, 1)))
        split_words = int(self._lib.spp_split_scratch_words(self._sz.max_nodes)) if cfg.distributed else 0
        sz = self._sz
        self._pool_key = (self._device.index, int(sz.max_nodes), int(sz.max_targets), int(sz.table_slots),
                          int(sz.table_direct), int(sz.tile_words), int(sz.cand_words), max_bs, split_words)
        pool = _SLOT_POOL.setdefault(self._pool_key, [])
        self._slots = [pool.pop() if pool else _Slot(sz, self._device, max_bs, split_words) for _ in range(depth)]
        self._released = False
        # arena layout (int64 words) of one batch's structure outputs, at their upper bounds
        self._y_in_arena = self._y is not None and self._y.dtype == torch.int64
        self._max_bs = max_bs
        self._lay = self._layout(None)
        self._executor = None
        if not self._full and os.environ.get("SPP_EXECUTOR", "1") != "0":
            self._executor = _executor(self._device.index)
        # The per-batch bookkeeping (output allocation, job fill, submit, view cutting) runs in the
        # native host path (csrc/host_session.cpp) whenever batches go through the executor;
        # SPP_NATIVE_HOST=0 keeps the interpreter implementation below (_enqueue / _finalize).
        use_native = self._executor is not None and os.environ.get("SPP_NATIVE_HOST", "1") != "0"
        if not self._full:
            for s_ in self._slots:
                self._init_job(s_)
                if not use_native:
                    self._prime_allocator(s_)
        self._free = deque(self._slots)
        self._pending: deque = deque()
        self._freq = None
        self._freq_reduced = False
        self.remote_frequency_tensor = torch.empty(0, dtype=torch.int64)
        self.remote_vertices_ordered_by_freq = torch.empty(0, dtype=torch.int64)
        self._slice_result = None
        self._slice_pending = None
        self._slice_stream = None
        # the side streams must see set-up work (uploads, cache map) issued on the current stream
        cur = torch.cuda.current_stream()
        for s in self._slots:
            s.stream.wait_stream(cur)
        if _t: _t.append(time.perf_counter())
        if use_native:
            self._native = _lib.load_host().HostSession(self._native_spec())
            for s_ in self._slots:  # the native path allocates ONE block per batch: prime that size
                self._prime_allocator(s_, block_bytes=int(self._native.block_bytes))
            self._native.fill()
        else:
            while self._free and self._next < self._num_total:
                self._enqueue()
        if _t:
            _t.append(time.perf_counter())
            print("[spp] Session init us: graph/idx %.0f features %.0f slots/jobs %.0f first enqueues %.0f" % tuple(
                (b - a) * 1e6 for a, b in zip(_t[:-1], _t[1:])), flush=True)

    # -- set-up -------------------------------------------------------------------------------
    def _leading_hop_edges(self, idx: torch.Tensor) -> List[int]:
        """Per-batch edge count of a full-neighbourhood first hop: every seed position is a
        target (duplicates included, sample_cpu.hpp:36-64), so a batch's count is the degree sum of
        its slice of idx.  One device pass + one read-back at Session set-up."""
        g = self._g
        deg = g.degrees()
        c = torch.cumsum(deg[idx.to(self._device)], 0)
        c = torch.cat([c.new_zeros(1), c])
        st = torch.tensor([a for a, _ in self._ranges], dtype=torch.int64, device=self._device)
        en = torch.tensor([b for _, b in self._ranges], dtype=torch.int64, device=self._device)
        return (c[en] - c[st]).tolist()

    def _setup_features(self):
        cfg = self._config
        self._fm = None
        self._x_table = None
        self._x_cpu_dev = None
        if not cfg.distributed:
            x = cfg.x_cpu
            if x is not None and x.dim() == 2 and x.numel() > 0:
                self._x_table = feature_table(x)
            self._feat_shape = (x.size(-1), x.dtype) if x is not None and x.dim() == 2 else (0, torch.float16)
            return
        book = cfg.partition_book
        if book is None:
            raise RuntimeError("distributed Session needs a partition_book")
        off, _ = _offsets_host(book.partition_offsets)
        self._off = off
        self._P = len(off) - 1
        self._rank = int(book.rank)
        if not (0 <= self._rank < self._P):
            raise RuntimeError("partition_book.rank out of range")
        xg = cfg.x_gpu if cfg.x_gpu is not None else torch.empty((0, 0))
        xc = cfg.x_cpu if cfg.x_cpu is not None else torch.empty((0, 0))
        self._x_gpu_rows = xg.size(0) if xg.dim() == 2 else 0
        has_g = xg.dim() == 2 and xg.numel() > 0
        has_c = xc.dim() == 2 and xc.numel() > 0
        if has_c:
            self._x_cpu_dev = _resident(xc, None, "x_cpu")
        if has_g and has_c:
            key = ("xlocal", xg.data_ptr(), xc.data_ptr(), xg.size(0), xc.size(0))
            hit = _RESIDENT.get(key)
            if hit is None:
                local = torch.cat([_resident(xg, None, "x_gpu"), self._x_cpu_dev], dim=0)
                _RESIDENT[key] = ((xg, xc), local)
            else:
                local = hit[1]
        elif has_g:
            local = _resident(xg, None, "x_gpu")
        elif has_c:
            local = self._x_cpu_dev
        else:
            local = None
        self._x_local = local
        ref = xg if has_g else xc
        self._feat_shape = (ref.size(-1), ref.dtype) if ref.dim() == 2 else (0, torch.float16)
        # an empty cache makes the cache branch (fast_sampler.cpp:1108-1260) produce exactly what the
        # no-cache branch (:1031-1107) produces, so it takes the kernel path without a cache map
        self._use_cache = bool(cfg.use_cache) and cfg.cache.cached_vertices.numel() > 0
        self._cache_map = cfg.cache.device_index(self._g.num_nodes) if self._use_cache else None
        self._cache_feats = cfg.cache.device_features() if self._use_cache else None
        # book-only map for the split kernel
        local_parts = [int(p) for p in (getattr(cfg, "local_parts", None) or ())]
        self._split_fm = make_feature_map(off, self._rank, [None] * self._P, local_parts=local_parts)
        if self._use_cache:
            self._split_fm.cache_index = self._cache_map[0].data_ptr()
            self._split_fm.cache_index_nodes = self._cache_map[1]
        # full map (tables of every partition) for the fused gather, when reachable.  The tables
        # are _FeatureTable objects: every rank derives the same row pitch from (F, dtype).
        if cfg.fused_gather and local is not None:
            ltab = feature_table(local)
            tables = [None] * self._P
            ptrs = [0] * self._P
            if cfg.partition_tables is not None:
                for p, t in enumerate(cfg.partition_tables):
                    if t is not None and p != self._rank:
                        ft = feature_table(t)
                        if ft.pitch != ltab.pitch:
                            raise RuntimeError("partition tables disagree on the feature row size")
                        tables[p] = ft.storage
            if cfg.peer_table_ptrs is not None:
                ptrs = [int(v or 0) for v in cfg.peer_table_ptrs]
                pitch = int(getattr(cfg, "peer_table_pitch", 0) or 0) or ltab.row_bytes
                if pitch != ltab.pitch:
                    raise RuntimeError(f"peer tables have row pitch {pitch}, local table {ltab.pitch}")
            tables[self._rank] = ltab.storage
            ptrs[self._rank] = 0
            reachable = all((tables[p] is not None) or ptrs[p] or off[p + 1] == off[p] for p in range(self._P))
            if not reachable:
                from . import peer
                got = peer.exchange_partition_tables(ltab.storage, self._rank, self._P)
                if got is not None:
                    ptrs = got
                    ptrs[self._rank] = 0
                    reachable = True
            if reachable:
                self._part_tables = (tables, ltab)  # keep alive
                ctab = cfg.cache.device_table() if self._use_cache else None
                self._fm = make_feature_map(off, self._rank, tables, ctab.storage if ctab else None,
                                            self._cache_map, ptrs, ltab.pitch, ctab.pitch if ctab else 0,
                                            local_parts=local_parts)

    def _layout(self, edges0: Optional[int]):
        """(per-hop (rowptr, col) offsets, n_id offset, y offset, total words, node bound) of one
        batch's int64 arena.  ``edges0``: exact edge count of a leading full-neighbourhood hop (the
        arena and x of such a batch are sized by it instead of by the Session-wide bound)."""
        sz, cfg = self._sz, self._config
        off, o = [], 0
        m = int(sz.max_nodes)
        for h in range(len(self._sizes)):
            T, E = int(sz.hop_targets[h]), max(int(sz.hop_edges[h]), 1)
            if h == 0 and edges0 is not None:
                # rounded up so that the per-batch outputs fall into a handful of size classes the caching
                # allocator can re-use (an exact size per batch made every few batches a cudaMalloc: ms)
                E = min(max(-(-int(edges0) // 16384) * 16384, 16384), max(int(sz.hop_edges[h]), 1))
                E = max(E, int(edges0), 1)
                m = min(m, T + E)
            off.append((o, o + T + 1))
            o += T + 1 + E
        nid = o
        if cfg.distributed:
            o += 3 * m  # n_id, bucket_ids, perm
        yo = o
        if self._y_in_arena:
            o += self._max_bs * self._y.size(-1)
        return off, nid, yo, max(o, 1), m

    # -- enqueue / finalise ---------------------------------------------------------------------
    def _init_job(self, slot: "_Slot"):
        """Static part of the slot's spp_batch_job (graph, workspace, tables, fan-outs)."""
        cfg, sz, j = self._config, self._sz, slot.cjob
        ctypes.memset(ctypes.byref(j), 0, ctypes.sizeof(j))
        j.graph = self._g.c
        j.ws = slot.ws.c
        L = len(self._sizes)
        j.n_hops = L
        for h in range(L):
            j.sizes[h] = self._sizes[h]
            j.out_col_cap[h] = int(sz.hop_edges[h])
        j.replace = 0
        j.row_bytes = self._row_bytes
        j.seeds_dev = slot.seeds.data_ptr()
        j.meta_host = slot.meta_host.data_ptr()
        j.stream = slot.stream.cuda_stream
        # one CUDA graph per slot replays the whole launch sequence (SPP_GRAPH=0: plain launches)
        j.job_dev = slot.job_dev.data_ptr()
        j.job_host = slot.job_host.data_ptr()
        j.seeds_stage_host = slot.seeds_host.data_ptr()
        j.batch_size_cap = max(self._max_bs, 1)
        if self._edge_bound is not None:
            j.out_col_bound[0] = int(sz.hop_edges[0])
        if self._y is not None:
            j.y_table = self._y.data_ptr()
            j.y_row_bytes = self._y.size(-1) * self._y.element_size()
        if not cfg.distributed:
            if self._x_table is not None:
                j.feature_mode = 1
                j.table = self._x_table.ptr
                j.table_pitch = self._x_table.pitch
        else:
            j.do_split = 1
            j.use_cache = int(self._use_cache)
            j.bucket_counts = slot.counts.data_ptr()
            j.split_scratch = slot.split_scratch.data_ptr()
            if self._fm is not None:
                j.feature_mode = 2
                j.fmap = self._fm
            else:
                j.fmap = self._split_fm
        # capture the slot's graph now (only the static fields matter; whether seeds are staged from
        # the host and whether n_id is exported are part of them)
        j.seeds_host = self._idx_host_ptr if self._idx_host is not None else None
        j.n_id_out = slot.job_dev.data_ptr() if cfg.distributed else None   # placeholder, non-NULL
        j.batch_size = j.batch_size_cap
        check(self._lib.spp_batch_prepare(ctypes.byref(j)), "spp_batch_prepare")

    def _native_spec(self) -> dict:
        """Everything csrc/host_session.cpp needs to run this Session's batches (see host_session_spec)."""
        lib, cfg = self._lib, self._config
        fdim, fdtype = self._feat_shape
        return host_session_spec(
            device=self._device, n_hops=len(self._sizes), num_parts=self._P if cfg.distributed else -1, layout=self._lay,
            has_x=bool(self._slots[0].cjob.feature_mode), feat_dim=fdim, feat_dtype=fdtype, y=self._y,
            y_in_arena=self._y_in_arena, ranges=self._ranges, batch_edges=self._batch_edges,
            idx_host_ptr=self._idx_host_ptr if self._idx_host is not None else 0,
            idx_dev_ptr=self._idx.data_ptr() if self._idx is not None else 0, executor=self._executor,
            entry_points=(lib.spp_executor_submit, lib.spp_executor_poll, lib.spp_executor_wait, lib.spp_last_error),
            slots=self._slots, e_id=_empty_eid(self._device))

    def _prime_allocator(self, slot: "_Slot", blocks: int = 3, block_bytes: Optional[int] = None):
        """Per-batch outputs are allocated at their upper bounds from PyTorch's caching allocator
        on the slot's stream.  A cache miss there is a cudaMalloc of a few hundred MB (2-30 ms,
        measured), so the first time a slot sees a given output size its pool is primed with the
        blocks a steady-state pipeline needs (in flight + held by the consumer + prefetched).
        ``block_bytes``: the native host path's single block per batch (arena + x + y)."""
        fdim, fdtype = self._feat_shape
        x_rows = slot.ws.max_nodes if slot.cjob.feature_mode else 0
        key = ("block", int(block_bytes)) if block_bytes is not None else (self._lay[3], x_rows, fdim, fdtype)
        primed = getattr(slot, "primed", None)
        if primed is None:
            primed = slot.primed = set()
        if key in primed:
            return
        primed.add(key)
        with torch.cuda.stream(slot.stream):
            keep = []
            for _ in range(blocks):
                if block_bytes is not None:
                    keep.append(torch.empty(int(block_bytes), dtype=torch.uint8, device=self._device))
                    continue
                keep.append(torch.empty(self._lay[3], dtype=torch.int64, device=self._device))
                if x_rows:
                    keep.append(torch.empty((x_rows, fdim), dtype=fdtype, device=self._device))
            del keep

    def _enqueue(self):
        slot = self._free.popleft()
        start, stop = self._ranges[self._next]
        # layer-wise batches too are allocated at the Session-wide bound (the edge count of the largest
        # batch): sizing them batch by batch made the caching allocator fall through to cudaMalloc /
        # cudaFree in steady state (stalls of 3-80 ms); the exact per-batch capacity still travels to
        # the kernels (out_col_cap) and the views are cut to the exact sizes
        lay = self._lay
        self._next += 1
        bs = stop - start
        cfg = self._config
        rng_seed = (stop * 17 + 5) & 0xFFFFFFFF  # fast_sampler.cpp:994
        L = len(self._sizes)
        job = {"range": (start, stop), "bs": bs}
        ws = slot.ws
        fdim, fdtype = self._feat_shape
        if self._full:
            # data-dependent sizes: stepwise ABI with host synchronisation (layer-wise inference)
            with torch.cuda.stream(slot.stream):
                sp = slot.stream.cuda_stream
                if self._idx_host is not None:
                    seeds = slot.seeds[:bs]
                    if bs:
                        slot.seeds_host[:bs].copy_(self._idx_host[start:stop])
                        seeds.copy_(slot.seeds_host[:bs], non_blocking=True)
                else:
                    seeds = self._idx[start:stop]
                nb, adjs = _sample_stepwise(self._g, ws, seeds, self._sizes, False, rng_seed, self._device)
                job["ready"] = (nb, adjs)
                n_dev = ws.meta_ptr(L)
                if not cfg.distributed:
                    x = torch.empty((nb if self._x_table is not None else 0, fdim), dtype=fdtype, device=self._device)
                    if self._x_table is not None and nb:
                        check(self._lib.spp_gather_rows_pitched(self._x_table.ptr, self._x_table.pitch, self._row_bytes,
                                                                ws.n_ids.data_ptr(), 0, nb, None, x.data_ptr(), nb, sp),
                              "spp_gather_rows_pitched")
                    job["x"] = x
                else:
                    arena = torch.empty(3 * max(nb, 1), dtype=torch.int64, device=self._device)
                    cap = max(nb, 1)
                    check(self._lib.spp_sample_export_nids(ctypes.byref(ws.c), L, arena.data_ptr(), 1, cap, sp),
                          "spp_sample_export_nids")
                    job["n_id"], job["bucket_ids"], job["perm"] = arena[:cap], arena[cap:2 * cap], arena[2 * cap:]
                    check(self._lib.spp_split_by_owner(ctypes.byref(self._split_fm), int(self._use_cache),
                                                       ws.n_ids.data_ptr(), 0, nb, None, job["bucket_ids"].data_ptr(),
                                                       job["perm"].data_ptr(), slot.counts.data_ptr(),
                                                       slot.split_scratch.data_ptr(), sp), "spp_split_by_owner")
                    if self._fm is not None:
                        x = torch.empty((nb, fdim), dtype=fdtype, device=self._device)
                        if nb:
                            check(self._lib.spp_gather_partitioned(ctypes.byref(self._fm), self._row_bytes,
                                                                   ws.n_ids.data_ptr(), 0, nb, None,
                                                                   slot.split_scratch.data_ptr(), x.data_ptr(), nb,
                                                                   None, sp), "spp_gather_partitioned")
                        job["x"] = x
                    slot.meta_host[SPP_META_WORDS:].copy_(slot.counts, non_blocking=True)
                job["y"] = self._labels_for(seeds, bs, sp)
                slot.event.record(slot.stream)
            slot.ticket = None
        else:
            j = slot.cjob
            # the outputs belong to the slot's stream
            if _FAST_STREAMS:
                prev = _get_stream_raw(self._device.index)
                _set_stream_raw(stream_id=slot.stream_raw[0], device_index=slot.stream_raw[1],
                                device_type=slot.stream_raw[2])
            else:
                prev = torch.cuda.current_stream(self._device)
                torch.cuda.set_stream(slot.stream)
            try:
                # one allocation for every structure output of the batch (rowptr / col per hop, and
                # n_id / bucket ids / perm in distributed mode); exact-size views are cut in
                # _finalize once the meta block has arrived
                arena = torch.empty(lay[3], dtype=torch.int64, device=self._device)
                x = None
                if j.feature_mode:
                    x = torch.empty((lay[4], fdim), dtype=fdtype, device=self._device)
                y = None
                if self._y is not None and not self._y_in_arena:
                    y = torch.empty((bs, self._y.size(-1)), dtype=self._y.dtype, device=self._device)
            finally:
                if _FAST_STREAMS:
                    _set_stream_raw(stream_id=prev[0], device_index=prev[1], device_type=prev[2])
                else:
                    torch.cuda.set_stream(prev)
            if self._y is not None and self._y_in_arena:  # int64 labels live at the tail of the arena
                yo = lay[2]
                y = arena[yo:yo + bs * self._y.size(-1)].view(bs, self._y.size(-1))
            base = arena.data_ptr()
            for h, (ro, co) in enumerate(lay[0]):
                j.out_rowptr[h] = base + 8 * ro
                j.out_col[h] = base + 8 * co
            if self._batch_edges is not None:
                j.out_col_cap[0] = max(int(self._batch_edges[self._next - 1]), 0)
            if self._idx_host is not None:
                # the executor thread copies straight out of the caller's idx memory (an 8 KB
                # cudaMemcpyAsync; pageable sources are staged by the driver); the Session keeps
                # idx alive.  Without the executor the pinned per-slot staging buffer is used.
                if self._executor is not None:
                    j.seeds_host = self._idx_host_ptr + 8 * start if bs else None
                else:
                    if bs:
                        slot.seeds_host[:bs].copy_(self._idx_host[start:stop])
                    j.seeds_host = slot.seeds_host.data_ptr() if bs else None
                j.seeds_dev = slot.seeds.data_ptr()
            else:
                j.seeds_host = None
                j.seeds_dev = self._idx.data_ptr() + 8 * start
            j.batch_size = bs
            j.rng_seed = rng_seed
            j.x_out = x.data_ptr() if x is not None else None
            j.y_out = y.data_ptr() if (y is not None and bs) else None
            if cfg.distributed:
                o, m = lay[1], lay[4]
                j.n_id_out = base + 8 * o
                j.bucket_ids = base + 8 * (o + m)
                j.perm = base + 8 * (o + 2 * m)
            job["arena"], job["y"], job["lay"] = arena, y, lay
            if x is not None:
                job["x"] = x
            elif not cfg.distributed:
                job["x"] = torch.empty((0, fdim), dtype=fdtype, device=self._device)
            if self._executor is not None:
                t = self._lib.spp_executor_submit(self._executor, ctypes.byref(j))
                if not t:
                    raise SalientB200Error("spp_executor_submit failed: " + self._lib.spp_last_error().decode())
                slot.ticket = t
            else:
                check(self._lib.spp_batch_enqueue(ctypes.byref(j)), "spp_batch_enqueue")
                slot.event.record(slot.stream)
                slot.ticket = None
        slot.job = job
        self._pending.append(slot)

    def _labels_for(self, seeds: torch.Tensor, bs: int, sp: int):
        if self._y is None:
            return None
        y = torch.empty((bs, self._y.size(-1)), dtype=self._y.dtype, device=self._device)
        if bs:
            check(self._lib.spp_gather_rows(self._y.data_ptr(), self._y.size(-1) * self._y.element_size(),
                                            seeds.data_ptr(), 1, bs, None, y.data_ptr(), bs, sp), "spp_gather_rows(y)")
        return y

    def _finalize(self, slot: _Slot):
        job = slot.job
        slot.job = None
        cfg = self._config
        L = len(self._sizes)
        if "ready" in job:
            nb, adjs = job["ready"]
        else:
            m = slot.meta_host.tolist()
            if m[META_OVERFLOW]:
                raise SalientB200Error("sampler buffer bound exceeded on the device (SPP_META_OVERFLOW)")
            arena, e_id = job["arena"], _empty_eid(self._device)
            # every exact-size view of the structure part of the arena in ONE split call: per hop
            # [rowptr | slack | col | slack] (the slack pieces are dropped)
            v = arena.split_with_sizes(_arena_cuts(job["lay"], m, L, self._P if cfg.distributed else None))
            adjs = [(v[4 * h], v[4 * h + 2], e_id, (m[h], m[h + 1])) for h in range(L - 1, -1, -1)]  # reversed like fast_sampler.cpp:224
            nb = m[L]
        start, stop = job["range"]
        if not cfg.distributed:
            x = job["x"]
            out = OwnedSample((x[:nb] if x.size(0) > nb else x, job["y"], adjs, (start, stop)))
            if "arena" in job:
                out.owners = (x, job["arena"]) if (self._y_in_arena or job["y"] is None) else (x, job["arena"], job["y"])
        else:
            P = self._P
            b = ProtoDistributedBatch()
            if "ready" in job:
                counts = slot.meta_host[SPP_META_WORDS:].tolist()
                ids = job["bucket_ids"]
                pos = 0
                for p in range(P):
                    b.partition_nids.append(ids[pos:pos + counts[p]])
                    pos += counts[p]
                b.cached_nids = ids[pos:pos + counts[P]]
                b.perm_partition_to_mfg = job["perm"][:nb]
                b.n_id = job["n_id"][:nb]
            else:
                q = 4 * L  # first distributed piece of the split above
                b.n_id = v[q]
                b.partition_nids = list(v[q + 2:q + 2 + P])
                b.cached_nids = v[q + 2 + P]
                b.perm_partition_to_mfg = v[q + 4 + P]
                b.owners = (job["arena"],) + ((job["x"],) if "x" in job else ()) + \
                    (() if (self._y_in_arena or job["y"] is None) else (job["y"],))
            b.adjs = adjs
            b.idx_range = (start, stop)
            b.sliced_cpu_labels = job["y"]
            b.x = (job["x"][:nb] if job["x"].size(0) > nb else job["x"]) if "x" in job else None
            self._attach_host_rows(b)
            out = b
        self._num_consumed += 1
        self._free.append(slot)
        while self._free and self._next < self._num_total:
            self._enqueue()
        if self._num_consumed == self._num_total:
            self._release_slots()
        return out

    def _release_slots(self):
        if self._released:
            return
        self._released = True
        if self._native is not None:
            abandoned = self._native.in_flight > 0  # Session dropped before its last batch
            self._native.release()                  # waits for that work and drops its outputs
            if abandoned:
                for s in self._slots:
                    s.stream.synchronize()
        for s in self._pending:  # abandoned in-flight work (Session dropped early)
            if s.ticket is not None:
                self._lib.spp_executor_wait(self._executor, s.ticket)
            s.stream.synchronize()
            s.job = None
        self._pending.clear()
        _SLOT_POOL.setdefault(self._pool_key, []).extend(self._slots)
        self._slots = []
        self._free = deque()

    def __del__(self):
        try:
            self._release_slots()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _count_remote(self, b: ProtoDistributedBatch):
        # fast_sampler.cpp:1093-1103 (set-up-time statistics for cache_strategy=simulation)
        if self._freq is None:
            self._freq = torch.zeros(self._g.num_nodes, dtype=torch.int64, device=self._device)
        for p in range(self._P):
            if p != self._rank and b.partition_nids[p].numel():
                self._freq.index_add_(0, b.partition_nids[p],
                                      torch.ones_like(b.partition_nids[p]))

    # -- consumer API ---------------------------------------------------------------------------
    @property
    def config(self) -> Config:
        return self._config

    def _slot_done(self, slot: "_Slot") -> bool:
        if slot.ticket is None:
            return slot.event.query()
        r = self._lib.spp_executor_poll(self._executor, slot.ticket)
        if r < 0:
            check(r, "spp_executor_poll")
        return r == 1

    def _slot_wait(self, slot: "_Slot") -> None:
        if slot.ticket is None:
            slot.event.synchronize()
        else:
            check(self._lib.spp_executor_wait(self._executor, slot.ticket), "spp_executor_wait")

    def _get(self, blocking: bool):
        if self._native is not None:
            return self._get_native(blocking)
        if self._num_consumed == self._num_total:
            return None
        slot = self._pending[0]
        if not self._slot_done(slot):
            if not blocking:
                return None
            t0 = time.perf_counter()
            self._slot_wait(slot)
            if os.environ.get("SPP_DEBUG_TIMING") and slot.ticket is not None and self._num_consumed < 2:
                tt = (ctypes.c_double * 3)()
                self._lib.spp_executor_times(self._executor, slot.ticket, tt)
                now = time.clock_gettime(time.CLOCK_MONOTONIC)
                print("[spp] batch %d: queued %.0f us, issuing %.0f us, issue->complete %.0f us (waited %.0f us)" % (
                    self._num_consumed, (tt[1] - tt[0]) * 1e6, (tt[2] - tt[1]) * 1e6, (now - tt[2]) * 1e6,
                    (time.perf_counter() - t0) * 1e6), flush=True)
            self._blocked_dur += datetime.timedelta(microseconds=int((time.perf_counter() - t0) * 1e6))
            self._blocked_occasions += 1
        self._pending.popleft()
        return self._finalize(slot)

    def try_get_batch(self):
        if self._config.distributed:
            raise RuntimeError("try_get_batch called on a distributed Session")
        return self._get(False)

    def blocking_get_batch(self):
        if self._config.distributed:
            raise RuntimeError("blocking_get_batch called on a distributed Session")
        return self._get(True)

    def try_get_batch_distributed(self):
        if not self._config.distributed:
            raise RuntimeError("try_get_batch_distributed called on a non-distributed Session")
        return self._get(False)

    def blocking_get_batch_distributed(self):
        if not self._config.distributed:
            raise RuntimeError("blocking_get_batch_distributed called on a non-distributed Session")
        return self._get(True)

    def _get_native(self, blocking: bool):
        """One batch from the native host path, wrapped in the containers the Python API returns."""
        nat = self._native
        r = nat.get(blocking)
        if r is None:
            return None
        cfg = self._config
        if not cfg.distributed:
            out = OwnedSample(r[:4])
            out.owners, out.y_flat = r[4], r[5]
        else:
            b = ProtoDistributedBatch()
            (b.n_id, b.partition_nids, b.cached_nids, b.perm_partition_to_mfg, b.adjs, b.idx_range,
             b.sliced_cpu_labels, b.x, b.owners, b.y_flat) = r
            self._attach_host_rows(b)
            out = b
        self._num_consumed = nat.consumed
        if self._num_consumed == self._num_total:
            self._release_slots()
        return out

    def _attach_host_rows(self, b: "ProtoDistributedBatch") -> None:
        """``sliced_cpu_features`` (+ the remote-frequency statistics) of a distributed batch."""
        fdim, fdtype = self._feat_shape
        if self._x_cpu_dev is not None:
            # compatibility with gpu_percent < 1 (fast_sampler.cpp:1041-1052,1142-1155): rows of
            # local nodes whose local id lies in the x_cpu tail, in partition_nids[rank] order
            loc = b.partition_nids[self._rank] - self._off[self._rank]
            sel = loc[loc >= self._x_gpu_rows] - self._x_gpu_rows
            b.sliced_cpu_features = serial_index(self._x_cpu_dev, sel)
        else:
            b.sliced_cpu_features = torch.empty((0, fdim), dtype=fdtype, device=self._device)
        if self._config.count_remote_frequency and not self._config.use_cache:
            self._count_remote(b)

    @property
    def total_blocked_dur(self) -> datetime.timedelta:
        if self._native is not None:
            return datetime.timedelta(microseconds=self._native.blocked_us)
        return self._blocked_dur

    @property
    def total_blocked_occasions(self) -> int:
        return self._native.blocked_occasions if self._native is not None else self._blocked_occasions

    num_consumed_batches = property(lambda self: self._num_consumed)
    num_total_batches = property(lambda self: self._num_total)

    @property
    def approx_num_complete_batches(self) -> int:
        if self._native is not None:
            return self._native.complete_count()
        return self._num_consumed + sum(1 for s in self._pending if self._slot_done(s))

    # -- async_slice_tensors (fast_sampler.cpp:720-775): serve other ranks' requests for rows that
    #    the reference keeps on the host; here those rows are in HBM too -----------------------
    def async_slice_tensors(self, ids: List[torch.Tensor], my_rank: int):
        """For every requesting rank i: positions of the requested local ids that are host rows in
        the reference (id >= 0, already offset by the GPU cutoff, fast_trainer/transferers.py:545)
        and of those that are GPU rows (id < 0), plus the host rows themselves (not for
        ``my_rank``).  Issued on a side stream; ``wait_slice_tensors`` joins it."""
        if self._slice_stream is None:
            self._slice_stream = torch.cuda.Stream(self._device)
            self._slice_event = torch.cuda.Event()
        st = self._slice_stream
        st.wait_stream(torch.cuda.current_stream(self._device))
        res = []
        with torch.cuda.stream(st):
            for i, t in enumerate(ids):
                t = t.to(self._device, non_blocking=True)
                host = t >= 0
                cpu_pos = torch.nonzero(host).view(-1)
                gpu_pos = torch.nonzero(~host).view(-1)
                if i != my_rank and self._x_cpu_dev is not None:
                    x_s = serial_index(self._x_cpu_dev, t[cpu_pos])
                elif i != my_rank:
                    fdim, fdtype = self._feat_shape
                    x_s = torch.empty((0, fdim), dtype=fdtype, device=self._device)
                else:
                    x_s = torch.empty(0, dtype=torch.int64, device=self._device)
                res.append([x_s, cpu_pos, gpu_pos])
            self._slice_event.record(st)
        self._slice_pending = res

    def wait_slice_tensors(self):
        if self._slice_pending is None:
            return None
        self._slice_event.synchronize()
        torch.cuda.current_stream(self._device).wait_stream(self._slice_stream)
        self._slice_result, self._slice_pending = self._slice_pending, None
        return None

    def get_slice_tensors(self):
        return self._slice_result

    # -- remote frequency statistics (fast_sampler.cpp:835-880) -------------------------------------
    def reduce_multithreaded_frequency_counts(self):
        if self._freq_reduced:
            return
        if self._freq is not None:
            verts = torch.nonzero(self._freq).view(-1)
            f = self._freq[verts]
            order = torch.argsort(f, descending=True, stable=True)
            self.remote_frequency_tensor = f[order].cpu()
            self.remote_vertices_ordered_by_freq = verts[order].cpu()
        self._freq_reduced = True

    def get_n_most_freq_remote_vertices(self, n: int) -> torch.Tensor:
        self.reduce_multithreaded_frequency_counts()
        return self.remote_vertices_ordered_by_freq[:n].clone()
