"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle: ctypes bindings for ``oracle/salient_oracle.c`` (sampling,
feature slicing) plus numpy restatements of the reference's partition-book, cache and
distributed-binning arithmetic.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; nothing
under ``salient_plusplus_b200/`` does.

Parity status: PINNED against the compiled, unmodified reference (``oracle/_ref``), see
``tests/test_oracle_vs_ref.py`` and the fixtures in ``tests/golden/``.

All citations are relative to ``/root/reference/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "salient_oracle.c")
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libsalient_oracle.so")

RNG_REFERENCE = 0  # std::mt19937 + the reference's (biased) Floyd variant
RNG_COUNTER = 1    # counter-based generator shared with the CUDA kernels, exact Floyd

_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (a second or two)."""
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", _LIB, _SRC])
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, i64, i32, u64, u32 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                  ctypes.c_uint64, ctypes.c_uint32)
        L.spo_state_new.restype = vp
        L.spo_state_new.argtypes = [vp, i64, ctypes.c_int, u64]
        L.spo_state_free.argtypes = [vp]
        L.spo_hop.restype = i64
        L.spo_hop.argtypes = [vp, vp, vp, vp, i32, ctypes.c_int]
        L.spo_state_num_nodes.restype = i64
        L.spo_state_num_nodes.argtypes = [vp]
        L.spo_state_num_adjs.restype = ctypes.c_int
        L.spo_state_num_adjs.argtypes = [vp]
        L.spo_state_adj_sizes.argtypes = [vp, ctypes.c_int, vp, vp, vp]
        L.spo_state_copy_adj.argtypes = [vp, ctypes.c_int, vp, vp]
        L.spo_state_copy_nids.argtypes = [vp, vp]
        L.spo_serial_index.argtypes = [vp, i64, vp, i64, i64, vp]
        L.spo_rand64.restype = u64
        L.spo_rand64.argtypes = [u64, u32, u64, u32]
        L.spo_minibatch.restype = i64
        L.spo_minibatch.argtypes = [vp, vp, vp, i64, vp, ctypes.c_int, u64, vp, i64, vp, i64]
        L.spo_mt_seed.argtypes = [vp, u32]
        L.spo_mt_next.restype = u32
        L.spo_mt_next.argtypes = [vp]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


Adj = Tuple[np.ndarray, np.ndarray, np.ndarray, Tuple[int, int]]


class SamplerState:
    """The per-batch state of ``multilayer_sample`` (fast_sampler/fast_sampler.cpp:191-227)."""

    def __init__(self, seeds, rng_mode: int = RNG_REFERENCE, rng_seed: int = 5489):
        self._seeds = _i64(seeds)
        self._h = lib().spo_state_new(_p(self._seeds), self._seeds.size, rng_mode, rng_seed)

    def hop(self, rowptr: np.ndarray, col: np.ndarray, num_neighbors: int, replace: bool = False) -> int:
        rowptr = _i64(rowptr)
        col = np.ascontiguousarray(col)
        c64 = col if col.dtype == np.int64 else None
        c32 = col if col.dtype == np.int32 else None
        assert (c64 is None) != (c32 is None), "col must be int64 or int32"
        return lib().spo_hop(self._h, _p(rowptr), _p(c64), _p(c32), int(num_neighbors), int(bool(replace)))

    @property
    def n_id(self) -> np.ndarray:
        n = lib().spo_state_num_nodes(self._h)
        out = np.empty(n, dtype=np.int64)
        lib().spo_state_copy_nids(self._h, _p(out))
        return out

    def adj(self, i: int) -> Adj:
        T, E, S = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lib().spo_state_adj_sizes(self._h, i, ctypes.byref(T), ctypes.byref(E), ctypes.byref(S))
        rp = np.empty(T.value + 1, dtype=np.int64)
        cl = np.empty(E.value, dtype=np.int64)
        lib().spo_state_copy_adj(self._h, i, _p(rp), _p(cl))
        return rp, cl, np.empty(0, dtype=np.int64), (T.value, S.value)

    @property
    def num_adjs(self) -> int:
        return lib().spo_state_num_adjs(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().spo_state_free(self._h)
            self._h = None


def sample_adj(rowptr, col, idx, num_neighbors: int, replace: bool = False, *,
               rng_mode: int = RNG_REFERENCE, rng_seed: int = 5489):
    """``fast_sampler.sample_adj`` (fast_sampler/sample_cpu.hpp:154-165):
    returns ``(rowptr, col, n_id int32, e_id)``."""
    st = SamplerState(idx, rng_mode, rng_seed)
    st.hop(rowptr, col, num_neighbors, replace)
    rp, cl, e_id, _ = st.adj(0)
    return rp, cl, st.n_id.astype(np.int32), e_id


def multilayer_sample(idx, sizes: Sequence[int], rowptr, col, *, rng_mode: int = RNG_REFERENCE,
                      rng_seed: int = 5489) -> Tuple[np.ndarray, List[Adj]]:
    """``fast_sampler.multilayer_sample`` (fast_sampler/fast_sampler.cpp:191-236): hops applied
    seed-side first, adjacency list reversed so the outermost hop comes first (:224)."""
    st = SamplerState(idx, rng_mode, rng_seed)
    for k in sizes:
        st.hop(rowptr, col, int(k), False)
    adjs = [st.adj(i) for i in range(st.num_adjs)][::-1]
    return st.n_id, adjs


def session_rng_seed(stop: int) -> int:
    """fast_sampler/fast_sampler.cpp:994: ``gen.seed(pair.second * 17 + 5)`` (int32 arithmetic,
    converted to the 32-bit result_type of std::mt19937)."""
    return (int(stop) * 17 + 5) & 0xFFFFFFFF


def serial_index(x: np.ndarray, idx, n: Optional[int] = None) -> np.ndarray:
    """``fast_sampler.serial_index`` (fast_sampler/fast_sampler.cpp:238-279)."""
    x = np.ascontiguousarray(x)
    x2 = x.reshape(x.shape[0], -1)
    idx = _i64(idx)
    n = idx.size if n is None else int(n)
    out = np.zeros((n, x2.shape[1]), dtype=x.dtype)
    lib().spo_serial_index(_p(x2), x2.shape[1] * x.dtype.itemsize, _p(idx), idx.size, n, _p(out))
    return out


def batch_ranges(n: int, batch_size: int, skip_nonfull_batch: bool = False,
                 force_exact_num_batches: bool = False, exact_num_batches: int = 0) -> List[Tuple[int, int]]:
    """Batch ranges the Session enqueues (fast_sampler/fast_sampler.cpp:587-627)."""
    out: List[Tuple[int, int]] = []
    if force_exact_num_batches:
        B = int(exact_num_batches)
        avg = n // B - 1                                  # :595 (sic: one less than the mean)
        sizes = [avg] * B
        rem = n - avg * B
        while rem > 0:                                    # :602-608 round-robin remainder
            for i in range(B):
                if rem <= 0:
                    break
                sizes[i] += 1
                rem -= 1
        s = 0
        for b in sizes:
            out.append((s, s + b))
            s += b
    else:
        for i in range(0, n, batch_size):                 # :618-626
            this = min(n, i + batch_size) - i
            if skip_nonfull_batch and this < batch_size:
                continue
            out.append((i, i + this))
    return out


# ---------------------------------------------------------------------------------------------
# RangePartitionBook (fast_sampler/range_partition_book.cpp:85-112)
# ---------------------------------------------------------------------------------------------
def nid2partid(offsets, nids) -> np.ndarray:
    """``searchsorted(offsets, nids, right=True) - 1`` (:98-100)."""
    return np.searchsorted(_i64(offsets), _i64(nids), side="right").astype(np.int64) - 1


def nid2localnid(offsets, nids, partition_idx: int) -> np.ndarray:
    """``nids - offsets[partition_idx]`` (:89-96)."""
    return _i64(nids) - _i64(offsets)[partition_idx]


def nid_is_local(offsets, rank: int, nids) -> np.ndarray:
    """(:105-107)"""
    o = _i64(offsets)
    n = _i64(nids)
    return (n >= o[rank]) & (n < o[rank + 1])


def partid2nids(offsets, partition_idx: int) -> np.ndarray:
    """(:109-112)"""
    o = _i64(offsets)
    return np.arange(o[partition_idx], o[partition_idx + 1], dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# Cache (fast_sampler/range_partition_book.cpp:116-195): dense id -> cache-row map.
# ---------------------------------------------------------------------------------------------
class Cache:
    def __init__(self, cached_vertices, num_nodes: int):
        cv = _i64(cached_vertices)
        self.cached_vertices = cv
        self.map = np.zeros(num_nodes, dtype=np.int32)      # :152 (uninitialised there)
        self.isin = np.zeros(num_nodes, dtype=bool)         # :153
        for i, v in enumerate(cv):                          # :154-158 later duplicates overwrite
            self.map[v] = i
            self.isin[v] = True

    def nid_is_cached(self, nids) -> np.ndarray:            # :161-183
        return self.isin[_i64(nids)]

    def nid2cachenid(self, nids) -> np.ndarray:             # :185-195
        return self.map[_i64(nids)].astype(np.int64)


# ---------------------------------------------------------------------------------------------
# Distributed binning (fast_sampler/fast_sampler.cpp:1017-1262)
# ---------------------------------------------------------------------------------------------
class ProtoDistributedBatch:
    __slots__ = ("partition_nids", "sliced_cpu_features", "sliced_cpu_labels", "cached_nids",
                 "perm_partition_to_mfg", "adjs", "idx_range", "local_on_cpu")


def distributed_binning(n_id, offsets, rank: int, world_size: int, x_gpu_rows: int,
                        use_cache: bool, cache: Optional[Cache] = None):
    """Returns ``(partition_nids[P] (global ids, n_id order), cached_nids (cache rows),
    perm_partition_to_mfg, local_on_cpu)`` such that
    ``cat(partition_nids + [cached_global])[perm] == n_id``.

    no-cache branch: fast_sampler.cpp:1031-1107; cache branch: :1108-1260.
    ``local_on_cpu`` are the host-resident local rows ``local_id - x_gpu_rows`` (:1041-1051,
    :1142-1155) that the reference slices from ``x_cpu``."""
    n_id = _i64(n_id)
    off = _i64(offsets)
    local_bool = nid_is_local(off, rank, n_id)
    local = n_id[local_bool]
    lloc = nid2localnid(off, local, rank)
    local_on_cpu = (lloc[lloc >= x_gpu_rows] - x_gpu_rows).astype(np.int64)
    perm = np.empty(n_id.size, dtype=np.int64)
    if not use_cache:
        machine = nid2partid(off, n_id)                                  # :1063
        counts = np.bincount(machine, minlength=world_size)              # :1065
        starts = np.concatenate([[0], np.cumsum(counts)])                # :1075-1078
        partition_nids = []
        for m in range(world_size):
            sel = np.nonzero(machine == m)[0]
            partition_nids.append(n_id[sel])                             # :1082-1087 (stable)
            perm[sel] = starts[m] + np.arange(sel.size)
        cached_nids = np.empty(0, dtype=np.int64)                        # :1106
        return partition_nids, cached_nids, perm, local_on_cpu
    assert cache is not None
    local_indices = np.nonzero(local_bool)[0]                            # :1130
    remote_indices = np.nonzero(~local_bool)[0]                          # :1131
    remote = n_id[remote_indices]                                        # :1160
    cached_bool = cache.nid_is_cached(remote)                            # :1170
    cached = remote[cached_bool]                                         # :1182
    remote_nc = remote[~cached_bool]                                     # :1184
    cached_indices = remote_indices[cached_bool]                         # :1195
    remote_nc_indices = remote_indices[~cached_bool]                     # :1196
    pid = nid2partid(off, remote_nc)                                     # :1202
    partition_nids, part_idx = [], []
    for r in range(world_size):                                          # :1214-1235
        if r == rank:
            partition_nids.append(local)
            part_idx.append(local_indices)
        else:
            sel = pid == r
            partition_nids.append(remote_nc[sel])
            part_idx.append(remote_nc_indices[sel])
    flipped = np.concatenate(part_idx + [cached_indices])                # :1241-1247
    perm[flipped] = np.arange(n_id.size, dtype=np.int64)                 # :1249-1252
    cached_nids = cache.nid2cachenid(cached)                             # :1256
    return partition_nids, cached_nids, perm, local_on_cpu


# ---------------------------------------------------------------------------------------------
# VIP analytic model, fp64 numpy restatement.
#   exact=True : caching/vip.py:123-180 (vip_analytical, log-product form).  PINNED: the
#                reference's own function was run here (torch_scatter.segment_csr replaced by a
#                stand-in with its published semantics) -> tests/golden/vip.npz, checked in
#                tests/test_oracle_golden.py at the reference's fp32 precision.
#   exact=False: driver/drivers/ddp.py:134-239 (get_frequency_tensors_fast, first-order form the
#                driver actually uses).  PINNED: the function's own source text was executed here
#                (tests/golden/make_golden_vip_driver.py: stand-ins for torch_scatter.segment_csr,
#                dist.get_rank and the CUDA stream plumbing only, asserts stripped like the
#                reference's PYTHONOPTIMIZE=1 launch) -> tests/golden/vip_driver.npz (fp64).
# ---------------------------------------------------------------------------------------------
def vip_probabilities(rowptr, col, train_idx, batch_size: int, fanouts, exact: bool = False) -> np.ndarray:
    rowptr, col = _i64(rowptr), _i64(col)
    n = rowptr.size - 1
    deg = (rowptr[1:] - rowptr[:-1]).astype(np.float64)            # ddp.py:153-154
    p = np.zeros(n, dtype=np.float64)
    p[_i64(train_idx)] = float(batch_size) / float(len(train_idx))  # :160
    not_total = np.ones(n, dtype=np.float64)
    nonempty = rowptr[1:] > rowptr[:-1]
    for fanout in fanouts:                                          # :193
        with np.errstate(divide="ignore"):
            w = np.minimum(1.0, float(fanout) / deg)                # :221
        t = w * p
        if exact:                                                   # vip.py:166-172
            with np.errstate(divide="ignore"):
                t = -np.log(1.0 - t)
        wp = t[col]
        s = np.zeros(n, dtype=np.float64)
        if col.size:
            starts = np.minimum(rowptr[:-1], col.size - 1)
            s = np.add.reduceat(wp, starts)                         # segment_csr(..., reduce='add') :222
            s[~nonempty] = 0.0
        p = 1.0 - np.exp(-s)                                        # :224
        not_total *= (1.0 - p)                                      # :229-231
    return 1.0 - not_total                                          # :233


def select_cache_vertices(vip, offsets, rank: int, num_to_cache: int) -> np.ndarray:
    """ddp.py:433-439 (with a stable sort) + owner-major bucketing :504-509,555."""
    off = _i64(offsets)
    score = np.array(vip, dtype=np.float64, copy=True)
    score[off[rank]:off[rank + 1]] = 0.0
    k = min(int(num_to_cache), int(np.count_nonzero(score)))
    order = np.argsort(-score, kind="stable")[:k]
    owner = nid2partid(off, order)
    return np.concatenate([order[owner == p] for p in range(off.size - 1)]) if k else np.empty(0, dtype=np.int64)
