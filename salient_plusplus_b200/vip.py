"""VIP cache policy on the GPU.

``vip_probabilities``   = ``get_frequency_tensors_fast`` (driver/drivers/ddp.py:134-239): the
                          analytic vertex-inclusion probabilities of one rank's mini-batches,
                          fp64, Taylor form ``1 - exp(-sum)`` exactly as the driver computes it.
``select_cache_vertices`` = the ranking part of ``create_vip_cache`` (ddp.py:417-446,504-509,555):
                          top ``int(N / P * pct / 100)`` *remote* vertices with non-zero VIP, laid
                          out owner-major, VIP-descending inside each owner.  The reference sorts
                          with an unstable ``argsort``; a stable one is used here so the layout is
                          deterministic (ties keep ascending vertex id).
``create_vip_cache``    = the rest of ``create_vip_cache`` (:510-560): the rows are pulled from
                          their owners by the P2P gather kernel instead of three blocking
                          all_to_alls, and wrapped in a ``Cache``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .fast_sampler import Cache, _DeviceGraph, _device, _stream_ptr, feature_table, make_feature_map


def vip_probabilities(rowptr: torch.Tensor, col: torch.Tensor, train_idx: torch.Tensor, batch_size: int,
                      fanouts: Sequence[int], exact: bool = False) -> torch.Tensor:
    """fp64 device tensor [N]: probability that a vertex appears in a mini-batch whose seeds are
    ``batch_size`` uniform draws from ``train_idx`` (ddp.py:157-234).  ``exact=True`` uses the
    log-product form of ``caching/vip.py:vip_analytical`` (:166-172) instead of the driver's
    first-order form."""
    dev = _device()
    g = _DeviceGraph.get(rowptr, col)
    n = g.num_nodes
    p = torch.zeros(n, dtype=torch.float64, device=dev)
    ids = train_idx.to(device=dev, dtype=torch.int64)
    p[ids] = float(batch_size) / float(ids.numel())          # ddp.py:160
    not_total = torch.ones(n, dtype=torch.float64, device=dev)
    scratch = torch.empty(n, dtype=torch.float64, device=dev)
    nxt = torch.empty_like(p)
    L = _lib.load()
    for fanout in fanouts:                                     # ddp.py:193 (given order)
        check(L.spp_vip_hop(ctypes.byref(g.c), float(fanout), int(bool(exact)), p.data_ptr(), nxt.data_ptr(), not_total.data_ptr(),
                            scratch.data_ptr(), _stream_ptr()), "spp_vip_hop")
        p, nxt = nxt, p
    return 1.0 - not_total                                     # ddp.py:229-233


def select_cache_vertices(vip: torch.Tensor, partition_offsets: torch.Tensor, rank: int, num_to_cache: int,
                          local_parts: Sequence[int] = ()) -> torch.Tensor:
    """Global ids to replicate on ``rank``: owner-major, VIP-descending inside each owner.
    ``local_parts``: further partitions resident on the same GPU (never cached either)."""
    dev = vip.device
    off = partition_offsets.to(dev)
    n = vip.numel()
    score = vip.clone()
    for p in sorted(set([int(rank)] + [int(q) for q in local_parts])):
        score[int(off[p]):int(off[p + 1])] = 0.0              # local vertices are never cached (ddp.py:433-434,512)
    k = min(int(num_to_cache), int(torch.count_nonzero(score).item()))   # ddp.py:436-437
    if k <= 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    order = torch.sort(score, descending=True, stable=True).indices[:k]
    owner = torch.searchsorted(off, order, right=True) - 1
    # stable grouping by owner keeps the VIP-descending order inside each bucket (ddp.py:504-509)
    grp = torch.sort(owner, stable=True).indices
    return order[grp]


def create_vip_cache(rowptr: torch.Tensor, col: torch.Tensor, train_idx_local: torch.Tensor, batch_size: int,
                     fanouts: Sequence[int], partition_offsets: torch.Tensor, rank: int, cache_pct: float,
                     local_features: torch.Tensor, partition_tables: Optional[Sequence[Optional[torch.Tensor]]] = None,
                     peer_table_ptrs: Optional[Sequence[int]] = None, vip: Optional[torch.Tensor] = None,
                     local_parts: Sequence[int] = ()) -> Cache:
    """Builds the replicated cache of ``rank``.  The feature rows come straight out of the owners'
    partitions (local tensors in ``partition_tables`` and/or IPC-mapped ``peer_table_ptrs``; with a
    process group of matching size they are exchanged automatically)."""
    dev = _device()
    P = partition_offsets.numel() - 1
    n = rowptr.numel() - 1
    if vip is None:
        vip = vip_probabilities(rowptr, col, train_idx_local, batch_size, fanouts)
    # `vip` may be any per-vertex score (e.g. degrees for cache_strategy == "degree", ddp.py:487-495)
    num = int(n / P * (cache_pct / 100.0))                     # ddp.py:421
    cv = select_cache_vertices(vip, partition_offsets, rank, num, local_parts)
    ltab = feature_table(local_features)
    tables = [None] * P
    ptrs = [0] * P
    if partition_tables is not None:
        for p, t in enumerate(partition_tables):
            if t is not None and p != rank:
                tables[p] = feature_table(t).storage
    if peer_table_ptrs is not None:
        ptrs = [int(v or 0) for v in peer_table_ptrs]
    tables[rank] = ltab.storage
    ptrs[rank] = 0
    off = [int(v) for v in partition_offsets.tolist()]
    if not all(tables[p] is not None or ptrs[p] or off[p + 1] == off[p] for p in range(P)):
        from . import peer
        got = peer.exchange_partition_tables(ltab.storage, rank, P)
        if got is None:
            raise _lib.SalientB200Error("create_vip_cache: peer partitions are not reachable")
        ptrs = got
        ptrs[rank] = 0
    fm = make_feature_map(off, rank, tables, None, None, ptrs, ltab.pitch, 0, local_parts=local_parts)
    feats = torch.empty((cv.numel(), ltab.dim), dtype=ltab.dtype, device=dev)
    if cv.numel():
        check(_lib.load().spp_gather_partitioned(ctypes.byref(fm), ltab.row_bytes, cv.data_ptr(), 1, cv.numel(), None,
                                                 None, feats.data_ptr(), cv.numel(), None, _stream_ptr()),
              "spp_gather_partitioned")
    return Cache(rank, P, cv, feats)
