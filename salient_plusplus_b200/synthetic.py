"""Seeded synthetic graphs, features, range partitions and cache lists of the BASELINE shapes.

The reference trains on OGB datasets (driver/dataset.py:68-82 symmetrises the edge list and
builds a CSR with int64 ``rowptr``/``col`` and row-major fp16 features); there is no network
here, so every test and benchmark uses graphs from this generator.  Works on CPU (tests) and
on the GPU (bench-scale graphs are generated directly in HBM).

Generator: Chung-Lu style power-law graph.  Node weight ``w_i = (i + head_offset)^(-1/(gamma-1))``
with ``gamma = 2.5`` and ``head_offset = 100`` (flattens the head so the maximum degree stays in
the 1e4 range for a products-sized graph); both endpoints of each directed edge are drawn
proportionally to ``w``; ids are randomly relabelled; the edge list is symmetrised and
de-duplicated (self loops kept out) like ``to_undirected``.  State (gamma, head_offset, seed)
with every reported number.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

SHAPES = {
    # name: (num_nodes, directed_edges, feature_dim, feature dtype)   -- BASELINE.json configs
    "arxiv": (169_343, 1_166_243, 128, torch.float32),
    "products": (2_449_029, 61_859_140, 100, torch.float16),
    "papers100M": (111_059_956, 1_615_685_872, 128, torch.float16),
    "mag240m": (244_160_499, 1_700_000_000, 768, torch.float16),
}


def powerlaw_graph(num_nodes: int, num_directed_edges: int, *, gamma: float = 2.5,
                   head_offset: float = 100.0, seed: int = 1, device="cpu",
                   symmetrize: bool = True, chunk: int = 1 << 27, locality: float = 0.0,
                   locality_parts: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(rowptr int64[N+1], col int64[nnz])`` with ascending columns in each row.

    ``locality`` > 0 plants partition structure: with that probability the destination of a
    directed edge is folded into the source's block of the ``locality_parts`` equal id ranges
    (what a METIS partition followed by the reference's contiguous relabelling gives,
    driver/dataset.py:309-323).  0 (default) is the locality-free worst case for range-partitioned
    features."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    alpha = 1.0 / (gamma - 1.0)
    w = (torch.arange(num_nodes, device=dev, dtype=torch.float64) + head_offset).pow_(-alpha)
    cdf = torch.cumsum(w, 0)
    total = cdf[-1].item()
    del w
    relabel = torch.randperm(num_nodes, generator=g, device=dev)
    keys: List[torch.Tensor] = []
    done = 0
    while done < num_directed_edges:
        m = min(chunk, num_directed_edges - done)
        u = torch.rand(m, generator=g, device=dev, dtype=torch.float64).mul_(total)
        src = torch.searchsorted(cdf, u).clamp_(max=num_nodes - 1)
        u = torch.rand(m, generator=g, device=dev, dtype=torch.float64).mul_(total)
        dst = torch.searchsorted(cdf, u).clamp_(max=num_nodes - 1)
        del u
        src = relabel[src]
        dst = relabel[dst]
        if locality > 0.0:
            bsz = (num_nodes + locality_parts - 1) // locality_parts
            fold = torch.rand(m, generator=g, device=dev) < locality
            blk = torch.div(src, bsz, rounding_mode="floor")
            folded = (blk * bsz + dst % bsz).clamp_(max=num_nodes - 1)
            dst = torch.where(fold, folded, dst)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        keys.append(src * num_nodes + dst)
        if symmetrize:
            keys.append(dst * num_nodes + src)
        done += m
    del cdf, relabel
    key = torch.cat(keys) if len(keys) > 1 else keys[0]
    del keys
    key = torch.unique(key)  # sorted + de-duplicated
    row = torch.div(key, num_nodes, rounding_mode="floor")
    col = key - row * num_nodes
    del key
    counts = torch.bincount(row, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr, col.contiguous()


def features(num_nodes: int, dim: int, dtype=torch.float16, *, seed: int = 2, device="cpu",
             chunk_rows: int = 1 << 22) -> torch.Tensor:
    """Row-major ``[N, dim]`` random features (``randn`` cast to ``dtype``), generated in row
    chunks so a papers100M-sized table never needs an fp32 staging copy of the whole thing."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((num_nodes, dim), dtype=dtype, device=dev)
    for s in range(0, num_nodes, chunk_rows):
        e = min(num_nodes, s + chunk_rows)
        out[s:e] = torch.randn((e - s, dim), generator=g, device=dev, dtype=torch.float32).to(dtype)
    return out


def _id_pattern(ids: torch.Tensor, dim: int, dtype: torch.dtype) -> torch.Tensor:
    """Bit pattern of feature element (id, j) as a pure function of the GLOBAL vertex id:
    32-bit wrap-around integer hash, masked so that the pattern is a finite positive number
    (fp16/bf16: < 2, fp32: < 2).  int16 / int32 tensor [len(ids), dim]."""
    es = torch.empty(0, dtype=dtype).element_size()
    if es not in (2, 4):
        raise ValueError("features_by_id supports 2- and 4-byte element types")
    i = ids.to(torch.int32).view(-1, 1)
    j = torch.arange(dim, device=ids.device, dtype=torch.int32).view(1, -1)
    v = i * 1103515245 + j * 40503 + 12345          # wraps in int32 (two's complement) on every backend
    v = v ^ (v >> 13)
    v = v * 1664525 + 1013904223
    v = v ^ (v >> 11)
    if es == 2:
        return (v & 0x3BFF).to(torch.int16)
    return v & 0x3F7FFFFF


def features_by_id(lo: int, hi: int, dim: int, dtype=torch.float16, *, device="cpu",
                   chunk_rows: int = 1 << 21) -> torch.Tensor:
    """Rows ``[lo, hi)`` of the synthetic feature matrix whose element (i, j) is a pure function of
    the global vertex id i -- any rank can regenerate the rows of any vertex, so a gathered batch
    can be verified (``x == expected_features(n_id)``) without holding the other partitions."""
    dev = torch.device(device)
    out = torch.empty((hi - lo, dim), dtype=dtype, device=dev)
    it = torch.int16 if out.element_size() == 2 else torch.int32
    ov = out.view(it)
    for s in range(lo, hi, chunk_rows):
        e = min(hi, s + chunk_rows)
        ov[s - lo:e - lo] = _id_pattern(torch.arange(s, e, device=dev), dim, dtype)
    return out


def expected_features(ids: torch.Tensor, dim: int, dtype=torch.float16) -> torch.Tensor:
    """``features_by_id`` evaluated at arbitrary vertex ids (same device as ``ids``)."""
    pat = _id_pattern(ids, dim, dtype)
    return pat.view(dtype)


def labels_by_id(ids: torch.Tensor, num_classes: int = 47) -> torch.Tensor:
    """int64 [n, 1] labels as a pure function of the vertex id."""
    return ((ids.to(torch.int64) * 2654435761 + 12345) % 1000003 % num_classes).view(-1, 1)


def labels(num_nodes: int, num_classes: int = 47, *, seed: int = 3, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    return torch.randint(0, num_classes, (num_nodes, 1), generator=g, device=device, dtype=torch.int64)


def seeds(num_nodes: int, count: int, *, seed: int = 7, device="cpu", lo: int = 0,
          hi: Optional[int] = None) -> torch.Tensor:
    """``count`` distinct seed ids from ``[lo, hi)`` (a prefix of a seeded permutation)."""
    hi = num_nodes if hi is None else hi
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    return (torch.randperm(hi - lo, generator=g, device=device)[:count] + lo).to(torch.int64)


def equal_partition_offsets(num_nodes: int, parts: int) -> torch.Tensor:
    """Contiguous, near-equal vertex ranges (driver/dataset.py:349-353 produces contiguous
    ranges after reordering; sizes there come from METIS, here they are equal)."""
    base, rem = divmod(num_nodes, parts)
    sizes = [base + (1 if p < rem else 0) for p in range(parts)]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    return torch.tensor(off, dtype=torch.int64)


def degree_cache_vertices(rowptr: torch.Tensor, offsets: torch.Tensor, rank: int,
                          rows: int) -> torch.Tensor:
    """Stand-in for the VIP ranking (driver/drivers/ddp.py:417-570): the ``rows`` highest-degree
    *remote* vertices, laid out owner-major and score-descending inside each owner exactly like
    ``cached_vertices = cat(buckets)`` (ddp.py:504-509,555).  Degree is the 1-hop VIP proxy
    (caching/vip.py:294-330); the analytic multi-hop VIP model is a setup-time component
    (SURVEY.md section 8f)."""
    deg = (rowptr[1:] - rowptr[:-1]).clone()
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    deg[lo:hi] = -1
    rows = min(rows, deg.numel() - (hi - lo))
    # stable descending order => deterministic ties
    order = torch.sort(deg, descending=True, stable=True).indices[:rows]
    owner = torch.searchsorted(offsets.to(order.device), order, right=True) - 1
    out = []
    for p in range(offsets.numel() - 1):
        if p == rank:
            continue
        out.append(order[owner == p])
    return torch.cat(out) if out else order[:0]


@dataclass
class SyntheticDataset:
    name: str
    rowptr: torch.Tensor
    col: torch.Tensor
    x: torch.Tensor
    y: torch.Tensor
    num_nodes: int

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


def make_dataset(name: str, *, scale: float = 1.0, device="cpu", seed: int = 1,
                 feature_dtype: Optional[torch.dtype] = None, with_features: bool = True) -> SyntheticDataset:
    """A graph of one of the BASELINE shapes; ``scale < 1`` shrinks nodes and edges together
    (same average degree) for tests."""
    n, e, f, dt = SHAPES[name]
    n = max(16, int(n * scale))
    e = max(32, int(e * scale))
    dt = feature_dtype or dt
    rowptr, col = powerlaw_graph(n, e, seed=seed, device=device)
    x = features(n, f, dt, seed=seed + 1, device=device) if with_features else torch.empty((0, f), dtype=dt)
    y = labels(n, seed=seed + 2, device=device)
    return SyntheticDataset(name, rowptr, col, x, y, n)
