"""Shared helpers of the test-suite (graphs, comparisons)."""
import numpy as np
import torch

from salient_plusplus_b200 import synthetic as S


def small_graph(n=3000, e=40000, seed=1):
    rowptr, col = S.powerlaw_graph(n, e, seed=seed, head_offset=5.0)
    return rowptr, col


def bounded_degree_graph(n=2000, max_deg=5, seed=3):
    """Every row has between 0 and max_deg distinct neighbours."""
    rng = np.random.default_rng(seed)
    rowptr = [0]
    col = []
    for _ in range(n):
        d = int(rng.integers(0, max_deg + 1))
        nb = np.sort(rng.choice(n, size=d, replace=False)) if d else np.empty(0, dtype=np.int64)
        col.extend(nb.tolist())
        rowptr.append(len(col))
    return torch.tensor(rowptr, dtype=torch.int64), torch.tensor(col, dtype=torch.int64)


def star_graph(num_leaves, num_centers):
    """Leaves 0..D-1 (empty rows); centres D..D+T-1 each adjacent to every leaf."""
    D, T = num_leaves, num_centers
    rowptr = np.zeros(D + T + 1, dtype=np.int64)
    rowptr[D + 1:] = np.arange(1, T + 1, dtype=np.int64) * D
    col = np.tile(np.arange(D, dtype=np.int64), T)
    return torch.from_numpy(rowptr), torch.from_numpy(col)


def adjs_equal(got, want):
    """got: list of (rowptr, col, e_id, (T,S)) torch tensors; want: numpy tuples from the oracle."""
    if len(got) != len(want):
        return False
    for a, b in zip(got, want):
        if tuple(int(v) for v in a[3]) != tuple(int(v) for v in b[3]):
            return False
        if not np.array_equal(a[0].cpu().numpy(), b[0]):
            return False
        if not np.array_equal(a[1].cpu().numpy(), b[1]):
            return False
        if a[2].numel() != 0:
            return False
    return True
