"""On-disk partitioned dataset format of the reference.

``DisjointPartFeatReorderedDataset`` mirrors ``driver/dataset.py:127-432``: the same files
(``name.pt num_parts.pt rowptr.pt col.pt split_idx.pt split_idx_parts.pt part_offsets.pt y.pt
meta_info.pt x{rank}.pt``, written by ``reorder_and_save`` :270-369, read by
``from_path_if_exists`` :183-215), the same fields, ``get_RangePartitionBook`` (:371-372) and
``get_num_iterations`` (:374-392), so data prepared by the reference's partitioners loads
straight into the GPU path and data written here loads into the reference.

``reorder_and_save`` relabels vertices so that every partition is a contiguous id range and,
inside a partition, vertices are in descending order of access probability (VIP); the CSR
permutation (relabel rows and columns, re-sort) is done with device-agnostic torch ops -- run it
on the GPU for full-size graphs.  Set-up time only.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Mapping, NamedTuple, Optional

import torch


def csr_permute_symmetric(rowptr: torch.Tensor, col: torch.Tensor, invperm: torch.Tensor):
    """New CSR of the graph with vertex v renamed invperm[v] (driver/dataset.py:290-297:
    relabel rows and columns of the COO form, coalesce, back to CSR)."""
    n = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=rowptr.device), deg)
    key = invperm[rows] * n + invperm[col.to(torch.int64)]
    key = torch.unique(key)  # coalesce(): sorted by (row, col), duplicates merged
    new_rows = torch.div(key, n, rounding_mode="floor")
    new_col = key - new_rows * n
    new_rowptr = torch.zeros(n + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(torch.bincount(new_rows, minlength=n), 0, out=new_rowptr[1:])
    return new_rowptr, new_col


def partition_permutation(partition_labels: torch.Tensor, probability_of_access: Optional[torch.Tensor] = None):
    """``perm`` (new position -> old vertex) and ``invperm`` (old vertex -> new id): ascending
    partition id globally, descending access probability inside a partition
    (driver/dataset.py:302-323).  A stable sort replaces the reference's ``argsort`` so ties keep
    ascending vertex id."""
    labels = partition_labels.to(torch.int64)
    vals = 2.0 * (labels.max() - labels).to(torch.float64)          # :309
    if probability_of_access is not None:
        p = probability_of_access
        if p.dim() == 1:
            vals = vals + p.to(torch.float64)                       # :311-312
        elif p.dim() == 2:
            own = p.to(torch.float64).gather(0, labels.view(1, -1)).view(-1)   # :313-316
            vals = vals + own
        else:
            raise ValueError("probability_of_access must be 1-D or 2-D")
    perm = torch.sort(vals, descending=True, stable=True).indices   # :322
    invperm = torch.empty_like(perm)
    invperm[perm] = torch.arange(perm.numel(), device=perm.device)  # :323
    return perm, invperm


class DisjointPartFeatReorderedDataset(NamedTuple):
    name: str
    rank: int
    num_parts: int
    x: torch.Tensor
    y: torch.Tensor
    rowptr: torch.Tensor
    col: torch.Tensor
    split_idx: Mapping[str, torch.Tensor]
    split_idx_parts: Mapping[int, Mapping[str, torch.Tensor]]
    part_offsets: torch.Tensor
    meta_info: Mapping[str, Any]

    # -- loading (driver/dataset.py:183-215) --------------------------------------------------------
    @classmethod
    def from_path(cls, _path, name, rank):
        path = Path(_path).joinpath(name)
        if not (path.exists() and path.is_dir()):
            raise ValueError("ERROR dataset does not exist at specified path.")
        return cls.from_path_if_exists(_path, name, rank)

    @classmethod
    def from_path_if_exists(cls, path, name, rank):
        path = Path(path).joinpath(name)
        assert path.exists() and path.is_dir()
        fields = [f for f in cls._fields if f not in ("x", "rank")]
        data = {f: torch.load(path.joinpath(f + ".pt"), weights_only=False) for f in fields}
        data["y"] = data["y"].long()
        data["x"] = torch.load(path.joinpath("x" + str(rank) + ".pt"), weights_only=False).to(torch.float16)
        data["rank"] = rank
        assert data["name"] == name
        return cls(**data)

    # -- writing (driver/dataset.py:270-369) ----------------------------------------------------------
    @staticmethod
    def reorder_and_save(name: str, rowptr: torch.Tensor, col: torch.Tensor, x: torch.Tensor, y: torch.Tensor,
                         split_idx: Mapping[str, torch.Tensor], meta_info: Mapping[str, Any],
                         partition_labels: torch.Tensor, probability_of_access: Optional[torch.Tensor], dir) -> Path:
        dev = rowptr.device
        labels = partition_labels.to(dev)
        num_parts = int(labels.max()) + 1
        sizes = torch.bincount(labels, minlength=num_parts)
        perm, invperm = partition_permutation(labels, None if probability_of_access is None
                                              else probability_of_access.to(dev))
        rowptr_p, col_p = csr_permute_symmetric(rowptr, col.to(dev), invperm)
        split_idx_parts = {r: {} for r in range(num_parts)}
        for k, v in split_idx.items():                                # :333-346
            v = v.to(dev)
            pid = labels[v]
            cnt = torch.bincount(pid, minlength=num_parts)
            offs = torch.cat((torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cnt, 0)))
            relabeled = invperm[v][torch.sort(pid, stable=True).indices]
            for r in range(num_parts):
                split_idx_parts[r][k] = relabeled[int(offs[r]):int(offs[r + 1])].cpu()
        x_p, y_p = x.to(dev)[perm], y.to(dev)[perm]                   # :348-349
        part_offsets = torch.cat((torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(sizes, 0))).cpu()
        prefix = Path(dir) / f"metis-reordered-k{num_parts}" / name   # :359
        prefix.mkdir(parents=True, exist_ok=False)
        torch.save(num_parts, prefix / "num_parts.pt")
        torch.save(rowptr_p.cpu(), prefix / "rowptr.pt")
        torch.save(col_p.cpu(), prefix / "col.pt")
        torch.save(dict(), prefix / "split_idx.pt")                   # :331,364 (the reference saves an empty dict)
        torch.save(split_idx_parts, prefix / "split_idx_parts.pt")
        torch.save(part_offsets, prefix / "part_offsets.pt")
        torch.save(y_p.cpu(), prefix / "y.pt")
        torch.save(dict(meta_info), prefix / "meta_info.pt")
        torch.save(name, prefix / "name.pt")
        for r in range(num_parts):
            lo, hi = int(part_offsets[r]), int(part_offsets[r + 1])
            torch.save(x_p[lo:hi].to(torch.float16).cpu().clone(), prefix / f"x{r}.pt")
        return prefix

    # -- accessors ---------------------------------------------------------------------------------------
    def get_RangePartitionBook(self):
        from .fast_sampler import RangePartitionBook
        return RangePartitionBook(self.rank, self.num_parts, self.part_offsets)

    def get_num_iterations(self, minibatch_size: int):
        """Iterations per split such that every rank runs the same number (:374-392)."""
        out = {}
        for split in ("train", "valid", "test"):
            total = sum(self.split_idx_parts[i][split].numel() for i in range(self.num_parts))
            out[split] = int(max(1, total // minibatch_size))
        return out

    @property
    def num_nodes(self):
        return self.rowptr.numel() - 1

    @property
    def num_features(self):
        return self.x.size(1)

    @property
    def num_classes(self):
        return int(self.meta_info["num classes"] if "num classes" in self.meta_info else self.meta_info["num_classes"])
