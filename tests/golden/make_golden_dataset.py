"""Generates tests/golden/dataset_reorder.npz by running the reference's own
``DisjointPartFeatReorderedDataset.reorder_and_save`` (driver/dataset.py:270-369, imported
read-only from /root/reference) on a small seeded graph.  The reference module needs
``torch_sparse``, ``ogb`` and its compiled ``fast_sampler``, none of which is installed here, so
stand-ins are injected before the import: ``torch_sparse.SparseTensor`` restated with its
published semantics for the three calls the writer makes (``coo()``, ``coalesce()`` = sort by
(row, col) and merge duplicates, ``csr()``), an empty ``ogb.nodeproppred`` and the compiled
reference module (oracle/_ref) as ``fast_sampler``.

    python tests/golden/make_golden_dataset.py
"""
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from salient_plusplus_b200 import synthetic as S  # noqa: E402


class SparseTensor:
    def __init__(self, row=None, rowptr=None, col=None, value=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        if row is None:
            n = rowptr.numel() - 1
            row = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
        self.row, self.col, self.value = row, col, value
        self.n = int(max(int(row.max()) if row.numel() else -1, int(col.max()) if col.numel() else -1)) + 1 \
            if sparse_sizes is None else int(sparse_sizes[0])

    def coo(self):
        return self.row, self.col, self.value

    def coalesce(self):
        key = torch.unique(self.row * self.n + self.col)
        out = SparseTensor(row=torch.div(key, self.n, rounding_mode="floor"), col=key % self.n, sparse_sizes=(self.n, self.n))
        return out

    def csr(self):
        rowptr = torch.zeros(self.n + 1, dtype=torch.int64)
        torch.cumsum(torch.bincount(self.row, minlength=self.n), 0, out=rowptr[1:])
        return rowptr, self.col, self.value


ts = types.ModuleType("torch_sparse")
ts.SparseTensor = SparseTensor
sys.modules["torch_sparse"] = ts
ogb = types.ModuleType("ogb")
ogbn = types.ModuleType("ogb.nodeproppred")
ogbn.PygNodePropPredDataset = object
ogb.nodeproppred = ogbn
sys.modules["ogb"], sys.modules["ogb.nodeproppred"] = ogb, ogbn
sys.modules["fast_sampler"] = ref.load_reference()
sys.path.insert(0, "/root/reference")
spec = importlib.util.spec_from_file_location("ref_dataset", "/root/reference/driver/dataset.py")
RD = importlib.util.module_from_spec(spec)
spec.loader.exec_module(RD)


def main():
    N, P = 1200, 4
    rowptr, col = S.powerlaw_graph(N, 9000, seed=11, head_offset=5.0)
    g = torch.Generator().manual_seed(12)
    x = torch.randn((N, 8), generator=g)
    y = torch.randint(0, 7, (N,), generator=g)
    labels = torch.randint(0, P, (N,), generator=g)
    # distinct access probabilities: the reference's argsort is unstable, ties would be implementation defined
    probs1 = (torch.randperm(N, generator=g).double() + 0.5) / (N + 1)
    probs2 = torch.stack([(torch.randperm(N, generator=g).double() + 0.5) / (N + 1) for _ in range(P)])
    perm = torch.randperm(N, generator=g)
    split = {"train": perm[:600], "valid": perm[600:800], "test": perm[800:1100]}
    out = {"rowptr": rowptr.numpy(), "col": col.numpy(), "x": x.numpy(), "y": y.numpy(), "labels": labels.numpy(),
           "probs1": probs1.numpy(), "probs2": probs2.numpy(), "num_parts": np.int64(P)}
    for k, v in split.items():
        out["split_" + k] = v.numpy()
    for tag, probs in (("p1", probs1), ("p2", probs2)):
        ds = RD.FastDataset("tiny", x, y, rowptr, col, split, {"num classes": 7})
        with tempfile.TemporaryDirectory() as d:
            RD.DisjointPartFeatReorderedDataset.reorder_and_save(ds, labels, probs, Path(d))
            prefix = Path(d) / f"metis-reordered-k{P}" / "tiny"
            for f in ("rowptr", "col", "part_offsets", "y"):
                out[f"{tag}_{f}"] = torch.load(prefix / f"{f}.pt", weights_only=False).numpy()
            sip = torch.load(prefix / "split_idx_parts.pt", weights_only=False)
            for r in range(P):
                out[f"{tag}_x{r}"] = torch.load(prefix / f"x{r}.pt", weights_only=False).view(torch.int16).numpy()
                for k in split:
                    out[f"{tag}_split_{r}_{k}"] = sip[r][k].numpy()
            assert torch.load(prefix / "split_idx.pt", weights_only=False) == dict()
            assert torch.load(prefix / "num_parts.pt", weights_only=False) == P
            # the reference's own loader reads what it wrote
            back = RD.DisjointPartFeatReorderedDataset.from_path(Path(d) / f"metis-reordered-k{P}", "tiny", 1)
            assert back.x.dtype == torch.float16 and back.num_parts == P
    path = os.path.join(ROOT, "tests", "golden", "dataset_reorder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
