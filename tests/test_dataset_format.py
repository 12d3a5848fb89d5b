"""On-disk partitioned dataset format (driver/dataset.py:183-215,270-369): write with
reorder_and_save, read back per rank, check the relabelling invariants.  CPU only (the same
torch ops run on the GPU for full-size graphs)."""
import os

import numpy as np
import torch

from salient_plusplus_b200 import synthetic as S
from salient_plusplus_b200.dataset import (DisjointPartFeatReorderedDataset, csr_permute_symmetric,
                                           partition_permutation)


def test_reorder_save_load_roundtrip(tmp_path):
    rowptr, col = S.powerlaw_graph(500, 4000, seed=5, head_offset=5.0)
    N = rowptr.numel() - 1
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 6, generator=g)
    y = torch.randint(0, 7, (N,), generator=g)
    labels = torch.randint(0, 3, (N,), generator=g)
    prob = torch.rand(N, generator=g, dtype=torch.float64)
    split = {"train": torch.randperm(N, generator=g)[:120], "valid": torch.arange(0, 30), "test": torch.arange(400, 450)}
    prefix = DisjointPartFeatReorderedDataset.reorder_and_save("toy", rowptr, col, x, y, split, {"num classes": 7}, labels,
                                                               prob, tmp_path)
    assert prefix == tmp_path / "metis-reordered-k3" / "toy"
    perm, invperm = partition_permutation(labels, prob)
    parts = [DisjointPartFeatReorderedDataset.from_path(tmp_path / "metis-reordered-k3", "toy", r) for r in range(3)]
    d0 = parts[0]
    off = d0.part_offsets
    assert d0.num_parts == 3 and d0.num_nodes == N and d0.num_classes == 7 and d0.num_features == 6
    assert off.tolist() == [0] + torch.cumsum(torch.bincount(labels, minlength=3), 0).tolist()
    # partitions are contiguous id ranges, descending probability inside each
    new_labels = labels[perm]
    assert bool((new_labels[1:] >= new_labels[:-1]).all())
    for r in range(3):
        pr = prob[perm][int(off[r]):int(off[r + 1])]
        assert bool((pr[1:] <= pr[:-1]).all())
        # features of rank r are exactly the rows of its range, fp16
        assert parts[r].x.dtype == torch.float16
        assert torch.equal(parts[r].x, x[perm][int(off[r]):int(off[r + 1])].to(torch.float16))
        # split ids are relabelled and belong to the partition
        for k in split:
            ids = parts[r].split_idx_parts[r][k]
            assert bool(((ids >= off[r]) & (ids < off[r + 1])).all())
    for k, v in split.items():
        got = torch.cat([d0.split_idx_parts[r][k] for r in range(3)])
        assert sorted(got.tolist()) == sorted(invperm[v].tolist())
    assert torch.equal(d0.y, y[perm])
    # the graph is the same graph under the relabelling: edge (u, v) <-> (invperm[u], invperm[v])
    deg = rowptr[1:] - rowptr[:-1]
    src = torch.repeat_interleave(torch.arange(N), deg)
    old = set(zip(invperm[src].tolist(), invperm[col].tolist()))
    ndeg = d0.rowptr[1:] - d0.rowptr[:-1]
    nsrc = torch.repeat_interleave(torch.arange(N), ndeg)
    assert set(zip(nsrc.tolist(), d0.col.tolist())) == old
    assert d0.get_num_iterations(32) == {"train": 3, "valid": 1, "test": 1}
    book = d0.get_RangePartitionBook()
    assert book.rank == 0 and book.world_size == 3


def test_partitionwise_probabilities_and_identity_permutation():
    labels = torch.tensor([1, 0, 1, 0, 2])
    p2 = torch.tensor([[0.1, 0.9, 0.1, 0.2, 0.0], [0.5, 0.0, 0.7, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0, 0.3]], dtype=torch.float64)
    perm, invperm = partition_permutation(labels, p2)
    assert perm.tolist() == [1, 3, 2, 0, 4]          # part 0: v1 (0.9), v3 (0.2); part 1: v2 (0.7), v0 (0.5); part 2: v4
    assert invperm[perm].tolist() == [0, 1, 2, 3, 4]
    rowptr = torch.tensor([0, 2, 3, 3, 4, 4])
    col = torch.tensor([1, 4, 0, 2])
    ident = torch.arange(5)
    rp, cl = csr_permute_symmetric(rowptr, col, ident)
    assert torch.equal(rp, rowptr) and torch.equal(cl, col)


def test_reorder_and_save_matches_the_reference_writer(tmp_path):
    """tests/golden/dataset_reorder.npz holds what the reference's own
    DisjointPartFeatReorderedDataset.reorder_and_save (driver/dataset.py:270-369, run by
    tests/golden/make_golden_dataset.py) wrote for a seeded graph: every file we write for the
    same inputs must be identical (1-D and 2-D access probabilities, distinct values so the
    reference's unstable argsort is well defined)."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_reorder.npz"))
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    P = int(g["num_parts"])
    split = {k: t("split_" + k) for k in ("train", "valid", "test")}
    for tag, probs in (("p1", t("probs1")), ("p2", t("probs2"))):
        prefix = DisjointPartFeatReorderedDataset.reorder_and_save(
            "tiny", t("rowptr"), t("col"), t("x"), t("y"), split, {"num classes": 7}, t("labels"), probs, tmp_path / tag)
        for f in ("rowptr", "col", "part_offsets", "y"):
            got = torch.load(prefix / f"{f}.pt", weights_only=False)
            assert np.array_equal(got.numpy(), g[f"{tag}_{f}"]), (tag, f)
        sip = torch.load(prefix / "split_idx_parts.pt", weights_only=False)
        for r in range(P):
            xr = torch.load(prefix / f"x{r}.pt", weights_only=False)
            assert xr.dtype == torch.float16 and np.array_equal(xr.view(torch.int16).numpy(), g[f"{tag}_x{r}"])
            for k in split:
                # the reference orders a part's seeds with an UNSTABLE argsort over partition ids that are
                # all tied (driver/dataset.py:341): the order inside a part is implementation defined (it
                # is shuffled every epoch anyway), the membership is the contract
                assert np.array_equal(np.sort(sip[r][k].numpy()), np.sort(g[f"{tag}_split_{r}_{k}"])), (tag, r, k)
        assert torch.load(prefix / "split_idx.pt", weights_only=False) == dict()
        assert torch.load(prefix / "num_parts.pt", weights_only=False) == P
        ds = DisjointPartFeatReorderedDataset.from_path(prefix.parent, "tiny", 2)
        assert ds.rank == 2 and ds.x.dtype == torch.float16 and ds.num_parts == P
