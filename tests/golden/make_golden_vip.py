"""Generates tests/golden/vip.npz by importing the reference's own ``caching/vip.py`` (read-only,
from /root/reference) and calling ``vip_analytical`` (caching/vip.py:123-180) on a small seeded
graph.  The reference depends on ``torch_scatter`` and PyG, which are not installed here, so two
stand-in modules are injected before the import: ``torch_scatter.segment_csr`` restated with its
published semantics (out[i] = sum(src[indptr[i]:indptr[i+1]])) and an empty
``torch_geometric.data.NeighborSampler`` (not used by ``vip_analytical``).

    python tests/golden/make_golden_vip.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from salient_plusplus_b200 import synthetic as S  # noqa: E402


def segment_csr(src, indptr, reduce="add"):
    assert reduce in ("add", "sum")
    out = torch.zeros(indptr.numel() - 1, dtype=src.dtype)
    seg = torch.repeat_interleave(torch.arange(indptr.numel() - 1), indptr[1:] - indptr[:-1])
    out.index_add_(0, seg, src)
    return out


ts = types.ModuleType("torch_scatter")
ts.segment_csr = segment_csr
ts.gather_csr = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
sys.modules["torch_scatter"] = ts
tg = types.ModuleType("torch_geometric")
tg.__path__ = []  # a package without a `loader` sub-module -> the reference falls back to `.data`
tgd = types.ModuleType("torch_geometric.data")
tgd.NeighborSampler = object
tg.data = tgd
sys.modules["torch_geometric"] = tg
sys.modules["torch_geometric.data"] = tgd
sys.path.insert(0, "/root/reference")
import caching.vip as ref_vip  # noqa: E402

rowptr, col = S.powerlaw_graph(800, 9000, seed=41, head_offset=5.0)
N = rowptr.numel() - 1
P = 4
off = S.equal_partition_offsets(N, P)
train = [S.seeds(N, 60, seed=50 + p, lo=int(off[p]), hi=int(off[p + 1])) for p in range(P)]
out = dict(rowptr=rowptr.numpy(), col=col.numpy(), offsets=off.numpy())
for fi, fanouts in enumerate(([15, 10, 5], [25, 15])):
    probs = ref_vip.vip_analytical(rowptr, col, train, 32, fanouts, verbose=False)
    out[f"fanouts{fi}"] = np.array(fanouts)
    for p in range(P):
        out[f"vip{fi}_{p}"] = probs[p].numpy()
for p in range(P):
    out[f"train{p}"] = train[p].numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vip.npz"), **out)
print("vip.npz written", {k: v.shape for k, v in out.items() if k.startswith("vip0")})
