"""Generates tests/golden/vip_driver.npz: the output of the reference DRIVER's own VIP routine,
``get_frequency_tensors_fast`` (driver/drivers/ddp.py:133-239, the first-order form
1 - exp(-sum) the training driver actually uses), on a small seeded graph.

The module cannot be imported here (relative imports into the driver package, torch_sparse,
torch_scatter, a CUDA device and a process group), so the FUNCTION'S OWN SOURCE TEXT is read from
/root/reference (read-only, not copied into the repo) and executed unmodified in a namespace that
supplies stand-ins for its plumbing only: ``torch_scatter.segment_csr`` restated with its
published semantics, ``dist.get_rank() -> my_rank``, and -- with ``device='cpu'`` -- no-op
``torch.cuda.Stream`` / ``default_stream`` / ``stream`` / ``Tensor.record_stream``.  Every
arithmetic statement that runs is the reference's.

    python tests/golden/make_golden_vip_driver.py
"""
import contextlib
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from salient_plusplus_b200 import synthetic as S  # noqa: E402

SRC = "/root/reference/driver/drivers/ddp.py"
lines = open(SRC).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("def get_frequency_tensors_fast("))
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("def ") or lines[i].startswith("class ") or lines[i].startswith("@"))
func_src = "\n".join(lines[start:end])


def segment_csr(src, indptr, reduce="add"):
    assert reduce in ("add", "sum")
    out = torch.zeros(indptr.numel() - 1, dtype=src.dtype)
    seg = torch.repeat_interleave(torch.arange(indptr.numel() - 1), indptr[1:] - indptr[:-1])
    out.index_add_(0, seg, src)
    return out


class _FakeStream:
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def synchronize(self):
        pass


def run_reference(rowptr, col, train_by_part, partition_tensor, fanouts, my_rank, batch_size):
    fake_cuda = types.SimpleNamespace(Stream=_FakeStream, default_stream=lambda device=None: _FakeStream(),
                                      stream=lambda s: contextlib.nullcontext())
    fake_torch = types.ModuleType("torch_standin")
    fake_torch.__dict__.update(torch.__dict__)
    fake_torch.cuda = fake_cuda
    dist = types.SimpleNamespace(get_rank=lambda: my_rank)
    ns = {"torch": fake_torch, "dist": dist, "segment_csr": segment_csr, "time": time, "print": lambda *a, **k: None}
    # optimize=1: the reference launches its driver with PYTHONOPTIMIZE=1 (utils/exp_driver.py:154); the
    # routine's own assert at ddp.py:145 only holds for i == rank and would fire otherwise
    exec(compile(func_src, SRC, "exec", optimize=1), ns)
    dataset = types.SimpleNamespace(adj_t=lambda: types.SimpleNamespace(csr=lambda: (rowptr, col, None)),
                                    split_idx_parts={p: {"train": t} for p, t in enumerate(train_by_part)})
    orig = torch.Tensor.record_stream
    torch.Tensor.record_stream = lambda self, stream: None
    try:
        return ns["get_frequency_tensors_fast"](dataset, fanouts, partition_tensor, "cpu", my_rank, batch_size)
    finally:
        torch.Tensor.record_stream = orig


def main():
    rowptr, col = S.powerlaw_graph(900, 11000, seed=43, head_offset=5.0)
    N = rowptr.numel() - 1
    P = 4
    off = S.equal_partition_offsets(N, P)
    part = torch.searchsorted(off, torch.arange(N), right=True) - 1
    train = [S.seeds(N, 70, seed=60 + p, lo=int(off[p]), hi=int(off[p + 1])) for p in range(P)]
    out = dict(rowptr=rowptr.numpy(), col=col.numpy(), offsets=off.numpy())
    for fi, fanouts in enumerate(([15, 10, 5], [25, 15])):
        out[f"fanouts{fi}"] = np.array(fanouts)
        for p in range(P):
            probs = run_reference(rowptr, col, train, part, fanouts, p, 32)
            assert probs.dtype == torch.float64 and probs.numel() == N
            out[f"vip{fi}_{p}"] = probs.numpy()
    for p in range(P):
        out[f"train{p}"] = train[p].numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vip_driver.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; max prob", float(max(out[f"vip0_{p}"].max() for p in range(P))))


if __name__ == "__main__":
    main()
