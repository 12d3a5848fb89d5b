"""``Adj`` / ``SparseTensor`` as the consumer sees them (fast_trainer/monkeypatch.py:25-69,
fast_trainer/samplers.py:22-30).  When torch_sparse / PyG are installed the real classes are
used so batches plug straight into PyG models; otherwise minimal stand-ins with the same
attribute surface (``adj_t``, ``e_id``, ``size``, ``to``, ``record_stream``, ``pin_memory``) are
provided -- this image has neither package."""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch

try:  # pragma: no cover - not installed in this image
    from torch_sparse import SparseTensor  # type: ignore
    HAVE_TORCH_SPARSE = True
except Exception:  # noqa: BLE001
    HAVE_TORCH_SPARSE = False

    class SparseTensor:  # type: ignore
        """CSR holder with the constructor signature fast_trainer/samplers.py:25-27 uses."""

        __slots__ = ("_rowptr", "_row", "_col", "_value", "_sparse_sizes")

        def __init__(self, rowptr=None, row=None, col=None, value=None, sparse_sizes=None, is_sorted=False,
                     trust_data=False):
            self._rowptr, self._row, self._col, self._value = rowptr, row, col, value
            self._sparse_sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))

        @classmethod
        def from_csr(cls, rowptr, col, sparse_sizes):
            """Same object as ``SparseTensor(rowptr=..., col=..., sparse_sizes=..., is_sorted=True,
            trust_data=True)`` without the keyword parsing (one per hop and mini-batch)."""
            st = cls.__new__(cls)
            st._rowptr, st._row, st._col, st._value, st._sparse_sizes = rowptr, None, col, None, sparse_sizes
            return st

        def csr(self):
            return self._rowptr, self._col, self._value

        def sparse_sizes(self):
            return self._sparse_sizes

        def sparse_size(self, dim):
            return self._sparse_sizes[dim]

        def nnz(self):
            return int(self._col.numel())

        @property
        def storage(self):
            return self

        def to(self, device=None, non_blocking=False, **kw):
            mv = lambda t: None if t is None else t.to(device=device, non_blocking=non_blocking)
            return SparseTensor(rowptr=mv(self._rowptr), row=mv(self._row), col=mv(self._col), value=mv(self._value),
                                sparse_sizes=self._sparse_sizes, is_sorted=True, trust_data=True)

        def pin_memory(self):
            pm = lambda t: None if t is None else (t if t.is_cuda else t.pin_memory())
            return SparseTensor(rowptr=pm(self._rowptr), row=pm(self._row), col=pm(self._col), value=pm(self._value),
                                sparse_sizes=self._sparse_sizes, is_sorted=True, trust_data=True)


def sparse_record_stream(st, stream) -> None:
    """fast_trainer/monkeypatch.py:36-63"""
    s = st.storage if HAVE_TORCH_SPARSE else st
    for name in ("_row", "_rowptr", "_col", "_value", "_rowcount", "_colptr", "_colcount", "_csr2csc", "_csc2csr"):
        t = getattr(s, name, None)
        if t is not None and t.is_cuda:
            t.record_stream(stream)


class Adj(NamedTuple):
    adj_t: SparseTensor
    e_id: Optional[torch.Tensor]
    size: Tuple[int, int]

    def to(self, *args, **kwargs):
        adj_t = self.adj_t.to(*args, **kwargs)
        e_id = self.e_id.to(*args, **kwargs) if self.e_id is not None else None
        return Adj(adj_t, e_id, self.size)

    def pin_memory(self, *args, **kwargs):
        e_id = self.e_id
        if e_id is not None and not e_id.is_cuda:
            e_id = e_id.pin_memory()
        return Adj(self.adj_t.pin_memory(), e_id, self.size)

    def record_stream(self, stream):
        sparse_record_stream(self.adj_t, stream)
        if self.e_id is not None and self.e_id.is_cuda:
            self.e_id.record_stream(stream)


_new_adj = tuple.__new__


def Adj__from_fast_sampler(adj) -> Adj:
    """fast_trainer/samplers.py:22-30"""
    rowptr, col, e_id, sparse_sizes = adj
    if HAVE_TORCH_SPARSE:
        adj_t = SparseTensor(rowptr=rowptr, row=None, col=col, value=None, sparse_sizes=tuple(sparse_sizes),
                             is_sorted=True, trust_data=True)
    else:
        adj_t = SparseTensor.from_csr(rowptr, col, (int(sparse_sizes[0]), int(sparse_sizes[1])))
    return _new_adj(Adj, (adj_t, e_id, (sparse_sizes[1], sparse_sizes[0])))
