"""Device-resident mini-batch pipeline: the same C-ABI launch sequence the Session issues
(sampler -> [owner split] -> feature gather -> label gather), but with static per-slot output
buffers and no host synchronisation between batches.  bench.py uses it for the kernel-only
throughput (`value`) and the per-kernel roofline pass; the Session (fast_sampler.py) is the
public, reference-shaped API on top of the same calls."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import SPP_META_WORDS, SPP_MAX_PARTS, BatchJob, FeatureMap, check
from .fast_sampler import _DeviceGraph, _Workspace, _sampler_sizes

c_vp = ctypes.c_void_p


class _PipeSlot:
    pass


class MiniBatchPipeline:
    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, sizes: Sequence[int], batch_size: int,
                 x_table: Optional[torch.Tensor] = None, y_table: Optional[torch.Tensor] = None,
                 feature_map: Optional[FeatureMap] = None, feat_dim: int = 0, feat_dtype=torch.float16,
                 split: bool = False, use_cache: bool = False, depth: int = 4, device=None):
        assert all(int(s) >= 0 for s in sizes), "the static pipeline covers sampled hops only"
        self.lib = _lib.load()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.g = _DeviceGraph.get(rowptr, col)
        self.sizes = [int(s) for s in sizes]
        self.L = len(self.sizes)
        self.bs = int(batch_size)
        self.sz = _sampler_sizes(self.bs, self.sizes, self.g)
        # x_table: a fast_sampler._FeatureTable (possibly row-padded) or a dense tensor
        if x_table is not None and isinstance(x_table, torch.Tensor):
            from .fast_sampler import feature_table
            x_table = feature_table(x_table)
        self.x_table, self.y_table, self.fm = x_table, y_table, feature_map
        if x_table is not None:
            feat_dim, feat_dtype = x_table.dim, x_table.dtype
        self.feat_dim, self.feat_dtype = feat_dim, feat_dtype
        self.row_bytes = feat_dim * torch.empty(0, dtype=feat_dtype).element_size()
        self.split, self.use_cache = split, use_cache
        import os
        self.split_gather = os.environ.get("SPP_GATHER_SPLIT", "0") != "0"
        self.sz_arr = (ctypes.c_int32 * max(self.L, 1))(*self.sizes)
        self.caps = (ctypes.c_int64 * max(self.L, 1))(*[int(self.sz.hop_edges[h]) for h in range(self.L)])
        self.slots: List[_PipeSlot] = []
        for _ in range(depth):
            s = _PipeSlot()
            s.ws = _Workspace(self.sz, self.device)
            s.stream = torch.cuda.Stream(self.device)
            s.rowptrs = [torch.empty(int(self.sz.hop_targets[h]) + 1, dtype=torch.int64, device=self.device)
                         for h in range(self.L)]
            s.cols = [torch.empty(max(int(self.sz.hop_edges[h]), 1), dtype=torch.int64, device=self.device)
                      for h in range(self.L)]
            s.rp = (c_vp * max(self.L, 1))(*[t.data_ptr() for t in s.rowptrs])
            s.cp = (c_vp * max(self.L, 1))(*[t.data_ptr() for t in s.cols])
            s.x = (torch.empty((s.ws.max_nodes, feat_dim), dtype=feat_dtype, device=self.device)
                   if (x_table is not None or feature_map is not None) else None)
            s.y = (torch.empty((self.bs, y_table.size(-1)), dtype=y_table.dtype, device=self.device)
                   if y_table is not None else None)
            if split:
                words = int(self.lib.spp_split_scratch_words(s.ws.max_nodes))
                s.scratch = torch.empty(words, dtype=torch.int32, device=self.device)
                s.bucket_ids = torch.empty(s.ws.max_nodes, dtype=torch.int64, device=self.device)
                s.perm = torch.empty(s.ws.max_nodes, dtype=torch.int64, device=self.device)
                s.counts = torch.zeros(SPP_MAX_PARTS + 2, dtype=torch.int64, device=self.device)
            s.counters = torch.zeros(3, dtype=torch.int64, device=self.device)  # local / cache / peer rows
            s.meta_host = torch.empty(SPP_META_WORDS, dtype=torch.int64).pin_memory()
            s.job_dev = torch.zeros(ctypes.sizeof(_lib.DeviceJob) // 8, dtype=torch.int64, device=self.device)
            s.job_host = torch.zeros(ctypes.sizeof(_lib.DeviceJob) // 8, dtype=torch.int64).pin_memory()
            s.done = torch.cuda.Event()
            s.in_flight = False
            s.job = self._make_job(s)
            self.slots.append(s)
        self.gather_events = None

    def _make_job(self, s: _PipeSlot, graph: bool = True) -> BatchJob:
        """The slot's spp_batch_job: the very descriptor a Session submits, with static outputs."""
        j = BatchJob()
        ctypes.memset(ctypes.byref(j), 0, ctypes.sizeof(j))
        j.graph, j.ws = self.g.c, s.ws.c
        j.n_hops, j.replace = self.L, 0
        for h in range(self.L):
            j.sizes[h] = self.sizes[h]
            j.out_col_cap[h] = int(self.sz.hop_edges[h])
            j.out_rowptr[h] = s.rowptrs[h].data_ptr()
            j.out_col[h] = s.cols[h].data_ptr()
        j.row_bytes = self.row_bytes
        j.stream = s.stream.cuda_stream
        if graph:   # one CUDA graph per slot replays the launch sequence (what a Session's slot does)
            j.job_dev, j.job_host, j.batch_size_cap = s.job_dev.data_ptr(), s.job_host.data_ptr(), self.bs
        if s.x is not None:
            j.x_out = s.x.data_ptr()
            if self.fm is not None:
                j.feature_mode, j.fmap = 2, self.fm
            else:
                j.feature_mode, j.table, j.table_pitch = 1, self.x_table.ptr, self.x_table.pitch
        if self.y_table is not None:
            j.y_table = self.y_table.data_ptr()
            j.y_row_bytes = self.y_table.size(-1) * self.y_table.element_size()
            j.y_out = s.y.data_ptr()
        if self.split:
            j.do_split, j.use_cache = 1, int(self.use_cache)
            if j.feature_mode != 2:
                j.fmap = self.fm
            j.bucket_ids, j.perm = s.bucket_ids.data_ptr(), s.perm.data_ptr()
            j.bucket_counts, j.split_scratch = s.counts.data_ptr(), s.scratch.data_ptr()
        if graph:
            j.batch_size = self.bs
            check(self.lib.spp_batch_prepare(ctypes.byref(j)), "spp_batch_prepare")
        return j

    # -- individual stages (all asynchronous on the slot's stream) ------------------------------
    def sample(self, s: _PipeSlot, seeds_ptr: int, bs: int, rng_seed: int):
        check(self.lib.spp_sample_minibatch(ctypes.byref(self.g.c), seeds_ptr, bs, self.sz_arr, self.L, 0,
                                            ctypes.c_uint64(rng_seed), ctypes.byref(s.ws.c), s.rp, s.cp, self.caps,
                                            None, s.stream.cuda_stream), "spp_sample_minibatch")

    def remote_mask(self) -> int:
        """Buckets (bit p: partition p) whose rows live on other GPUs (or are treated as such)."""
        fm = self.fm
        if fm is None:
            return 0
        local = (1 << fm.rank) | int(fm.local_parts)
        return sum(1 << p for p in range(fm.num_parts) if not (local >> p) & 1 and fm.offsets[p + 1] > fm.offsets[p])

    def gather(self, s: _PipeSlot, count_rows: bool = False):
        n_dev = s.ws.meta_ptr(self.L)
        cnt = s.counters.data_ptr() if count_rows else None
        if self.fm is not None and self.split and self.remote_mask() and self.split_gather:
            # what a batch does (session.cu): rows by source class, peer buckets and local buckets as
            # two launches (here back to back on one stream so that CUDA events time both)
            peer = self.remote_mask()
            local = ((1 << (self.fm.num_parts + 1)) - 1) & ~peer
            for mask in (peer, local):
                check(self.lib.spp_gather_by_class(ctypes.byref(self.fm), self.row_bytes, s.bucket_ids.data_ptr(),
                                                   s.scratch.data_ptr(), s.ws.max_nodes, mask, s.x.data_ptr(), cnt,
                                                   s.stream.cuda_stream), "spp_gather_by_class")
        elif self.fm is not None:
            check(self.lib.spp_gather_partitioned(ctypes.byref(self.fm), self.row_bytes, s.ws.n_ids.data_ptr(), 0,
                                                  s.ws.max_nodes, n_dev, s.scratch.data_ptr() if self.split else None,
                                                  s.x.data_ptr(), s.ws.max_nodes, cnt,
                                                  s.stream.cuda_stream), "spp_gather_partitioned")
        elif self.x_table is not None:
            check(self.lib.spp_gather_rows_pitched(self.x_table.ptr, self.x_table.pitch, self.row_bytes,
                                                   s.ws.n_ids.data_ptr(), 0, s.ws.max_nodes, n_dev, s.x.data_ptr(),
                                                   s.ws.max_nodes, s.stream.cuda_stream), "spp_gather_rows_pitched")

    def labels(self, s: _PipeSlot, seeds_ptr: int, bs: int):
        if self.y_table is not None and bs > 0:
            yb = self.y_table.size(-1) * self.y_table.element_size()
            check(self.lib.spp_gather_rows(self.y_table.data_ptr(), yb, seeds_ptr, 1, bs, None, s.y.data_ptr(), bs,
                                           s.stream.cuda_stream), "spp_gather_rows(y)")

    def owner_split(self, s: _PipeSlot):
        if self.split:
            check(self.lib.spp_split_by_owner(ctypes.byref(self.fm), int(self.use_cache), s.ws.n_ids.data_ptr(), 0,
                                              s.ws.max_nodes, s.ws.meta_ptr(self.L), s.bucket_ids.data_ptr(),
                                              s.perm.data_ptr(), s.counts.data_ptr(), s.scratch.data_ptr(),
                                              s.stream.cuda_stream), "spp_split_by_owner")

    def launch(self, slot: int, seeds_ptr: int, bs: int, rng_seed: int, time_gather: bool = False,
               count_rows: bool = False):
        """One mini-batch on slot ``slot``; returns the (start, end) events around the feature
        gather when ``time_gather`` (``count_rows``: add the rows served local / cache / peer to
        the slot's ``counters``)."""
        s = self.slots[slot]
        if not time_gather:  # one C call, exactly what the Session's executor issues
            j = s.job
            if s.in_flight:
                # the pinned staging copy of the job block is read by the GPU when the slot's previous
                # graph launch executes: wait for that batch before overwriting it (a Session only
                # re-uses a slot after its batch was delivered)
                s.done.synchronize()
            j.seeds_dev, j.batch_size, j.rng_seed = seeds_ptr, bs, rng_seed
            check(self.lib.spp_batch_enqueue(ctypes.byref(j)), "spp_batch_enqueue")
            s.done.record(s.stream)
            s.in_flight = True
            return None
        self.sample(s, seeds_ptr, bs, rng_seed)
        self.owner_split(s)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(s.stream)
        self.gather(s, count_rows)
        ev[1].record(s.stream)
        self.labels(s, seeds_ptr, bs)
        return ev

    def read_meta(self, slot: int) -> List[int]:
        s = self.slots[slot]
        with torch.cuda.stream(s.stream):
            s.meta_host.copy_(s.ws.meta, non_blocking=True)
        s.stream.synchronize()
        return s.meta_host.tolist()
