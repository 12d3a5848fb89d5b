#!/usr/bin/env python
"""bench.py -- sampled-and-gathered mini-batches/s (+ gathered feature GB/s) of the mini-batch
generation path, on synthetic graphs of the BASELINE shapes.

    python bench.py --gpus 1 --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU sampler

A "step" is one mini-batch: 1024 seeds -> multi-hop neighbour sampling with dedup/relabel ->
owner split (partition book + VIP cache) -> feature gather of every sampled node (+ label gather).

Default workload at every N: BASELINE.json configs[3], the configuration the metric is quoted on --
ogbn-papers100M-shaped graph, features range-partitioned 8 ways with a 15 % VIP cache
(docs/INSTALL.md:186), fan-out (15,10,5), batch 1024.  With N < 8 GPUs every GPU hosts 8/N
consecutive partitions (N = 1: all eight, every table local); rows of partitions on other GPUs
that are not in the replicated cache are read over NVLink by the fused P2P gather.  Weak scaling:
every rank runs K batches whose seeds are local vertices (the reference's federated split,
driver/drivers/ddp.py:324-328).  `--workload products` (configs[1]), `arxiv`, `mag240m` and
`products-layerwise` (configs[2]) select the other BASELINE shapes.

Features and labels are pure functions of the GLOBAL vertex id (synthetic.features_by_id), so
every rank can verify the rows it gathered from anywhere: the `parity` object of the JSON line is
the result of checking, outside the timed regions and through the public Session API,
x == f(n_id), y == g(seeds), n_id[:bs] == seeds and cat(partition_nids, cached)[perm] == n_id on
>= 3 batches per rank; a mismatch on any rank makes the run exit non-zero.

Printed JSON (rank 0, one line): see the contract in the task statement; `value` is the
device-timed throughput with every input resident in HBM, `e2e` goes through the public
FastSampler -> DevicePrefetcher / DeviceDistributedPrefetcher API with the seeds in host memory
(H2D of the seeds and D2H of the batch's size block inside the timed region, every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (dataset shape, fanout, batch, default partitions (0 = one per GPU), description)
    "papers100M": ("papers100M", [15, 10, 5], 1024, 8,
                   "ogbn-papers100M-shaped synthetic (111M nodes, 1.6B CSR entries = 0.8B directed edges symmetrised, "
                   "128-d fp16), fanout (15,10,5), batch 1024"),
    "products": ("products", [15, 10, 5], 1024, 0,
                 "ogbn-products-shaped synthetic (2.45M nodes, 61.9M directed edges symmetrised, 100-d fp16), "
                 "fanout (15,10,5), batch 1024"),
    "arxiv": ("arxiv", [15, 10, 5], 1024, 0,
              "ogbn-arxiv-shaped synthetic (169K nodes, 1.17M directed edges symmetrised, 128-d fp32), "
              "fanout (15,10,5), batch 1024"),
    "mag240m": ("mag240m", [25, 15], 1024, 8,
                "MAG240M-shaped homogeneous synthetic (244M nodes, 1.7B CSR entries = 0.85B directed edges symmetrised, "
                "768-d fp16), fanout (25,15), batch 1024"),
    "products-layerwise": ("products", [-1], 1024, 1,
                           "ogbn-products-shaped synthetic (2.45M nodes, 61.9M directed edges symmetrised, 100-d fp16), "
                           "layer-wise full-neighbourhood inference batches (sizes [-1], driver/models.py:441-495), "
                           "1024 consecutive vertices per batch, vertices split by rank"),
}
DEFAULT_WORKLOAD = "papers100M"
METRIC = "sampled_and_gathered_minibatches_per_sec"
UNIT = "batches/s"
GRAPH_GEN = "Chung-Lu power law gamma=2.5 head_offset=100 seed=1, symmetrised, deduplicated"


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def log(msg):
    if os.environ.get("SPP_BENCH_VERBOSE") and int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench +{time.perf_counter() - _T0:.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


FEAT_DTYPE = None   # --feat-dtype: overrides the shape's feature dtype ("products ... also report fp32")


def ref_threads(args) -> int:
    """Worker threads of the CPU reference legs: every host core, or --ref-threads (the reference's
    launcher default is 15 workers per GPU, utils/exp_driver.py:48)."""
    return int(args.ref_threads) if getattr(args, "ref_threads", 0) else (os.cpu_count() or 1)


def make_graph(shape: str, scale: float, device, locality: float = 0.0, parts: int = 8):
    from salient_plusplus_b200 import synthetic as S
    n, e, f, dt = S.SHAPES[shape]
    dt = FEAT_DTYPE or dt
    if shape in ("papers100M", "mag240m"):
        e //= 2  # BASELINE's 1.6B is taken as the CSR entry count ("say which")
    n, e = max(1024, int(n * scale)), max(4096, int(e * scale))
    rowptr, col = S.powerlaw_graph(n, e, seed=1, device=device, locality=locality, locality_parts=parts)
    return n, f, dt, rowptr, col


def num_parts(args, world):
    P = args.parts if args.parts > 0 else WORKLOADS[args.workload][3]
    P = P if P > 0 else world
    if P % world != 0:
        raise SystemExit(f"{P} feature partitions cannot be spread evenly over {world} GPUs")
    return P


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the box's host cores
# ------------------------------------------------------------------------------------------------
def run_reference_cpu(rowptr, col, x, y, idx, sizes, bs, threads, warmup, steps, max_seconds=None, preroll=None):
    """Drain a reference fast_sampler.Session (oracle/_ref, built from /root/reference by
    oracle/build_ref.sh).  `preroll` batches (default 2 x threads) are consumed before the `warmup`
    ones so that the timed batches are steady state, not the first wave of the thread pool.
    Returns (batches/s, gathered GB/s, batches timed, kind, total batches)."""
    from oracle import ref
    os.environ.setdefault("OMP_NUM_THREADS", "1")  # utils/exp_driver.py:154
    if ref.available():
        R = ref.load_reference()
        cfg = R.Config()
        cfg.x_cpu, cfg.x_gpu, cfg.y = x, torch.empty(0), y
        cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
        cfg.batch_size, cfg.sizes = bs, list(sizes)
        cfg.skip_nonfull_batch = True
        cfg.pin_memory = bool(torch.cuda.is_available())
        cfg.distributed = False
        cfg.force_exact_num_batches, cfg.exact_num_batches = False, 0
        cfg.count_remote_frequency = cfg.use_cache = False
        sess = R.Session(threads, 100, cfg)
        total = sess.num_total_batches
        pre = (2 * threads if preroll is None else preroll) + warmup
        pre = max(0, min(pre, total - steps))
        got, nodes = 0, 0
        t0 = None
        t_start = time.perf_counter()
        while True:
            if got == pre:
                t0 = time.perf_counter()
                nodes = 0
            b = sess.blocking_get_batch()
            if b is None:
                break
            got += 1
            if got > pre:
                nodes += b[0].size(0)
            if got >= pre + steps:
                break
            if max_seconds is not None and t0 is not None and time.perf_counter() - t0 > max_seconds and got > pre + 8:
                break
        t1 = time.perf_counter()
        timed = got - pre
        # drain so the worker threads go back to the pool cleanly
        while sess.blocking_get_batch() is not None:
            pass
        del sess
        dt = t1 - (t0 if t0 is not None else t_start)
        gbs = nodes * x.size(1) * x.element_size() / dt / 1e9
        return timed / dt, gbs, timed, "reference", total
    # fallback: the single-threaded C port (oracle/salient_oracle.c)
    import ctypes
    import numpy as np
    from oracle import oracle as O
    L = O.lib()
    rp, cl = rowptr.numpy(), col.numpy()
    xs = x.view(torch.int16).numpy() if x.dtype == torch.float16 else x.numpy()
    row_bytes = x.size(1) * x.element_size()
    out = np.empty((1100000, xs.shape[1]), dtype=xs.dtype)
    sz = (ctypes.c_int32 * len(sizes))(*sizes)
    ids = idx.numpy()
    t0, nodes, timed = None, 0, 0
    for b in range(warmup + steps):
        if b == warmup:
            t0 = time.perf_counter()
        seeds = np.ascontiguousarray(ids[b * bs:(b + 1) * bs])
        nb = L.spo_minibatch(rp.ctypes.data, cl.ctypes.data, seeds.ctypes.data, seeds.size, sz, len(sizes),
                             ((b + 1) * bs * 17 + 5) & 0xFFFFFFFF, xs.ctypes.data, row_bytes, out.ctypes.data,
                             out.shape[0])
        if b >= warmup:
            nodes += nb
            timed += 1
            if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
                break
    dt = time.perf_counter() - t0
    return timed / dt, nodes * row_bytes / dt / 1e9, timed, "port", warmup + steps


def run_reference_distributed(R, rowptr, col, x_blocks, y, idxs, sizes, bs, off, part_ranks, caches, threads_each,
                              warmup, steps, preroll):
    """The reference's DISTRIBUTED CPU path: one Session per GPU of
    the run, each with cores / N worker threads, `distributed=True`, the same RangePartitionBook and
    a Cache, drained through `try_get_batch_distributed` (fast_sampler.cpp:802-828) -- sampling +
    owner binning + the slice of host-resident local rows (gpu_percent 0.999, docs/INSTALL.md:193).
    The Sessions live in ONE process (their worker threads are native and come from the reference's
    process-wide pool; the Python thread only pops finished batches round-robin), which loads the
    host cores exactly like N processes would.  Returns (aggregate batches/s, batches timed)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    P = off.numel() - 1
    sessions = []
    for g, pr in enumerate(part_ranks):
        xb = x_blocks[g]
        cut = int(xb.size(0) * 0.999)
        cfg = R.Config()
        cfg.x_cpu = xb[cut:].contiguous()
        cfg.x_gpu = torch.empty((cut, 0), dtype=xb.dtype)   # only its row count is read (fast_sampler.cpp:1044,1144)
        cfg.y = y
        cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idxs[g]
        cfg.batch_size, cfg.sizes = bs, list(sizes)
        cfg.skip_nonfull_batch = True
        cfg.pin_memory = bool(torch.cuda.is_available())
        cfg.distributed = True
        cfg.partition_book = R.RangePartitionBook(pr, P, off)
        cfg.cache = caches[g]
        cfg.use_cache = True
        cfg.force_exact_num_batches, cfg.exact_num_batches = False, 0
        cfg.count_remote_frequency = False
        sessions.append(R.Session(threads_each, 100, cfg))
    n = len(sessions)
    pre_total, stop_total = n * (preroll + warmup), n * (preroll + warmup + steps)
    got, t0, t1 = 0, None, None
    done = [False] * n
    while got < stop_total and not all(done):
        for g, s in enumerate(sessions):
            if done[g]:
                continue
            if s.num_consumed_batches == s.num_total_batches:
                done[g] = True
                continue
            b = s.try_get_batch_distributed()
            if b is None:
                continue
            got += 1
            if got == pre_total:
                t0 = time.perf_counter()
            if got == stop_total:
                t1 = time.perf_counter()
                break
    if t1 is None:
        t1 = time.perf_counter()
    for s in sessions:  # drain so the worker threads go back to the pool cleanly
        while s.blocking_get_batch_distributed() is not None:
            pass
    timed = got - pre_total
    return timed / (t1 - t0), timed


def reference_arm(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    from salient_plusplus_b200 import synthetic as S
    from salient_plusplus_b200.peer import hosted_partitions
    shape, sizes, bs, _, desc = WORKLOADS[args.workload]
    if args.feat_dtype != "default":
        desc += f" [features stored as {args.feat_dtype} for this run]"
    N = max(args.gpus, 1)
    P = num_parts(args, N)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n, f, dt, rowptr, col = make_graph(shape, args.scale, dev, args.locality, max(P, 1))
    deg = (rowptr[1:] - rowptr[:-1]) if N > 1 else None
    rowptr, col = rowptr.cpu(), col.cpu()
    threads = ref_threads(args)
    W, K = args.warmup, args.steps
    layerwise = args.workload == "products-layerwise"
    if N == 1 or layerwise:
        # non-distributed Session over the whole graph: sampling + CPU feature slice
        x = S.features_by_id(0, n, f, dt, device=dev).cpu()
        y = S.labels_by_id(torch.arange(n))
        preroll = 2 * threads
        # a reference "step" is a bundle of m mini-batches so that the K timed steps cover >= 20 waves of
        # the thread pool (20 single batches on 16-32 threads are ~1 wave: a +-2.7x error)
        m = max(1, -(-20 * threads // max(K, 1)))
        need = (preroll + W + K * m) * bs
        if layerwise:
            idx = torch.arange(n, dtype=torch.int64)[:min(n, need)]
        else:
            idx = S.seeds(n, min(n, need), seed=7)
        if idx.numel() < need:
            idx = idx.repeat((need + idx.numel() - 1) // idx.numel())[:need]
        bps, gbs, timed, kind, _ = run_reference_cpu(rowptr, col, x, y, idx, sizes, bs, threads, W, K * m, preroll=preroll)
        sample = (f"{timed} mini-batches ({K} steps of {m}) after {preroll} pre-roll + {W} warm-up batches (steady state of the thread pool), "
                  f"reference fast_sampler.Session with {threads} worker threads (sampling + CPU feature slice, pinned outputs)")
        kind_out = kind
    else:
        # distributed Sessions, one per GPU of the run (cores / N threads each), same book + cache
        from oracle import ref
        # without a CUDA driver (build container) the distributed Session cannot pin its outputs: the
        # variant of the same sources with the pinned_memory(true) literals switched off is used there
        nopin = not torch.cuda.is_available()
        if not ref.available(nopin):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fast_sampler.so missing"}), flush=True)
            return
        R = ref.load_reference(nopin)
        off = S.equal_partition_offsets(n, P)
        threads_each = max(1, threads // N)
        preroll = 2 * threads_each
        m = max(1, -(-20 * threads_each // max(K, 1)))   # batches per step and Session (>= 20 waves timed)
        need = (preroll + W + K * m + 4) * bs
        y = S.labels_by_id(torch.arange(n))
        x_blocks, idxs, caches, part_ranks = [], [], [], []
        num_cache = int(n / P * (float(args.cache_pct) / 100.0)) if str(args.cache_pct) != "auto" else int(n / P * 0.15)
        for g in range(N):
            hosted = hosted_partitions(g, N, P)
            blo, bhi = int(off[hosted[0]]), int(off[hosted[-1] + 1])
            x_blocks.append(S.features_by_id(blo, bhi, f, dt, device=dev).cpu())
            ids = S.seeds(n, min(bhi - blo, need), seed=7 + g, lo=blo, hi=bhi)
            if ids.numel() < need:
                ids = ids.repeat((need + ids.numel() - 1) // ids.numel())[:need]
            idxs.append(ids)
            # the reference's degree-ranked policy (driver/drivers/ddp.py:487-495), torch ops only:
            # highest-degree vertices outside this GPU's block, owner-major
            score = deg.clone()
            score[blo:bhi] = -1
            order = torch.sort(score, descending=True, stable=True).indices[:min(num_cache, n - (bhi - blo))]
            owner = torch.searchsorted(off.to(order.device), order, right=True) - 1
            cv = order[torch.sort(owner, stable=True).indices].cpu()
            del score, order, owner
            cf = S.expected_features(cv, f, dt)
            caches.append(R.Cache(hosted[0], P, cv, cf))
            part_ranks.append(hosted[0])
        bps, timed = run_reference_distributed(R, rowptr, col, x_blocks, y, idxs, sizes, bs, off, part_ranks, caches,
                                               threads_each, W, K * m, preroll)
        gbs = None
        kind_out = "reference-distributed"
        sample = (f"{timed} mini-batches in aggregate ({K} steps of {m} per Session) after {preroll} pre-roll + {W} warm-up batches per Session; {N} distributed "
                  f"reference Sessions (one per GPU of the run) x {threads_each} worker threads, RangePartitionBook({P} parts) + "
                  f"degree-ranked Cache ({num_cache} rows), gpu_percent 0.999: sampling + owner binning + host-resident row slice")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(bps, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": round(1000.0 * max(N, 1) / bps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc + (f"; features partitioned {P}-way" if P > 1 else ""), "graph_generator": GRAPH_GEN,
                   "scale": args.scale, "nnz": int(col.numel()), "feature_partitions": P},
        "gathered_GBps": None if gbs is None else round(gbs, 3),
        "cpu_baseline": {"value": round(bps, 3), "unit": UNIT, "cores": threads, "kind": kind_out, "sample": sample},
        "e2e": {"value": round(bps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------
def check_batch(b, idx_host, cache, feat_dim, feat_dtype, S, dev, distributed, resample=None):
    """Parity of one public-API batch against the id-function ground truth (device ops only).
    A non-distributed batch carries no n_id (the reference's PreparedSample does not either):
    `resample(seeds, stop)` re-derives it with the same deterministic generator."""
    errs = []
    if distributed:
        st, en = b.idx_range
        n_id, x, y = b.n_id, b.x, b.sliced_cpu_labels
    else:
        x, y, adjs, (st, en) = b
        n_id = None
    seeds = idx_host[st:en].to(dev)
    if n_id is None and resample is not None:
        n_id = resample(seeds, en)
    if n_id is not None:
        if not torch.equal(n_id[:en - st], seeds):
            errs.append("n_id[:bs] != seeds")
        if x is None or x.size(0) != n_id.numel() or not torch.equal(
                x.view(torch.int16 if x.element_size() == 2 else torch.int32), S._id_pattern(n_id, feat_dim, feat_dtype)):
            errs.append("x != f(n_id)")
    if distributed:
        parts = list(b.partition_nids)
        cached_global = cache.cached_vertices.to(dev)[b.cached_nids] if b.cached_nids.numel() else b.cached_nids
        cat = torch.cat(parts + [cached_global])
        if not torch.equal(cat[b.perm_partition_to_mfg], n_id):
            errs.append("cat(partition_nids, cached)[perm] != n_id")
        adjs = b.adjs
    if y is None or not torch.equal(y.view(-1, 1), S.labels_by_id(seeds)):
        errs.append("y != g(seeds)")
    # structure invariants (the sampled structure itself is pinned against the oracle in tests/)
    S_prev = None
    for rp, cl, _e, (T, Sz) in adjs:
        if rp.numel() != T + 1 or int(rp[-1]) != cl.numel() or (cl.numel() and int(cl.max()) >= Sz):
            errs.append("adjacency sizes inconsistent")
        if S_prev is not None and Sz != S_prev:
            errs.append("hop sizes do not chain")
        S_prev = T
    if not distributed and x is not None and adjs and x.size(0) != adjs[0][3][1]:
        errs.append("x rows != |n_id|")
    return errs


def ours(args):
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from salient_plusplus_b200 import _lib, fast_sampler as fs, peer, synthetic as S
    from salient_plusplus_b200.pipeline import MiniBatchPipeline
    from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
    from salient_plusplus_b200.transferers import DeviceDistributedPrefetcher, DevicePrefetcher
    lib = _lib.load()

    shape, sizes, bs, _, desc = WORKLOADS[args.workload]
    if args.feat_dtype != "default":
        desc += f" [features stored as {args.feat_dtype} for this run]"
    K, W = args.steps, args.warmup
    P = num_parts(args, world)
    n, f, dt, rowptr, col = make_graph(shape, args.scale, dev, args.locality, max(P, 1))
    col32 = col.to(torch.int32)
    del col
    log(f"graph ready: {n} nodes, {col32.numel()} CSR entries")
    y = S.labels_by_id(torch.arange(n, device=dev))
    row_bytes = f * torch.empty(0, dtype=dt).element_size()

    off = S.equal_partition_offsets(n, P)
    offl = [int(v) for v in off.tolist()]
    hosted = peer.hosted_partitions(rank, world, P)       # consecutive partitions resident on this GPU
    emulate = bool(args.emulate_peers) and world == 1 and P > 1
    if emulate:
        # profiling aid (ncu sees one GPU): this GPU plays partition 0 only; the other partitions' tables
        # also live here but are treated as peers (book search, cache probe, replicated cache) -- every
        # code path of an N-GPU rank except the NVLink hop
        hosted = [0]
    prank = hosted[0]                                      # the partition this rank acts as in the book
    blo, bhi = offl[hosted[0]], offl[hosted[-1] + 1]
    need = (W + K) * bs
    idx = S.seeds(n, min(bhi - blo, need), seed=7 + rank, device=dev, lo=blo, hi=bhi)  # federated: local seeds
    if emulate:
        blo, bhi = 0, n                                    # ... but every row is materialised here
    if idx.numel() < need:
        idx = idx.repeat((need + idx.numel() - 1) // idx.numel())[:need]
    idx_host = idx.cpu().pin_memory()

    # every rank materialises ONLY the rows of the partitions it hosts (a MAG240M-shaped table is
    # 375 GB in total); element (i, j) is a pure function of the global id i
    x_block = S.features_by_id(blo, bhi, f, dt, device=dev)
    log("features ready")
    fm = None
    cache = fs.Cache()
    part_ptrs = None
    pitch = row_bytes
    if P > 1:
        from salient_plusplus_b200 import vip as V
        btab = fs.feature_table(x_block)   # resident copy with a 128-byte-multiple row pitch when rows need one
        pitch = btab.pitch
        if world > 1:
            rank_ptrs = peer.exchange_device_tables(btab.storage)
            if rank_ptrs is None:
                raise SystemExit("peer feature tables are not reachable (no P2P between the GPUs of this box?)")
        else:
            rank_ptrs = [btab.storage.data_ptr()]
        if emulate:
            part_ptrs = [btab.storage.data_ptr() + offl[p] * pitch for p in range(P)]
        else:
            part_ptrs = peer.partition_pointers(rank_ptrs, offl, world, pitch)
        x_rank = x_block[:offl[prank + 1] - blo]
        if args.cache_policy == "vip":
            # the reference's policy (driver/drivers/ddp.py:417-446): analytic vertex-inclusion
            # probabilities of this rank's mini-batches (federated: every local vertex can be a seed)
            probs = V.vip_probabilities(rowptr, col32, torch.arange(offl[hosted[0]], offl[hosted[-1] + 1], device=dev), bs, sizes)
        else:  # degree ranking (ddp.py:487-495)
            probs = (rowptr[1:] - rowptr[:-1]).to(torch.float64)
        if str(args.cache_pct).lower() == "auto":
            # B200-first replication factor: as many of the hottest remote rows as fit in a quarter of
            # the HBM that is still free, capped at "everything" (alpha in the reference's units: ddp.py:421)
            free_b, _ = torch.cuda.mem_get_info()
            part_rows = max(1, n // P)
            args.cache_pct = round(min((P - 1) * 100.0, 100.0 * (free_b // 4) / (part_rows * pitch)), 2)
        else:
            args.cache_pct = float(args.cache_pct)
        # the replicated cache is filled by pulling the chosen rows out of the owners' partitions with
        # the P2P gather kernel (replaces the three blocking all_to_alls of ddp.py:524-551)
        ptrs_for_cache = list(part_ptrs)
        cache = V.create_vip_cache(rowptr, col32, None, bs, sizes, off, prank, args.cache_pct, x_rank,
                                   peer_table_ptrs=ptrs_for_cache, vip=probs, local_parts=hosted)
        del probs
        ctab = cache.device_table()
        fm = fs.make_feature_map(offl, prank, None, ctab.storage if ctab else None,
                                 cache.device_index(n) if ctab else None, part_ptrs, pitch,
                                 ctab.pitch if ctab else 0, local_parts=hosted)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        log(f"cache ready: {cache.cached_vertices.numel()} rows")
    use_cache = P > 1 and cache.cached_vertices.numel() > 0

    if args.no_features and args.device_only:
        pipe = MiniBatchPipeline(rowptr, col32, sizes, bs, depth=args.depth, device=dev)
    else:
        pipe = MiniBatchPipeline(rowptr, col32, sizes, bs, x_table=None if fm is not None else x_block, y_table=y,
                                 feature_map=fm, feat_dim=f, feat_dtype=dt, split=P > 1, use_cache=use_cache,
                                 depth=args.depth, device=dev)
    D = len(pipe.slots)
    main = torch.cuda.current_stream()

    def seed_ptr(b):
        return idx.data_ptr() + 8 * b * bs

    def rng_of(b):
        return ((b + 1) * bs * 17 + 5) & 0xFFFFFFFF

    def run_batches(first, count, time_gather=False, single_stream=False, count_rows=False):
        evs = []
        for i in range(count):
            b = first + i
            ev = pipe.launch(0 if single_stream else b % D, seed_ptr(b), bs, rng_of(b), time_gather, count_rows)
            if ev:
                evs.append((b, ev))
        return evs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident in HBM -------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()  # sampled from here to the end of the end-to-end loop (both timed regions)
    run_batches(0, W)
    barrier()
    launches0 = lib.spp_launch_count()
    replays0 = int(lib.spp_graph_replays())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in pipe.slots:
        s.stream.wait_event(e0)
    t_issue0 = time.perf_counter()
    run_batches(W, K)
    t_issue = time.perf_counter() - t_issue0  # host time spent issuing the K batches
    for s in pipe.slots:
        main.wait_stream(s.stream)
    e1.record(main)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.spp_launch_count() - launches0)
    replays = int(lib.spp_graph_replays()) - replays0
    if args.device_only:
        clocks.stop()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank == 0:
            print(json.dumps({"device_only": True, "batches_per_s": round(world * K / (ms * 1e-3), 1),
                              "us_per_batch": round(1000.0 * ms / K, 2), "launches": launches, "graph_replays": replays,
                              "host_issue_us_per_batch": round(1e6 * t_issue / K, 2),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("SPP_")}}), flush=True)
        if world > 1:
            dist.barrier()
        return

    # ---- roofline pass: the feature gather timed with CUDA events on its own stream -----------
    evs = run_batches(W, min(K, 64), time_gather=True, single_stream=True)
    torch.cuda.synchronize()
    g_ms = sum(ev[0].elapsed_time(ev[1]) for _b, ev in evs)
    # node counts and the local / cache / peer split of the rows, batch by batch (re-sampling is deterministic)
    nodes_total, shares = 0, [0, 0, 0]
    probe = [b for b, _ in evs[:8]]
    for b in probe:
        s0 = pipe.slots[0]
        s0.counters.zero_()
        pipe.launch(0, seed_ptr(b), bs, rng_of(b), True, fm is not None)
        nodes_total += pipe.read_meta(0)[len(sizes)]
        if fm is not None:
            c = s0.counters.tolist()
            shares = [a + int(v) for a, v in zip(shares, c)]
    mean_nodes = nodes_total / max(1, len(probe))
    rows_share = None
    if fm is not None and sum(shares) > 0:
        rows_share = {"local": round(shares[0] / len(probe), 1), "cache": round(shares[1] / len(probe), 1),
                      "peer": round(shares[2] / len(probe), 1), "unit": "rows per mini-batch (mean over %d)" % len(probe)}
    idx_bytes = 4 + (4 if P > 1 else 0)  # int32 node id (+ the int32 source descriptor of the owner split)
    alg_bytes_per_launch = mean_nodes * (2 * row_bytes + idx_bytes)
    g_ms_avg = g_ms / max(1, len(evs))
    peak, peak_src = measured_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_gather_traffic.json")
    if os.path.exists(tp) and world == 1 and args.feat_dtype == "default":
        try:  # ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum per launch) of the same workload
            traffic = json.load(open(tp)).get(f"{args.workload}_p{P}")
        except Exception:  # noqa: BLE001
            traffic = None
    achieved = alg_bytes_per_launch / (g_ms_avg * 1e-3) / 1e9 if g_ms_avg > 0 else 0.0

    # ---- public API: parity check (untimed), then e2e ------------------------------------------
    def make_iter(first, count, raw=False):
        x_rank = x_block[:offl[prank + 1] - blo] if P > 1 else x_block
        cfg = FastSamplerConfig(
            x_cpu=x_block if P == 1 else torch.empty((0, f), dtype=dt), x_gpu=x_rank if P > 1 else torch.empty((0, f), dtype=dt),
            y=y, rowptr=rowptr, col=col32, idx=idx_host[first * bs:(first + count) * bs], batch_size=bs,
            sizes=list(sizes), skip_nonfull_batch=False, pin_memory=True, distributed=P > 1,
            partition_book=fs.RangePartitionBook(prank, P, off) if P > 1 else None, cache=cache,
            force_exact_num_batches=False, exact_num_batches=0, count_remote_frequency=False, use_cache=P > 1)
        if P > 1:
            cfg.peer_table_ptrs = part_ptrs
            cfg.peer_table_pitch = pitch
            cfg.local_parts = hosted
        if raw:
            return fs.Session(16, max(args.depth, 4), cfg.to_fast_sampler())
        sampler = FastSampler(16, max(args.depth, 4), cfg)
        it = iter(sampler)
        return DeviceDistributedPrefetcher([dev], it) if P > 1 else DevicePrefetcher([dev], it)

    n_check = max(3, min(args.parity_batches, W + K))
    errs, checked = [], 0
    sess = make_iter(0, n_check, raw=True)
    while True:
        b = sess.blocking_get_batch_distributed() if P > 1 else sess.blocking_get_batch()
        if b is None:
            break
        errs += check_batch(b, idx_host, cache, f, dt, S, dev, P > 1,
                            lambda sd, stop: fs.multilayer_sample(sd, sizes, rowptr, col32, seed=(stop * 17 + 5) & 0xFFFFFFFF)[0])
        checked += 1
    del sess
    ok_local = not errs and checked == n_check
    okt = torch.tensor([1 if ok_local else 0, checked], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    parity_ok = bool(int(okt[0].item()) == 1)
    if errs:
        print(f"[bench rank {rank}] PARITY FAILED: {sorted(set(errs))}", file=sys.stderr, flush=True)
    log("parity checked")

    import gc
    gc.collect()
    gc.freeze()  # keep a generation-2 collection (10-20 ms with torch imported) out of the timed loop
    barrier()
    prof = None
    if args.profile_e2e:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    # ONE iterator over W + K batches; the clock starts when warm-up batch W has been delivered --
    # the same way the reference arm is timed (clock started on an already running Session)
    t_c0 = time.perf_counter()
    e2e_iter = make_iter(0, W + K)
    t_setup = time.perf_counter() - t_c0
    got, e2e_nodes, lat = 0, 0, []
    t0 = tl = None
    first_us = None
    for (batch,) in e2e_iter:
        got += 1
        tn = time.perf_counter()
        if got == 1:
            first_us = (tn - t_c0) * 1e6
        if got == W:
            torch.cuda.current_stream().synchronize()
            t0 = tl = time.perf_counter()
        elif got > W:
            e2e_nodes += batch.x.size(0)
            lat.append(tn - tl)
            tl = tn
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    try:  # time the consumer spent waiting for a batch that was not ready (GPU- or issue-bound share)
        st_ = e2e_iter.it.get_stats()
        blocked_us, blocked_n = st_.total_blocked_dur.total_seconds() * 1e6, int(st_.total_blocked_occasions)
    except Exception:  # noqa: BLE001
        blocked_us, blocked_n = None, None
    clk = clocks.stop()
    top3 = sorted(range(len(lat)), key=lambda i: -lat[i])[:3]
    top3 = [(i, round(lat[i] * 1e6)) for i in top3]
    lat.sort()
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    if world > 1:
        t = torch.tensor([ms, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = t.tolist()
    assert got == W + K

    # ---- NVLink roofline (N > 1): the P2P miss fetch alone, every row owned by a peer GPU --------
    nvlink = None
    if world > 1:
        import ctypes as _ct
        rows_n = 1_000_000
        gq = torch.Generator(device=dev).manual_seed(99 + rank)
        remote = torch.tensor([p for p in range(P) if p not in hosted], device=dev)
        owner = remote[torch.randint(0, remote.numel(), (rows_n,), generator=gq, device=dev)]
        offd = off.to(dev)
        span = (offd[1:] - offd[:-1])[owner]
        r = (torch.rand(rows_n, generator=gq, device=dev, dtype=torch.float64) * span.double()).long()
        ids = (offd[owner] + torch.minimum(r, span - 1).clamp_(min=0)).to(torch.int64)
        fm_nc = fs.make_feature_map(offl, prank, None, None, None, part_ptrs, pitch, 0, local_parts=hosted)  # no cache
        outb = torch.empty((rows_n, f), dtype=dt, device=dev)
        spn = main.cuda_stream
        for _ in range(3):
            _lib.check(lib.spp_gather_partitioned(_ct.byref(fm_nc), row_bytes, ids.data_ptr(), 1, rows_n, None, None,
                                                  outb.data_ptr(), rows_n, None, spn))
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        n0.record(main)
        for _ in range(reps):
            _lib.check(lib.spp_gather_partitioned(_ct.byref(fm_nc), row_bytes, ids.data_ptr(), 1, rows_n, None, None,
                                                  outb.data_ptr(), rows_n, None, spn))
        n1.record(main)
        torch.cuda.synchronize()
        nv_ms = n0.elapsed_time(n1) / reps
        nv_ok = bool(torch.equal(outb.view(torch.int16 if outb.element_size() == 2 else torch.int32),
                                 S._id_pattern(ids, f, dt)))
        tms = torch.tensor([nv_ms, 0.0 if nv_ok else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        nv_ms = float(tms[0].item())
        parity_ok = parity_ok and float(tms[1].item()) == 0.0
        gbs = rows_n * row_bytes / (nv_ms * 1e-3) / 1e9
        nvlink = {"bound": "nvlink", "kernel": "k_gather partitioned, all rows on peer GPUs (P2P miss fetch)",
                  "achieved": round(gbs, 1), "peak": 900.0, "unit": "GB/s inbound per GPU", "frac": round(gbs / 900.0, 4),
                  "measured_peer_copy_peak": 770.0, "frac_of_measured": round(gbs / 770.0, 4),
                  "rows": rows_n, "row_bytes": row_bytes, "ms": round(nv_ms, 4), "timing": "CUDA events, max over ranks",
                  "rows_verified": nv_ok}
        del outb, ids

    # ---- CPU baseline beside it (rank 0, N = 1): the reference fast_sampler on the host cores --
    cpu = None
    oracle_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = ref_threads(args)
        cb = max(96, 10 * threads)   # >= 10 waves of the thread pool, bounded by max_seconds below
        pre = 2 * threads
        rp_h, col_h = rowptr.cpu(), col32.to(torch.int64).cpu()
        x_h = x_block.cpu()
        log("host copies ready")
        # the checker: one batch of the timed workload against the CPU restatement of the algorithm
        # (counter-RNG mode), bit for bit -- n_id, every hop's rowptr / col
        try:
            import numpy as np
            from oracle import oracle as O
            bq = W  # first timed batch
            on, oa = O.multilayer_sample(idx_host[bq * bs:(bq + 1) * bs].numpy(), sizes, rp_h.numpy(), col_h.numpy(),
                                         rng_mode=O.RNG_COUNTER, rng_seed=rng_of(bq))
            pipe.launch(0, seed_ptr(bq), bs, rng_of(bq))
            meta = pipe.read_meta(0)
            s0 = pipe.slots[0]
            L_ = len(sizes)
            good = meta[L_] == on.size and np.array_equal(s0.ws.n_ids[:meta[L_]].cpu().numpy(), on)
            for h in range(L_):
                T_, E_ = meta[h], meta[12 + h]
                w = oa[L_ - 1 - h]
                good = good and np.array_equal(s0.rowptrs[h][:T_ + 1].cpu().numpy(), w[0]) and \
                    np.array_equal(s0.cols[h][:E_].cpu().numpy(), w[1])
            oracle_check = bool(good)
            parity_ok = parity_ok and oracle_check
        except Exception as e:  # noqa: BLE001
            oracle_check = f"not run: {e}"
        bps, gbs, timed, kind, _ = run_reference_cpu(rp_h, col_h, x_h, y.cpu(), idx_host[:(cb + pre + 8) * bs].clone()
                                                     if (cb + pre + 8) * bs <= idx_host.numel() else
                                                     idx_host.repeat(((cb + pre + 8) * bs + idx_host.numel() - 1) // idx_host.numel())[:(cb + pre + 8) * bs],
                                                     sizes, bs, threads, 8, cb, max_seconds=25.0, preroll=pre)
        cpu = {"value": round(bps, 3), "unit": UNIT, "cores": threads, "kind": kind, "gathered_GBps": round(gbs, 3),
               "sample": f"{timed} mini-batches of the same workload after {pre} pre-roll + 8 warm-up batches, reference "
                         f"fast_sampler.Session (non-distributed: every row local), {threads} worker threads, pinned outputs "
                         f"(sampling + CPU feature slice)"}

    if rank == 0:
        value = world * K / (ms * 1e-3)
        e2e_v = world * K / t_e2e
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": desc + (f"; features partitioned {P}-way ({len(hosted)} partition(s) per GPU), "
                                           f"{args.cache_pct}% replicated {args.cache_policy}-ranked cache, P2P gather over NVLink"
                                           if P > 1 else ""),
                       "graph_generator": GRAPH_GEN + (f", partition locality {args.locality}" if args.locality > 0 else ""),
                       "features": "element (i, j) = pure function of the global vertex id (synthetic.features_by_id)",
                       "nnz": int(col32.numel()), "mean_nodes_per_batch": round(mean_nodes, 1),
                       "streams_in_flight": D, "scale": args.scale, "feature_partitions": P,
                       "partitions_per_gpu": len(hosted), "emulated_peers": emulate, "cache_rows": int(cache.cached_vertices.numel()),
                       "l2": "inputs larger than L2 (feature table + CSR >> 126 MB, random rows)"},
            "gathered_GBps": round(value * mean_nodes * row_bytes / 1e9, 2),
            "parity": {"ok": parity_ok, "batches_checked_per_rank": int(okt[1].item()),
                       "checks": "x == f(n_id), y == g(seeds), n_id[:bs] == seeds, cat(partition_nids, cached)[perm] == n_id, "
                                 "adjacency sizes (public Session API, every rank)"
                                 + ("; all-remote P2P rows == f(id)" if world > 1 else ""),
                       "oracle_batch_bit_exact": oracle_check},
            "e2e": {"value": round(e2e_v, 2), "unit": UNIT, "h2d_bytes_per_step": bs * 8,
                    "d2h_bytes_per_step": 8 * 32 + (8 * 18 if P > 1 else 0),
                    "gathered_GBps": round(world * e2e_nodes * row_bytes / t_e2e / 1e9, 2),
                    "api": "FastSampler -> " + ("DeviceDistributedPrefetcher" if P > 1 else "DevicePrefetcher"),
                    "timing": f"one iterator over {W}+{K} batches, clock from the delivery of warm-up batch {W}",
                    "per_batch_us": {"p50": round(lat[len(lat) // 2] * 1e6, 1), "p90": round(lat[int(len(lat) * 0.9)] * 1e6, 1),
                                     "max": round(lat[-1] * 1e6, 1), "first_batch_after_iter_creation": round(first_us, 1),
                                     "setup": round(t_setup * 1e6, 1), "slowest_iters": top3,
                                     "consumer_blocked_us_total": None if blocked_us is None else round(blocked_us, 1),
                                     "consumer_blocked_batches": blocked_n}},
            "gpu_launches": launches,
            "cuda_graph_replays": replays,
            "host_issue_us_per_batch": round(1e6 * t_issue / K, 2),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_gather (feature gather" + (", partitioned: local + cache + peer rows)" if P > 1 else ")"),
                         "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "avg_launch_ms": round(g_ms_avg, 5),
                         "algorithmic_bytes_per_launch": int(alg_bytes_per_launch),
                         "bytes_model": f"N_b * (2*row_bytes + {idx_bytes})", "launches_timed": len(evs),
                         "frac_of_nominal_8TBps": round(achieved / 8000.0, 4)},
        }
        if rows_share is not None:
            line["rows_served"] = rows_share
        if nvlink is not None:
            line["nvlink_roofline"] = nvlink
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


def layerwise(args):
    """BASELINE configs[2]: layer-wise full-neighbourhood inference batches (sizes [-1],
    driver/models.py:441-495) over the products-shaped graph at 1 / 2 GPUs -- every batch is 1024
    CONSECUTIVE vertices, one full-neighbourhood hop, feature gather of every node; the vertex
    range is split by rank, graph and features are replicated (0.5 GB).  Deterministic, so rank 0
    checks whole batches bit for bit against the oracle at N = 1."""
    import torch.distributed as dist
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from salient_plusplus_b200 import _lib, fast_sampler as fs, synthetic as S
    from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
    from salient_plusplus_b200.transferers import DevicePrefetcher
    lib = _lib.load()
    shape, sizes, bs, _, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    n, f, dt, rowptr, col = make_graph(shape, args.scale, dev, 0.0, 1)
    col32 = col.to(torch.int32)
    x = S.features_by_id(0, n, f, dt, device=dev)
    y = S.labels_by_id(torch.arange(n, device=dev))
    row_bytes = f * x.element_size()
    lo, hi = n * rank // world, n * (rank + 1) // world
    need = (W + K) * bs
    idx = torch.arange(lo, min(hi, lo + need), dtype=torch.int64, device=dev)
    if idx.numel() < need:
        idx = idx.repeat((need + idx.numel() - 1) // idx.numel())[:need]
    idx_host = idx.cpu().pin_memory()

    def make_iter(seeds, raw=False):
        cfg = FastSamplerConfig(x_cpu=x, x_gpu=torch.empty((0, f), dtype=dt), y=y, rowptr=rowptr, col=col32, idx=seeds,
                                batch_size=bs, sizes=list(sizes), skip_nonfull_batch=False, pin_memory=True,
                                distributed=False)
        if raw:
            return fs.Session(16, max(args.depth, 4), cfg.to_fast_sampler())
        return DevicePrefetcher([dev], iter(FastSampler(16, max(args.depth, 4), cfg)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(seeds):
        """One iterator over W + K batches, clock (host and CUDA events) from the delivery of batch W."""
        t_c0 = time.perf_counter()
        it = make_iter(seeds)
        got, nodes, edges, lat = 0, 0, 0, []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = tl = first = None
        for (batch,) in it:
            got += 1
            tn = time.perf_counter()
            if got == 1:
                first = (tn - t_c0) * 1e6
            if got == W:
                torch.cuda.current_stream().synchronize()
                e0.record()
                t0 = tl = time.perf_counter()
            elif got > W:
                nodes += batch.x.size(0)
                edges += batch.adjs[0].adj_t.nnz() if hasattr(batch.adjs[0], "adj_t") else 0
                lat.append(tn - tl)
                tl = tn
        e1.record()
        torch.cuda.synchronize()
        return time.perf_counter() - t0, e0.elapsed_time(e1), nodes, edges, lat, first, it

    clocks = ClockSampler(local)
    clocks.start()
    for _ in make_iter(idx[:max(W, 3) * bs]):   # allocator / slot pool warm-up
        pass
    barrier()
    launches0 = lib.spp_launch_count()
    caps0, mem0 = int(lib.spp_graph_captures()), torch.cuda.memory_stats()
    # value: seeds resident in HBM (the Session uses a device idx in place: no H2D)
    _, ms, nodes_v, edges_v, lat_v, _, _ = timed_loop(idx)
    mem1 = torch.cuda.memory_stats()
    diag = {"graph_captures_in_timed_loop": int(lib.spp_graph_captures()) - caps0,
            "device_allocs_in_timed_loop": int(mem1.get("num_device_alloc", 0) - mem0.get("num_device_alloc", 0)),
            "slowest_batch_us": round(max(lat_v) * 1e6, 1) if lat_v else None}
    launches = int(lib.spp_launch_count() - launches0) * K // (W + K)
    barrier()
    # per-kernel time of the dominant kernel: event trace of a few batches at depth 1
    os.environ["SPP_SESSION_DEPTH"] = "1"
    _lib.trace_begin(64 * 16)
    for _ in make_iter(idx[W * bs:(W + 16) * bs]):
        pass
    marks = _lib.trace_end(64 * 16)
    os.environ.pop("SPP_SESSION_DEPTH", None)
    dur, prev = {}, {}
    for lab, _hop, st_, t_ms in marks:
        if lab != "batch_begin" and st_ in prev:
            dur.setdefault(lab, []).append(t_ms - prev[st_])
        prev[st_] = t_ms
    k_ms = {k_: sum(v) / len(v) for k_, v in dur.items()}
    # parity: x == f(n_id), y == g(seeds) through the public API on every rank
    errs, checked = [], 0
    sess = make_iter(idx_host[:3 * bs], raw=True)
    while True:
        b = sess.blocking_get_batch()
        if b is None:
            break
        errs += check_batch(b, idx_host, None, f, dt, S, dev, False,
                            lambda sd, stop: fs.multilayer_sample(sd, sizes, rowptr, col32)[0])
        checked += 1
    del sess
    okt = torch.tensor([0 if errs else 1, checked], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    parity_ok = bool(int(okt[0].item()) == 1)
    if errs:
        print(f"[bench rank {rank}] PARITY FAILED: {sorted(set(errs))}", file=sys.stderr, flush=True)
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    # e2e: seeds in (pinned) host memory, H2D per batch + D2H of the size block
    t_e2e, _, nodes_e, _, lat, first_us, it = timed_loop(idx_host)
    try:  # time the consumer spent waiting for a batch that was not ready (GPU-bound share of the loop)
        st_ = it.it.get_stats()
        blocked_us, blocked_n = round(st_.total_blocked_dur.total_seconds() * 1e6, 1), int(st_.total_blocked_occasions)
    except Exception:  # noqa: BLE001
        blocked_us, blocked_n = None, None
    clk = clocks.stop()
    lat.sort()
    if world > 1:
        t = torch.tensor([ms, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = t.tolist()
    cpu, oracle_check = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import numpy as np
        from oracle import oracle as O
        threads = ref_threads(args)
        rp_h, col_h = rowptr.cpu(), col32.to(torch.int64).cpu()
        good = True
        for q in range(3):   # whole batches, bit for bit (deterministic path)
            sd = idx_host[(W + q) * bs:(W + q + 1) * bs]
            on, oa = O.multilayer_sample(sd.numpy(), sizes, rp_h.numpy(), col_h.numpy())
            n_id, adjs = fs.multilayer_sample(sd, sizes, rowptr, col32)
            good = good and np.array_equal(n_id.cpu().numpy(), on) and \
                np.array_equal(adjs[0][0].cpu().numpy(), oa[0][0]) and np.array_equal(adjs[0][1].cpu().numpy(), oa[0][1])
        oracle_check = bool(good)
        parity_ok = parity_ok and oracle_check
        pre = 2 * threads
        cb = max(256, 10 * threads)
        seeds_c = torch.arange(0, min(n, (cb + pre + 8) * bs), dtype=torch.int64)
        bps, gbs, timed, kind, _ = run_reference_cpu(rp_h, col_h, x.cpu(), y.cpu(), seeds_c, sizes, bs, threads, 8, cb,
                                                     max_seconds=25.0, preroll=pre)
        cpu = {"value": round(bps, 3), "unit": UNIT, "cores": threads, "kind": kind, "gathered_GBps": round(gbs, 3),
               "sample": f"{timed} layer-wise batches (consecutive vertices from 0) after {pre} pre-roll + 8 warm-up batches, "
                         f"reference fast_sampler.Session, {threads} worker threads, pinned outputs"}
    if rank == 0:
        value = world * K / (ms * 1e-3)
        mean_nodes, mean_edges = nodes_v / K, edges_v / K
        peak, peak_src = measured_peaks()
        # dominant kernel of this workload: the edge-parallel full-neighbourhood sampler
        # (k_hop_sample_edges): per candidate a 4-byte col read + an 8-byte slot write, plus the
        # rowptr pair and the scanned out_rowptr per target
        s_ms = k_ms.get("sample", 0.0)
        alg = mean_edges * (4 + 8) + bs * (16 + 8)
        ach = alg / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": desc, "graph_generator": GRAPH_GEN, "nnz": int(col32.numel()),
                       "features": "element (i, j) = pure function of the global vertex id (synthetic.features_by_id)",
                       "mean_nodes_per_batch": round(mean_nodes, 1), "mean_edges_per_batch": round(mean_edges, 1),
                       "scale": args.scale, "l2": "inputs larger than L2 (feature table + CSR >> 126 MB)",
                       "value_timing": "public Session API with the seeds resident in HBM, CUDA events from the delivery of "
                                       "warm-up batch W (edge counts are data dependent, so there is no static-buffer pipeline)"},
            "gathered_GBps": round(value * mean_nodes * row_bytes / 1e9, 2),
            "parity": {"ok": parity_ok, "batches_checked_per_rank": int(okt[1].item()),
                       "checks": "x == f(n_id), y == g(seeds), adjacency sizes (public Session API, every rank)",
                       "oracle_batches_bit_exact": oracle_check},
            "e2e": {"value": round(world * K / t_e2e, 2), "unit": UNIT, "h2d_bytes_per_step": bs * 8, "d2h_bytes_per_step": 8 * 32,
                    "gathered_GBps": round(world * nodes_e * row_bytes / t_e2e / 1e9, 2), "api": "FastSampler -> DevicePrefetcher",
                    "timing": f"one iterator over {W}+{K} batches, clock from the delivery of warm-up batch {W}",
                    "per_batch_us": {"p50": round(lat[len(lat) // 2] * 1e6, 1), "p90": round(lat[int(len(lat) * 0.9)] * 1e6, 1),
                                     "max": round(lat[-1] * 1e6, 1), "first_batch_after_iter_creation": round(first_us, 1),
                                     "consumer_blocked_us_total": blocked_us, "consumer_blocked_batches": blocked_n}},
            "gpu_launches": launches, "clocks": clk,
            "kernel_ms": {k_: round(v, 5) for k_, v in sorted(k_ms.items())},
            "value_loop_diagnostics": diag,
            "roofline": {"bound": "hbm", "kernel": "k_hop_sample_edges (edge-parallel full-neighbourhood sampling + id-table insert)",
                         "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                         "peak_source": peak_src, "avg_launch_ms": round(s_ms, 5), "algorithmic_bytes_per_launch": int(alg),
                         "bytes_model": "E * (4 + 8) + T * 24; latency / L2-atomic bound, not bandwidth bound (DESIGN.md section 4)"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (testing only)")
    ap.add_argument("--depth", type=int, default=6, help="mini-batches in flight (CUDA streams)")
    ap.add_argument("--cache-pct", default="15.0",
                    help="replicated rows per GPU in %% of a partition (the reference's alpha, default its documented "
                         "15); 'auto' = whatever fits a quarter of the free HBM, up to every remote row")
    ap.add_argument("--parts", type=int, default=0,
                    help="feature partitions (default: 8 for the papers100M / MAG240M shapes, else one per GPU)")
    ap.add_argument("--cache-policy", default="vip", choices=["vip", "degree"])
    ap.add_argument("--locality", type=float, default=0.0,
                    help="probability that an edge stays inside its source's partition block (0 = locality-free "
                         "Chung-Lu graph, the worst case for range-partitioned features)")
    ap.add_argument("--parity-batches", type=int, default=3, help="batches per rank verified outside the timed regions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-peers", action="store_true",
                    help="one GPU only, profiling aid: act as partition 0 of --parts and treat the other partitions "
                         "(resident on the same GPU) as peers, with the replicated cache in front of them")
    ap.add_argument("--no-features", action="store_true",
                    help="A/B experiments with --device-only: sampler alone (no feature / label gather)")
    ap.add_argument("--device-only", action="store_true",
                    help="A/B experiments: print the device-timed batches/s and exit (not a bench line)")
    ap.add_argument("--profile-e2e", action="store_true", help="cProfile the public-API loop (stderr)")
    ap.add_argument("--feat-dtype", default="default", choices=["default", "fp16", "fp32"],
                    help="feature dtype of the synthetic table (default: the shape's own; the reference stores "
                         "ogbn-products as fp16, driver/dataset.py:70)")
    ap.add_argument("--ref-threads", type=int, default=0,
                    help="worker threads of the CPU reference (cpu_baseline leg and --impl reference); 0 = every host "
                         "core, 15 = the reference launcher's default per GPU (utils/exp_driver.py:48)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global FEAT_DTYPE
    FEAT_DTYPE = {"default": None, "fp16": torch.float16, "fp32": torch.float32}[args.feat_dtype]
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "products-layerwise":
        layerwise(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
