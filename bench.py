#!/usr/bin/env python
"""bench.py -- sampled-and-gathered mini-batches/s (+ gathered feature GB/s) of the mini-batch
generation path, on synthetic graphs of the BASELINE shapes.

    python bench.py --gpus 1 --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU sampler

A "step" is one mini-batch: 1024 seeds -> 3-hop (15,10,5) neighbour sampling with dedup/relabel
-> feature gather of every sampled node (+ label gather).  At N=1 the workload is BASELINE.json
configs[1] (ogbn-products-shaped, 100-d fp16).  At N>1 (torchrun, one rank per GPU) the graph
is replicated, the features are range-partitioned N ways with a replicated hot-vertex cache and
the rows of other partitions are read over NVLink by the fused P2P gather (weak scaling: every
rank runs K batches).

Printed JSON (rank 0, one line): see the contract in the task statement; `value` is the
device-timed throughput with every input resident in HBM, `e2e` goes through the public
FastSampler/DevicePrefetcher API with the seeds in pinned host memory (H2D of the seeds and
D2H of the batch's size block inside the timed region, every step).
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
    # name: (dataset shape, fanout, batch, description)
    "products": ("products", [15, 10, 5], 1024,
                 "ogbn-products-shaped synthetic (2.45M nodes, 61.9M directed edges symmetrised, 100-d fp16), "
                 "fanout (15,10,5), batch 1024"),
    "arxiv": ("arxiv", [15, 10, 5], 1024,
              "ogbn-arxiv-shaped synthetic (169K nodes, 1.17M directed edges symmetrised, 128-d fp32), "
              "fanout (15,10,5), batch 1024"),
    "mag240m": ("mag240m", [25, 15], 1024,
                "MAG240M-shaped homogeneous synthetic (244M nodes, 1.7B CSR entries = 0.85B directed edges symmetrised, "
                "768-d fp16), fanout (25,15), batch 1024"),
    "papers100M": ("papers100M", [15, 10, 5], 1024,
                   "ogbn-papers100M-shaped synthetic (111M nodes, 1.6B CSR entries = 0.8B directed edges symmetrised, "
                   "128-d fp16), fanout (15,10,5), batch 1024"),
}
METRIC = "sampled_and_gathered_minibatches_per_sec"
UNIT = "batches/s"


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


def make_graph(shape: str, scale: float, device, locality: float = 0.0, parts: int = 8):
    from salient_plusplus_b200 import synthetic as S
    n, e, f, dt = S.SHAPES[shape]
    if shape in ("papers100M", "mag240m"):
        e //= 2  # BASELINE's 1.6B is taken as the CSR entry count ("say which")
    n, e = max(1024, int(n * scale)), max(4096, int(e * scale))
    rowptr, col = S.powerlaw_graph(n, e, seed=1, device=device, locality=locality, locality_parts=parts)
    return n, f, dt, rowptr, col


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the box's host cores
# ------------------------------------------------------------------------------------------------
def run_reference_cpu(rowptr, col, x, y, idx, sizes, bs, threads, warmup, steps, max_seconds=None):
    """Drain a reference fast_sampler.Session (oracle/_ref, built from /root/reference by
    oracle/build_ref.sh).  Returns (batches/s, gathered GB/s, batches timed, kind)."""
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
        got, nodes = 0, 0
        t0 = None
        t_start = time.perf_counter()
        while True:
            if got == warmup:
                t0 = time.perf_counter()
                nodes = 0
            b = sess.blocking_get_batch()
            if b is None:
                break
            got += 1
            if got > warmup:
                nodes += b[0].size(0)
            if got >= warmup + steps:
                break
            if max_seconds is not None and t0 is not None and time.perf_counter() - t0 > max_seconds and got > warmup + 8:
                break
        t1 = time.perf_counter()
        timed = got - warmup
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


def reference_arm(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    from salient_plusplus_b200 import synthetic as S
    shape, sizes, bs, desc = WORKLOADS[args.workload]
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n, f, dt, rowptr, col = make_graph(shape, args.scale, dev, args.locality, max(args.gpus, args.parts, 1))
    rowptr, col = rowptr.cpu(), col.cpu()
    x = S.features(n, f, dt, seed=2, device=dev).cpu()
    y = S.labels(n, seed=3)
    need = (args.warmup + args.steps) * bs
    idx = S.seeds(n, min(n, need), seed=7)
    if idx.numel() < need:
        idx = idx.repeat((need + idx.numel() - 1) // idx.numel())[:need]
    threads = os.cpu_count() or 1
    bps, gbs, timed, kind, _ = run_reference_cpu(rowptr, col, x, y, idx, sizes, bs, threads, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(bps, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": timed, "warmup": args.warmup, "ms_per_step": round(1000.0 / bps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc, "graph_generator": "Chung-Lu power law gamma=2.5 head_offset=100 seed=1",
                   "scale": args.scale, "nnz": int(col.numel())},
        "gathered_GBps": round(gbs, 3),
        "cpu_baseline": {"value": round(bps, 3), "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{timed} mini-batches after {args.warmup} warm-up, reference fast_sampler.Session "
                                   f"with {threads} worker threads (sampling + CPU feature slice, pinned outputs)"},
        "e2e": {"value": round(bps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from salient_plusplus_b200 import _lib, fast_sampler as fs, synthetic as S
    from salient_plusplus_b200.pipeline import MiniBatchPipeline
    from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
    from salient_plusplus_b200.transferers import DeviceDistributedPrefetcher, DevicePrefetcher
    lib = _lib.load()

    shape, sizes, bs, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    n, f, dt, rowptr, col = make_graph(shape, args.scale, dev, args.locality, max(world, args.parts, 1))
    col32 = col.to(torch.int32)
    del col
    y = S.labels(n, seed=3, device=dev)
    row_bytes = f * torch.empty(0, dtype=dt).element_size()

    P = max(world, args.parts)
    if P % world != 0:
        raise SystemExit("--parts must be a multiple of the number of GPUs")
    if P > world and world > 1:
        raise SystemExit("more partitions than GPUs is only emulated on a single GPU")
    off = S.equal_partition_offsets(n, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    need = (W + K) * bs
    idx = S.seeds(n, min(hi - lo, need), seed=7 + rank, device=dev, lo=lo, hi=hi)  # federated: local seeds
    if idx.numel() < need:
        idx = idx.repeat((need + idx.numel() - 1) // idx.numel())[:need]
    idx_host = idx.cpu().pin_memory()

    fm = None
    cache = fs.Cache()
    part_tensors = None
    ptrs = None
    if world > 1:
        # one process per GPU: every rank materialises ONLY its own feature partition (a
        # MAG240M-shaped table is 375 GB in total) ...
        x_local = S.features(hi - lo, f, dt, seed=2 + rank, device=dev)
    else:
        x_full = S.features(n, f, dt, seed=2, device=dev)
        x_local = x_full if P == 1 else x_full[lo:hi].clone()
    if P > 1:
        from salient_plusplus_b200 import peer, vip as V
        ltab = fs.feature_table(x_local)      # resident copy, 128-byte-multiple row pitch
        if world > 1:
            ptrs = peer.exchange_partition_tables(ltab.storage, rank, P)
            ptrs[rank] = 0
        else:  # every partition lives on this GPU (single-GPU point of a partitioned config)
            part_tensors = [x_full[int(off[p]):int(off[p + 1])] if p != rank else None for p in range(P)]
        probs = None
        if args.cache_policy == "vip":
            # the reference's policy (driver/drivers/ddp.py:417-446): analytic vertex-inclusion
            # probabilities of this rank's mini-batches (federated: every local vertex can be a seed)
            probs = V.vip_probabilities(rowptr, col32, torch.arange(lo, hi, device=dev), bs, sizes)
        else:  # degree ranking (ddp.py:487-495)
            probs = (rowptr[1:] - rowptr[:-1]).to(torch.float64)
        # ... and fills its replicated cache by pulling the chosen rows out of the owners'
        # partitions with the P2P gather kernel (replaces the three blocking all_to_alls of
        # ddp.py:524-551)
        if str(args.cache_pct).lower() == "auto":
            # B200-first replication factor: as many of the hottest remote rows as fit in a fixed share
            # (a quarter) of the HBM that is still free, capped at "everything" (alpha = (P-1) * 100 %
            # of a partition, in the reference's units: ddp.py:421)
            free_b, _ = torch.cuda.mem_get_info()
            part_rows = max(1, n // P)
            args.cache_pct = round(min((P - 1) * 100.0, 100.0 * (free_b // 4) / (part_rows * ltab.pitch)), 2)
        else:
            args.cache_pct = float(args.cache_pct)
        cache = V.create_vip_cache(rowptr, col32, None, bs, sizes, off, rank, args.cache_pct, x_local,
                                   partition_tables=part_tensors, peer_table_ptrs=ptrs, vip=probs)
        del probs
        ctab = cache.device_table()
        tables = [None] * P
        tables[rank] = ltab.storage
        if world == 1:
            ptrs = [0] * P
            for p in range(P):
                if p != rank:
                    tables[p] = fs.feature_table(part_tensors[p]).storage
        fm = fs.make_feature_map(off.tolist(), rank, tables, ctab.storage if ctab else None,
                                 cache.device_map(n) if ctab else None, ptrs, ltab.pitch, ctab.pitch if ctab else 0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    if args.no_features and args.device_only:
        pipe = MiniBatchPipeline(rowptr, col32, sizes, bs, depth=args.depth, device=dev)
    else:
        pipe = MiniBatchPipeline(rowptr, col32, sizes, bs, x_table=None if fm is not None else x_local, y_table=y,
                                 feature_map=fm, feat_dim=f, feat_dtype=dt, split=False, depth=args.depth, device=dev)
    D = len(pipe.slots)
    main = torch.cuda.current_stream()

    def seed_ptr(b):
        return idx.data_ptr() + 8 * b * bs

    def run_batches(first, count, time_gather=False, single_stream=False):
        evs = []
        for i in range(count):
            b = first + i
            ev = pipe.launch(0 if single_stream else b % D, seed_ptr(b), bs, ((b + 1) * bs * 17 + 5) & 0xFFFFFFFF,
                             time_gather)
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
    if args.device_only:
        clocks.stop()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank == 0:
            print(json.dumps({"device_only": True, "batches_per_s": round(world * K / (ms * 1e-3), 1),
                              "us_per_batch": round(1000.0 * ms / K, 2), "launches": launches,
                              "host_issue_us_per_batch": round(1e6 * t_issue / K, 2),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("SPP_")}}), flush=True)
        if world > 1:
            dist.barrier()
        return

    # ---- roofline pass: the feature gather timed with CUDA events on its own stream -----------
    evs = run_batches(W, min(K, 64), time_gather=True, single_stream=True)
    torch.cuda.synchronize()
    g_ms, g_bytes, nodes_total = 0.0, 0, 0
    # N_b of every batch of the pass: re-sample is deterministic, so read N_b batch by batch
    for b, ev in evs:
        g_ms += ev[0].elapsed_time(ev[1])
    for b, _ in evs[:8]:
        pipe.launch(0, seed_ptr(b), bs, ((b + 1) * bs * 17 + 5) & 0xFFFFFFFF)
        nodes_total += pipe.read_meta(0)[len(sizes)]
    mean_nodes = nodes_total / max(1, min(8, len(evs)))
    idx_bytes = 4  # the gather reads the sampler's int32 node list
    alg_bytes_per_launch = mean_nodes * (2 * row_bytes + idx_bytes)
    g_ms_avg = g_ms / max(1, len(evs))
    peak, peak_src = measured_peaks()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_gather_traffic.json")
    if args.workload == "products" and P == 1 and os.path.exists(tp):
        try:
            traffic = int(json.load(open(tp))["traffic_bytes_per_launch"])  # ncu --set full, dram rd + wr
        except Exception:  # noqa: BLE001
            traffic = None
    achieved = alg_bytes_per_launch / (g_ms_avg * 1e-3) / 1e9 if g_ms_avg > 0 else 0.0

    # ---- e2e: public API, seeds in pinned host memory, per-step H2D + D2H ---------------------
    def make_iter(first, count):
        tc0 = time.perf_counter()
        cfg = FastSamplerConfig(
            x_cpu=x_local if P == 1 else torch.empty((0, f), dtype=dt), x_gpu=x_local if P > 1 else torch.empty((0, f), dtype=dt),
            y=y, rowptr=rowptr, col=col32, idx=idx_host[first * bs:(first + count) * bs], batch_size=bs,
            sizes=list(sizes), skip_nonfull_batch=False, pin_memory=True, distributed=P > 1,
            partition_book=fs.RangePartitionBook(rank, P, off) if P > 1 else None, cache=cache,
            force_exact_num_batches=False, exact_num_batches=0, count_remote_frequency=False, use_cache=P > 1)
        if P > 1 and world > 1:
            cfg.peer_table_ptrs = ptrs
            cfg.peer_table_pitch = ltab.pitch
        elif P > 1:
            cfg.partition_tables = part_tensors
        ta = time.perf_counter()
        sampler = FastSampler(16, max(args.depth, 4), cfg)
        it = iter(sampler)
        tb = time.perf_counter()
        r = (DeviceDistributedPrefetcher([dev], it) if P > 1 else DevicePrefetcher([dev], it))
        if os.environ.get("SPP_DEBUG_TIMING"):
            print("[bench] make_iter: cfg %.0f us, sampler+session %.0f us, prefetcher(first batch) %.0f us" % (
                (ta - tc0) * 1e6, (tb - ta) * 1e6, (time.perf_counter() - tb) * 1e6), file=sys.stderr, flush=True)
        return r

    import gc
    gc.collect()
    gc.freeze()  # keep a generation-2 collection (10-20 ms with torch imported) out of the timed loop
    for _ in make_iter(0, max(W, 3)):  # warm-up through the same public API, right before the timed loop
        pass
    barrier()
    prof = None
    if args.profile_e2e:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    got, e2e_nodes = 0, 0
    lat, tl = [], t0
    e2e_iter = make_iter(W, K)
    t_setup = time.perf_counter() - t0
    for (batch,) in e2e_iter:
        got += 1
        e2e_nodes += batch.x.size(0)
        tn = time.perf_counter()
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
    first_us = lat[0] * 1e6
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
    assert got == K

    # ---- NVLink roofline (N > 1): the P2P miss fetch alone, every row owned by a peer ------------
    nvlink = None
    if world > 1:
        import ctypes as _ct
        rows_n = 1_000_000
        gq = torch.Generator(device=dev).manual_seed(99 + rank)
        owner = (rank + 1 + torch.randint(0, P - 1, (rows_n,), generator=gq, device=dev)) % P
        offd = off.to(dev)
        span = (offd[1:] - offd[:-1])[owner]
        r = (torch.rand(rows_n, generator=gq, device=dev, dtype=torch.float64) * span.double()).long()
        ids = (offd[owner] + torch.minimum(r, span - 1).clamp_(min=0)).to(torch.int64)
        fm_nc = fs.make_feature_map(off.tolist(), rank, tables, None, None, ptrs, ltab.pitch, 0)  # no cache
        outb = torch.empty((rows_n, f), dtype=dt, device=dev)
        spn = main.cuda_stream
        for _ in range(3):
            _lib.check(lib.spp_gather_partitioned(_ct.byref(fm_nc), row_bytes, ids.data_ptr(), 1, rows_n, None,
                                                  outb.data_ptr(), rows_n, None, spn))
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        n0.record(main)
        for _ in range(reps):
            _lib.check(lib.spp_gather_partitioned(_ct.byref(fm_nc), row_bytes, ids.data_ptr(), 1, rows_n, None,
                                                  outb.data_ptr(), rows_n, None, spn))
        n1.record(main)
        torch.cuda.synchronize()
        nv_ms = n0.elapsed_time(n1) / reps
        tms = torch.tensor([nv_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        nv_ms = float(tms.item())
        gbs = rows_n * row_bytes / (nv_ms * 1e-3) / 1e9
        nvlink = {"bound": "nvlink", "kernel": "k_gather partitioned, all rows on peers (P2P miss fetch)",
                  "achieved": round(gbs, 1), "peak": 900.0, "unit": "GB/s inbound per GPU", "frac": round(gbs / 900.0, 4),
                  "measured_peer_copy_peak": 770.0, "frac_of_measured": round(gbs / 770.0, 4),
                  "rows": rows_n, "row_bytes": row_bytes, "ms": round(nv_ms, 4), "timing": "CUDA events, max over ranks"}
        del outb, ids

    # ---- CPU baseline beside it (rank 0, N = 1): the reference fast_sampler on the host cores --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = min(K + W, 96)
        bps, gbs, timed, kind, _ = run_reference_cpu(rowptr.cpu(), col32.to(torch.int64).cpu(), x_local.cpu(), y.cpu(),
                                                     idx_host[:(cb + 8) * bs].clone(), sizes, bs, threads, 8, cb,
                                                     max_seconds=25.0)
        cpu = {"value": round(bps, 3), "unit": UNIT, "cores": threads, "kind": kind, "gathered_GBps": round(gbs, 3),
               "sample": f"{timed} mini-batches of the same workload after 8 warm-up, reference fast_sampler.Session, "
                         f"{threads} worker threads, pinned outputs (sampling + CPU feature slice)"}

    if rank == 0:
        value = world * K / (ms * 1e-3)
        e2e_v = world * K / t_e2e
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": desc + (f"; features partitioned {P}-way, {args.cache_pct}% replicated "
                                           f"{args.cache_policy}-ranked cache, P2P gather over NVLink" if P > 1 else ""),
                       "graph_generator": "Chung-Lu power law gamma=2.5 head_offset=100 seed=1, symmetrised, deduplicated"
                                          + (f", partition locality {args.locality}" if args.locality > 0 else ""),
                       "nnz": int(col32.numel()), "mean_nodes_per_batch": round(mean_nodes, 1),
                       "streams_in_flight": D, "scale": args.scale, "feature_partitions": P,
                       "l2": "inputs larger than L2 (feature table + CSR >> 126 MB, random rows)"},
            "gathered_GBps": round(value * mean_nodes * row_bytes / 1e9, 2),
            "e2e": {"value": round(e2e_v, 2), "unit": UNIT, "h2d_bytes_per_step": bs * 8,
                    "d2h_bytes_per_step": 8 * 32 + (8 * 18 if P > 1 else 0),
                    "gathered_GBps": round(world * e2e_nodes * row_bytes / t_e2e / 1e9, 2),
                    "api": "FastSampler -> " + ("DeviceDistributedPrefetcher" if P > 1 else "DevicePrefetcher"),
                    "per_batch_us": {"p50": round(lat[len(lat) // 2] * 1e6, 1), "p90": round(lat[int(len(lat) * 0.9)] * 1e6, 1),
                                     "max": round(lat[-1] * 1e6, 1), "first": round(first_us, 1), "setup": round(t_setup * 1e6, 1), "slowest_iters": top3,
                                     "consumer_blocked_us_total": None if blocked_us is None else round(blocked_us, 1),
                                     "consumer_blocked_batches": blocked_n}},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_gather (feature gather)", "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "avg_launch_ms": round(g_ms_avg, 5),
                         "algorithmic_bytes_per_launch": int(alg_bytes_per_launch),
                         "bytes_model": "N_b * (2*row_bytes + 4)", "launches_timed": len(evs),
                         "frac_of_nominal_8TBps": round(achieved / 8000.0, 4)},
        }
        if nvlink is not None:
            line["nvlink_roofline"] = nvlink
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (testing only)")
    ap.add_argument("--depth", type=int, default=6, help="mini-batches in flight (CUDA streams)")
    ap.add_argument("--cache-pct", default="15.0",
                    help="replicated rows per GPU in %% of a partition (the reference's alpha, default its documented "
                         "15); 'auto' = whatever fits a quarter of the free HBM, up to every remote row")
    ap.add_argument("--parts", type=int, default=0, help="feature partitions (default: one per GPU)")
    ap.add_argument("--cache-policy", default="vip", choices=["vip", "degree"])
    ap.add_argument("--locality", type=float, default=0.0,
                    help="probability that an edge stays inside its source's partition block (0 = locality-free "
                         "Chung-Lu graph, the worst case for range-partitioned features)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-features", action="store_true",
                    help="A/B experiments with --device-only: sampler alone (no feature / label gather)")
    ap.add_argument("--device-only", action="store_true",
                    help="A/B experiments: print the device-timed batches/s and exit (not a bench line)")
    ap.add_argument("--profile-e2e", action="store_true", help="cProfile the public-API loop (stderr)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
