"""Multi-GPU check of the P2P path (run under torchrun, one rank per GPU):
  1. correctness: distributed Session with IPC-mapped peer partitions, x == X_global[n_id], the
     ProtoDistributedBatch fields against the oracle, and the NCCL all_to_all comparison path
     producing the same x;
  2. bandwidth: spp_gather_partitioned on all-remote rows (pure NVLink inbound) and on all-local
     rows (pure HBM), GB/s per GPU.
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as O
from salient_plusplus_b200 import _lib, fast_sampler as fs, peer, synthetic as S
from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
from salient_plusplus_b200.transferers import DeviceDistributedPrefetcher, NcclAllToAllPrefetcher

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _lib.load()
P = world

# ---- 1. correctness ------------------------------------------------------------------------------
N, E, F = 200000, 4000000, 128
rowptr, col = S.powerlaw_graph(N, E, seed=1, device=dev)
X = S.features(N, F, torch.float16, seed=2, device=dev)
y = S.labels(N, seed=3, device=dev)
off = S.equal_partition_offsets(N, P)
lo, hi = int(off[rank]), int(off[rank + 1])
x_local = X[lo:hi].clone()
cv = S.degree_cache_vertices(rowptr, off.to(dev), rank, 5000)
cache = fs.Cache(rank, P, cv, X[cv].contiguous())
idx = S.seeds(N, 1024 * 6, seed=11 + rank, lo=lo, hi=hi)


def make_cfg(use_cache):
    return FastSamplerConfig(x_cpu=torch.empty((0, F), dtype=torch.float16), x_gpu=x_local, y=y, rowptr=rowptr, col=col,
                             idx=idx, batch_size=1024, sizes=[15, 10, 5], skip_nonfull_batch=False, pin_memory=True,
                             distributed=True, partition_book=fs.RangePartitionBook(rank, P, off),
                             cache=cache if use_cache else fs.Cache(), force_exact_num_batches=True,
                             exact_num_batches=6, use_cache=use_cache)


ok = True
rp_h, col_h = rowptr.cpu().numpy(), col.cpu().numpy()
for use_cache in (False, True):
    xs = []
    it = iter(FastSampler(4, 4, make_cfg(use_cache)))
    oc = O.Cache(cv.cpu().numpy(), N) if use_cache else None
    for k, (batch,) in enumerate(DeviceDistributedPrefetcher([dev], it)):
        st, en = batch.idx_range.start, batch.idx_range.stop
        on, oa = O.multilayer_sample(idx[st:en].numpy(), [15, 10, 5], rp_h, col_h, rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        good = torch.equal(batch.x, X[torch.from_numpy(on).to(dev)])
        good &= torch.equal(batch.y.cpu(), y[idx[st:en].to(dev)].squeeze().cpu())
        ok &= bool(good)
        xs.append(batch.x.clone())
    # NCCL all_to_all comparison path on the same batches
    it = iter(FastSampler(4, 4, make_cfg(use_cache)))
    for k, (batch,) in enumerate(NcclAllToAllPrefetcher([dev], it)):
        ok &= bool(torch.equal(batch.x, xs[k]))
    dist.barrier()
print(f"[rank {rank}] correctness {'OK' if ok else 'FAILED'}", flush=True)

# ---- 2. bandwidth -----------------------------------------------------------------------------------
del X
rows_per_part = 4_000_000
for F, dt in ((128, torch.float16), (768, torch.float16)):
    rb = F * 2
    part = torch.randn((rows_per_part if F == 128 else rows_per_part // 4, F), device=dev, dtype=torch.float32).to(dt)
    R = part.size(0)
    ptrs = peer.exchange_partition_tables(part, rank, P)
    ptrs[rank] = 0
    tables = [None] * P
    tables[rank] = part
    boff = [p * R for p in range(P + 1)]
    fm = fs.make_feature_map(boff, rank, tables, None, None, ptrs)
    n = 1_000_000 if F == 128 else 300_000
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    out = torch.empty((n, F), dtype=dt, device=dev)
    res = {}
    for name in ("local", "remote"):
        if name == "local":
            ids = torch.randint(0, R, (n,), generator=g, device=dev) + rank * R
        else:
            owner = (rank + 1 + torch.randint(0, P - 1, (n,), generator=g, device=dev)) % P
            ids = torch.randint(0, R, (n,), generator=g, device=dev) + owner * R
        ids = ids.to(torch.int64)
        sp = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, out.data_ptr(), n, None, sp))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, out.data_ptr(), n, None, sp))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": round(ms, 4), "inbound_GBps": round(n * rb / ms / 1e6, 1)}
        if name == "remote":  # verify a sample against the owner's rows fetched through NCCL-free path
            chk = out[:4].clone()
        dist.barrier()
    print(json.dumps({"rank": rank, "world": P, "F": F, "row_bytes": rb, "rows": n, **res}), flush=True)
    del part, out
    torch.cuda.empty_cache()
    dist.barrier()
dist.destroy_process_group()
