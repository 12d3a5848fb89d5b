"""Multi-GPU measurements of the P2P path (run under torchrun, one rank per GPU; the parity check
against the oracle lives in tests/multigpu_check.py):
  1. comparison: the same distributed mini-batches through the fused P2P gather
     (DeviceDistributedPrefetcher) and through the reference's protocol of three NCCL all_to_alls
     per batch (NcclAllToAllPrefetcher, fast_trainer/transferers.py:507-766); identical x required,
     batches/s of both reported;
  2. bandwidth: spp_gather_partitioned on all-remote rows (pure NVLink inbound) and on all-local
     rows (pure HBM), GB/s per GPU.
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from salient_plusplus_b200 import _lib, fast_sampler as fs, peer, synthetic as S
from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
from salient_plusplus_b200.transferers import DeviceDistributedPrefetcher, NcclAllToAllPrefetcher

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _lib.load()
P = world

# ---- 1. fused P2P gather vs the NCCL all_to_all protocol ----------------------------------------
N, E, F = 2_000_000, 50_000_000, 128
B = 48
rowptr, col = S.powerlaw_graph(N, E, seed=1, device=dev)
X = S.features(N, F, torch.float16, seed=2, device=dev)
y = S.labels(N, seed=3, device=dev)
off = S.equal_partition_offsets(N, P)
lo, hi = int(off[rank]), int(off[rank + 1])
x_local = X[lo:hi].clone()
del X
from salient_plusplus_b200 import vip as V  # noqa: E402
# 15 % replicated cache, degree ranked; the rows are pulled out of the owners' partitions over P2P
cache = V.create_vip_cache(rowptr, col, None, 1024, [15, 10, 5], off, rank, 15.0, x_local,
                           vip=(rowptr[1:] - rowptr[:-1]).to(torch.float64))
idx = S.seeds(N, 1024 * B, seed=11 + rank, lo=lo, hi=hi)


def make_cfg():
    return FastSamplerConfig(x_cpu=torch.empty((0, F), dtype=torch.float16), x_gpu=x_local, y=y, rowptr=rowptr, col=col,
                             idx=idx, batch_size=1024, sizes=[15, 10, 5], skip_nonfull_batch=False, pin_memory=True,
                             distributed=True, partition_book=fs.RangePartitionBook(rank, P, off), cache=cache,
                             force_exact_num_batches=True, exact_num_batches=B, use_cache=True)


def drain(cls, keep):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, nodes = [], 0
    for (batch,) in cls([dev], iter(FastSampler(4, 6, make_cfg()))):
        nodes += batch.x.size(0)
        if keep and len(out) < 4:
            out.append(batch.x.clone())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), nodes, out


drain(DeviceDistributedPrefetcher, False)  # warm-up (slot pools, allocator, NCCL communicators)
drain(NcclAllToAllPrefetcher, False)
t_p2p, nodes, xs = drain(DeviceDistributedPrefetcher, True)
t_nccl, nodes2, xs2 = drain(NcclAllToAllPrefetcher, True)
same = nodes == nodes2 and all(torch.equal(a, b) for a, b in zip(xs, xs2))
if rank == 0:
    print(json.dumps({"comparison": "fused P2P gather vs 3x NCCL all_to_all per batch", "world": P, "batches_per_rank": B,
                      "graph": f"power law {N} nodes {E} edges, {F}-d fp16, 15% degree-ranked cache",
                      "identical_x": bool(same), "p2p_batches_per_s": round(P * B / t_p2p, 1),
                      "nccl_all_to_all_batches_per_s": round(P * B / t_nccl, 1),
                      "speedup": round(t_nccl / t_p2p, 2)}), flush=True)
del x_local, cache, rowptr, col
torch.cuda.empty_cache()
dist.barrier()

# ---- 2. bandwidth -----------------------------------------------------------------------------------
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
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n, None, sp))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n, None, sp))
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
