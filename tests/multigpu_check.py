"""Multi-GPU parity check of the P2P path (test infrastructure; run under torchrun, one rank per
GPU -- tests/test_gpu_multi.py launches it when the box has >= 2 GPUs): a distributed Session with
IPC-mapped peer partitions must give x == X_global[n_id] with n_id from the oracle on the same
seeds, the labels of the seeds, and the NCCL all_to_all comparison path must produce the same x.
Exit code 0 = every rank OK."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as O
from salient_plusplus_b200 import _lib, fast_sampler as fs, synthetic as S
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

flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
