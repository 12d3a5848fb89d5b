"""cProfile of the public-API path (FastSampler -> DevicePrefetcher) to find host overheads."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from salient_plusplus_b200 import fast_sampler as fs, synthetic as S
from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
from salient_plusplus_b200.transferers import DevicePrefetcher

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda", 0)
n, e, f, dt = S.SHAPES["products"]
n, e = int(n * scale), int(e * scale)
rowptr, col = S.powerlaw_graph(n, e, seed=1, device=dev)
col = col.to(torch.int32)
x = S.features(n, f, dt, seed=2, device=dev)
y = S.labels(n, seed=3, device=dev)
bs, K = 1024, 200
idx = S.seeds(n, bs * K, seed=7).pin_memory()


def run(k):
    cfg = FastSamplerConfig(x_cpu=x, x_gpu=torch.empty((0, f), dtype=dt), y=y, rowptr=rowptr, col=col,
                            idx=idx[:k * bs], batch_size=bs, sizes=[15, 10, 5], skip_nonfull_batch=False,
                            pin_memory=True, distributed=False)
    it = DevicePrefetcher([dev], iter(FastSampler(16, 8, cfg)))
    c = 0
    for (b,) in it:
        c += b.x.size(0)
    torch.cuda.synchronize()
    return c


run(20)
t = time.perf_counter(); run(K); dt_ = time.perf_counter() - t
print(f"{K / dt_:.1f} batches/s ({dt_ / K * 1e6:.0f} us/batch)")
pr = cProfile.Profile()
pr.enable(); run(K); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
