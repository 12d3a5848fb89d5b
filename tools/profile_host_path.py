"""Per-batch latency distribution of the public-API path (FastSampler -> DevicePrefetcher)."""
import os
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
    t0 = time.perf_counter()
    it = DevicePrefetcher([dev], iter(FastSampler(16, 8, cfg)))
    t_init = time.perf_counter() - t0
    lat = []
    t = time.perf_counter()
    for (b,) in it:
        t2 = time.perf_counter()
        lat.append(t2 - t)
        t = t2
    torch.cuda.synchronize()
    return t_init, lat, time.perf_counter() - t0


run(20)
for rep in range(6):
    t_init, lat, tot = run(K)
    lat.sort()
    print(f"rep {rep}: {K / tot:7.1f} batches/s  init {t_init * 1e3:6.2f} ms  per-batch us: p10 {lat[20] * 1e6:6.0f} "
          f"p50 {lat[100] * 1e6:6.0f} p90 {lat[180] * 1e6:6.0f} max {lat[-1] * 1e6:7.0f}", flush=True)
