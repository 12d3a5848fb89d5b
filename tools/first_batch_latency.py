"""Why does the first batch of a new Session sometimes take tens of ms inside bench.py?"""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from salient_plusplus_b200 import fast_sampler as fs, synthetic as S
from salient_plusplus_b200.pipeline import MiniBatchPipeline
from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
from salient_plusplus_b200.transferers import DevicePrefetcher

dev = torch.device("cuda", 0)
n, e, f, dt = S.SHAPES["products"]
rowptr, col = S.powerlaw_graph(n, e, seed=1, device=dev)
col = col.to(torch.int32)
x = S.features(n, f, dt, seed=2, device=dev)
y = S.labels(n, seed=3, device=dev)
bs, K = 1024, 220
idx_dev = S.seeds(n, bs * K, seed=7, device=dev)
idx = idx_dev.cpu().pin_memory()
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"


def session_run(k, first):
    cfg = FastSamplerConfig(x_cpu=x, x_gpu=torch.empty((0, f), dtype=dt), y=y, rowptr=rowptr, col=col,
                            idx=idx[first * bs:(first + k) * bs], batch_size=bs, sizes=[15, 10, 5],
                            skip_nonfull_batch=False, pin_memory=True, distributed=False)
    t0 = time.perf_counter()
    it = DevicePrefetcher([dev], iter(FastSampler(16, 4, cfg)))
    t1 = time.perf_counter()
    for _ in it:
        pass
    torch.cuda.synchronize()
    return (t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3


if mode in ("pipe", "all"):
    pipe = MiniBatchPipeline(rowptr, col, [15, 10, 5], bs, x_table=x, y_table=y, depth=4, device=dev)
    for b in range(220):
        pipe.launch(b % 4, idx_dev.data_ptr() + 8 * b * bs, bs, b)
    torch.cuda.synchronize()
    for b in range(64):
        pipe.launch(0, idx_dev.data_ptr() + 8 * b * bs, bs, b, True)
    torch.cuda.synchronize()
    for b in range(8):
        pipe.read_meta(0)
if mode in ("smi", "all2"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler
    cs = ClockSampler(0); cs.start(); time.sleep(0.5); print(cs.stop())
if mode in ("gc", "all"):
    gc.collect(); gc.freeze()
out = []
for rep in range(8):
    a, b = session_run(20, 0)
    torch.cuda.synchronize()
    c, d = session_run(200, 20)
    out.append(f"{a:.1f}/{c:.1f}")
print(mode, "first-batch ms (warm-up session / timed session):", " ".join(out), flush=True)
