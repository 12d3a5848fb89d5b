"""Per-tile phase timeline of k_hop_compact_fused (hop 3 of a products-shaped batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from salient_plusplus_b200 import _lib, synthetic as S
from salient_plusplus_b200.pipeline import MiniBatchPipeline

dev = torch.device("cuda", 0)
n, e, f, dt = S.SHAPES["products"]
rowptr, col = S.powerlaw_graph(n, e, seed=1, device=dev)
col = col.to(torch.int32)
x = S.features(n, f, dt, seed=2, device=dev)
idx = S.seeds(n, 1024 * 16, seed=7, device=dev)
pipe = MiniBatchPipeline(rowptr, col, [15, 10, 5], 1024, x_table=x, depth=1, device=dev)
L = _lib.load()
for b in range(8):
    pipe.launch(0, idx.data_ptr() + 8 * b * 1024, 1024, b)
torch.cuda.synchronize()
buf = torch.zeros(8 * 1024, dtype=torch.int64, device=dev)
L.spp_debug_set_timeline(buf.data_ptr())
pipe.launch(0, idx.data_ptr() + 8 * 9 * 1024, 1024, 9)
torch.cuda.synchronize()
L.spp_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0][:, :5].astype(np.float64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["tile start", "entries loaded", "aggregate published", "predecessors summed", "outputs written"]
print(f"{len(t)} tiles; times in us relative to the first tile start")
for i, nm in enumerate(names):
    c = t[:, i]
    print(f"  {nm:22s} min {c.min():7.2f}  p50 {np.median(c):7.2f}  p90 {np.percentile(c, 90):7.2f}  max {c.max():7.2f}")
d = np.diff(t, axis=1)
for i in range(4):
    print(f"  phase {names[i]} -> {names[i + 1]}: p50 {np.median(d[:, i]):6.2f}  max {d[:, i].max():6.2f}")
order = np.argsort(t[:, 0])
print("  start time of tiles (sorted), every 40th:", np.round(t[order, 0][::40], 2))
