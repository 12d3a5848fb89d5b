"""Event trace of the mini-batch pipeline under load (diagnostics, B200).

Runs the products-shaped workload through MiniBatchPipeline with `--depth` batches in flight,
arms the library's event trace (spp_trace_begin) for a window of steady-state batches and prints,
per operation, the time its stream spent on it (difference of consecutive marks of one stream:
includes waiting for SM slots behind the other in-flight batches) next to the per-batch
latency.  Usage: python tools/trace_pipeline.py [--depth 6] [--batches 24] [--scale 1.0]
"""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from salient_plusplus_b200 import _lib, synthetic as S  # noqa: E402
from salient_plusplus_b200.pipeline import MiniBatchPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--batches", type=int, default=24)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shape, sizes, bs, _, desc = B.WORKLOADS[a.workload]
    n, f, dt, rowptr, col = B.make_graph(shape, a.scale, dev, 0.0, 1)
    col = col.to(torch.int32)
    x = S.features(n, f, dt, seed=2, device=dev)
    y = S.labels(n, seed=3, device=dev)
    warm = 4 * a.depth
    idx = S.seeds(n, min(n, (warm + a.batches) * bs), seed=7, device=dev)
    pipe = MiniBatchPipeline(rowptr, col, sizes, bs, x_table=x, y_table=y, depth=a.depth, device=dev)

    def launch(b):
        pipe.launch(b % a.depth, idx.data_ptr() + 8 * b * bs, bs, ((b + 1) * bs * 17 + 5) & 0xFFFFFFFF)

    for b in range(warm):
        launch(b)
    # no synchronisation: the traced window starts with the pipeline full
    _lib.trace_begin(64 * a.batches)
    for b in range(warm, warm + a.batches):
        launch(b)
    marks = _lib.trace_end(64 * a.batches)
    torch.cuda.synchronize()

    # main stream of a mark: side streams are attributed to the batch whose begin mark precedes them
    per_stream = collections.defaultdict(list)
    for lab, hop, st, ms in marks:
        per_stream[st].append((lab, hop, ms))
    dur = collections.defaultdict(list)
    lat = []
    for st, ms_list in per_stream.items():
        prev = None
        t_begin = None
        for lab, hop, ms in ms_list:
            if lab == "batch_begin":
                t_begin = ms
                prev = ms
                continue
            if prev is not None:
                dur[(lab, hop)].append(ms - prev)
            prev = ms
            if lab in ("label_gather", "meta_d2h") and t_begin is not None:
                lat.append(ms - t_begin)
                t_begin = None
    t_all = [m[3] for m in marks]
    window_ms = max(t_all) - min(t_all)
    rows = []
    for (lab, hop), v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        v.sort()
        rows.append({"op": lab, "hop": hop, "n": len(v), "mean_us": round(1e3 * sum(v) / len(v), 1),
                     "p50_us": round(1e3 * v[len(v) // 2], 1), "max_us": round(1e3 * v[-1], 1)})
    out = {"workload": desc, "depth": a.depth, "batches": a.batches, "window_ms": round(window_ms, 3),
           "us_per_batch": round(1e3 * window_ms / a.batches, 1),
           "batch_latency_us_mean": round(1e3 * sum(lat) / max(1, len(lat)), 1), "ops": rows,
           "env": {k: v for k, v in os.environ.items() if k.startswith("SPP_")}}
    print(json.dumps(out))
    print("%-16s %3s %4s %9s %9s %9s" % ("op", "hop", "n", "mean_us", "p50_us", "max_us"), file=sys.stderr)
    for r in rows:
        print("%-16s %3d %4d %9.1f %9.1f %9.1f" % (r["op"], r["hop"], r["n"], r["mean_us"], r["p50_us"], r["max_us"]),
              file=sys.stderr)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump({"summary": out, "marks": marks}, fh)


if __name__ == "__main__":
    main()
