"""Layer-wise inference sampling (BASELINE config 3): one full-neighbourhood hop (sizes [-1]) per
batch of 1024 consecutive vertices + feature gather, through the public API
(FastSampler -> DevicePrefetcher), batches/s; `--ref` times the reference's CPU fast_sampler
Session (oracle/_ref) on the same batches with all host cores instead (run it as a separate
process: the reference arm is the only part of this tool that touches oracle/).
usage: python tools/layerwise_bench.py [--batches 300] [--ref]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from salient_plusplus_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--batches", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--profile", action="store_true", help="cProfile the consumer thread (stderr)")
    ap.add_argument("--trace", action="store_true", help="event trace of the timed batches (use SPP_SESSION_DEPTH=1)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shape, _, bs, _, desc = B.WORKLOADS[a.workload]
    n, f, dt, rowptr, col = B.make_graph(shape, a.scale, dev, 0.0, 1)
    x = S.features(n, f, dt, seed=2, device=dev)
    y = S.labels(n, seed=3, device=dev)
    total = a.warmup + a.batches
    idx_all = torch.arange(min(n, total * bs), dtype=torch.int64)  # layer-wise inference walks every vertex in order
    if a.ref:
        bps, gbs, timed, kind, _ = B.run_reference_cpu(rowptr.cpu(), col.cpu(), x.cpu(), y.cpu(), idx_all, [-1], bs,
                                                       os.cpu_count() or 1, a.warmup, a.batches, max_seconds=25.0)
        print(json.dumps({"impl": "reference", "kind": kind, "workload": desc, "sizes": [-1], "batches_per_s": round(bps, 1),
                          "gathered_GBps": round(gbs, 2), "batches": timed, "cores": os.cpu_count()}))
        return
    from salient_plusplus_b200.samplers import FastSampler, FastSamplerConfig
    from salient_plusplus_b200.transferers import DevicePrefetcher
    col32 = col.to(torch.int32)

    def run(first, count):
        cfg = FastSamplerConfig(x_cpu=x, x_gpu=torch.empty((0, f), dtype=dt), y=y, rowptr=rowptr, col=col32,
                                idx=idx_all[first * bs:(first + count) * bs], batch_size=bs, sizes=[-1],
                                skip_nonfull_batch=False, pin_memory=True, distributed=False)
        nodes, edges, got = 0, 0, 0
        for (batch,) in DevicePrefetcher([dev], iter(FastSampler(16, 6, cfg))):
            nodes += batch.x.size(0)
            edges += batch.adjs[0].adj_t.nnz()
            got += 1
        torch.cuda.synchronize()
        return got, nodes, edges

    run(0, a.warmup)
    if a.trace:
        from salient_plusplus_b200 import _lib
        _lib.trace_begin(64 * a.batches)
    prof = None
    if a.profile:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    got, nodes, edges = run(a.warmup, a.batches)
    dt_s = time.perf_counter() - t0
    st1 = torch.cuda.memory_stats()
    alloc = {k: st1[k] - st0[k] for k in ("segment.all.allocated", "segment.all.freed", "allocation.all.allocated",
                                          "num_alloc_retries", "num_device_alloc", "num_device_free") if k in st1}
    alloc["reserved_GB"] = round(st1["reserved_bytes.all.current"] / 1e9, 2)
    print("allocator:", json.dumps(alloc), file=sys.stderr)
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(16)
    if a.trace:
        import collections
        marks = _lib.trace_end(64 * a.batches)
        dur, prev = collections.defaultdict(list), {}
        for lab, hop, st, ms in marks:
            if lab != "batch_begin" and st in prev:
                dur[lab].append(ms - prev[st])
            prev[st] = ms
        for lab, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
            v.sort()
            print("%-18s n=%4d mean %8.1f us  p50 %8.1f  p90 %8.1f  max %8.1f" % (
                lab, len(v), 1e3 * sum(v) / len(v), 1e3 * v[len(v) // 2], 1e3 * v[int(len(v) * 0.9)], 1e3 * v[-1]),
                file=sys.stderr)
    rb = f * x.element_size()
    print(json.dumps({"impl": "ours", "workload": desc, "sizes": [-1], "batches_per_s": round(got / dt_s, 1),
                      "us_per_batch": round(1e6 * dt_s / got, 1), "mean_nodes": round(nodes / got, 1),
                      "mean_edges": round(edges / got, 1), "gathered_GBps": round(nodes * rb / dt_s / 1e9, 2),
                      "batches": got, "env": {k: v for k, v in os.environ.items() if k.startswith("SPP_")}}))


if __name__ == "__main__":
    main()
