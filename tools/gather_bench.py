"""Feature-gather micro-benchmark (B200): spp_gather_rows_pitched alone on a products- / papers- /
MAG-shaped table with unique random row ids, CUDA-event timed, checked against torch indexing.
Variants are selected with the library's SPP_GATHER_* environment switches (read once per process).
usage: python tools/gather_bench.py [--shape products|papers|mag] [--reps 50]"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salient_plusplus_b200 import _lib, fast_sampler as fs  # noqa: E402

SHAPES = {"products": (2_449_029, 100, torch.float16, 572_000), "papers": (8_000_000, 128, torch.float16, 460_000),
          "mag": (1_000_000, 768, torch.float16, 300_000), "arxiv": (169_343, 128, torch.float32, 116_000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    rows, dim, dt, n = SHAPES[a.shape]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((rows, dim), generator=g, device=dev, dtype=torch.float32).to(dt)
    tab = fs.feature_table(x)
    idx = torch.randperm(rows, generator=g, device=dev)[:n].to(torch.int32)
    out = torch.empty((n, dim), dtype=dt, device=dev)
    L = _lib.load()
    rb = dim * x.element_size()
    sp = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.check(L.spp_gather_rows_pitched(tab.ptr, tab.pitch, rb, idx.data_ptr(), 0, n, None, out.data_ptr(), n, sp))

    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ok = bool(torch.equal(out, x[idx.long()]))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for _ in range(a.reps):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    med = times[len(times) // 2]
    # back-to-back launches (no flush), like the pipeline's roofline pass
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / a.reps
    alg = n * (2 * rb + 4)
    peak = 6544.7
    print(json.dumps({"shape": a.shape, "rows": n, "row_bytes": rb, "pitch": tab.pitch, "bit_exact": ok,
                      "us_flushed_median": round(med * 1e3, 2), "us_back_to_back": round(b2b * 1e3, 2),
                      "GBps_alg_flushed": round(alg / med / 1e6, 1), "frac_flushed": round(alg / med / 1e6 / peak, 4),
                      "GBps_alg_b2b": round(alg / b2b / 1e6, 1), "frac_b2b": round(alg / b2b / 1e6 / peak, 4),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("SPP_")}}))


if __name__ == "__main__":
    main()
