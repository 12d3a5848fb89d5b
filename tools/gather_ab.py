"""A/B of the gather kernel flavours (LDG.128 tiles vs bulk-copy / 1-D TMA pipeline) on all-local
rows (HBM bound) and, under torchrun with >= 2 GPUs, on all-remote rows (NVLink bound).  Every
variant's output is verified against the id-function ground truth (x == f(ids)), so this is also
the parity check of the bulk kernel over real NVLink.
    python tools/gather_ab.py                      # one GPU: local rows only
    torchrun --nproc-per-node 2 tools/gather_ab.py # + remote rows
One JSON line per (row shape, variant)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from salient_plusplus_b200 import _lib, fast_sampler as fs, peer, synthetic as S

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L = _lib.load()
P = world

VARIANTS = [("ldg", dict(gather_bulk=0)),
            ("bulk t4096 s6 c2", dict(gather_bulk=1, bulk_tile=4096, bulk_stages=6, bulk_ctas_per_sm=2)),
            ("bulk t8192 s4 c2", dict(gather_bulk=1, bulk_tile=8192, bulk_stages=4, bulk_ctas_per_sm=2)),
            ("bulk t4096 s8 c2", dict(gather_bulk=1, bulk_tile=4096, bulk_stages=8, bulk_ctas_per_sm=2)),
            ("bulk t2048 s8 c4", dict(gather_bulk=1, bulk_tile=2048, bulk_stages=8, bulk_ctas_per_sm=4)),
            ("bulk t4096 s6 c1", dict(gather_bulk=1, bulk_tile=4096, bulk_stages=6, bulk_ctas_per_sm=1))]
SHAPES = [(128, torch.float16, 4_000_000, 1_000_000), (768, torch.float16, 1_000_000, 300_000),
          (100, torch.float16, 4_000_000, 1_000_000)]


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


for F, dt, R, n in SHAPES:
    rb = F * 2
    part = S.features_by_id(rank * R, (rank + 1) * R, F, dt, device=dev)
    tab = fs.feature_table(part)          # 200-byte rows get a 256-byte pitch
    boff = [p * R for p in range(P + 1)]
    if world > 1:
        ptrs = peer.exchange_device_tables(tab.storage)
    else:
        ptrs = [tab.storage.data_ptr()]
    fm = fs.make_feature_map(boff, rank, None, None, None, ptrs, tab.pitch, 0)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    out = torch.empty((n, F), dtype=dt, device=dev)
    cases = {"local": (torch.randint(0, R, (n,), generator=g, device=dev) + rank * R).to(torch.int64)}
    if world > 1:
        owner = (rank + 1 + torch.randint(0, P - 1, (n,), generator=g, device=dev)) % P
        cases["remote"] = (torch.randint(0, R, (n,), generator=g, device=dev) + owner * R).to(torch.int64)
    sp = torch.cuda.current_stream().cuda_stream
    for vname, opts in VARIANTS:
        if opts.get("gather_bulk") == 1 and rb % 16 != 0 and tab.pitch == rb:
            continue  # dense rows that are not multiples of 16 bytes cannot be bulk-copied
        for k_, v_ in opts.items():
            _lib.tune(k_, v_)
        res = {}
        for cname, ids in cases.items():
            out.zero_()
            for _ in range(3):
                _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n, None, sp))
            barrier()
            ok = bool(torch.equal(out.view(torch.int16), S._id_pattern(ids, F, dt)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n, None, sp))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if world > 1:
                t = torch.tensor([ms, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, ok = float(t[0]), float(t[1]) == 0.0
            res[cname] = {"ms": round(ms, 4), "GBps_in": round(n * rb / ms / 1e6, 1), "rows_ok": ok}
            barrier()
        if rank == 0:
            print(json.dumps({"world": world, "row_bytes": rb, "pitch": tab.pitch, "rows": n, "variant": vname, **res}), flush=True)
    _lib.tune("gather_bulk", -1)
    del part, tab, out, cases
    fs._FEATURE_TABLES.clear()
    fs.clear_resident_cache()
    torch.cuda.empty_cache()
    barrier()
if world > 1:
    dist.destroy_process_group()
