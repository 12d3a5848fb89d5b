"""Single-process, two-GPU run of the P2P miss fetch for ncu (ncu must never wrap a multi-rank
command): the feature table lives on cuda:1, the gather kernel runs on cuda:0 and reads every row
over NVLink through plain peer access (same loads as through a CUDA-IPC mapping).
    python tools/ncu_peer_gather.py [--row-bytes 256] [--rows 1000000] [--bulk 0|1]
    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum.per_second,gpu__time_duration.sum -k regex:k_gather ...
Prints the CUDA-event GB/s of the same launches and verifies the rows."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from salient_plusplus_b200 import _lib, fast_sampler as fs, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--row-bytes", type=int, default=256)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--table-rows", type=int, default=4_000_000)
ap.add_argument("--bulk", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
assert torch.cuda.device_count() >= 2, "needs two GPUs"
L = _lib.load()
F = a.row_bytes // 2
R = a.table_rows
remote = S.features_by_id(R, 2 * R, F, torch.float16, device="cuda:1")   # partition 1 on the peer GPU
rtab_pitch = a.row_bytes
if a.row_bytes % 128 != 0 and a.row_bytes >= 96:   # same padded layout fast_sampler.feature_table gives
    rtab_pitch = (a.row_bytes + 127) // 128 * 128
    padded = torch.zeros((R, rtab_pitch // 2), dtype=torch.float16, device="cuda:1")
    padded[:, :F] = remote
    remote = padded
torch.cuda.set_device(0)
_lib.check(L.spp_enable_peer_access(1), "spp_enable_peer_access")
local = S.features_by_id(0, 1024, F, torch.float16, device="cuda:0")
fm = fs.make_feature_map([0, R, 2 * R], 0, None, None, None, [local.data_ptr(), remote.data_ptr()], rtab_pitch, 0)
g = torch.Generator(device="cuda:0").manual_seed(1)
ids = (torch.randint(0, R, (a.rows,), generator=g, device="cuda:0") + R).to(torch.int64)
out = torch.empty((a.rows, F), dtype=torch.float16, device="cuda:0")
_lib.tune("gather_bulk", a.bulk)
_lib.tune("gather_ctas_per_sm", 2 if not a.bulk else 0)
sp = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps + 2):
    if i == 2:
        e0.record()
    _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), a.row_bytes, ids.data_ptr(), 1, a.rows, None, None, out.data_ptr(),
                                        a.rows, None, sp))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
ok = bool(torch.equal(out.view(torch.int16), S._id_pattern(ids, F, torch.float16)))
print(json.dumps({"row_bytes": a.row_bytes, "pitch": rtab_pitch, "rows": a.rows, "bulk": a.bulk, "ms": round(ms, 4),
                  "GBps_in": round(a.rows * a.row_bytes / ms / 1e6, 1), "rows_ok": ok}))
sys.exit(0 if ok else 1)
