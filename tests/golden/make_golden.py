"""Generates tests/golden/*.npz by running the compiled, UNMODIFIED reference
(oracle/_ref/fast_sampler.so, and the no-pin build for the distributed Session which otherwise
needs a CUDA driver) on small seeded inputs.  Run in the build container:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through the deterministic cases,
the CUDA path (tests/test_gpu_golden.py) on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from salient_plusplus_b200 import synthetic as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def adjs_to_dict(prefix, adjs, d):
    d[f"{prefix}_n"] = np.array(len(adjs))
    for i, (rp, cl, e_id, sz) in enumerate(adjs):
        d[f"{prefix}_{i}_rowptr"] = rp.numpy()
        d[f"{prefix}_{i}_col"] = cl.numpy()
        d[f"{prefix}_{i}_eid_len"] = np.array(e_id.numel())
        d[f"{prefix}_{i}_size"] = np.array(sz, dtype=np.int64)


def base_config(R, x, y, rowptr, col, idx, sizes, bs):
    cfg = R.Config()
    cfg.x_cpu, cfg.x_gpu, cfg.y = x, torch.empty(0), y
    cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
    cfg.batch_size, cfg.sizes = bs, sizes
    cfg.skip_nonfull_batch = False
    cfg.pin_memory = False
    cfg.distributed = False
    cfg.force_exact_num_batches = False
    cfg.exact_num_batches = 0
    cfg.count_remote_frequency = False
    cfg.use_cache = False
    return cfg


def make_inputs():
    rowptr, col = S.powerlaw_graph(600, 6000, seed=21, head_offset=5.0)
    N = rowptr.numel() - 1
    x = S.features(N, 12, torch.float16, seed=22)
    y = S.labels(N, seed=23)
    return rowptr, col, N, x, y


def make_sampling():
    R = ref.load_reference()
    rowptr, col, N, x, y = make_inputs()
    _unused = S.powerlaw_graph = S.powerlaw_graph(600, 6000, seed=21, head_offset=5.0)
    idx = S.seeds(N, 100, seed=24)
    idx[7] = idx[3]  # a duplicated seed

    d = dict(rowptr=rowptr.numpy(), col=col.numpy(), x=x.view(torch.int16).numpy(), y=y.numpy(), idx=idx.numpy())

    # deterministic: full neighbourhood, 1..3 hops (multilayer_sample, fast_sampler.cpp:191-236)
    for L in (1, 2, 3):
        n_id, adjs = R.multilayer_sample(idx[:40], [-1] * L, rowptr, col)
        d[f"full{L}_n_id"] = n_id.numpy()
        adjs_to_dict(f"full{L}", adjs, d)
    # sample_adj free function (sample_cpu.hpp:154-165): int32 n_id
    rp, cl, nid32, e_id = R.sample_adj(rowptr, col, idx[:40], -1, False)
    d["sa_rowptr"], d["sa_col"], d["sa_n_id"] = rp.numpy(), cl.numpy(), nid32.numpy()
    assert nid32.dtype == torch.int32 and e_id.numel() == 0

    # stochastic through a Session (per-batch seed stop*17+5, fast_sampler.cpp:994)
    cfg = base_config(R, x, y, rowptr, col, idx, [15, 10, 5], 32)
    s = R.Session(2, 10, cfg)
    nb = 0
    while True:
        b = s.blocking_get_batch()
        if b is None:
            break
        xb, yb, adjs, (st, en) = b
        p = f"sess_{st}"
        d[f"{p}_stop"] = np.array(en)
        d[f"{p}_x"] = xb.view(torch.int16).numpy()
        d[f"{p}_y"] = yb.numpy()
        adjs_to_dict(p, adjs, d)
        nb += 1
    d["sess_num_batches"] = np.array(nb)

    # exact-num-batches ranges (fast_sampler.cpp:592-615)
    cfg = base_config(R, x, y, rowptr, col, idx, [2], 32)
    cfg.force_exact_num_batches, cfg.exact_num_batches = True, 7
    s = R.Session(1, 10, cfg)
    rngs = []
    while True:
        b = s.blocking_get_batch()
        if b is None:
            break
        rngs.append(b[3])
    d["exact7_ranges"] = np.array(sorted(rngs), dtype=np.int64)

    # serial_index with the n argument (fast_sampler.cpp:238-279)
    sel = torch.tensor([5, 0, 599, 17, 17, 3], dtype=torch.int64)
    d["si_idx"] = sel.numpy()
    d["si_out"] = R.serial_index(x, sel).view(torch.int16).numpy()
    d["si_out_n4"] = R.serial_index(x, sel, 4).view(torch.int16).numpy()
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), **d)


def make_distributed():
    # Two pybind builds of the same C++ types cannot live in one process, so this half runs in
    # its own interpreter with only the no-pin build loaded (identical code except that host
    # buffers are not pinned, which needs a CUDA driver).
    RN = ref.load_reference(nopin=True)
    R = RN
    rowptr, col, N, x, y = make_inputs()
    # partition book / cache / distributed binning (range_partition_book.cpp, fast_sampler.cpp:1017-1262)
    P, rank = 4, 1
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    book = R.RangePartitionBook(rank, P, off)
    probe = torch.cat([off.clamp(max=N - 1), S.seeds(N, 64, seed=31)])
    cv = S.degree_cache_vertices(rowptr, off, rank, 60)
    cache = R.Cache(rank, P, cv, x[cv].contiguous())
    e = dict(offsets=off.numpy(), rank=np.array(rank), probe=probe.numpy(),
             partid=book.nid2partid(probe).numpy(), localnid=book.nid2localnid(probe, rank).numpy(),
             partid2nids=book.partid2nids(2).numpy(), cached_vertices=cv.numpy(),
             is_cached=cache.nid_is_cached(probe).numpy(),
             cachenid=RN.Cache(rank, P, cv, x[cv].contiguous()).nid2cachenid(cv[:20]).numpy())
    x_local = x[lo:hi].contiguous()
    cut = (hi - lo) // 2
    lidx = S.seeds(N, 90, seed=33, lo=lo, hi=hi)
    e["lidx"] = lidx.numpy()
    e["cut"] = np.array(cut)
    for use_cache in (False, True):
        cfg = base_config(RN, x_local[cut:].contiguous(), y, rowptr, col, lidx, [15, 10, 5], 32)
        cfg.x_gpu = x_local[:cut].contiguous()
        cfg.distributed = True
        cfg.partition_book = RN.RangePartitionBook(rank, P, off)
        cfg.cache = RN.Cache(rank, P, cv, x[cv].contiguous()) if use_cache else RN.Cache()
        cfg.use_cache = use_cache
        cfg.force_exact_num_batches, cfg.exact_num_batches = True, 3
        s = RN.Session(2, 10, cfg)
        tag = "c" if use_cache else "nc"
        k = 0
        while True:
            b = s.blocking_get_batch_distributed()
            if b is None:
                break
            p = f"{tag}{k}"
            e[f"{p}_range"] = np.array(b.idx_range, dtype=np.int64)
            for q, t in enumerate(b.partition_nids):
                e[f"{p}_part{q}"] = t.numpy()
            e[f"{p}_cached_nids"] = b.cached_nids.numpy()
            e[f"{p}_perm"] = b.perm_partition_to_mfg.numpy()
            e[f"{p}_cpu_feats"] = b.sliced_cpu_features.view(torch.int16).numpy()
            e[f"{p}_labels"] = b.sliced_cpu_labels.numpy()
            adjs_to_dict(p, b.adjs, e)
            k += 1
        e[f"{tag}_num_batches"] = np.array(k)
    np.savez_compressed(os.path.join(OUT, "distributed.npz"), **e)


if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1:
        {"sampling": make_sampling, "distributed": make_distributed}[sys.argv[1]]()
    else:
        for part in ("sampling", "distributed"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), part])
            f = part + ".npz"
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
