"""GPU parity tests through the reference-facing PUBLIC API (FastSampler -> DeviceIterator), at
full size, and of the round-2 structures (cache index, source descriptors, several partitions on
one GPU).  Everything is compared with the oracle / the compiled reference on the same inputs."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from salient_plusplus_b200 import synthetic as S
from tests.util import adjs_equal, small_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs():
    from salient_plusplus_b200 import fast_sampler
    return fast_sampler


@pytest.fixture(scope="module")
def data():
    rowptr, col = small_graph(n=8000, e=200000)
    N = rowptr.numel() - 1
    x = S.features_by_id(0, N, 128, torch.float16)
    y = S.labels_by_id(torch.arange(N))
    return rowptr, col, x, y, N


def _raw(adj):
    """Adj record of samplers.py -> (rowptr, col, e_id, (T, S)) like the module-level tuples."""
    a = adj
    if hasattr(a, "adj_t"):
        rp, cl, _ = a.adj_t.csr()
        return rp, cl, a.e_id, (a.size[1], a.size[0])
    return a


def _cfg(x, y, rowptr, col, idx, **kw):
    from salient_plusplus_b200.samplers import FastSamplerConfig
    base = dict(x_cpu=x, x_gpu=torch.empty((0, x.size(1)), dtype=x.dtype), y=y, rowptr=rowptr, col=col, idx=idx,
                batch_size=64, sizes=[15, 10, 5], skip_nonfull_batch=False, pin_memory=True, distributed=False)
    base.update(kw)
    return FastSamplerConfig(**base)


def test_features_by_id_same_on_cpu_and_gpu():
    a = S.features_by_id(1000, 9000, 100, torch.float16, chunk_rows=3000)
    b = S.features_by_id(1000, 9000, 100, torch.float16, device="cuda")
    assert torch.equal(a.view(torch.int16), b.cpu().view(torch.int16))
    ids = torch.tensor([2 ** 31 - 5, 123456789, 1000, 8999])
    assert torch.equal(S.expected_features(ids, 100).view(torch.int16),
                       S.expected_features(ids.cuda(), 100).cpu().view(torch.int16))
    c = S.features_by_id(0, 500, 128, torch.float32, device="cuda")
    assert torch.equal(c.cpu().view(torch.int32), S.features_by_id(0, 500, 128, torch.float32).view(torch.int32))
    assert torch.equal(S.labels_by_id(torch.arange(777)), S.labels_by_id(torch.arange(777, device="cuda")).cpu())


def test_device_prefetcher_against_oracle(fs, data):
    """FastSampler -> DevicePrefetcher (fast_trainer/transferers.py:890-970) on one GPU: every
    PreparedBatch against the oracle (counter-RNG mode) and x == X[n_id], y == Y[seeds]."""
    from salient_plusplus_b200.samplers import FastSampler
    from salient_plusplus_b200.transferers import DevicePrefetcher
    rowptr, col, x, y, N = data
    idx = S.seeds(N, 64 * 9 + 17)
    it = iter(FastSampler(4, 4, _cfg(x, y, rowptr, col, idx)))
    dev = torch.device("cuda", 0)
    seen = []
    for (batch,) in DevicePrefetcher([dev], it):
        st, en = batch.idx_range.start, batch.idx_range.stop
        seen.append((st, en))
        on, oa = O.multilayer_sample(idx[st:en].numpy(), [15, 10, 5], rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal([_raw(a) for a in batch.adjs], oa)
        assert batch.x.is_cuda and torch.equal(batch.x.cpu(), x[torch.from_numpy(on)])
        assert torch.equal(batch.y.cpu(), y[idx[st:en]].squeeze())
    assert sorted(seen) == O.batch_ranges(idx.numel(), 64, False, False, 0)
    assert it.get_stats().total_blocked_occasions >= 0


@pytest.mark.parametrize("use_cache", [False, True])
@pytest.mark.parametrize("prefetcher", ["p2p", "nccl"])
def test_distributed_prefetchers_against_oracle(fs, data, use_cache, prefetcher):
    """FastSampler -> DeviceDistributedPrefetcher (fused partition-book + cache + peer gather,
    replaces fast_trainer/transferers.py:33-887) and -> NcclAllToAllPrefetcher (the reference's
    all_to_all protocol, :507-766, here with a world-size-1 NCCL group so that it runs on one GPU):
    same batches, x == X[n_id] with n_id from the oracle."""
    import torch.distributed as dist
    from salient_plusplus_b200.samplers import FastSampler
    from salient_plusplus_b200.transferers import DeviceDistributedPrefetcher, NcclAllToAllPrefetcher
    rowptr, col, x, y, N = data
    dev = torch.device("cuda", 0)
    if prefetcher == "nccl":
        P, rank = 1, 0
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    else:
        P, rank = 4, 2
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    idx = S.seeds(N, 64 * 5, lo=lo, hi=hi)
    cv = S.degree_cache_vertices(rowptr, off, rank, 700) if P > 1 else torch.empty(0, dtype=torch.int64)
    use_cache = use_cache and P > 1
    cfg = _cfg(torch.empty((0, x.size(1)), dtype=x.dtype), y, rowptr, col, idx, distributed=True, use_cache=use_cache,
               partition_book=fs.RangePartitionBook(rank, P, off),
               cache=fs.Cache(rank, P, cv, x[cv].contiguous()) if use_cache else fs.Cache())
    cfg.x_gpu = x[lo:hi].contiguous()
    cfg.partition_tables = [x[int(off[p]):int(off[p + 1])].contiguous() if p != rank else None for p in range(P)]
    it = iter(FastSampler(4, 4, cfg))
    pf = (NcclAllToAllPrefetcher([dev], it) if prefetcher == "nccl" else DeviceDistributedPrefetcher([dev], it))
    n = 0
    for (batch,) in pf:
        st, en = batch.idx_range.start, batch.idx_range.stop
        on, oa = O.multilayer_sample(idx[st:en].numpy(), [15, 10, 5], rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal([_raw(a) for a in batch.adjs], oa)
        assert torch.equal(batch.x.cpu(), x[torch.from_numpy(on)])
        assert torch.equal(batch.y.cpu(), y[idx[st:en]].squeeze())
        n += 1
    assert n == 5
    if prefetcher == "nccl":
        dist.destroy_process_group()


def test_several_partitions_on_one_gpu(fs, data):
    """Fewer GPUs than partitions: a GPU hosts a block of consecutive partitions
    (Config.local_parts).  Rows of every hosted partition count as local, the cache only holds rows
    of the others; the reference-shaped outputs (partition_nids per PARTITION, perm) are unchanged."""
    from salient_plusplus_b200 import peer, vip as V
    rowptr, col, x, y, N = data
    P, world, grank = 8, 2, 1
    hosted = peer.hosted_partitions(grank, world, P)
    assert hosted == [4, 5, 6, 7]
    off = S.equal_partition_offsets(N, P)
    offl = off.tolist()
    prank = hosted[0]
    blo, bhi = offl[hosted[0]], offl[hosted[-1] + 1]
    xd = x.cuda()
    block = xd[blo:bhi]
    other = xd[0:blo].contiguous()           # stands for the peer GPU's block
    pitch = 256
    ptrs = peer.partition_pointers([other.data_ptr(), block.data_ptr()], offl, world, pitch)
    assert ptrs[4] == block.data_ptr() and ptrs[1] == other.data_ptr() + (offl[1] - offl[0]) * pitch
    deg = (rowptr[1:] - rowptr[:-1]).double().cuda()
    cv = V.select_cache_vertices(deg, off, prank, 600, hosted)
    assert cv.numel() == 600 and bool((cv < blo).all())       # nothing hosted here is ever cached
    owner = torch.searchsorted(off.cuda(), cv, right=True) - 1
    assert bool((owner[1:] >= owner[:-1]).all())               # owner-major layout
    cache = fs.Cache(prank, P, cv, xd[cv].contiguous())
    idx = S.seeds(N, 64 * 4, lo=blo, hi=bhi)
    cfg = fs.Config()
    cfg.x_cpu, cfg.x_gpu, cfg.y = torch.empty((0, 128), dtype=x.dtype), block[:offl[prank + 1] - blo], y
    cfg.rowptr, cfg.col, cfg.idx, cfg.batch_size, cfg.sizes = rowptr, col, idx, 64, [15, 10, 5]
    cfg.distributed, cfg.use_cache, cfg.cache = True, True, cache
    cfg.partition_book = fs.RangePartitionBook(prank, P, off)
    cfg.peer_table_ptrs, cfg.peer_table_pitch, cfg.local_parts = ptrs, pitch, hosted
    sess = fs.Session(2, 4, cfg)
    oc = O.Cache(cv.cpu().numpy(), N)
    for _ in range(4):
        b = sess.blocking_get_batch_distributed()
        st, en = b.idx_range
        on, oa = O.multilayer_sample(idx[st:en].numpy(), cfg.sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal(b.adjs, oa) and np.array_equal(b.n_id.cpu().numpy(), on)
        assert torch.equal(b.x.cpu(), x[torch.from_numpy(on)])
        # per-partition buckets: hosted partitions keep ALL their nodes (never diverted to the cache)
        part = np.searchsorted(np.array(offl), on, side="right") - 1
        cached = oc.nid_is_cached(on) & ~np.isin(part, hosted)
        for p in range(P):
            want = on[(part == p) & ~cached]
            assert np.array_equal(b.partition_nids[p].cpu().numpy(), want)
        assert np.array_equal(b.cached_nids.cpu().numpy(), oc.nid2cachenid(on[cached]))
        cat = torch.cat(list(b.partition_nids) + [cv[b.cached_nids]])
        assert torch.equal(cat[b.perm_partition_to_mfg], b.n_id)


def test_cache_index_duplicates_and_out_of_range(fs):
    """Cache lookups (fast_sampler/range_partition_book.cpp:116-195): the LAST position of a
    repeated vertex wins, ids the index does not cover are "not cached"."""
    N = 100000
    g = torch.Generator().manual_seed(5)
    cv = torch.randint(0, N, (30000,), generator=g)           # with repeats
    c = fs.Cache(0, 2, cv, torch.zeros(cv.numel(), 2).half())
    oc = O.Cache(cv.numpy(), N)
    q = torch.cat([torch.arange(N), torch.randint(0, N, (5000,), generator=g)])
    assert np.array_equal(c.nid_is_cached(q).numpy(), oc.nid_is_cached(q.numpy()))
    hit = q[c.nid_is_cached(q)]
    assert np.array_equal(c.nid2cachenid(hit).numpy(), oc.nid2cachenid(hit.numpy()))
    beyond = torch.tensor([N, N + 223, N + 224, 2 ** 31 - 1])
    assert not bool(c.nid_is_cached(beyond).any())
    assert c.nid2cachenid(beyond).tolist() == [-1, -1, -1, -1]
    # edge of the 224-id blocks
    cv2 = torch.tensor([0, 223, 224, 447, 448, 99999])
    c2 = fs.Cache(0, 2, cv2, torch.zeros(6, 2).half())
    assert c2.nid2cachenid(cv2).tolist() == [0, 1, 2, 3, 4, 5]
    assert c2.nid_is_cached(torch.tensor([1, 222, 225, 446, 449, 99998])).tolist() == [False] * 6


def test_fullsize_products_batch_against_oracle(fs):
    """One mini-batch of the full ogbn-products-shaped graph (2.45 M nodes, direct-mapped id table),
    fan-out (15,10,5), batch 1024, against the oracle in counter-RNG mode: bit-exact n_id, rowptr and
    col of every hop, x == f(n_id).  (The CPU oracle needs well under a second per batch.)"""
    n, e, f, dt = S.SHAPES["products"]
    rowptr, col = S.powerlaw_graph(n, e, seed=1, device="cuda")
    col32 = col.to(torch.int32)
    x = S.features_by_id(0, n, f, dt, device="cuda")
    y = S.labels_by_id(torch.arange(n, device="cuda"))
    idx = S.seeds(n, 2048, seed=7)
    cfg = fs.Config()
    cfg.x_cpu, cfg.y, cfg.rowptr, cfg.col, cfg.idx = x, y, rowptr, col32, idx
    cfg.batch_size, cfg.sizes = 1024, [15, 10, 5]
    sess = fs.Session(1, 4, cfg)
    rp_h, col_h = rowptr.cpu().numpy(), col.cpu().numpy()
    for _ in range(2):
        xb, yb, adjs, (st, en) = sess.blocking_get_batch()
        on, oa = O.multilayer_sample(idx[st:en].numpy(), cfg.sizes, rp_h, col_h, rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal(adjs, oa) and on.size > 400000
        assert torch.equal(xb.view(torch.int16), S._id_pattern(torch.from_numpy(on).cuda(), f, dt))
        assert torch.equal(yb, S.labels_by_id(idx[st:en].cuda()))


def test_hashed_table_with_ids_beyond_2_to_30(fs):
    """papers100M / MAG240M-scale ids: a graph with 1.2e9 vertices (ids up to 2^30 + 2^27, so the
    HASHED id table and the int32 id paths are live), whose edges connect 3 M active vertices spread
    over the whole id range; fan-out (15,10,5), batch 1024, bit-exact against the oracle in
    counter-RNG mode.  Also a full-neighbourhood hop on the same graph (reference semantics)."""
    N = 2 ** 30 + 2 ** 27
    A, E = 3_000_000, 40_000_000
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    active = torch.unique(torch.randint(0, N, (A,), generator=g, device=dev, dtype=torch.int64))
    A = active.numel()
    # power-law-ish endpoints among the active vertices
    u = torch.rand(E, generator=g, device=dev, dtype=torch.float64)
    src = active[(u * u * A).long().clamp_(max=A - 1)]
    dst = active[torch.randint(0, A, (E,), generator=g, device=dev)]
    keep = src != dst
    key = torch.unique(torch.cat([src[keep] * N + dst[keep], dst[keep] * N + src[keep]]))
    row = torch.div(key, N, rounding_mode="floor")
    col = (key - row * N).contiguous()
    del key, src, dst, u
    rows_u, cnt = torch.unique_consecutive(row, return_counts=True)
    del row
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    rowptr[rows_u + 1] = cnt
    torch.cumsum(rowptr, 0, out=rowptr)
    assert int(rowptr[-1]) == col.numel() and int(col.max()) >= 2 ** 30
    col32 = col.to(torch.int32)
    seeds = active[torch.randperm(A, generator=g, device=dev)[:1024]]
    rp_h, col_h = rowptr.cpu().numpy(), col.cpu().numpy()
    sz = fs._sampler_sizes(1024, [15, 10, 5], fs._DeviceGraph.get(rowptr, col32))
    assert int(sz.table_direct) == 0                           # hashed table
    n_id, adjs = fs.multilayer_sample(seeds, [15, 10, 5], rowptr, col32, seed=12345)
    on, oa = O.multilayer_sample(seeds.cpu().numpy(), [15, 10, 5], rp_h, col_h, rng_mode=O.RNG_COUNTER, rng_seed=12345)
    assert np.array_equal(n_id.cpu().numpy(), on) and adjs_equal(adjs, oa)
    assert int(n_id.max()) >= 2 ** 30 and on.size > 100000
    n_id, adjs = fs.multilayer_sample(seeds[:256], [-1, -1], rowptr, col32)
    on, oa = O.multilayer_sample(seeds[:256].cpu().numpy(), [-1, -1], rp_h, col_h)
    assert np.array_equal(n_id.cpu().numpy(), on) and adjs_equal(adjs, oa)
    fs._DeviceGraph._cache.clear()
    fs.clear_resident_cache()


def test_async_slice_tensors_against_compiled_reference(fs, data):
    """Session.async_slice_tensors / get_slice_tensors (fast_sampler/fast_sampler.cpp:720-775)
    against the UNMODIFIED reference module on the same requests (needs a CUDA driver for its pinned
    outputs, which the GPU box has)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/fast_sampler.so not built")
    R = ref.load_reference()
    rowptr, col, x, y, N = data
    P, rank = 4, 1
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    cut = (hi - lo) // 2
    idx = S.seeds(N, 128, lo=lo, hi=hi)
    g = torch.Generator().manual_seed(8)
    host_rows = hi - lo - cut
    # requests as transferers.py:545 prepares them: local id - gpu cutoff (negative = row lives on the GPU)
    reqs = [torch.randint(-cut, host_rows, (n_,), generator=g) for n_ in (300, 0, 57, 1000)]

    def run(mod, is_ref):
        cfg = mod.Config()
        cfg.x_cpu = x[lo + cut:hi].contiguous()
        cfg.x_gpu = x[lo:lo + cut].contiguous() if not is_ref else torch.empty((cut, 0), dtype=x.dtype)
        cfg.y = y
        cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
        cfg.batch_size, cfg.sizes = 64, [5, 5]
        cfg.skip_nonfull_batch, cfg.pin_memory, cfg.distributed = False, True, True
        cfg.partition_book = mod.RangePartitionBook(rank, P, off)
        cfg.cache = mod.Cache()
        cfg.force_exact_num_batches, cfg.exact_num_batches = False, 0
        cfg.count_remote_frequency = cfg.use_cache = False
        sess = mod.Session(2, 4, cfg)
        sess.async_slice_tensors([r.clone() for r in reqs], rank)
        sess.wait_slice_tensors()
        out = sess.get_slice_tensors()
        while sess.blocking_get_batch_distributed() is not None:
            pass
        return out

    want = run(R, True)
    got = run(fs, False)
    assert len(got) == len(want) == len(reqs)
    for i, (g_, w_) in enumerate(zip(got, want)):
        assert len(g_) == len(w_) == 3
        if i != rank:
            assert torch.equal(g_[0].cpu(), w_[0]), f"rows of request {i}"
        assert torch.equal(g_[1].cpu(), w_[1]) and torch.equal(g_[2].cpu(), w_[2]), f"positions of request {i}"


@pytest.mark.parametrize("dim,dtype", [(100, torch.float16), (50, torch.float32), (62, torch.float16), (49, torch.float16)])
@pytest.mark.parametrize("tile,stages", [(4096, 6), (1024, 4)])
def test_bulk_copy_gather_of_pitched_rows(fs, dim, dtype, tile, stages):
    """Rows that are not multiples of 16 bytes in a 128-byte-multiple pitch (ogbn-products: 200 bytes
    in 256): the bulk flavour copies round_up(row, 16) bytes per row into shared memory and the warp
    stores the dense rows (8- or 4-byte words; 98-byte rows are not eligible and take the LDG path)."""
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    N, P, rank = 30000, 4, 3
    X = S.features_by_id(0, N, dim, dtype, device="cuda")
    tab = fs.feature_table(X)
    rb = dim * X.element_size()
    assert tab.pitch % 128 == 0 and tab.pitch > rb
    off = S.equal_partition_offsets(N, P).tolist()
    it = torch.int16 if X.element_size() == 2 else torch.int32
    g = torch.Generator().manual_seed(dim)
    sp = torch.cuda.current_stream().cuda_stream
    try:
        for k_, v_ in (("gather_bulk", 1), ("bulk_tile", tile), ("bulk_stages", stages)):
            _lib.tune(k_, v_)
        for n in (1, 19, 20, 4097, 25013):
            ids = torch.randint(0, N, (n,), generator=g).cuda()
            out = torch.zeros((n + 1, dim), dtype=dtype, device="cuda")
            _lib.check(L.spp_gather_rows_pitched(tab.ptr, tab.pitch, rb, ids.data_ptr(), 1, n, None, out.data_ptr(), n, sp))
            assert torch.equal(out[:n].view(it), S._id_pattern(ids, dim, dtype)) and not bool(out[n:].any())
            cv = torch.randperm(N, generator=g)[:4000]
            cv = cv[(cv < off[rank]) | (cv >= off[rank + 1])]
            cache = fs.Cache(rank, P, cv, X[cv.cuda()].contiguous())
            ctab = cache.device_table()
            ptrs = [tab.ptr + off[p] * tab.pitch for p in range(P)]
            fm = make_feature_map(off, rank, None, ctab.storage, cache.device_index(N), ptrs, tab.pitch, ctab.pitch)
            out.zero_()
            cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n,
                                                cnt.data_ptr(), sp))
            torch.cuda.synchronize()
            assert torch.equal(out[:n].view(it), S._id_pattern(ids, dim, dtype)) and int(cnt.sum()) == n
            assert int(cnt[1]) == int(torch.isin(ids.cpu(), cv).sum())
    finally:
        for k_, v_ in (("gather_bulk", -1), ("bulk_tile", 4096), ("bulk_stages", 6)):
            _lib.tune(k_, v_)


@pytest.mark.parametrize("dim,dtype", [(128, torch.float16), (768, torch.float16), (128, torch.float32), (8, torch.float16)])
@pytest.mark.parametrize("tile,stages", [(4096, 6), (8192, 3), (2048, 8), (256, 4)])
def test_bulk_copy_gather_is_bit_exact(fs, dim, dtype, tile, stages):
    """The bulk-copy (cp.async.bulk + mbarrier) flavour of the gather, forced on through spp_tune:
    single table and partitioned (book search + cache index, and through the owner split's source
    descriptors), ragged row counts, device-side row count."""
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    N, P, rank = 30000, 4, 1
    X = S.features_by_id(0, N, dim, dtype, device="cuda")
    off = S.equal_partition_offsets(N, P).tolist()
    g = torch.Generator().manual_seed(dim + tile)
    rb = dim * X.element_size()
    sp = torch.cuda.current_stream().cuda_stream
    it = torch.int16 if X.element_size() == 2 else torch.int32
    try:
        for k_, v_ in (("gather_bulk", 1), ("bulk_tile", tile), ("bulk_stages", stages)):
            _lib.tune(k_, v_)
        launches0 = _lib.launch_count()
        for n in (1, 31, 33, 4097, 20011):
            ids = torch.randint(0, N, (n,), generator=g).cuda()
            out = torch.zeros((n + 3, dim), dtype=dtype, device="cuda")
            n_dev = torch.tensor([n], dtype=torch.int64, device="cuda")
            _lib.check(L.spp_gather_rows(X.data_ptr(), rb, ids.data_ptr(), 1, n + 3, n_dev.data_ptr(), out.data_ptr(), n + 3, sp))
            assert torch.equal(out[:n].view(it), S._id_pattern(ids, dim, dtype)) and not bool(out[n:].any())
            # partitioned, with a cache holding remote rows
            cv = torch.randperm(N, generator=g)[:5000]
            cv = cv[(cv < off[rank]) | (cv >= off[rank + 1])]
            cache = fs.Cache(rank, P, cv, X[cv.cuda()].contiguous())
            parts = [X[off[p]:off[p + 1]] for p in range(P)]
            fm = make_feature_map(off, rank, parts, cache.device_features(), cache.device_index(N))
            ids32 = ids.to(torch.int32)
            out.zero_()
            cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids32.data_ptr(), 0, n, None, None, out.data_ptr(), n,
                                                cnt.data_ptr(), sp))
            assert torch.equal(out[:n].view(it), S._id_pattern(ids, dim, dtype)) and int(cnt.sum()) == n
            scratch = torch.empty(int(L.spp_split_scratch_words(n)), dtype=torch.int32, device="cuda")
            b_ids, b_perm = torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, dtype=torch.int64, device="cuda")
            b_cnt = torch.zeros(P + 2, dtype=torch.int64, device="cuda")
            _lib.check(L.spp_split_by_owner(ctypes.byref(fm), 1, ids32.data_ptr(), 0, n, None, b_ids.data_ptr(), b_perm.data_ptr(),
                                            b_cnt.data_ptr(), scratch.data_ptr(), sp))
            out.zero_()
            cnt2 = torch.zeros(3, dtype=torch.int64, device="cuda")
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids32.data_ptr(), 0, n, None, scratch.data_ptr(),
                                                out.data_ptr(), n, cnt2.data_ptr(), sp))
            torch.cuda.synchronize()
            assert torch.equal(out[:n].view(it), S._id_pattern(ids, dim, dtype)) and cnt2.tolist() == cnt.tolist()
            assert int(cnt2[1]) == int(b_cnt[P])
        assert _lib.launch_count() > launches0
    finally:
        for k_, v_ in (("gather_bulk", -1), ("bulk_tile", 4096), ("bulk_stages", 6)):
            _lib.tune(k_, v_)


def test_session_batches_are_cuda_graph_replays(fs, data):
    """Every mini-batch of a Session is ONE CUDA-graph launch (the per-batch pointers travel through
    the device job block); ragged last batch and a second Session on the pooled slots replay the
    same graphs; results are bit-exact against the oracle (same kernels as the plain path)."""
    from salient_plusplus_b200 import _lib
    if os.environ.get("SPP_GRAPH", "1") == "0" or os.environ.get("SPP_FORK", "0") != "0":
        pytest.skip("graph replay switched off")
    L = _lib.load()
    rowptr, col, x, y, N = data
    for rep in range(2):
        idx = S.seeds(N, 64 * 7 + 13, seed=21 + rep)
        cfg = fs.Config()
        cfg.x_cpu, cfg.y, cfg.rowptr, cfg.col, cfg.idx = x, y, rowptr, col, idx
        cfg.batch_size, cfg.sizes = 64, [15, 10, 5]
        r0, k0 = int(L.spp_graph_replays()), _lib.launch_count()
        sess = fs.Session(2, 4, cfg)
        n = 0
        while True:
            b = sess.blocking_get_batch()
            if b is None:
                break
            xb, yb, adjs, (st, en) = b
            on, oa = O.multilayer_sample(idx[st:en].numpy(), cfg.sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                         rng_seed=O.session_rng_seed(en))
            assert adjs_equal(adjs, oa) and torch.equal(xb.cpu(), x[torch.from_numpy(on)])
            assert torch.equal(yb.cpu(), y[idx[st:en]])
            n += 1
        assert n == 8
        assert int(L.spp_graph_replays()) - r0 == 8, "batches were not issued as graph launches"
        assert _lib.launch_count() - k0 >= 8 * 12          # kernels inside the graphs are accounted for


@pytest.mark.parametrize("P,hosted", [(4, [1]), (8, [4, 5, 6, 7]), (2, [0])])
@pytest.mark.parametrize("dim,dtype", [(128, torch.float16), (100, torch.float16), (7, torch.float32)])
def test_gather_by_class_equals_fused_gather(fs, P, hosted, dim, dtype):
    """spp_gather_by_class (rows of one set of buckets, walked in bucket order through the owner
    split's inverse permutation): the peer launch + the local launch together write exactly what the
    fused spp_gather_partitioned writes, i.e. x == f(n_id); counters add up class by class."""
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    N = 50000
    X = S.features_by_id(0, N, dim, dtype, device="cuda") if dtype != torch.float32 or dim % 1 == 0 else None
    off = S.equal_partition_offsets(N, P).tolist()
    rank = hosted[0]
    g = torch.Generator().manual_seed(P * 7 + dim)
    n = 30011
    n_id = torch.randperm(N, generator=g)[:n].cuda()
    cv = torch.randperm(N, generator=g)[:6000]
    own = torch.zeros(N, dtype=torch.bool)
    for p in hosted:
        own[off[p]:off[p + 1]] = True
    cv = cv[~own[cv]]
    cache = fs.Cache(rank, P, cv, X[cv.cuda()].contiguous())
    ctab = cache.device_table()
    tabs = [fs.feature_table(X[off[p]:off[p + 1]]) for p in range(P)]
    fm = make_feature_map(off, rank, [t.storage for t in tabs], ctab.storage, cache.device_index(N), None, tabs[0].pitch,
                          ctab.pitch, local_parts=hosted)
    rb = dim * X.element_size()
    sp = torch.cuda.current_stream().cuda_stream
    scratch = torch.empty(int(L.spp_split_scratch_words(n)), dtype=torch.int32, device="cuda")
    b_ids, b_perm = torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, dtype=torch.int64, device="cuda")
    b_cnt = torch.zeros(P + 2, dtype=torch.int64, device="cuda")
    _lib.check(L.spp_split_by_owner(ctypes.byref(fm), 1, n_id.data_ptr(), 1, n, None, b_ids.data_ptr(), b_perm.data_ptr(),
                                    b_cnt.data_ptr(), scratch.data_ptr(), sp))
    inv = scratch[n:2 * n]
    assert torch.equal(b_perm[inv.long()], torch.arange(n, device="cuda"))        # inverse permutation
    peer = sum(1 << p for p in range(P) if p not in hosted)
    local = ((1 << (P + 1)) - 1) & ~peer
    out = torch.zeros((n, dim), dtype=dtype, device="cuda")
    c_peer, c_local = torch.zeros(3, dtype=torch.int64, device="cuda"), torch.zeros(3, dtype=torch.int64, device="cuda")
    _lib.check(L.spp_gather_by_class(ctypes.byref(fm), rb, b_ids.data_ptr(), scratch.data_ptr(), n, peer, out.data_ptr(),
                                     c_peer.data_ptr(), sp))
    it = torch.int16 if X.element_size() == 2 else torch.int32
    want = S._id_pattern(n_id, dim, dtype)
    torch.cuda.synchronize()
    part = torch.searchsorted(torch.tensor(off, device="cuda"), n_id, right=True) - 1
    is_peer = ~own.cuda()[n_id] & ~torch.isin(n_id, cv.cuda())
    got = out.view(it)
    assert torch.equal(got[is_peer], want[is_peer]) and not bool(got[~is_peer].any())   # only peer rows written so far
    _lib.check(L.spp_gather_by_class(ctypes.byref(fm), rb, b_ids.data_ptr(), scratch.data_ptr(), n, local, out.data_ptr(),
                                     c_local.data_ptr(), sp))
    torch.cuda.synchronize()
    assert torch.equal(out.view(it), want)
    assert c_peer.tolist() == [0, 0, int(is_peer.sum())]
    assert c_local.tolist() == [int(own.cuda()[n_id].sum()), int(torch.isin(n_id, cv.cuda()).sum()), 0]
    assert int(part.max()) < P


@pytest.mark.parametrize("tile_rows", [64, 128, 256])
@pytest.mark.parametrize("dim,dtype", [(128, torch.float16), (100, torch.float16), (768, torch.float16), (3, torch.float32)])
def test_gather_tile_heights_are_bit_exact(fs, tile_rows, dim, dtype):
    """k_gather with 64 / 128 / 256 rows per tile (the deep tiles are what a map with peer tables
    uses): single table and partitioned, ragged counts."""
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    N, P, rank = 20000, 4, 2
    X = S.features_by_id(0, N, dim, dtype, device="cuda")
    tab = fs.feature_table(X)
    off = S.equal_partition_offsets(N, P).tolist()
    g = torch.Generator().manual_seed(dim * 3 + tile_rows)
    rb = dim * X.element_size()
    it = torch.int16 if X.element_size() == 2 else torch.int32
    sp = torch.cuda.current_stream().cuda_stream
    try:
        _lib.tune("gather_tile_rows", tile_rows)
        for n in (1, 63, 257, 5000, 12345):
            ids = torch.randint(0, N, (n,), generator=g).cuda()
            out = torch.zeros((n, dim), dtype=dtype, device="cuda")
            _lib.check(L.spp_gather_rows_pitched(tab.ptr, tab.pitch, rb, ids.data_ptr(), 1, n, None, out.data_ptr(), n, sp))
            assert torch.equal(out.view(it), S._id_pattern(ids, dim, dtype))
            ptrs = [tab.ptr + off[p] * tab.pitch for p in range(P)]
            fm = make_feature_map(off, rank, None, None, None, ptrs, tab.pitch, 0)
            out.zero_()
            _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), rb, ids.data_ptr(), 1, n, None, None, out.data_ptr(), n, None, sp))
            torch.cuda.synchronize()
            assert torch.equal(out.view(it), S._id_pattern(ids, dim, dtype))
    finally:
        _lib.tune("gather_tile_rows", 0)
