"""GPU parity tests: every kernel, called through the C ABI (ctypes) and through the
reference-facing Python surface, against the CPU oracle on the same seeded inputs.

Bar: bit-exact for everything (this path is integer / byte work).  The stochastic sampler is
checked bit-for-bit against the oracle's counter-RNG mode, for validity, and with the chi-square
test stated in SURVEY.md section 8(c).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle as O
from salient_plusplus_b200 import synthetic as S
from tests.util import adjs_equal, bounded_degree_graph, small_graph, star_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs():
    from salient_plusplus_b200 import fast_sampler
    return fast_sampler


# ------------------------------------------------------------------------------------------------
# K4 feature gather
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,dtype", [(100, torch.float16), (128, torch.float16), (128, torch.float32),
                                       (768, torch.float16), (1, torch.int64), (7, torch.uint8),
                                       (3, torch.int16), (5, torch.float32), (4097, torch.float16)])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_gather_rows_bitexact(fs, dim, dtype, idx_dtype):
    g = torch.Generator().manual_seed(11)
    n = 5000
    if dtype.is_floating_point:
        x = torch.randn((n, dim), generator=g).to(dtype)
    else:
        x = torch.randint(0, 100, (n, dim), generator=g).to(dtype)
    idx = torch.randint(0, n, (7777,), generator=g).to(idx_dtype)
    out = fs.serial_index(x, idx)
    want = O.serial_index(x.numpy() if dtype != torch.float16 else x.view(torch.int16).numpy(), idx.numpy())
    got = out.cpu()
    got = got.view(torch.int16).numpy() if dtype == torch.float16 else got.numpy()
    assert np.array_equal(got, want)


def test_gather_rows_n_argument_and_empty(fs):
    x = torch.arange(60, dtype=torch.float32).view(20, 3)
    idx = torch.tensor([5, 1, 19, 0], dtype=torch.int64)
    out = fs.serial_index(x, idx, 2)                        # n < len(idx): first two rows only
    assert out.shape == (2, 3) and torch.equal(out.cpu(), x[idx[:2]])
    out = fs.serial_index(x, idx, 6)                        # n > len(idx): tail rows unspecified
    assert out.shape == (6, 3) and torch.equal(out[:4].cpu(), x[idx])
    out = fs.serial_index(x, torch.empty(0, dtype=torch.int64))
    assert out.shape == (0, 3)
    with pytest.raises(RuntimeError):
        fs.serial_index(x.t(), idx)                         # not row-major -> TORCH_CHECK analogue


def test_gather_device_count(fs):
    """Row count taken from device memory (the way the Session chains sampler -> gather)."""
    from salient_plusplus_b200 import _lib
    L = _lib.load()
    x = torch.randn(1000, 64, device="cuda").half()
    idx = torch.randint(0, 1000, (500,), device="cuda", dtype=torch.int32)
    out = torch.zeros(500, 64, device="cuda", dtype=torch.float16)
    n_dev = torch.tensor([123], device="cuda", dtype=torch.int64)
    _lib.check(L.spp_gather_rows(x.data_ptr(), 128, idx.data_ptr(), 0, 500, n_dev.data_ptr(), out.data_ptr(), 500,
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out[:123], x[idx[:123].long()])
    assert not out[123:].any()


# ------------------------------------------------------------------------------------------------
# K1 sampling: deterministic paths are bit-exact against the reference semantics
# ------------------------------------------------------------------------------------------------
def test_hand_kat(fs):
    rows = {0: [5, 3, 1], 1: [0, 5], 2: [], 3: [0], 4: [4], 5: [0, 1]}
    rowptr, col = [0], []
    for i in range(6):
        col += rows[i]
        rowptr.append(len(col))
    rowptr, col = torch.tensor(rowptr), torch.tensor(col)
    n_id, adjs = fs.multilayer_sample(torch.tensor([1, 0, 2, 4]), [-1, -1], rowptr, col)
    assert n_id.tolist() == [1, 0, 2, 4, 5, 3]
    assert adjs[0][0].tolist() == [0, 2, 5, 5, 6, 8, 9] and adjs[0][1].tolist() == [1, 4, 0, 4, 5, 3, 0, 1, 1]
    assert tuple(adjs[0][3]) == (6, 6)
    assert adjs[1][0].tolist() == [0, 2, 5, 5, 6] and adjs[1][1].tolist() == [1, 4, 0, 4, 5, 3]
    assert tuple(adjs[1][3]) == (4, 6)


@pytest.mark.parametrize("sizes", [[-1], [-1, -1], [-1, -1, -1]])
def test_full_neighbourhood_bitexact(fs, sizes):
    rowptr, col = small_graph()
    idx = S.seeds(rowptr.numel() - 1, 257)
    idx[5] = idx[100]                                        # duplicated seed: last position wins
    n_id, adjs = fs.multilayer_sample(idx, sizes, rowptr, col)
    on, oa = O.multilayer_sample(idx.numpy(), sizes, rowptr.numpy(), col.numpy())
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


def test_fanout_at_least_degree_matches_reference_rng_mode(fs):
    """deg <= fanout everywhere: no randomness is consumed, so the output must equal the
    reference algorithm (mt19937 mode of the oracle) bit for bit."""
    rowptr, col = bounded_degree_graph(max_deg=5)
    idx = S.seeds(rowptr.numel() - 1, 300)
    n_id, adjs = fs.multilayer_sample(idx, [15, 10, 5], rowptr, col)
    on, oa = O.multilayer_sample(idx.numpy(), [15, 10, 5], rowptr.numpy(), col.numpy(), rng_mode=O.RNG_REFERENCE)
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


def test_sample_adj_single_hop_and_int32_nid(fs):
    rowptr, col = small_graph()
    idx = S.seeds(rowptr.numel() - 1, 64)
    rp, cl, n_id, e_id = fs.sample_adj(rowptr, col, idx, -1, False)
    orp, ocl, on, _ = O.sample_adj(rowptr.numpy(), col.numpy(), idx.numpy(), -1, False)
    assert n_id.dtype == torch.int32 and e_id.numel() == 0
    assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(cl.cpu().numpy(), ocl)
    assert np.array_equal(n_id.cpu().numpy(), on)


def test_long_rows_sort_paths(fs):
    """Rows longer than a warp (shared-memory warp sort), than 1024 (CTA sort in shared memory)
    and than 48 K (CTA sort in global memory)."""
    degs = [0, 1, 31, 32, 33, 100, 1024, 1025, 5000, 50000, 3]
    n = 60000
    rng = np.random.default_rng(5)
    rowptr, col = [0], []
    for d in degs:
        col.append(rng.permutation(n)[:d])                  # unsorted neighbour order
        rowptr.append(rowptr[-1] + d)
    rowptr += [rowptr[-1]] * (n - len(degs))
    rowptr = torch.tensor(rowptr, dtype=torch.int64)
    col = torch.from_numpy(np.concatenate(col).astype(np.int64))
    idx = torch.arange(len(degs), dtype=torch.int64)
    n_id, adjs = fs.multilayer_sample(idx, [-1], rowptr, col)
    on, oa = O.multilayer_sample(idx.numpy(), [-1], rowptr.numpy(), col.numpy())
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


def _rows_graph(rows, n):
    rowptr, col = [0], []
    for r in rows:
        col.append(np.asarray(r, dtype=np.int64))
        rowptr.append(rowptr[-1] + len(r))
    rowptr += [rowptr[-1]] * (n - len(rows))
    return torch.tensor(rowptr, dtype=torch.int64), torch.from_numpy(np.concatenate(col))


def test_full_rows_with_repeated_neighbours(fs):
    """Multigraph rows: a repeated neighbour defeats the bitmap row sort, so those rows must reach
    the comparison sorter (second work list) while clean rows of the same hop stay on the bitmap."""
    n = 20000
    rng = np.random.default_rng(11)
    rows = [rng.permutation(n)[:300],                                   # clean, > 32
            np.concatenate([rng.permutation(n)[:200], [7, 7, 7]]),       # one id three times
            rng.integers(0, 50, size=400),                               # heavy repetition
            rng.permutation(n)[:3000],                                   # clean, > 1024
            np.concatenate([rng.permutation(n)[:2000], rng.permutation(n)[:2000]]),  # many pairs
            rng.permutation(n)[:20]]                                     # register path
    rowptr, col = _rows_graph(rows, n)
    idx = torch.arange(len(rows), dtype=torch.int64)
    n_id, adjs = fs.multilayer_sample(idx, [-1], rowptr, col)
    on, oa = O.multilayer_sample(idx.numpy(), [-1], rowptr.numpy(), col.numpy())
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


@pytest.mark.parametrize("n,row_len,n_rows", [(400_000, 30_000, 9), (2_600_000, 45_000, 42)])
def test_full_rows_large_node_counts(fs, n, row_len, n_rows):
    """Node counts beyond the small bitmap (128 K ids: second launch with the 200 KB bitmap) and
    beyond both bitmaps (1.6 M ids: every long row falls back to the comparison sorter)."""
    rng = np.random.default_rng(13)
    perm = rng.permutation(n)
    rows = [perm[i * row_len:(i + 1) * row_len][rng.permutation(row_len)] for i in range(n_rows)]
    rows.append(perm[:40])
    rowptr, col = _rows_graph(rows, n)
    idx = torch.arange(len(rows), dtype=torch.int64)
    n_id, adjs = fs.multilayer_sample(idx, [-1], rowptr, col)
    on, oa = O.multilayer_sample(idx.numpy(), [-1], rowptr.numpy(), col.numpy())
    assert n_id.numel() > (128 * 1024 if n < 1_000_000 else 1_638_400)
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


def test_empty_and_isolated(fs):
    rowptr, col = bounded_degree_graph(n=50, max_deg=0)
    n_id, adjs = fs.multilayer_sample(torch.tensor([3, 4, 5]), [15, 10], rowptr, col)
    assert n_id.tolist() == [3, 4, 5]
    assert all(a[1].numel() == 0 for a in adjs) and adjs[0][0].tolist() == [0, 0, 0, 0]
    rowptr, col = small_graph()
    n_id, adjs = fs.multilayer_sample(torch.empty(0, dtype=torch.int64), [5, 5], rowptr, col)
    assert n_id.numel() == 0 and all(a[0].tolist() == [0] and a[1].numel() == 0 for a in adjs)


# ------------------------------------------------------------------------------------------------
# K1 sampling: stochastic path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes", [[15, 10, 5], [25, 15], [40, 3], [100], [128, 2], [4, 4, 4, 4], [1], [32, 16, 8]])
def test_stochastic_bitexact_vs_counter_oracle(fs, sizes):
    rowptr, col = small_graph(n=4000, e=120000)
    idx = S.seeds(rowptr.numel() - 1, 200)
    n_id, adjs = fs.multilayer_sample(idx, sizes, rowptr, col, seed=12345)
    on, oa = O.multilayer_sample(idx.numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                 rng_seed=12345)
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)


def test_with_replacement_vs_counter_oracle(fs):
    rowptr, col = small_graph()
    idx = S.seeds(rowptr.numel() - 1, 128)
    for k in (3, 12, 40):
        rp, cl, n_id, _ = fs.sample_adj(rowptr, col, idx, k, True, seed=77)
        orp, ocl, on, _ = O.sample_adj(rowptr.numpy(), col.numpy(), idx.numpy(), k, True, rng_mode=O.RNG_COUNTER,
                                       rng_seed=77)
        assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(cl.cpu().numpy(), ocl)
        assert np.array_equal(n_id.cpu().numpy(), on)


def test_sampled_rows_are_valid(fs):
    rowptr, col = small_graph(n=4000, e=120000)
    N = rowptr.numel() - 1
    idx = S.seeds(N, 512)
    k = 10
    rp, cl, n_id, _ = fs.sample_adj(rowptr, col, idx, k, False, seed=5)
    rp, cl, n_id = rp.cpu().numpy(), cl.cpu().numpy(), n_id.cpu().numpy()
    rpn, coln = rowptr.numpy(), col.numpy()
    assert len(np.unique(n_id)) == len(n_id)                 # n_id holds each node once
    for i, v in enumerate(idx.numpy()):
        row = cl[rp[i]:rp[i + 1]]
        nbrs = coln[rpn[v]:rpn[v + 1]]
        assert len(row) == min(k, len(nbrs))
        assert np.all(np.diff(row) > 0)                       # ascending, distinct local ids
        assert np.all(np.isin(n_id[row], nbrs))               # all are real neighbours


@pytest.mark.parametrize("D,k", [(3, 2), (6, 2), (6, 5), (20, 5), (20, 15), (40, 5), (40, 15), (1000, 15), (1000, 2)])
def test_chi_square_uniform(fs, D, k):
    """T = 20 000 independent draws of k out of D neighbours; per-neighbour
    inclusion counts against T*k/D, Pearson chi-square with D-1 dof, p > 1e-3.  (The reference
    sampler fails this test by orders of magnitude; fast_sampler/sample_cpu.hpp:99.)"""
    from scipy.stats import chi2
    T = 20000
    rowptr, col = star_graph(D, T)
    idx = torch.arange(D, D + T, dtype=torch.int64)
    rp, cl, n_id, _ = fs.sample_adj(rowptr, col, idx, k, False, seed=2024 + D * 31 + k)
    rp, cl, n_id = rp.cpu().numpy(), cl.cpu().numpy(), n_id.cpu().numpy()
    assert np.all(np.diff(rp) == k)
    leaves = n_id[cl]
    assert leaves.min() >= 0 and leaves.max() < D
    counts = np.bincount(leaves, minlength=D).astype(np.float64)
    expected = T * k / D
    # Plain Pearson statistic (the stated test; conservative because the k inclusion indicators of
    # one draw are negatively correlated) ...
    pearson = ((counts - expected) ** 2 / expected).sum()
    assert chi2.sf(pearson, D - 1) > 1e-3, (D, k, pearson)
    # ... and the exactly calibrated one: Cov(X) = T*pi*(1-pi)*D/(D-1) * (I - J/D), pi = k/D, so
    # pearson * (D-1) / (D*(1-pi)) ~ chi2(D-1).  Stricter than the stated test.
    stat = pearson * (D - 1) / (D * (1.0 - k / D))
    p = chi2.sf(stat, D - 1)
    assert p > 1e-3, (D, k, stat, p)


# ------------------------------------------------------------------------------------------------
# K2 partition book / cache
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P", [1, 2, 4, 8, 16])
def test_partition_book(fs, P):
    N = 100003
    off = S.equal_partition_offsets(N, P)
    g = torch.Generator().manual_seed(P)
    nids = torch.randint(0, N, (20000,), generator=g)
    nids[:P + 1] = off.clamp(max=N - 1)                       # boundaries
    rank = P // 2
    book = fs.RangePartitionBook(rank, P, off)
    assert np.array_equal(book.nid2partid(nids).numpy(), O.nid2partid(off.numpy(), nids.numpy()))
    assert np.array_equal(book.nid2localnid(nids, rank).numpy(), O.nid2localnid(off.numpy(), nids.numpy(), rank))
    assert np.array_equal(book.nid_is_local(nids).numpy(), O.nid_is_local(off.numpy(), rank, nids.numpy()))
    assert np.array_equal(book.partid2nids(rank).numpy(), O.partid2nids(off.numpy(), rank))
    d = book.nid2partid(nids.cuda())
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), O.nid2partid(off.numpy(), nids.numpy()))


def test_cache_lookup(fs):
    N = 50000
    g = torch.Generator().manual_seed(3)
    cv = torch.randperm(N, generator=g)[:4000]
    cv[17] = cv[3]                                            # duplicate: later index wins
    feats = torch.randn(4000, 16).half()
    c = fs.Cache(0, 2, cv, feats)
    oc = O.Cache(cv.numpy(), N)
    nids = torch.randint(0, N, (30000,), generator=g)
    assert np.array_equal(c.nid_is_cached(nids).numpy(), oc.nid_is_cached(nids.numpy()))
    hit = nids[c.nid_is_cached(nids)]
    assert np.array_equal(c.nid2cachenid(hit).numpy(), oc.nid2cachenid(hit.numpy()))
    assert c.cached_vertices is cv and c.cached_features is feats and c.rank == 0 and c.world_size == 2
    e = fs.Cache()
    assert e.cached_vertices.numel() == 0 and e.cached_features.dtype == torch.float16


# ------------------------------------------------------------------------------------------------
# K3 split, K4+K5 fused gather (single-GPU emulation of P partitions)
# ------------------------------------------------------------------------------------------------
def _split_via_abi(n_id, off, rank, use_cache, cache_map):
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    P = len(off) - 1
    n = n_id.numel()
    fm = make_feature_map(off, rank, [None] * P)
    if use_cache:
        fm.cache_index, fm.cache_index_nodes = cache_map[0].data_ptr(), cache_map[1]
    words = int(L.spp_split_scratch_words(n))
    scratch = torch.empty(words, dtype=torch.int32, device="cuda")
    ids = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    perm = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    counts = torch.zeros(P + 2, dtype=torch.int64, device="cuda")
    _lib.check(L.spp_split_by_owner(ctypes.byref(fm), int(use_cache), n_id.data_ptr(), int(n_id.dtype == torch.int64),
                                    n, None, ids.data_ptr(), perm.data_ptr(), counts.data_ptr(), scratch.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    # the per-node source descriptors the split leaves for the fused gather: p >= 0 -> partition, < 0 -> ~cache row
    desc = scratch[:n].cpu().numpy()
    pos = perm[:n].cpu().numpy()
    bounds = np.cumsum([0] + counts.cpu().numpy()[:P + 1].tolist())
    owner = np.searchsorted(bounds, pos, side="right") - 1
    assert np.array_equal(np.where(desc < 0, P, desc), owner)
    assert np.array_equal((~desc)[desc < 0], ids[:n].cpu().numpy()[pos[desc < 0]])
    return ids[:n].cpu().numpy(), perm[:n].cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize("P,rank", [(1, 0), (2, 1), (4, 2), (8, 0), (8, 7), (16, 5)])
@pytest.mark.parametrize("use_cache", [False, True])
@pytest.mark.parametrize("idt", [torch.int64, torch.int32])
def test_split_by_owner_bitexact(fs, P, rank, use_cache, idt):
    N = 40000
    off = S.equal_partition_offsets(N, P).tolist()
    g = torch.Generator().manual_seed(P * 10 + rank)
    n_id = torch.randperm(N, generator=g)[:9001]
    cv = torch.randperm(N, generator=g)[:3000]
    lo, hi = off[rank], off[rank + 1]
    cv = cv[(cv < lo) | (cv >= hi)]                           # caches hold remote vertices only (ddp.py:512)
    cache = fs.Cache(rank, P, cv, torch.zeros(cv.numel(), 4).half())
    cmap = cache.device_index(N)
    ids, perm, counts = _split_via_abi(n_id.to(idt).cuda(), off, rank, use_cache, cmap)
    pn, cn, operm, _ = O.distributed_binning(n_id.numpy(), np.array(off), rank, P, 10 ** 9, use_cache,
                                             O.Cache(cv.numpy(), N) if use_cache else None)
    want_ids = np.concatenate(pn + [cn])
    assert counts[:P].tolist() == [len(p) for p in pn] and counts[P] == len(cn) and counts[P + 1] == n_id.numel()
    assert np.array_equal(ids, want_ids)
    assert np.array_equal(perm, operm)


def test_split_empty(fs):
    ids, perm, counts = _split_via_abi(torch.empty(0, dtype=torch.int64, device="cuda"), [0, 10, 20], 0, False, None)
    assert counts.tolist() == [0, 0, 0, 0]


@pytest.mark.parametrize("P,rank,use_cache", [(1, 0, False), (4, 1, False), (8, 3, True), (2, 0, True)])
@pytest.mark.parametrize("dim,dtype", [(128, torch.float16), (100, torch.float16), (768, torch.float16)])
def test_gather_partitioned_equals_global_gather(fs, P, rank, use_cache, dim, dtype):
    """x == X_global[n_id]: every row, whether served from the local
    partition, the replicated cache or a peer partition, is the node's global feature row."""
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    L = _lib.load()
    N = 20000
    X = S.features(N, dim, dtype, seed=5)
    off = S.equal_partition_offsets(N, P).tolist()
    parts = [X[off[p]:off[p + 1]].cuda().contiguous() for p in range(P)]
    g = torch.Generator().manual_seed(1)
    n_id = torch.randint(0, N, (15001,), generator=g)
    cmap = cfeat = None
    if use_cache:
        cv = torch.randperm(N, generator=g)[:4000]
        cv = cv[(cv < off[rank]) | (cv >= off[rank + 1])]
        cache = fs.Cache(rank, P, cv, X[cv].contiguous())
        cmap, cfeat = cache.device_index(N), cache.device_features()
        # poison the peer copies of cached rows: they must be served from the cache
        for p in range(P):
            if p != rank:
                sel = cv[(cv >= off[p]) & (cv < off[p + 1])] - off[p]
                parts[p][sel.cuda()] = 0
    fm = make_feature_map(off, rank, parts, cfeat, cmap)
    out = torch.empty((n_id.numel(), dim), dtype=dtype, device="cuda")
    counters = torch.zeros(3, dtype=torch.int64, device="cuda")
    ids = n_id.cuda()
    _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), dim * X.element_size(), ids.data_ptr(), 1, ids.numel(), None,
                                        None, out.data_ptr(), ids.numel(), counters.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), X[n_id])
    # same rows through the descriptors of the owner split (what a Session's batch does)
    words = int(L.spp_split_scratch_words(ids.numel()))
    scratch = torch.empty(words, dtype=torch.int32, device="cuda")
    b_ids, b_perm = torch.empty_like(ids), torch.empty_like(ids)
    b_counts = torch.zeros(P + 2, dtype=torch.int64, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    _lib.check(L.spp_split_by_owner(ctypes.byref(fm), int(use_cache), ids.data_ptr(), 1, ids.numel(), None, b_ids.data_ptr(),
                                    b_perm.data_ptr(), b_counts.data_ptr(), scratch.data_ptr(), sp))
    out2 = torch.zeros_like(out)
    counters2 = torch.zeros(3, dtype=torch.int64, device="cuda")
    _lib.check(L.spp_gather_partitioned(ctypes.byref(fm), dim * X.element_size(), ids.data_ptr(), 1, ids.numel(), None,
                                        scratch.data_ptr(), out2.data_ptr(), ids.numel(), counters2.data_ptr(), sp))
    torch.cuda.synchronize()
    assert torch.equal(out2, out) and counters2.tolist() == counters.tolist()
    local = ((n_id >= off[rank]) & (n_id < off[rank + 1])).sum().item()
    c = counters.tolist()
    assert c[0] == local and sum(c) == n_id.numel()
    if use_cache:
        assert c[1] == int(torch.isin(n_id, cv).sum())
    else:
        assert c[1] == 0


@pytest.mark.parametrize("dim,dtype,pitch_elems", [(100, torch.float16, 128), (50, torch.float32, 64), (7, torch.uint8, 16)])
def test_gather_rows_pitched(fs, dim, dtype, pitch_elems):
    """Source rows `pitch` bytes apart (the padded resident layout), dense output."""
    from salient_plusplus_b200 import _lib
    L = _lib.load()
    g = torch.Generator().manual_seed(3)
    n = 3000
    dense = (torch.randn((n, dim), generator=g) * 50).to(dtype)
    padded = torch.full((n, pitch_elems), 77, dtype=dtype)
    padded[:, :dim] = dense
    idx = torch.randint(0, n, (5001,), generator=g)
    pd, ix = padded.cuda(), idx.cuda()
    out = torch.empty((idx.numel(), dim), dtype=dtype, device="cuda")
    es = dense.element_size()
    _lib.check(L.spp_gather_rows_pitched(pd.data_ptr(), pitch_elems * es, dim * es, ix.data_ptr(), 1, ix.numel(), None,
                                         out.data_ptr(), ix.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), dense[idx])
    t = fs.feature_table(dense)
    assert t.row_bytes == dim * es and t.pitch % es == 0 and t.pitch >= t.row_bytes
    if dim * es == 200:
        assert t.pitch == 256
        assert torch.equal(t.storage[:, :dim].cpu(), dense)


@pytest.mark.parametrize("case", range(24))
def test_randomized_gather_alignment_and_shapes(fs, case):
    """Random row sizes (1..1500 bytes), misaligned table / output bases (exercises the 16/8/4/2/1
    byte vector paths), pitched sources, int32/int64 indices, device-side row count."""
    from salient_plusplus_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(7000 + case)
    row_bytes = int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16, 24, 100, 200, 256, 264, 512, 1000, 1500]))
    pitch = row_bytes + int(rng.choice([0, 0, 8, 56, 13])) if row_bytes > 1 else row_bytes
    n_rows, n_idx = int(rng.integers(1, 3000)), int(rng.integers(0, 5000))
    src_off, dst_off = int(rng.choice([0, 0, 1, 2, 4, 8, 16])), int(rng.choice([0, 0, 1, 2, 4, 8, 16]))
    tbuf = torch.from_numpy(rng.integers(0, 256, size=n_rows * pitch + 64, dtype=np.uint8)).cuda()
    table = tbuf[src_off:src_off + n_rows * pitch]
    idx64 = torch.from_numpy(rng.integers(0, n_rows, size=n_idx)).to(torch.int64)
    use64 = bool(rng.integers(0, 2))
    idx = (idx64 if use64 else idx64.to(torch.int32)).cuda()
    n_out = max(1, int(n_idx * rng.choice([0.5, 1.0, 1.5])))
    obuf = torch.full((n_out * row_bytes + 64,), 0xAB, dtype=torch.uint8, device="cuda")
    out = obuf[dst_off:dst_off + n_out * row_bytes]
    n_dev = None
    n_eff = min(n_idx, n_out)
    if rng.integers(0, 2):
        n_eff = int(rng.integers(0, n_eff + 1))
        n_dev = torch.tensor([n_eff], dtype=torch.int64, device="cuda")
    _lib.check(L.spp_gather_rows_pitched(table.data_ptr(), pitch, row_bytes, idx.data_ptr(), int(use64), n_idx,
                                         n_dev.data_ptr() if n_dev is not None else None, out.data_ptr(), n_out,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    t2 = table.cpu().view(n_rows, pitch)[:, :row_bytes]
    want = t2[idx64[:n_eff]]
    got = out.cpu().view(n_out, row_bytes)
    assert torch.equal(got[:n_eff], want)
    assert bool((got[n_eff:] == 0xAB).all())                      # nothing written past the row count
    assert bool((obuf[:dst_off] == 0xAB).all()) and bool((obuf[dst_off + n_out * row_bytes:] == 0xAB).all())


@pytest.mark.parametrize("case", range(12))
def test_randomized_book_cache_and_sample_adj(fs, case):
    rng = np.random.default_rng(8000 + case)
    N = int(rng.integers(10, 5000))
    P = int(rng.choice([1, 2, 5, 8, 16]))
    cuts = np.sort(rng.integers(0, N + 1, size=P - 1)) if P > 1 else np.empty(0, dtype=np.int64)
    off = torch.tensor([0] + cuts.tolist() + [N], dtype=torch.int64)       # ragged / empty partitions
    nids = torch.from_numpy(rng.integers(0, N, size=int(rng.integers(0, 3000)))).to(torch.int64)
    rank = int(rng.integers(0, P))
    book = fs.RangePartitionBook(rank, P, off)
    assert np.array_equal(book.nid2partid(nids).numpy(), O.nid2partid(off.numpy(), nids.numpy()))
    assert np.array_equal(book.nid2localnid(nids, rank).numpy(), O.nid2localnid(off.numpy(), nids.numpy(), rank))
    assert np.array_equal(book.nid_is_local(nids).numpy(), O.nid_is_local(off.numpy(), rank, nids.numpy()))
    cv = torch.from_numpy(rng.integers(0, N, size=int(rng.integers(0, N)))).to(torch.int64)   # duplicates allowed
    c, oc = fs.Cache(rank, P, cv, torch.zeros(cv.numel(), 2).half()), O.Cache(cv.numpy(), N)
    assert np.array_equal(c.nid_is_cached(nids).numpy(), oc.nid_is_cached(nids.numpy()))
    hit = nids[c.nid_is_cached(nids)]
    assert np.array_equal(c.nid2cachenid(hit).numpy(), oc.nid2cachenid(hit.numpy()))
    # one-hop sample_adj in all three modes on a random multigraph
    n = int(rng.integers(5, 300))
    rowptr, col = [0], []
    for _ in range(n):
        d = int(rng.integers(0, 50))
        col.extend(rng.integers(0, n, size=d).tolist())
        rowptr.append(len(col))
    rowptr, col = torch.tensor(rowptr), torch.tensor(col, dtype=torch.int64)
    idx = torch.from_numpy(rng.integers(0, n, size=int(rng.integers(1, 100)))).to(torch.int64)
    for k, rep in ((-1, False), (int(rng.integers(1, 60)), False), (int(rng.integers(1, 60)), True)):
        rp, cl, n_id, _ = fs.sample_adj(rowptr, col, idx, k, rep, seed=99 + case)
        orp, ocl, on, _ = O.sample_adj(rowptr.numpy(), col.numpy(), idx.numpy(), k, rep, rng_mode=O.RNG_COUNTER,
                                       rng_seed=99 + case)
        assert np.array_equal(rp.cpu().numpy(), orp) and np.array_equal(cl.cpu().numpy(), ocl)
        assert np.array_equal(n_id.cpu().numpy(), on)


@pytest.mark.parametrize("mode", ["hash", "direct"])
@pytest.mark.parametrize("sizes", [[15, 10, 5], [-1, -1], [40, 3], [5, -1]])
def test_both_table_flavours(fs, monkeypatch, mode, sizes):
    """The id table is direct-mapped for small graphs and hashed (L2 resident) for large ones;
    both flavours must give identical, oracle-exact results (SPP_TABLE forces the choice)."""
    monkeypatch.setenv("SPP_TABLE", mode)
    rowptr, col = small_graph(n=4000, e=120000)
    idx = S.seeds(rowptr.numel() - 1, 200)
    idx[3] = idx[77]
    n_id, adjs = fs.multilayer_sample(idx, sizes, rowptr, col, seed=4242)
    on, oa = O.multilayer_sample(idx.numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER, rng_seed=4242)
    assert np.array_equal(n_id.cpu().numpy(), on)
    assert adjs_equal(adjs, oa)
