"""GPU tests of the reference-facing Session / FastSampler surface against the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from salient_plusplus_b200 import synthetic as S
from tests.util import adjs_equal, small_graph

pytestmark = pytest.mark.gpu


def _config(fs, ds_x, y, rowptr, col, idx, **kw):
    cfg = fs.Config()
    cfg.x_cpu = ds_x
    cfg.x_gpu = torch.empty((0, ds_x.size(1)), dtype=ds_x.dtype)
    cfg.y = y
    cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
    cfg.batch_size = 64
    cfg.sizes = [15, 10, 5]
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


@pytest.fixture(scope="module")
def fs():
    from salient_plusplus_b200 import fast_sampler
    return fast_sampler


@pytest.fixture(scope="module")
def data():
    rowptr, col = small_graph(n=6000, e=150000)
    N = rowptr.numel() - 1
    x = S.features(N, 100, torch.float16, seed=9)
    y = S.labels(N, seed=4)
    return rowptr, col, x, y, N


@pytest.mark.parametrize("sizes", [[15, 10, 5], [-1], [5, -1]])
@pytest.mark.parametrize("kw", [dict(), dict(skip_nonfull_batch=True), dict(force_exact_num_batches=True, exact_num_batches=7)])
def test_session_nondistributed(fs, data, sizes, kw):
    rowptr, col, x, y, N = data
    idx = S.seeds(N, 500)
    cfg = _config(fs, x, y, rowptr, col, idx, sizes=sizes, **kw)
    sess = fs.Session(4, 3, cfg)
    want_ranges = O.batch_ranges(idx.numel(), 64, cfg.skip_nonfull_batch, cfg.force_exact_num_batches,
                                 cfg.exact_num_batches)
    assert sess.num_total_batches == len(want_ranges)
    got = []
    while True:
        b = sess.blocking_get_batch()
        if b is None:
            break
        xb, yb, adjs, (st, en) = b
        got.append((st, en))
        on, oa = O.multilayer_sample(idx[st:en].numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal(adjs, oa)
        assert torch.equal(xb.cpu(), x[torch.from_numpy(on)])
        assert torch.equal(yb.cpu(), y[idx[st:en]])
    assert sorted(got) == want_ranges
    assert sess.num_consumed_batches == sess.num_total_batches
    assert sess.blocking_get_batch() is None
    assert sess.total_blocked_occasions >= 0 and sess.total_blocked_dur.total_seconds() >= 0


def test_session_without_labels_and_config_copy(fs, data):
    rowptr, col, x, y, N = data
    idx = S.seeds(N, 100)
    cfg = _config(fs, x, None, rowptr, col, idx)
    sess = fs.Session(1, 10, cfg)
    cfg.sizes = [1]                                          # mutating after construction has no effect
    xb, yb, adjs, rng = sess.blocking_get_batch()
    assert yb is None and len(adjs) == 3 and sess.config.sizes == [15, 10, 5]
    with pytest.raises(RuntimeError):
        fs.Session(1, 0, cfg)


@pytest.mark.parametrize("use_cache", [False, True])
@pytest.mark.parametrize("P,rank", [(4, 1), (8, 6)])
def test_session_distributed_single_process(fs, data, use_cache, P, rank):
    """All P partitions live on this GPU (Config.partition_tables): ProtoDistributedBatch fields
    against the oracle's restatement of fast_sampler.cpp:1017-1262 and x == X[n_id]."""
    rowptr, col, x, y, N = data
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    idx = S.seeds(N, 300, lo=lo, hi=hi)
    cut = (hi - lo) * 2 // 3
    cv = S.degree_cache_vertices(rowptr, off, rank, 500)
    cfg = _config(fs, x[lo + cut:hi].contiguous(), y, rowptr, col, idx, distributed=True, use_cache=use_cache,
                  force_exact_num_batches=True, exact_num_batches=4)
    cfg.x_gpu = x[lo:lo + cut].contiguous()
    cfg.partition_book = fs.RangePartitionBook(rank, P, off)
    cfg.cache = fs.Cache(rank, P, cv, x[cv].contiguous()) if use_cache else fs.Cache()
    cfg.partition_tables = [x[int(off[p]):int(off[p + 1])].contiguous() if p != rank else None for p in range(P)]
    sess = fs.Session(2, 8, cfg)
    oc = O.Cache(cv.numpy(), N) if use_cache else None
    ranges = O.batch_ranges(idx.numel(), 64, False, True, 4)
    for want in ranges:                                      # distributed batches arrive in order
        b = sess.blocking_get_batch_distributed()
        assert tuple(b.idx_range) == want
        st, en = want
        on, oa = O.multilayer_sample(idx[st:en].numpy(), cfg.sizes, rowptr.numpy(), col.numpy(),
                                     rng_mode=O.RNG_COUNTER, rng_seed=O.session_rng_seed(en))
        pn, cn, perm, loc_cpu = O.distributed_binning(on, off.numpy(), rank, P, cut, use_cache, oc)
        assert adjs_equal(b.adjs, oa)
        assert len(b.partition_nids) == P
        for a, w in zip(b.partition_nids, pn):
            assert np.array_equal(a.cpu().numpy(), w)
        assert np.array_equal(b.cached_nids.cpu().numpy(), cn)
        assert np.array_equal(b.perm_partition_to_mfg.cpu().numpy(), perm)
        assert torch.equal(b.sliced_cpu_labels.cpu(), y[idx[st:en]])
        assert torch.equal(b.sliced_cpu_features.cpu(), cfg.x_cpu[torch.from_numpy(loc_cpu)])
        assert np.array_equal(b.n_id.cpu().numpy(), on)
        assert torch.equal(b.x.cpu(), x[torch.from_numpy(on)])
    assert sess.blocking_get_batch_distributed() is None


def _random_multigraph(rng, n, max_deg):
    rowptr, col = [0], []
    for _ in range(n):
        d = int(rng.integers(0, max_deg + 1))
        col.extend(rng.integers(0, n, size=d).tolist())       # unsorted, duplicate neighbours, self loops
        rowptr.append(len(col))
    return torch.tensor(rowptr, dtype=torch.int64), torch.tensor(col, dtype=torch.int64)


@pytest.mark.parametrize("case", range(32))
def test_randomized_sessions_against_oracle(fs, case):
    """Seeded sweep mirroring tests/test_oracle_vs_ref.py: random multigraphs (isolated nodes, self
    loops, duplicate neighbours), duplicate seeds, every fan-out regime incl. > 32 and full
    neighbourhood, mixed hops, both batch-range modes, host or device idx -- bit-exact."""
    rng = np.random.default_rng(5000 + case)
    n = int(rng.integers(5, 600))
    rowptr, col = _random_multigraph(rng, n, int(rng.integers(0, 60)))
    L = int(rng.integers(1, 4))
    sizes = [int(rng.choice([-1, 1, 2, 5, 15, 25, 40, 100])) for _ in range(L)]
    idx = torch.from_numpy(rng.integers(0, n, size=int(rng.integers(1, 300)))).to(torch.int64)
    x = torch.from_numpy(rng.integers(0, 1000, size=(n, int(rng.choice([3, 50, 64]))))).to(torch.float16)
    y = torch.from_numpy(rng.integers(0, 9, size=(n, 1)))
    cfg = fs.Config()
    cfg.x_cpu, cfg.y, cfg.rowptr, cfg.col = x, y, rowptr, col
    cfg.idx = idx.cuda() if case % 3 == 0 else idx
    cfg.batch_size, cfg.sizes = int(rng.integers(1, 128)), sizes
    cfg.skip_nonfull_batch = bool(rng.integers(0, 2))
    exact = bool(rng.integers(0, 2)) and idx.numel() >= 8
    cfg.force_exact_num_batches, cfg.exact_num_batches = exact, (int(rng.integers(1, 5)) if exact else 0)
    sess = fs.Session(1, int(rng.integers(1, 9)), cfg)
    want_ranges = O.batch_ranges(idx.numel(), cfg.batch_size, cfg.skip_nonfull_batch, exact, cfg.exact_num_batches)
    assert sess.num_total_batches == len(want_ranges)
    got = []
    while True:
        b = sess.blocking_get_batch()
        if b is None:
            break
        xb, yb, adjs, (st, en) = b
        got.append((st, en))
        on, oa = O.multilayer_sample(idx[st:en].numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal(adjs, oa), (case, sizes, (st, en))
        assert torch.equal(xb.cpu(), x[torch.from_numpy(on)])
        assert torch.equal(yb.cpu(), y[idx[st:en]].view(-1, 1))
    assert got == want_ranges


@pytest.mark.parametrize("case", range(16))
def test_randomized_distributed_sessions_against_oracle(fs, case):
    """Random partition counts / ranks / caches / x_gpu-x_cpu cut-offs: every ProtoDistributedBatch
    field against the oracle's restatement of fast_sampler.cpp:1017-1262, and x == X[n_id]."""
    rng = np.random.default_rng(9000 + case)
    n = int(rng.integers(40, 800))
    rowptr, col = _random_multigraph(rng, n, int(rng.integers(1, 40)))
    P = int(rng.choice([1, 2, 3, 4, 8, 16]))
    cuts = np.sort(rng.integers(0, n + 1, size=P - 1)) if P > 1 else np.empty(0, dtype=np.int64)
    off = torch.tensor([0] + cuts.tolist() + [n], dtype=torch.int64)          # ragged, possibly empty partitions
    sizes_p = (off[1:] - off[:-1]).tolist()
    rank = int(rng.choice([p for p in range(P) if sizes_p[p] > 0]))
    lo, hi = int(off[rank]), int(off[rank + 1])
    x = torch.from_numpy(rng.integers(0, 1000, size=(n, int(rng.choice([4, 50, 64]))))).to(torch.float16)
    y = torch.from_numpy(rng.integers(0, 9, size=(n, 1)))
    idx = torch.from_numpy(rng.integers(lo, hi, size=int(rng.integers(1, 200)))).to(torch.int64)
    use_cache = bool(rng.integers(0, 2)) and P > 1
    remote = np.setdiff1d(np.arange(n), np.arange(lo, hi))
    cv = torch.from_numpy(rng.permutation(remote)[:int(rng.integers(0, max(1, remote.size // 2) + 1))]).to(torch.int64)
    cut = int(rng.integers(0, hi - lo + 1))
    sizes = [int(rng.choice([2, 5, 15, 40])) for _ in range(int(rng.integers(1, 4)))]
    cfg = fs.Config()
    cfg.x_gpu, cfg.x_cpu, cfg.y = x[lo:lo + cut].contiguous(), x[lo + cut:hi].contiguous(), y
    cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
    cfg.batch_size, cfg.sizes, cfg.distributed, cfg.use_cache = int(rng.integers(1, 96)), sizes, True, use_cache
    cfg.partition_book = fs.RangePartitionBook(rank, P, off)
    cfg.cache = fs.Cache(rank, P, cv, x[cv].contiguous()) if use_cache else fs.Cache()
    cfg.partition_tables = [x[int(off[p]):int(off[p + 1])].contiguous() if p != rank and sizes_p[p] > 0 else None
                            for p in range(P)]
    sess = fs.Session(1, 4, cfg)
    oc = O.Cache(cv.numpy(), n) if use_cache else None
    for st, en in O.batch_ranges(idx.numel(), cfg.batch_size):
        b = sess.blocking_get_batch_distributed()
        assert tuple(b.idx_range) == (st, en)
        on, oa = O.multilayer_sample(idx[st:en].numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        pn, cn, perm, loc_cpu = O.distributed_binning(on, off.numpy(), rank, P, cut, use_cache, oc)
        assert adjs_equal(b.adjs, oa)
        for a, w in zip(b.partition_nids, pn):
            assert np.array_equal(a.cpu().numpy(), w)
        assert np.array_equal(b.cached_nids.cpu().numpy(), cn)
        assert np.array_equal(b.perm_partition_to_mfg.cpu().numpy(), perm)
        assert torch.equal(b.sliced_cpu_features.cpu(), cfg.x_cpu[torch.from_numpy(loc_cpu)])
        assert torch.equal(b.sliced_cpu_labels.cpu(), y[idx[st:en]].view(-1, 1))
        assert torch.equal(b.x.cpu(), x[torch.from_numpy(on)])
    assert sess.blocking_get_batch_distributed() is None


def test_session_api_behaviour(fs, data):
    """try_get_batch, abandoning a Session mid-epoch, slot reuse across Sessions, the slice API,
    wrong-mode calls and unsupported fan-outs failing loudly."""
    rowptr, col, x, y, N = data
    idx = S.seeds(N, 640)
    cfg = _config(fs, x, y, rowptr, col, idx)
    s1 = fs.Session(8, 2, cfg)
    first = s1.blocking_get_batch()
    assert first[3] == (0, 64) and s1.num_consumed_batches == 1 and s1.num_total_batches == 10
    assert s1.approx_num_complete_batches >= 1
    with pytest.raises(RuntimeError):
        s1.blocking_get_batch_distributed()
    del s1                                                    # dropped with batches still in flight
    s2 = fs.Session(8, 2, cfg)                                # picks the pooled slots up again
    seen = 0
    import time as _t
    deadline = _t.time() + 30
    while seen < 10 and _t.time() < deadline:                 # non-blocking polling like the reference's try_get_batch
        b = s2.try_get_batch()
        if b is not None:
            assert b[3] == (64 * seen, 64 * seen + 64)
            seen += 1
    assert seen == 10 and s2.try_get_batch() is None and s2.blocking_get_batch() is None
    # fan-out beyond SPP_MAX_FANOUT without replacement is refused, not silently truncated
    with pytest.raises(RuntimeError):
        fs.multilayer_sample(idx[:8], [200], rowptr, col)
    with pytest.raises(RuntimeError):
        fs.multilayer_sample(idx[:8], [5] * 9, rowptr, col)   # more than SPP_MAX_HOPS hops
    # full_sample is a documented stub
    with pytest.raises(RuntimeError):
        fs.full_sample()


def test_async_slice_tensors_contract(fs, data):
    """fast_sampler.cpp:720-758: per requester [rows of x_cpu for ids >= 0, positions of ids >= 0,
    positions of ids < 0]; own rank gets an empty first element."""
    rowptr, col, x, y, N = data
    P, rank = 2, 0
    off = S.equal_partition_offsets(N, P)
    hi = int(off[1])
    cut = hi // 2
    cfg = _config(fs, x[cut:hi].contiguous(), y, rowptr, col, S.seeds(N, 64, lo=0, hi=hi), distributed=True)
    cfg.x_gpu = x[:cut].contiguous()
    cfg.partition_book = fs.RangePartitionBook(rank, P, off)
    cfg.partition_tables = [None, x[hi:].contiguous()]
    sess = fs.Session(1, 2, cfg)
    req = [torch.tensor([5, -3, 0, -1, 7]), torch.tensor([-2, 4])]
    sess.async_slice_tensors(req, rank)
    sess.wait_slice_tensors()
    res = sess.get_slice_tensors()
    assert res[0][0].numel() == 0 and res[0][1].tolist() == [0, 2, 4] and res[0][2].tolist() == [1, 3]
    assert torch.equal(res[1][0].cpu(), cfg.x_cpu[torch.tensor([4])]) and res[1][1].tolist() == [1] and res[1][2].tolist() == [0]


def test_remote_frequency_statistics(fs, data):
    """cache_strategy == 'simulation' support (fast_sampler.cpp:835-880,1093-1103): how often each
    remote vertex was needed over an epoch, most frequent first."""
    rowptr, col, x, y, N = data
    P, rank = 4, 1
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    idx = S.seeds(N, 256, lo=lo, hi=hi)
    cfg = _config(fs, torch.empty((0, x.size(1)), dtype=x.dtype), y, rowptr, col, idx, distributed=True,
                  count_remote_frequency=True, use_cache=False, sizes=[10, 5])
    cfg.x_gpu = x[lo:hi].contiguous()
    cfg.partition_book = fs.RangePartitionBook(rank, P, off)
    cfg.partition_tables = [x[int(off[p]):int(off[p + 1])].contiguous() if p != rank else None for p in range(P)]
    sess = fs.Session(1, 4, cfg)
    want = np.zeros(N, dtype=np.int64)
    while True:
        b = sess.blocking_get_batch_distributed()
        if b is None:
            break
        st, en = b.idx_range
        on, _ = O.multilayer_sample(idx[st:en].numpy(), [10, 5], rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                    rng_seed=O.session_rng_seed(en))
        remote = on[(on < lo) | (on >= hi)]
        np.add.at(want, remote, 1)
    sess.reduce_multithreaded_frequency_counts()
    f, v = sess.remote_frequency_tensor.numpy(), sess.remote_vertices_ordered_by_freq.numpy()
    assert np.all(np.diff(f) <= 0) and f.sum() == want.sum() and len(v) == np.count_nonzero(want)
    assert np.array_equal(want[v], f)
    top = sess.get_n_most_freq_remote_vertices(10).numpy()
    assert np.array_equal(top, v[:10]) and want[top].min() >= np.sort(want)[-10]


@pytest.mark.parametrize("async_full", ["1", "0"])
def test_session_single_full_hop_async_and_stepwise(fs, data, monkeypatch, async_full):
    """Layer-wise inference batches (sizes [-1]): the pipelined path sized by the seeds' degree sum
    and the stepwise path must both reproduce the oracle, duplicates among the seeds included."""
    monkeypatch.setenv("SPP_ASYNC_FULL", async_full)
    rowptr, col, x, y, N = data
    base = S.seeds(N, 400)
    idx = torch.cat([base, base[:37], torch.arange(64)])      # duplicates + a run of consecutive ids
    cfg = _config(fs, x, y, rowptr, col, idx, sizes=[-1], batch_size=96)
    sess = fs.Session(4, 6, cfg)
    assert (sess._edge_bound is not None) == (async_full == "1")
    deg = (rowptr[1:] - rowptr[:-1])
    n = 0
    while True:
        b = sess.blocking_get_batch()
        if b is None:
            break
        xb, yb, adjs, (st, en) = b
        on, oa = O.multilayer_sample(idx[st:en].numpy(), [-1], rowptr.numpy(), col.numpy())
        assert adjs_equal(adjs, oa)
        assert adjs[0][1].numel() == int(deg[idx[st:en]].sum())
        assert torch.equal(xb.cpu(), x[torch.from_numpy(on)])
        assert torch.equal(yb.cpu(), y[idx[st:en]])
        n += 1
    assert n == sess.num_total_batches == (idx.numel() + 95) // 96


def test_session_distributed_single_full_hop(fs, data):
    rowptr, col, x, y, N = data
    P, rank = 4, 2
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    idx = torch.arange(lo, hi, dtype=torch.int64)[:300]
    cfg = _config(fs, torch.empty((0, x.size(1)), dtype=x.dtype), y, rowptr, col, idx, distributed=True, use_cache=False,
                  sizes=[-1])
    cfg.x_gpu = x[lo:hi].contiguous()
    cfg.partition_book = fs.RangePartitionBook(rank, P, off)
    cfg.cache = fs.Cache()
    cfg.partition_tables = [x[int(off[p]):int(off[p + 1])].contiguous() if p != rank else None for p in range(P)]
    sess = fs.Session(2, 8, cfg)
    assert sess._edge_bound is not None
    for want in O.batch_ranges(idx.numel(), 64, False, False, 0):
        b = sess.blocking_get_batch_distributed()
        assert tuple(b.idx_range) == want
        st, en = want
        on, oa = O.multilayer_sample(idx[st:en].numpy(), [-1], rowptr.numpy(), col.numpy())
        pn, cn, perm, _ = O.distributed_binning(on, off.numpy(), rank, P, hi - lo, False, None)
        assert adjs_equal(b.adjs, oa)
        for a, w in zip(b.partition_nids, pn):
            assert np.array_equal(a.cpu().numpy(), w)
        assert np.array_equal(b.perm_partition_to_mfg.cpu().numpy(), perm)
        assert np.array_equal(b.n_id.cpu().numpy(), on)
        assert torch.equal(b.x.cpu(), x[torch.from_numpy(on)])
    assert sess.blocking_get_batch_distributed() is None


def test_side_stream_fork_is_bit_exact():
    """SPP_FORK=3 (relabel/sort kernels and the feature gather forked onto side streams; read once
    per process by the library) must not change any result: the Session tests above are re-run in
    a child process with the switch on."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPP_FORK="3")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_session.py"), "-q", "-x",
                        "-m", "gpu", "-k", "test_session_nondistributed or test_session_distributed_single_process"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_device_side_epoch_shuffle_feeds_a_session_without_host_seeds(fs, data):
    """shufflers.py with device="cuda": the epoch permutation is drawn on the GPU, the Session reads
    the seeds of every batch in place from HBM (no H2D), and the batches are bit-exact against the
    oracle on those seeds; rank slices of the common permutation are disjoint and cover it."""
    from salient_plusplus_b200.shufflers import DistributedShuffler, Shuffler
    rowptr, col, x, y, N = data
    train = torch.arange(0, N, 3)
    sh = Shuffler(train, device="cuda")
    sh.set_epoch(3)
    a, b = sh.get_idx(), sh.get_idx()
    assert a.is_cuda and torch.equal(a, b) and torch.equal(torch.sort(a).values.cpu(), train)
    sh.set_epoch(4)
    assert not torch.equal(sh.get_idx(), a)
    ds = DistributedShuffler(train, 4, device="cuda")
    parts = [ds.get_idx(r) for r in range(4)]
    assert torch.equal(torch.sort(torch.cat(parts)).values.cpu(), train)
    n = train.numel()
    assert [p.numel() for p in parts] == [(n * (r + 1)) // 4 - (n * r) // 4 for r in range(4)]
    idx = a[:64 * 3 + 5]
    cfg = _config(fs, x, y, rowptr, col, idx)
    sess = fs.Session(2, 4, cfg)
    assert sess._idx is not None and sess._idx_host is None      # device-resident seeds, used in place
    idx_h = idx.cpu()
    for _ in range(4):
        xb, yb, adjs, (st, en) = sess.blocking_get_batch()
        on, oa = O.multilayer_sample(idx_h[st:en].numpy(), cfg.sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_COUNTER,
                                     rng_seed=O.session_rng_seed(en))
        assert adjs_equal(adjs, oa) and torch.equal(xb.cpu(), x[torch.from_numpy(on)]) and torch.equal(yb.cpu(), y[idx_h[st:en]])
    assert sess.blocking_get_batch() is None


def test_plain_launches_match_graph_replay():
    """SPP_GRAPH=0 (every kernel launched on the stream, read once per process by the library) and
    the default graph replay produce the same batches: the Session tests are re-run in a child
    process with replay switched off (the parent process ran them with replay on)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPP_GRAPH="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_session.py"), "-q", "-x",
                        "-m", "gpu", "-k", "test_session_nondistributed or test_session_distributed_single_process"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_native_host_path_is_default_and_matches_interpreter(fs, data, monkeypatch):
    """Sessions whose batches go through the executor run their per-batch bookkeeping in the native
    host path (csrc/host_session.cpp); SPP_NATIVE_HOST=0 selects the interpreter implementation.
    Same batches bit for bit (sampling is a pure function of the seeds and the per-batch key), one
    block per batch as the owner of every tensor on the native side."""
    rowptr, col, x, y, N = data
    idx = S.seeds(N, 300)

    def drain(native):
        if native:
            monkeypatch.delenv("SPP_NATIVE_HOST", raising=False)
        else:
            monkeypatch.setenv("SPP_NATIVE_HOST", "0")
        sess = fs.Session(4, 3, _config(fs, x, y, rowptr, col, idx))
        assert (sess._native is not None) == native
        out = []
        while True:
            b = sess.blocking_get_batch()
            if b is None:
                break
            if native:
                assert len(b.owners) == 1 and b.owners[0].dtype == torch.uint8
                assert b.y_flat is not None and b.y_flat.shape == b[1].squeeze().shape
                base = b.owners[0].untyped_storage().data_ptr()
                assert all(t.untyped_storage().data_ptr() == base for t in (b[0], b[1], b[2][0][0], b[2][0][1]))
            out.append((b[3], b[0].cpu(), b[1].cpu(), [(a[0].cpu(), a[1].cpu(), a[3]) for a in b[2]]))
        assert sess.num_consumed_batches == sess.num_total_batches
        return out

    nat, ref = drain(True), drain(False)
    assert len(nat) == len(ref) == 5
    for (r1, x1, y1, a1), (r2, x2, y2, a2) in zip(nat, ref):
        assert r1 == r2 and torch.equal(x1, x2) and torch.equal(y1, y2)
        for (p1, c1, s1), (p2, c2, s2) in zip(a1, a2):
            assert torch.equal(p1, p2) and torch.equal(c1, c2) and tuple(s1) == tuple(s2)


def test_tunables_reject_unknown_keys():
    from salient_plusplus_b200 import _lib
    with pytest.raises(_lib.SalientB200Error):
        _lib.tune("no_such_switch", 1)
    _lib.tune("gather_tile_rows", 0)
