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
