"""Pins the oracle against the compiled, unmodified reference on larger seeded inputs than the
golden fixtures hold.  Needs oracle/_ref (built by oracle/build_ref.sh; it travels to the GPU
box with the snapshot); skipped when absent."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref
from salient_plusplus_b200 import synthetic as S

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def R():
    return ref.load_reference()


@pytest.fixture(scope="module")
def ds():
    return S.make_dataset("arxiv", scale=0.05)


def same(adjs, oa):
    return all(np.array_equal(a[0].numpy(), b[0]) and np.array_equal(a[1].numpy(), b[1]) and tuple(a[3]) == b[3]
               and a[2].numel() == 0 for a, b in zip(adjs, oa)) and len(adjs) == len(oa)


@pytest.mark.parametrize("sizes", [[-1], [-1, -1]])
def test_multilayer_full(R, ds, sizes):
    idx = S.seeds(ds.num_nodes, 300)
    n_id, adjs = R.multilayer_sample(idx, sizes, ds.rowptr, ds.col)
    on, oa = O.multilayer_sample(idx.numpy(), sizes, ds.rowptr.numpy(), ds.col.numpy())
    assert np.array_equal(n_id.numpy(), on) and same(adjs, oa)


@pytest.mark.parametrize("sizes", [[15, 10, 5], [25, 15], [3]])
def test_session_stochastic(R, ds, sizes):
    idx = S.seeds(ds.num_nodes, 300)
    cfg = R.Config()
    cfg.x_cpu, cfg.x_gpu, cfg.y = ds.x, torch.empty(0), ds.y
    cfg.rowptr, cfg.col, cfg.idx = ds.rowptr, ds.col, idx
    cfg.batch_size, cfg.sizes = 64, sizes
    cfg.skip_nonfull_batch = cfg.pin_memory = cfg.distributed = False
    cfg.force_exact_num_batches, cfg.exact_num_batches = False, 0
    cfg.count_remote_frequency = cfg.use_cache = False
    s = R.Session(2, 10, cfg)
    seen = 0
    while True:
        b = s.blocking_get_batch()
        if b is None:
            break
        x, y, adjs, (st, en) = b
        on, oa = O.multilayer_sample(idx[st:en].numpy(), sizes, ds.rowptr.numpy(), ds.col.numpy(),
                                     rng_mode=O.RNG_REFERENCE, rng_seed=O.session_rng_seed(en))
        assert same(adjs, oa)
        assert np.array_equal(x.numpy(), ds.x.numpy()[on])
        assert np.array_equal(y.numpy(), ds.y.numpy()[idx[st:en]])
        seen += 1
    assert seen == len(O.batch_ranges(300, 64))


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.int64, torch.uint8])
def test_serial_index(R, dtype):
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 200, (500, 9), generator=g).to(dtype)
    idx = torch.randint(0, 500, (333,), generator=g)
    want = R.serial_index(x, idx)
    xn = x.view(torch.int16).numpy() if dtype == torch.float16 else x.numpy()
    wn = want.view(torch.int16).numpy() if dtype == torch.float16 else want.numpy()
    assert np.array_equal(O.serial_index(xn, idx.numpy()), wn)


def test_partition_book(R):
    N = 12345
    for P in (1, 2, 8):
        off = S.equal_partition_offsets(N, P)
        book = R.RangePartitionBook(P - 1, P, off)
        nids = torch.cat([off.clamp(max=N - 1), S.seeds(N, 500)])
        assert np.array_equal(book.nid2partid(nids).numpy(), O.nid2partid(off.numpy(), nids.numpy()))
        assert np.array_equal(book.nid2localnid(nids, P - 1).numpy(), O.nid2localnid(off.numpy(), nids.numpy(), P - 1))
        assert np.array_equal(book.partid2nids(0).numpy(), O.partid2nids(off.numpy(), 0))


def _random_graph(rng, n, max_deg):
    rowptr = [0]
    col = []
    for _ in range(n):
        d = int(rng.integers(0, max_deg + 1))
        col.extend(rng.integers(0, n, size=d).tolist())   # unsorted, duplicates and self loops allowed
        rowptr.append(len(col))
    return torch.tensor(rowptr, dtype=torch.int64), torch.tensor(col, dtype=torch.int64)


@pytest.mark.parametrize("case", range(24))
def test_randomized_sessions_against_reference(R, case):
    """Seeded sweep: tiny to medium random multigraphs (isolated nodes, self loops, duplicate
    neighbours), duplicate seeds, every fan-out regime (full, fan-out >= degree, Floyd), both batch
    range modes -- oracle == compiled reference, field by field."""
    rng = np.random.default_rng(1000 + case)
    n = int(rng.integers(5, 400))
    rowptr, col = _random_graph(rng, n, int(rng.integers(0, 30)))
    L = int(rng.integers(1, 4))
    sizes = [int(rng.choice([-1, 1, 2, 5, 15, 40])) for _ in range(L)]
    idx = torch.from_numpy(rng.integers(0, n, size=int(rng.integers(1, 150)))).to(torch.int64)  # duplicates likely
    x = torch.from_numpy(rng.integers(0, 1000, size=(n, 3))).to(torch.float16)
    y = torch.from_numpy(rng.integers(0, 9, size=(n, 1)))
    cfg = R.Config()
    cfg.x_cpu, cfg.x_gpu, cfg.y = x, torch.empty(0), y
    cfg.rowptr, cfg.col, cfg.idx = rowptr, col, idx
    cfg.batch_size, cfg.sizes = int(rng.integers(1, 64)), sizes
    cfg.skip_nonfull_batch = bool(rng.integers(0, 2))
    cfg.pin_memory = cfg.distributed = False
    exact = bool(rng.integers(0, 2)) and idx.numel() >= 8
    cfg.force_exact_num_batches, cfg.exact_num_batches = exact, (int(rng.integers(1, 5)) if exact else 0)
    cfg.count_remote_frequency = cfg.use_cache = False
    s = R.Session(2, 8, cfg)
    want_ranges = O.batch_ranges(idx.numel(), cfg.batch_size, cfg.skip_nonfull_batch, exact, cfg.exact_num_batches)
    assert s.num_total_batches == len(want_ranges)
    got = []
    while True:
        b = s.blocking_get_batch()
        if b is None:
            break
        xb, yb, adjs, (st, en) = b
        got.append((st, en))
        on, oa = O.multilayer_sample(idx[st:en].numpy(), sizes, rowptr.numpy(), col.numpy(), rng_mode=O.RNG_REFERENCE,
                                     rng_seed=O.session_rng_seed(en))
        assert same(adjs, oa), (case, sizes)
        assert np.array_equal(xb.view(torch.int16).numpy(), x.view(torch.int16).numpy()[on])
        assert np.array_equal(yb.numpy(), y.numpy()[idx[st:en].numpy()])
    assert sorted(got) == want_ranges
