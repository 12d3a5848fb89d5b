"""The CUDA path against the committed golden vectors of the compiled reference
(tests/golden/*.npz): every deterministic output must be bit-identical to the reference's."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fs():
    from salient_plusplus_b200 import fast_sampler
    return fast_sampler


@pytest.fixture(scope="module")
def samp():
    return np.load(os.path.join(G, "sampling.npz"))


@pytest.fixture(scope="module")
def dist():
    return np.load(os.path.join(G, "distributed.npz"))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_full_neighbourhood_equals_reference(fs, samp, L):
    n_id, adjs = fs.multilayer_sample(t(samp["idx"][:40]), [-1] * L, t(samp["rowptr"]), t(samp["col"]))
    assert np.array_equal(n_id.cpu().numpy(), samp[f"full{L}_n_id"])
    assert len(adjs) == int(samp[f"full{L}_n"])
    for i, a in enumerate(adjs):
        assert np.array_equal(a[0].cpu().numpy(), samp[f"full{L}_{i}_rowptr"])
        assert np.array_equal(a[1].cpu().numpy(), samp[f"full{L}_{i}_col"])
        assert tuple(int(v) for v in a[3]) == tuple(samp[f"full{L}_{i}_size"].tolist())
        assert a[2].numel() == 0


def test_sample_adj_equals_reference(fs, samp):
    rp, cl, n_id, e_id = fs.sample_adj(t(samp["rowptr"]), t(samp["col"]), t(samp["idx"][:40]), -1, False)
    assert n_id.dtype == torch.int32 and e_id.numel() == 0
    assert np.array_equal(rp.cpu().numpy(), samp["sa_rowptr"]) and np.array_equal(cl.cpu().numpy(), samp["sa_col"])
    assert np.array_equal(n_id.cpu().numpy(), samp["sa_n_id"])


def test_serial_index_equals_reference(fs, samp):
    x = t(samp["x"]).view(torch.float16)
    assert np.array_equal(fs.serial_index(x, t(samp["si_idx"])).cpu().view(torch.int16).numpy(), samp["si_out"])
    assert np.array_equal(fs.serial_index(x, t(samp["si_idx"]), 4).cpu().view(torch.int16).numpy(), samp["si_out_n4"])


def test_session_ranges_equal_reference(fs, samp):
    cfg = fs.Config()
    cfg.x_cpu, cfg.y = t(samp["x"]).view(torch.float16), t(samp["y"])
    cfg.rowptr, cfg.col, cfg.idx = t(samp["rowptr"]), t(samp["col"]), t(samp["idx"])
    cfg.batch_size, cfg.sizes = 32, [2]
    cfg.force_exact_num_batches, cfg.exact_num_batches = True, 7
    s = fs.Session(1, 10, cfg)
    got = []
    while True:
        b = s.blocking_get_batch()
        if b is None:
            break
        got.append(b[3])
    assert np.array_equal(np.array(sorted(got)), samp["exact7_ranges"])


def test_partition_book_and_cache_equal_reference(fs, dist):
    off, rank, probe = t(dist["offsets"]), int(dist["rank"]), t(dist["probe"])
    book = fs.RangePartitionBook(rank, off.numel() - 1, off)
    assert np.array_equal(book.nid2partid(probe).numpy(), dist["partid"])
    assert np.array_equal(book.nid2localnid(probe, rank).numpy(), dist["localnid"])
    assert np.array_equal(book.partid2nids(2).numpy(), dist["partid2nids"])
    cv = t(dist["cached_vertices"])
    c = fs.Cache(rank, off.numel() - 1, cv, torch.zeros(cv.numel(), 4).half())
    assert np.array_equal(c.nid_is_cached(probe).numpy(), dist["is_cached"])
    assert np.array_equal(c.nid2cachenid(cv[:20]).numpy(), dist["cachenid"])


@pytest.mark.parametrize("tag", ["nc", "c"])
def test_distributed_binning_equals_reference(fs, samp, dist, tag):
    """The reference's ProtoDistributedBatch fields for the n_id the reference itself sampled
    (binning is deterministic given n_id): split kernel through the C ABI."""
    import ctypes
    from salient_plusplus_b200 import _lib
    from salient_plusplus_b200.fast_sampler import make_feature_map
    from oracle import oracle as O
    L = _lib.load()
    off, rank = dist["offsets"], int(dist["rank"])
    P = off.size - 1
    use_cache = tag == "c"
    cv = t(dist["cached_vertices"])
    cache = fs.Cache(rank, P, cv, torch.zeros(cv.numel(), 4).half())
    cmap = cache.device_index(int(off[-1]))
    lidx = dist["lidx"]
    for k in range(int(dist[f"{tag}_num_batches"])):
        p = f"{tag}{k}"
        st, en = dist[f"{p}_range"].tolist()
        # n_id exactly as the reference sampled it (mt19937 stream) -- reproduced by the pinned oracle
        n_id, _ = O.multilayer_sample(lidx[st:en], [15, 10, 5], samp["rowptr"], samp["col"], rng_mode=O.RNG_REFERENCE,
                                      rng_seed=O.session_rng_seed(en))
        ids_dev = t(n_id).cuda()
        n = ids_dev.numel()
        fm = make_feature_map(off.tolist(), rank, [None] * P)
        if use_cache:
            fm.cache_index, fm.cache_index_nodes = cmap[0].data_ptr(), cmap[1]
        scratch = torch.empty(int(L.spp_split_scratch_words(n)), dtype=torch.int32, device="cuda")
        ids = torch.empty(n, dtype=torch.int64, device="cuda")
        perm = torch.empty(n, dtype=torch.int64, device="cuda")
        counts = torch.zeros(P + 2, dtype=torch.int64, device="cuda")
        _lib.check(L.spp_split_by_owner(ctypes.byref(fm), int(use_cache), ids_dev.data_ptr(), 1, n, None, ids.data_ptr(),
                                        perm.data_ptr(), counts.data_ptr(), scratch.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        c = counts.cpu().tolist()
        pos = 0
        for q in range(P):
            assert np.array_equal(ids[pos:pos + c[q]].cpu().numpy(), dist[f"{p}_part{q}"])
            pos += c[q]
        assert np.array_equal(ids[pos:pos + c[P]].cpu().numpy(), dist[f"{p}_cached_nids"])
        assert np.array_equal(perm.cpu().numpy(), dist[f"{p}_perm"])
