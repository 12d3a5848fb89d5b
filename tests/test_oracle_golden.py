"""The oracle against golden vectors produced by the compiled, unmodified reference
(tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def samp():
    return np.load(os.path.join(G, "sampling.npz"))


@pytest.fixture(scope="module")
def dist():
    return np.load(os.path.join(G, "distributed.npz"))


def golden_adjs(d, prefix):
    out = []
    for i in range(int(d[f"{prefix}_n"])):
        assert int(d[f"{prefix}_{i}_eid_len"]) == 0
        out.append((d[f"{prefix}_{i}_rowptr"], d[f"{prefix}_{i}_col"], tuple(d[f"{prefix}_{i}_size"].tolist())))
    return out


def same_adjs(oa, ga):
    return len(oa) == len(ga) and all(
        np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and tuple(a[3]) == b[2] for a, b in zip(oa, ga))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_full_neighbourhood(samp, L):
    n_id, adjs = O.multilayer_sample(samp["idx"][:40], [-1] * L, samp["rowptr"], samp["col"])
    assert np.array_equal(n_id, samp[f"full{L}_n_id"])
    assert same_adjs(adjs, golden_adjs(samp, f"full{L}"))


def test_sample_adj(samp):
    rp, cl, n_id, e_id = O.sample_adj(samp["rowptr"], samp["col"], samp["idx"][:40], -1, False)
    assert n_id.dtype == np.int32 and e_id.size == 0
    assert np.array_equal(rp, samp["sa_rowptr"]) and np.array_equal(cl, samp["sa_col"])
    assert np.array_equal(n_id, samp["sa_n_id"])


def test_session_stochastic_stream(samp):
    """Same std::mt19937 stream, same (biased) Floyd variant, same per-batch seed."""
    idx, x, y = samp["idx"], samp["x"], samp["y"]
    ranges = O.batch_ranges(idx.size, 32)
    assert len(ranges) == int(samp["sess_num_batches"])
    for st, en in ranges:
        assert int(samp[f"sess_{st}_stop"]) == en
        n_id, adjs = O.multilayer_sample(idx[st:en], [15, 10, 5], samp["rowptr"], samp["col"],
                                         rng_mode=O.RNG_REFERENCE, rng_seed=O.session_rng_seed(en))
        assert same_adjs(adjs, golden_adjs(samp, f"sess_{st}"))
        assert np.array_equal(O.serial_index(x, n_id), samp[f"sess_{st}_x"])
        assert np.array_equal(O.serial_index(y, n_id, en - st), samp[f"sess_{st}_y"])


def test_exact_num_batches_ranges(samp):
    got = O.batch_ranges(samp["idx"].size, 32, False, True, 7)
    assert np.array_equal(np.array(got), samp["exact7_ranges"])


def test_serial_index(samp):
    assert np.array_equal(O.serial_index(samp["x"], samp["si_idx"]), samp["si_out"])
    assert np.array_equal(O.serial_index(samp["x"], samp["si_idx"], 4), samp["si_out_n4"])


def test_partition_book_and_cache(dist):
    off, rank, probe = dist["offsets"], int(dist["rank"]), dist["probe"]
    assert np.array_equal(O.nid2partid(off, probe), dist["partid"])
    assert np.array_equal(O.nid2localnid(off, probe, rank), dist["localnid"])
    assert np.array_equal(O.partid2nids(off, 2), dist["partid2nids"])
    c = O.Cache(dist["cached_vertices"], int(off[-1]))
    assert np.array_equal(c.nid_is_cached(probe), dist["is_cached"])
    assert np.array_equal(c.nid2cachenid(dist["cached_vertices"][:20]), dist["cachenid"])


@pytest.mark.parametrize("tag", ["nc", "c"])
def test_distributed_binning(samp, dist, tag):
    off, rank, cut, lidx = dist["offsets"], int(dist["rank"]), int(dist["cut"]), dist["lidx"]
    P = off.size - 1
    use_cache = tag == "c"
    cache = O.Cache(dist["cached_vertices"], int(off[-1])) if use_cache else None
    x_local = samp["x"][off[rank]:off[rank + 1]]
    ranges = O.batch_ranges(lidx.size, 32, False, True, 3)
    assert int(dist[f"{tag}_num_batches"]) == 3
    for k, (st, en) in enumerate(ranges):
        p = f"{tag}{k}"
        assert tuple(dist[f"{p}_range"].tolist()) == (st, en)
        n_id, adjs = O.multilayer_sample(lidx[st:en], [15, 10, 5], samp["rowptr"], samp["col"],
                                         rng_mode=O.RNG_REFERENCE, rng_seed=O.session_rng_seed(en))
        assert same_adjs(adjs, golden_adjs(dist, p))
        pn, cn, perm, loc_cpu = O.distributed_binning(n_id, off, rank, P, cut, use_cache, cache)
        for q in range(P):
            assert np.array_equal(pn[q], dist[f"{p}_part{q}"])
        assert np.array_equal(cn, dist[f"{p}_cached_nids"])
        assert np.array_equal(perm, dist[f"{p}_perm"])
        assert np.array_equal(x_local[cut:][loc_cpu], dist[f"{p}_cpu_feats"])
        assert np.array_equal(samp["y"][lidx[st:en]], dist[f"{p}_labels"])


def test_mt19937_known_answer():
    """std::mt19937 default-seeded: the 10000th output is 4123659995 (C++ standard, [rand.predef])."""
    import ctypes
    L = O.lib()
    buf = ctypes.create_string_buffer(624 * 4 + 8)
    L.spo_mt_seed(buf, 5489)
    v = 0
    for _ in range(10000):
        v = L.spo_mt_next(buf)
    assert v == 4123659995


@pytest.mark.parametrize("fi", [0, 1])
def test_vip_first_order_form_against_the_reference_driver(fi):
    """Golden vectors produced by executing the reference driver's own get_frequency_tensors_fast
    (driver/drivers/ddp.py:134-239, fp64; tests/golden/make_golden_vip_driver.py)."""
    g = np.load(os.path.join(G, "vip_driver.npz"))
    fanouts = g[f"fanouts{fi}"].tolist()
    for p in range(4):
        got = O.vip_probabilities(g["rowptr"], g["col"], g[f"train{p}"], 32, fanouts, exact=False)
        assert np.max(np.abs(got - g[f"vip{fi}_{p}"])) < 1e-14


@pytest.mark.parametrize("fi", [0, 1])
def test_vip_analytical_against_reference_module(fi):
    """Golden vectors produced by the reference's own caching/vip.py:vip_analytical (fp32)."""
    g = np.load(os.path.join(G, "vip.npz"))
    fanouts = g[f"fanouts{fi}"].tolist()
    for p in range(4):
        got = O.vip_probabilities(g["rowptr"], g["col"], g[f"train{p}"], 32, fanouts, exact=True)
        want = g[f"vip{fi}_{p}"]
        assert want.dtype == np.float32
        assert np.max(np.abs(got - want.astype(np.float64))) < 2e-5     # fp32 reference vs fp64 oracle
        assert np.argmax(got) == np.argmax(want) or abs(got.max() - want.max()) < 2e-5


def test_select_cache_vertices_layout():
    g = np.load(os.path.join(G, "vip.npz"))
    vip = O.vip_probabilities(g["rowptr"], g["col"], g["train1"], 32, [15, 10, 5])
    off = g["offsets"]
    cv = O.select_cache_vertices(vip, off, 1, 50)
    assert cv.size == 50 and not np.any((cv >= off[1]) & (cv < off[2]))
    owner = O.nid2partid(off, cv)
    assert np.all(np.diff(owner) >= 0)
    for p in range(4):
        v = vip[cv[owner == p]]
        assert np.all(np.diff(v) <= 0)
    # the 50 selected are the 50 largest remote VIP values
    remote = np.ones(vip.size, bool)
    remote[off[1]:off[2]] = False
    assert np.isclose(np.sort(vip[cv])[0], np.sort(vip[remote])[-50])
