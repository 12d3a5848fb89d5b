"""CPU tests of the native per-batch host path (csrc/host_session.cpp, module ``_spp_host``): output
allocation, job-descriptor fill, in-order delivery, view cutting, blocked statistics and error
behaviour, driven through a recording stand-in for ``spp_executor_*`` (tests/fake_executor.c) whose
"device" is host memory.  What the reference does at this level: Session's queueing and delivery
order, fast_sampler/fast_sampler.cpp:587-627 (batch ranges), :672-712 (in idx_range order),
:777-828 (try / blocking getters), :994 (per-batch RNG seed)."""
import ctypes
import os
import subprocess
import types

import pytest
import torch

from salient_plusplus_b200 import _lib
from salient_plusplus_b200._lib import SPP_MAX_PARTS, SPP_META_WORDS, BatchJob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fx(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fx") / "libfake_executor.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "fake_executor.c"), "-o", out])
    L = ctypes.CDLL(out)
    L.fx_create.restype = ctypes.c_void_p
    L.fx_create.argtypes = [ctypes.c_int]
    L.fx_destroy.argtypes = [ctypes.c_void_p]
    L.fx_submit.restype = ctypes.c_uint64
    L.fx_last_error.restype = ctypes.c_char_p
    L.fx_last_job.restype = ctypes.POINTER(BatchJob)
    for name in ("fx_fail_submit_at", "fx_overflow_at"):
        getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_int]
    for name in ("fx_submitted", "fx_waits"):
        getattr(L, name).restype = ctypes.c_int64
        getattr(L, name).argtypes = [ctypes.c_void_p]
    L.fx_last_job.argtypes = [ctypes.c_void_p]
    return L


@pytest.fixture(scope="module")
def host():
    _lib.build_host_extension()
    return _lib.load_host()


def expected_sizes(bs, caps):
    T, nodes, edges = bs, [bs], []
    for h, cap in enumerate(caps):
        E = min(2 * T + h, cap)
        edges.append(E)
        T += E // 2
        nodes.append(T)
    return nodes, edges


def make(fx, host, *, n_seeds=70, bs=16, L=2, P=-1, depth=3, polls=0, y_dtype=torch.int64, feat=True, idx_on_device=False,
         batch_edges=None, skip_last=False):
    """A HostSession over CPU tensors, laid out exactly like Session._layout does it."""
    from salient_plusplus_b200 import fast_sampler as fs
    T_b, E_b, t = [], [], bs
    for h in range(L):
        T_b.append(t)
        E_b.append(2 * t + h)
        t += E_b[-1] // 2
    node_bound = t
    off, o = [], 0
    for h in range(L):
        off.append((o, o + T_b[h] + 1))
        o += T_b[h] + 1 + E_b[h]
    nid = o
    if P >= 0:
        o += 3 * node_bound
    yo = o
    y_in_arena = y_dtype == torch.int64
    if y_in_arena:
        o += bs
    layout = (off, nid, yo, max(o, 1), node_bound)
    idx = torch.arange(100, 100 + n_seeds, dtype=torch.int64)
    ranges = [(i, min(i + bs, n_seeds)) for i in range(0, n_seeds, bs)]
    if skip_last and ranges[-1][1] - ranges[-1][0] < bs:
        ranges.pop()
    y_table = torch.zeros((4, 1), dtype=y_dtype)
    ex = fx.fx_create(polls)
    slots = []
    for _ in range(depth):
        s = types.SimpleNamespace()
        s.cjob = BatchJob()
        s.meta_host = torch.zeros(SPP_META_WORDS + SPP_MAX_PARTS + 2, dtype=torch.int64)
        s.seeds = torch.zeros(bs, dtype=torch.int64)
        s.stream_raw = (0, 0, 0)
        j = s.cjob
        j.n_hops = L
        for h in range(L):
            j.out_col_cap[h] = E_b[h]
        j.row_bytes = 8 * 2   # 8 fp16 features
        j.feature_mode = 1 if feat else 0
        j.meta_host = s.meta_host.data_ptr()
        j.y_row_bytes = y_table.element_size()
        if P >= 0:
            j.do_split = 1
            j.fmap.num_parts = P
        slots.append(s)
    spec = fs.host_session_spec(
        device=torch.device("cpu"), n_hops=L, num_parts=P, layout=layout, has_x=feat, feat_dim=8, feat_dtype=torch.float16,
        y=y_table, y_in_arena=y_in_arena, ranges=ranges, batch_edges=batch_edges,
        idx_host_ptr=0 if idx_on_device else idx.data_ptr(), idx_dev_ptr=idx.data_ptr() if idx_on_device else 0,
        executor=ex, entry_points=(fx.fx_submit, fx.fx_poll, fx.fx_wait, fx.fx_last_error), slots=slots,
        e_id=torch.empty(0, dtype=torch.int64))
    hs = host.HostSession(spec)
    keep = (idx, slots, y_table)
    return hs, ex, ranges, idx, [E_b, keep]


def inside(owners, *tensors):
    """Every tensor lies inside the batch's single block (the only thing record_stream must touch)."""
    assert len(owners) == 1 and owners[0].dtype == torch.uint8 and owners[0].dim() == 1
    lo, hi = owners[0].data_ptr(), owners[0].data_ptr() + owners[0].numel()
    for t in tensors:
        assert t._base is None and t.untyped_storage().data_ptr() == owners[0].untyped_storage().data_ptr()
        if t.numel():
            assert lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= hi
    return True


def check_structure(adjs, seed0, bs, caps):
    nodes, edges = expected_sizes(bs, caps)
    L = len(caps)
    assert len(adjs) == L
    for k, (rp, cl, e_id, size) in enumerate(adjs):     # outermost hop first (fast_sampler.cpp:224)
        h = L - 1 - k
        assert size == (nodes[h], nodes[h + 1])
        assert rp.tolist() == [1000 * (h + 1) + i for i in range(nodes[h] + 1)]
        assert cl.tolist() == [seed0 + 10 * (h + 1) + e for e in range(edges[h])]
        assert e_id.numel() == 0 and e_id.dtype == torch.int64
    return nodes[-1]


def test_module_matches_header(host):
    assert host.ABI_VERSION == _lib.ABI_VERSION
    assert host.BATCH_JOB_BYTES == ctypes.sizeof(BatchJob)


@pytest.mark.parametrize("polls", [0, 2])
def test_nondistributed_batches_in_order(fx, host, polls):
    hs, ex, ranges, idx, (caps, _keep) = make(fx, host, polls=polls)
    hs.fill()
    assert hs.total == len(ranges) == 5 and hs.issued == 3 and hs.in_flight == 3   # depth 3
    got = 0
    while True:
        r = hs.get(True)
        if r is None:
            break
        x, y, adjs, rng, owners, y_flat = r
        st, en = ranges[got]
        assert rng == (st, en)
        seed0 = int(idx[st])
        nb = check_structure(adjs, seed0, en - st, caps)
        assert x.shape == (nb, 8) and x.dtype == torch.float16
        assert x.view(torch.uint8)[:, 0].tolist() == [(seed0 + i) & 0xff for i in range(nb)]
        assert y.shape == (en - st, 1) and y.view(-1).tolist() == (3 * idx[st:en]).tolist()
        assert y_flat.shape == y.squeeze().shape and torch.equal(y_flat, y.squeeze()) and y_flat.data_ptr() == y.data_ptr()
        assert inside(owners, x, y, y_flat, *[t for a in adjs for t in a[:2]])   # one block owns every tensor of the batch
        got += 1
        assert hs.consumed == got and hs.issued == min(got + 3, 5)
    assert got == 5 and hs.get(True) is None and hs.get(False) is None
    assert fx.fx_submitted(ex) == 5
    # a batch that was not ready when asked for is waited for and counted (fast_sampler.cpp:789-799)
    assert hs.blocked_occasions == (5 if polls else 0) and fx.fx_waits(ex) == hs.blocked_occasions
    assert hs.blocked_us >= 0
    fx.fx_destroy(ex)


def test_try_get_returns_none_until_ready(fx, host):
    hs, ex, ranges, idx, keep = make(fx, host, polls=2, n_seeds=32)
    hs.fill()
    assert hs.get(False) is None and hs.get(False) is None     # two polls say "not yet"
    assert hs.complete_count() == 1                            # the third poll completes the oldest batch
    r = hs.get(False)
    assert r is not None and r[3] == ranges[0]
    assert hs.blocked_occasions == 0
    assert hs.get(True)[3] == ranges[1]
    assert hs.blocked_occasions == 1
    fx.fx_destroy(ex)


def test_job_fields_per_batch(fx, host):
    """Pointers and per-batch scalars written into the slot's spp_batch_job."""
    hs, ex, ranges, idx, (caps, keep) = make(fx, host, depth=1, n_seeds=40, P=-1)
    slots = keep[1]
    hs.fill()
    j = fx.fx_last_job(ex).contents
    assert j.batch_size == 16 and j.rng_seed == (16 * 17 + 5)                  # stop * 17 + 5, fast_sampler.cpp:994
    assert j.seeds_host == idx.data_ptr() and j.seeds_dev == slots[0].seeds.data_ptr()
    assert slots[0].seeds.tolist() == idx[:16].tolist()                       # the stand-in's "H2D copy"
    assert j.n_id_out is None and j.bucket_ids is None                         # not distributed
    assert j.out_col[0] - j.out_rowptr[0] == 8 * (16 + 1)
    hs.get(True)
    j = fx.fx_last_job(ex).contents
    assert j.batch_size == 16 and j.seeds_host == idx.data_ptr() + 8 * 16 and j.rng_seed == 32 * 17 + 5
    hs.get(True)
    j = fx.fx_last_job(ex).contents
    assert j.batch_size == 8 and j.rng_seed == 40 * 17 + 5                     # ragged last batch
    x, y, adjs, rng, _owners, _yf = hs.get(True)
    assert rng == (32, 40) and y.shape == (8, 1) and adjs[-1][3][0] == 8
    fx.fx_destroy(ex)


def test_device_resident_seeds_and_separate_labels(fx, host):
    hs, ex, ranges, idx, (caps, keep) = make(fx, host, idx_on_device=True, y_dtype=torch.int32, n_seeds=32, depth=2)
    hs.fill()
    j = fx.fx_last_job(ex).contents
    assert j.seeds_host is None and j.seeds_dev == idx.data_ptr() + 8 * 16     # used in place
    x, y, adjs, rng, owners, y_flat = hs.get(True)
    assert y.dtype == torch.int32 and y.shape == (16, 1) and y_flat.shape == (16,) and y_flat.data_ptr() == y.data_ptr()
    assert inside(owners, x, y) and y.data_ptr() > x.data_ptr()               # labels that are not int64 sit behind x
    assert y.data_ptr() % 256 == x.data_ptr() % 256 == owners[0].data_ptr() % 256
    fx.fx_destroy(ex)


def test_no_feature_table(fx, host):
    hs, ex, ranges, idx, keep = make(fx, host, feat=False, n_seeds=16)
    hs.fill()
    x, y, adjs, rng, owners, y_flat = hs.get(True)
    assert x.shape == (0, 8) and x.dtype == torch.float16
    assert fx.fx_last_job(ex).contents.x_out is None
    fx.fx_destroy(ex)


@pytest.mark.parametrize("P", [1, 3, 8])
def test_distributed_pieces(fx, host, P):
    hs, ex, ranges, idx, (caps, _keep) = make(fx, host, P=P, n_seeds=48, L=3)
    hs.fill()
    for k in range(3):
        n_id, parts, cached, perm, adjs, rng, y, x, owners, y_flat = hs.get(True)
        st, en = ranges[k]
        seed0 = int(idx[st])
        nb = check_structure(adjs, seed0, en - st, caps)
        assert rng == (st, en)
        assert n_id.tolist() == [seed0 + i for i in range(nb)]
        assert len(parts) == P and all(p.numel() == nb // (P + 1) for p in parts)
        assert cached.numel() == nb - P * (nb // (P + 1))
        assert torch.cat(list(parts) + [cached]).tolist() == [7 * pos + seed0 for pos in range(nb)]
        assert perm.tolist() == [nb - 1 - i for i in range(nb)]
        assert x.shape == (nb, 8) and y.view(-1).tolist() == (3 * idx[st:en]).tolist()
        assert inside(owners, n_id, cached, perm, y, x, *parts)
    assert hs.get(True) is None
    fx.fx_destroy(ex)


def test_layerwise_capacity_per_batch(fx, host):
    """Layer-wise batches carry their exact edge count as the capacity of hop 0 (Session._batch_edges)."""
    hs, ex, ranges, idx, keep = make(fx, host, L=1, n_seeds=48, batch_edges=[5, 0, 31], depth=1)
    hs.fill()
    for cap in (5, 0, 31):
        assert fx.fx_last_job(ex).contents.out_col_cap[0] == cap
        x, y, adjs, rng, _owners, _yf = hs.get(True)
        assert adjs[0][1].numel() == cap and adjs[0][3] == (16, 16 + cap // 2)
    fx.fx_destroy(ex)


def test_errors_surface_as_library_errors(fx, host):
    hs, ex, *keep = make(fx, host, depth=2)
    fx.fx_fail_submit_at(ex, 2)
    with pytest.raises(_lib.SalientB200Error, match="injected failure"):
        hs.fill()
    fx.fx_destroy(ex)
    hs, ex, *keep2 = make(fx, host, depth=2)
    fx.fx_overflow_at(ex, 1)
    hs.fill()
    with pytest.raises(_lib.SalientB200Error, match="SPP_META_OVERFLOW"):
        hs.get(True)
    fx.fx_destroy(ex)


def test_release_waits_for_abandoned_work(fx, host):
    hs, ex, ranges, idx, keep = make(fx, host, polls=5, depth=3)
    hs.fill()
    assert hs.in_flight == 3
    hs.get(True)
    assert hs.in_flight == 3           # the freed slot took the next batch
    hs.release()
    assert hs.in_flight == 0 and fx.fx_waits(ex) == 1 + 3
    assert hs.get(True) is None        # nothing more is delivered or enqueued
    hs.fill()
    assert fx.fx_submitted(ex) == 4
    fx.fx_destroy(ex)


def test_spec_validation(fx, host):
    from salient_plusplus_b200 import fast_sampler as fs
    with pytest.raises((ValueError, RuntimeError, KeyError)):
        host.HostSession({})
    s = types.SimpleNamespace(cjob=BatchJob(), meta_host=torch.zeros(50, dtype=torch.int64),
                              seeds=torch.zeros(4, dtype=torch.int64), stream_raw=(0, 0, 0))
    base = dict(device=torch.device("cpu"), n_hops=1, num_parts=-1, layout=([(0, 5)], 20, 20, 24, 12), has_x=False,
                feat_dim=4, feat_dtype=torch.float32, y=None, y_in_arena=False, ranges=[(0, 4)], batch_edges=None,
                idx_host_ptr=0, idx_dev_ptr=0, executor=1, entry_points=(fx.fx_submit, fx.fx_poll, fx.fx_wait, fx.fx_last_error),
                slots=[s], e_id=torch.empty(0, dtype=torch.int64))
    with pytest.raises((ValueError, RuntimeError)):        # neither host nor device seeds
        host.HostSession(fs.host_session_spec(**base))
    with pytest.raises((ValueError, RuntimeError)):        # batch_edges must cover every batch
        host.HostSession(fs.host_session_spec(**dict(base, idx_dev_ptr=8, batch_edges=[1, 2])))
    with pytest.raises((ValueError, RuntimeError)):        # one (rowptr, col) offset pair per hop
        host.HostSession(fs.host_session_spec(**dict(base, idx_dev_ptr=8, n_hops=2)))


def test_randomized_sweep(fx, host):
    """Seeded sweep over hop counts, partition counts, batch / slot / seed counts, label dtypes, seed
    residency and readiness patterns: every delivered piece is compared with the stand-in's pattern,
    mixing try / blocking getters the way a consumer does."""
    import random
    rnd = random.Random(20261018)
    for trial in range(40):
        L = rnd.choice([1, 2, 3, 4])
        P = rnd.choice([-1, -1, 1, 2, 5, 16])
        bs = rnd.choice([1, 3, 16, 33])
        n_seeds = rnd.randint(1, 6 * bs + 2)
        depth = rnd.choice([1, 2, 6])
        polls = rnd.choice([0, 0, 1, 3])
        y_dtype = rnd.choice([torch.int64, torch.int32, torch.float32])
        on_dev = rnd.random() < 0.5
        feat = rnd.random() < 0.8
        hs, ex, ranges, idx, (caps, keep) = make(fx, host, n_seeds=n_seeds, bs=bs, L=L, P=P, depth=depth, polls=polls,
                                                  y_dtype=y_dtype, idx_on_device=on_dev, feat=feat)
        hs.fill()
        assert hs.total == len(ranges) and hs.in_flight == min(depth, len(ranges))
        got = 0
        while got < len(ranges):
            r = hs.get(rnd.random() < 0.7)
            if r is None:
                continue                      # try-get on a batch that is not ready yet
            st, en = ranges[got]
            seed0 = int(idx[st])
            if P < 0:
                x, y, adjs, rng, owners, y_flat = r
                n_id = None
            else:
                n_id, parts, cached, perm, adjs, rng, y, x, owners, y_flat = r
            assert rng == (st, en)
            nb = check_structure(adjs, seed0, en - st, caps)
            if feat:
                assert x.shape == (nb, 8) and x.view(torch.uint8)[:, 0].tolist() == [(seed0 + i) & 0xff for i in range(nb)]
            elif P < 0:
                assert x.shape == (0, 8)
            else:
                assert x is None
            assert y.shape == (en - st, 1) and y.dtype == y_dtype
            assert y_flat.shape == y.squeeze().shape and y_flat.data_ptr() == y.data_ptr() and y_flat.dtype == y_dtype
            if y_dtype == torch.int64:       # the stand-in only writes 8-byte labels
                assert y.view(-1).tolist() == (3 * idx[st:en]).tolist()
            if P >= 0:
                assert n_id.tolist() == [seed0 + i for i in range(nb)] and perm.tolist() == [nb - 1 - i for i in range(nb)]
                assert [p.numel() for p in parts] == [nb // (P + 1)] * P and cached.numel() == nb - P * (nb // (P + 1))
                assert torch.cat(list(parts) + [cached]).tolist() == [7 * pos + seed0 for pos in range(nb)]
                assert inside(owners, n_id, cached, perm, y, *parts)
            assert inside(owners, y, *[t for a in adjs for t in a[:2]]) and (x is None or inside(owners, x))
            got += 1
        assert hs.get(True) is None and hs.consumed == hs.total and fx.fx_submitted(ex) == len(ranges)
        del hs
        fx.fx_destroy(ex)
