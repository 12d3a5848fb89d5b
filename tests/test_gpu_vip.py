"""VIP analytic model and VIP cache construction on the GPU against the fp64 numpy oracle
(tolerance 1e-12 absolute on probabilities: same arithmetic, different summation order)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from salient_plusplus_b200 import synthetic as S
from tests.util import small_graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fi", [0, 1])
def test_vip_exact_form_against_reference_golden(fi):
    """The kernel's log-product form against the reference's own vip_analytical output
    (tests/golden/vip.npz, fp32) and against the fp64 oracle."""
    import os
    from salient_plusplus_b200 import vip
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vip.npz"))
    fanouts = g[f"fanouts{fi}"].tolist()
    rowptr, col = torch.from_numpy(g["rowptr"]), torch.from_numpy(g["col"])
    for p in range(4):
        train = torch.from_numpy(g[f"train{p}"])
        got = vip.vip_probabilities(rowptr, col, train, 32, fanouts, exact=True).cpu().numpy()
        assert np.max(np.abs(got - g[f"vip{fi}_{p}"].astype(np.float64))) < 2e-5
        assert np.max(np.abs(got - O.vip_probabilities(g["rowptr"], g["col"], g[f"train{p}"], 32, fanouts, exact=True))) < 1e-12


@pytest.mark.parametrize("fi", [0, 1])
def test_vip_first_order_form_against_reference_driver_golden(fi):
    """The kernel's first-order form (what `bench.py` / create_vip_cache use) against the output of
    the reference driver's own get_frequency_tensors_fast (tests/golden/vip_driver.npz, fp64)."""
    import os
    from salient_plusplus_b200 import vip
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vip_driver.npz"))
    fanouts = g[f"fanouts{fi}"].tolist()
    rowptr, col = torch.from_numpy(g["rowptr"]), torch.from_numpy(g["col"])
    for p in range(4):
        got = vip.vip_probabilities(rowptr, col, torch.from_numpy(g[f"train{p}"]), 32, fanouts).cpu().numpy()
        assert np.max(np.abs(got - g[f"vip{fi}_{p}"])) < 1e-12


@pytest.mark.parametrize("fanouts", [[15, 10, 5], [25, 15], [5]])
def test_vip_probabilities(fanouts):
    from salient_plusplus_b200 import vip
    rowptr, col = small_graph(n=5000, e=90000)
    N = rowptr.numel() - 1
    train = S.seeds(N, 700, seed=3, lo=1000, hi=3000)
    got = vip.vip_probabilities(rowptr, col, train, 64, fanouts).cpu().numpy()
    want = O.vip_probabilities(rowptr.numpy(), col.numpy(), train.numpy(), 64, fanouts)
    assert got.dtype == np.float64 and got.shape == (N,)
    assert np.max(np.abs(got - want)) < 1e-12
    assert got.min() >= 0.0 and got.max() <= 1.0 and got[train.numpy()].min() >= 0.0
    # isolated vertices that are not seeds can never be sampled
    deg = (rowptr[1:] - rowptr[:-1]).numpy()
    iso = np.setdiff1d(np.nonzero(deg == 0)[0], train.numpy())
    assert np.all(got[iso] == 0.0)


def test_select_and_create_vip_cache():
    from salient_plusplus_b200 import fast_sampler as fs, vip
    rowptr, col = small_graph(n=6000, e=150000)
    N = rowptr.numel() - 1
    P, rank = 4, 2
    off = S.equal_partition_offsets(N, P)
    lo, hi = int(off[rank]), int(off[rank + 1])
    X = S.features(N, 100, torch.float16, seed=4)
    train = S.seeds(N, 400, seed=5, lo=lo, hi=hi)
    probs = vip.vip_probabilities(rowptr, col, train, 128, [15, 10, 5])
    num = int(N / P * 0.15)
    cv = vip.select_cache_vertices(probs, off, rank, num)
    want = O.select_cache_vertices(probs.cpu().numpy(), off.numpy(), rank, num)
    assert np.array_equal(cv.cpu().numpy(), want)              # same scores -> identical layout
    assert cv.numel() == num and not bool(((cv >= lo) & (cv < hi)).any())
    owner = (torch.searchsorted(off.cuda(), cv, right=True) - 1).cpu()
    assert bool((owner[1:] >= owner[:-1]).all())               # owner-major
    pv = probs[cv].cpu()
    same = owner[1:] == owner[:-1]
    assert bool((pv[1:][same] <= pv[:-1][same]).all())         # VIP-descending inside an owner
    parts = [X[int(off[p]):int(off[p + 1])].contiguous() if p != rank else None for p in range(P)]
    cache = vip.create_vip_cache(rowptr, col, train, 128, [15, 10, 5], off, rank, 15.0, X[lo:hi].contiguous(),
                                 partition_tables=parts)
    assert isinstance(cache, fs.Cache) and cache.rank == rank and cache.world_size == P
    assert torch.equal(cache.cached_vertices, cv)
    assert torch.equal(cache.cached_features.cpu(), X[cv.cpu()])   # rows pulled from the owners' partitions
    # a VIP cache must beat a degree-ranked cache of the same size on expected hits
    dv = S.degree_cache_vertices(rowptr, off, rank, num)
    assert float(probs[cv].sum()) >= float(probs[dv.cuda()].sum())
