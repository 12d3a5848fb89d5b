"""Host-side logic that needs neither a GPU nor the reference: batch ranges, config mirror,
feature-map marshalling, Adj wrappers."""
import ctypes

import pytest
import torch

from oracle import oracle as O
from salient_plusplus_b200 import _lib
from salient_plusplus_b200.adj import Adj, Adj__from_fast_sampler
from salient_plusplus_b200.fast_sampler import Cache, Config, RangePartitionBook, _batch_ranges, make_feature_map
from salient_plusplus_b200.samplers import FastSamplerConfig, PreparedBatch, ProtoDistributedBatch


@pytest.mark.parametrize("n,bs,skip,exact,B", [(1000, 64, False, False, 0), (1000, 64, True, False, 0), (1024, 64, True, False, 0),
                                              (1000, 64, False, True, 7), (10, 64, False, True, 3), (0, 64, False, False, 0),
                                              (63, 64, True, False, 0), (100, 1, False, True, 100)])
def test_batch_ranges_match_oracle(n, bs, skip, exact, B):
    cfg = Config()
    cfg.batch_size, cfg.skip_nonfull_batch, cfg.force_exact_num_batches, cfg.exact_num_batches = bs, skip, exact, B
    assert _batch_ranges(n, cfg) == O.batch_ranges(n, bs, skip, exact, B)


def test_fast_sampler_config_mirror():
    t = torch.zeros(4, 2)
    cfg = FastSamplerConfig(x_cpu=t, x_gpu=t, y=None, rowptr=torch.zeros(5, dtype=torch.int64), col=torch.zeros(0, dtype=torch.int64),
                            idx=torch.arange(130), batch_size=64, sizes=[15, 10, 5], skip_nonfull_batch=False,
                            pin_memory=True, distributed=False)
    assert cfg.get_num_batches() == 3
    cfg.skip_nonfull_batch = True
    assert cfg.get_num_batches() == 2
    cfg.force_exact_num_batches, cfg.exact_num_batches = True, 9
    assert cfg.get_num_batches() == 9
    c = cfg.to_fast_sampler()
    assert c.sizes == [15, 10, 5] and c.batch_size == 64 and isinstance(c.cache, Cache) and c.partition_book is None
    cfg.distributed = True
    cfg.partition_book = RangePartitionBook(1, 2, torch.tensor([0, 2, 4]))
    assert cfg.to_fast_sampler().partition_book.rank == 1


def test_feature_map_marshalling():
    fm = make_feature_map([0, 10, 25, 40], 1, [None, None, None], None, None, [111, 0, 333], 256, 128)
    assert fm.num_parts == 3 and fm.rank == 1 and list(fm.offsets)[:5] == [0, 10, 25, 40, 40]
    assert fm.tables[0] == 111 and fm.tables[1] is None and fm.tables[2] == 333
    assert fm.cache_table is None and fm.cache_index is None and fm.local_parts == 0 and fm.table_pitch == 256 and fm.cache_pitch == 128
    with pytest.raises(RuntimeError):
        RangePartitionBook(0, 1, torch.arange(40))._off()      # more than SPP_MAX_PARTS partitions


def test_adj_and_batches():
    rowptr, col = torch.tensor([0, 1, 3]), torch.tensor([2, 0, 1])
    a = Adj__from_fast_sampler((rowptr, col, torch.empty(0, dtype=torch.int64), (2, 3)))
    assert isinstance(a, Adj) and a.size == (3, 2) and a.adj_t.sparse_sizes() == (2, 3)
    pb = PreparedBatch.from_fast_sampler((torch.zeros(3, 4), torch.zeros(2, 1), [(rowptr, col, torch.empty(0), (2, 3))], (5, 7)))
    assert pb.batch_size == 2 and pb.num_total_nodes == 3 and pb.y.shape == (2,) and pb.idx_range == slice(5, 7)

    class B:
        partition_nids = [torch.tensor([1]), torch.tensor([2, 3])]
        sliced_cpu_features = torch.zeros(0, 4)
        sliced_cpu_labels = torch.zeros(2, 1)
        cached_nids = torch.zeros(0, dtype=torch.int64)
        perm_partition_to_mfg = torch.tensor([0, 1, 2])
        adjs = [(rowptr, col, torch.empty(0), (2, 3))]
        idx_range = (0, 2)
    p = ProtoDistributedBatch.from_fast_sampler(B())
    assert p.num_total_nodes == 3 and p.num_cached_nodes == 0 and p.idx_range == slice(0, 2) and p.x is None


def test_struct_sizes_stable():
    assert ctypes.sizeof(_lib.BatchJob) == 952 and ctypes.sizeof(_lib.DeviceJob) == 264


def test_prepared_batch_keeps_four_fields_and_owner_fast_path():
    """The reference unpacks a PreparedBatch into four values (driver/models.py:464), so `owners`
    must stay an attribute; record_stream must touch exactly the owners when they are known."""
    from salient_plusplus_b200.fast_sampler import OwnedSample
    from salient_plusplus_b200.samplers import OwnedPreparedBatch

    class Probe:
        def __init__(self):
            self.calls = 0
            self.is_cuda = True

        def record_stream(self, stream):
            self.calls += 1

    rowptr, col = torch.tensor([0, 1, 3]), torch.tensor([2, 0, 1])
    sample = OwnedSample((torch.zeros(3, 4), torch.zeros(2, 1), [(rowptr, col, torch.empty(0), (2, 3))], (5, 7)))
    a, b = Probe(), Probe()
    sample.owners = (a, b)
    pb = PreparedBatch.from_fast_sampler(sample)
    assert isinstance(pb, OwnedPreparedBatch) and isinstance(pb, PreparedBatch) and len(pb) == 4
    x, y, adjs, rng = pb                                     # four-way unpacking still works
    assert rng == slice(5, 7) and PreparedBatch._fields == ("x", "y", "adjs", "idx_range")
    pb.record_stream(object())
    assert (a.calls, b.calls) == (1, 1)
    plain = PreparedBatch.from_fast_sampler(tuple(sample))   # no owners: per-tensor path, CPU tensors are skipped
    assert type(plain) is PreparedBatch and plain.owners == ()
    plain.record_stream(object())
    assert pb.to("cpu").x.shape == (3, 4)


def test_trace_label_table_matches_header():
    """The label order of _lib.TRACE_LABELS is the enum order documented in the header."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "salient_b200.h")).read()
    doc = text[text.index("diagnostics: event trace"):text.index("int spp_trace_begin")]
    doc = doc[doc.index("issue order: label"):]
    ids = [int(v) for v in re.findall(r"(?<![\w.])(\d+) [a-zA-Z]", doc)]
    assert ids[:16] == list(range(16)) and len(_lib.TRACE_LABELS) == 16
    cuh = open(os.path.join(root, "salient_plusplus_b200", "csrc", "common.cuh")).read()
    enum = cuh[cuh.index("enum TraceLabel {"):cuh.index("};", cuh.index("enum TraceLabel {"))]
    assert len(re.findall(r"kTr\w+", enum)) == len(_lib.TRACE_LABELS)


def test_arena_cuts_cover_the_arena_and_land_on_the_layout():
    """One split call cuts a batch's arena: the pieces must tile the arena exactly and the
    tensors must start at the offsets the kernels were given."""
    from salient_plusplus_b200._lib import META_EDGES0, SPP_META_WORDS
    from salient_plusplus_b200.fast_sampler import _arena_cuts
    # layout of a 2-hop batch: bounds T=(4, 10), E=(12, 30), node bound 40, 3 label words at the tail
    a_off = [(0, 5), (17, 28)]
    a_nid = 58
    m = [0] * (SPP_META_WORDS + 18)
    m[0], m[1], m[2] = 4, 9, 21                       # |n_id| before hop 0 / 1, final
    m[META_EDGES0], m[META_EDGES0 + 1] = 7, 25
    # non-distributed
    lay = (a_off, a_nid, a_nid, a_nid + 3, 40)
    cuts = _arena_cuts(lay, m, 2, None)
    assert sum(cuts) == lay[3] and all(c >= 0 for c in cuts)
    starts = [sum(cuts[:i]) for i in range(len(cuts))]
    assert (starts[0], cuts[0]) == (0, 5) and (starts[2], cuts[2]) == (5, 7)          # hop 0 rowptr, col
    assert (starts[4], cuts[4]) == (17, 10) and (starts[6], cuts[6]) == (28, 25)      # hop 1 rowptr, col
    # distributed, P = 3: buckets 6 + 5 + 4, 6 cached
    m[SPP_META_WORDS:SPP_META_WORDS + 4] = [6, 5, 4, 6]
    lay = (a_off, a_nid, a_nid + 120, a_nid + 123, 40)
    cuts = _arena_cuts(lay, m, 2, 3)
    assert sum(cuts) == lay[3] and all(c >= 0 for c in cuts)
    starts = [sum(cuts[:i]) for i in range(len(cuts))]
    q = 8
    assert (starts[q], cuts[q]) == (58, 21)                                           # n_id
    assert [starts[q + 2 + p] for p in range(3)] == [98, 104, 109] and cuts[q + 2:q + 5] == [6, 5, 4]
    assert (starts[q + 5], cuts[q + 5]) == (113, 6)                                   # cached ids
    assert (starts[q + 7], cuts[q + 7]) == (138, 21)                                  # perm
    t = torch.arange(lay[3]).split_with_sizes(cuts)
    assert t[q + 7][0].item() == 138 and t[-1].numel() == 3
