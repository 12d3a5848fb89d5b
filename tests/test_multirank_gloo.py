"""world_size-2 tests on CPU (gloo) of the host logic behind the multi-GPU path: how seeds are
sharded, that every rank issues the same number of batches, and the peer-table exchange protocol
(the CUDA-IPC calls are replaced by recorders: the data path itself needs GPUs)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from salient_plusplus_b200 import peer, synthetic as S
        from salient_plusplus_b200.fast_sampler import Config, _batch_ranges
        from salient_plusplus_b200.shufflers import DistributedShuffler, FederatedDistributedShuffler

        N = 1000
        train = torch.arange(0, N, 3)
        # 1. DistributedShuffler: the ranks' slices are disjoint and cover the common permutation
        sh = DistributedShuffler(train, world)
        sh.set_epoch(5)
        mine = sh.get_idx(rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        allv = sum(gathered, [])
        assert sorted(allv) == train.tolist() and len(set(allv)) == len(allv)
        n = train.numel()
        assert mine.numel() == (n * (rank + 1)) // world - (n * rank) // world
        # 2. federated: local ids only; force_exact_num_batches gives every rank the same count
        off = S.equal_partition_offsets(N, world)
        lo, hi = int(off[rank]), int(off[rank + 1])
        local = FederatedDistributedShuffler(train[(train >= lo) & (train < hi)]).get_idx()
        assert bool(((local >= lo) & (local < hi)).all())
        cfg = Config()
        cfg.batch_size, cfg.force_exact_num_batches, cfg.exact_num_batches = 64, True, 3
        ranges = _batch_ranges(local.numel(), cfg)
        counts = [None] * world
        dist.all_gather_object(counts, len(ranges))
        assert counts == [3] * world and ranges[-1][1] == local.numel()
        # 3. peer-table exchange protocol with the IPC calls recorded
        exported, imported, closed = [], [], []
        peer.export_handle = lambda t: (exported.append(t.data_ptr()) or bytes([rank]) * 64, 4096 * rank)
        peer.import_handle = lambda h, o: (imported.append((h[0], o)) or 1_000_000 * (h[0] + 1) + o)
        peer.close_handle = lambda p_, o: closed.append((p_, o))
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.current_device = lambda: 0
        table = torch.zeros(8, 4)
        ptrs = peer.exchange_partition_tables(table, rank, world)
        assert len(ptrs) == world and ptrs[rank] == table.data_ptr()
        for p in range(world):
            if p != rank:
                assert ptrs[p] == 1_000_000 * (p + 1) + 4096 * p
        assert exported == [table.data_ptr()] and sorted(i[0] for i in imported) == [p for p in range(world) if p != rank]
        # second call: every rank takes part in the collective again (no rank may skip it), the
        # unchanged peer tables are NOT mapped a second time; refused when the group size mismatches
        assert peer.exchange_partition_tables(table, rank, world) == ptrs and len(imported) == world - 1 and not closed
        assert peer.exchange_partition_tables(table, rank, world + 1) is None
        # a peer re-exports a table that moved: the stale mapping is closed and replaced
        if rank == 0:
            peer.export_handle = lambda t: (bytes([rank]) * 64, 8192)
        ptrs2 = peer.exchange_partition_tables(table, rank, world)
        if rank == 1:
            assert ptrs2[0] == 1_000_000 + 8192 and closed == [(ptrs[0], 0)] and len(imported) == world
        else:
            assert ptrs2 == ptrs
        # a peer that cannot be mapped (other host): every rank agrees on None and the mappings opened
        # in the failed round are closed again -> the caller falls back to the all_to_all path
        import socket
        if rank == 1:
            socket.gethostname = lambda: "some-other-host"
        n_closed = len(closed)
        peer._IMPORTED.clear()
        assert peer.exchange_partition_tables(table, rank, world) is None
        assert not peer._IMPORTED and len(closed) >= n_closed
        assert peer.hosted_partitions(1, 2, 8) == [4, 5, 6, 7]
        assert peer.partition_pointers([1000, 5000], [0, 10, 20, 30, 40], 2, 4) == [1000, 1040, 5000, 5040]
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 400)
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert not p.is_alive(), "rank hung"
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
