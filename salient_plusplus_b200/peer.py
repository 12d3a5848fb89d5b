"""Peer mapping of feature partitions for the P2P gather (K5).

Every rank owns one range partition (or, with fewer GPUs than partitions, a block of consecutive
partitions) of the feature matrix in its own HBM.  Once, at set-up, the ranks exchange CUDA-IPC
handles of those tables over ``torch.distributed`` (control plane only) and map every peer's table
into their own address space; from then on the gather kernel reads peer rows with plain loads
that travel over NVLink / NVSwitch -- no per-batch collective.  This replaces the reference's
three per-batch ``all_to_all`` calls (fast_trainer/transferers.py:521,709,757).

Collective discipline: ``exchange_device_tables`` ALWAYS runs the same two small
``all_gather_object`` rounds on every rank (descriptor, then an "ok" flag), whatever is cached
locally, so ranks can never disagree about entering a collective.  A peer is only mapped when it
lives on the same host and ``cudaDeviceCanAccessPeer`` says the two GPUs are connected; if any
rank fails to map any peer, all ranks agree to return ``None``, the mappings opened by the failed
round are closed again, ``ProtoDistributedBatch.x`` stays ``None`` and the caller falls back to
the all_to_all protocol (``transferers.NcclAllToAllPrefetcher``), which also covers multi-node
runs -- the reference's main deployment.  Exported tables are pinned for the lifetime of the
process (a peer may still hold a mapping of them), re-exports of a table that moved are detected
by their (handle, offset) descriptor and remapped.
"""
from __future__ import annotations

import ctypes
import socket
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import SPP_MAX_PARTS, check

# tables this process has exported: kept alive for the process lifetime (peers map them)
_PINNED: Dict[int, torch.Tensor] = {}
# (group id, peer rank) -> (descriptor, mapped pointer, offset)
_IMPORTED: Dict[Tuple[int, int], Tuple[tuple, int, int]] = {}


def export_handle(t: torch.Tensor) -> Tuple[bytes, int]:
    """IPC handle (64 bytes) + offset of ``t`` inside its allocation."""
    h = (ctypes.c_uint8 * 64)()
    off = ctypes.c_int64(0)
    check(_lib.load().spp_ipc_export(t.data_ptr(), h, ctypes.byref(off)), "spp_ipc_export")
    return bytes(h), int(off.value)


def import_handle(handle: bytes, offset: int) -> int:
    h = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
    p = ctypes.c_void_p()
    check(_lib.load().spp_ipc_import(h, int(offset), ctypes.byref(p)), "spp_ipc_import")
    return int(p.value)


def close_handle(ptr: int, offset: int) -> None:
    _lib.load().spp_ipc_close(ctypes.c_void_p(ptr), int(offset))


def _can_access(peer_device: int) -> bool:
    if peer_device == torch.cuda.current_device():
        return True
    try:
        return bool(torch.cuda.can_device_access_peer(torch.cuda.current_device(), peer_device))
    except Exception:  # noqa: BLE001
        return False


def exchange_device_tables(local: torch.Tensor, group=None) -> Optional[List[int]]:
    """Collective over ``group``: returns, for every group rank, a device pointer to that rank's
    ``local`` table valid in THIS process (own entry = local pointer), or ``None`` on every rank
    when some rank could not map some peer."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world, me = dist.get_world_size(group), dist.get_rank(group)
    gid = id(group) if group is not None else 0
    torch.cuda.synchronize()
    _PINNED[local.data_ptr()] = local
    handle, offset = export_handle(local)
    mine = (socket.gethostname(), torch.cuda.current_device(), handle, offset, local.data_ptr())
    gathered: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    ptrs: List[int] = [0] * world
    opened: List[Tuple[int, tuple]] = []     # (peer, previous entry) of the mappings this round created
    ok = True
    for p, desc in enumerate(gathered):
        if p == me:
            ptrs[p] = local.data_ptr()
            continue
        host, dev, h, off, _ = desc
        key = (gid, p)
        have = _IMPORTED.get(key)
        if have is not None and have[0] == desc:
            ptrs[p] = have[1]                 # same allocation as last time: mapping still valid
            continue
        if host != mine[0] or not _can_access(dev):
            ok = False
            continue
        try:
            if have is not None:              # the peer re-exported a table that moved: drop the stale mapping
                close_handle(have[1], have[2])
                del _IMPORTED[key]
            ptr = import_handle(h, off)
        except _lib.SalientB200Error:
            ok = False
            continue
        _IMPORTED[key] = (desc, ptr, off)
        opened.append((p, have))
        ptrs[p] = ptr
    flags: List[Optional[bool]] = [None] * world
    dist.all_gather_object(flags, ok, group=group)   # doubles as the "everyone has mapped" barrier
    if not all(flags):
        for p, _prev in opened:
            _d, ptr, off = _IMPORTED.pop((gid, p))
            try:
                close_handle(ptr, off)
            except Exception:  # noqa: BLE001
                pass
        return None
    return ptrs


def exchange_partition_tables(local: torch.Tensor, rank: int, num_parts: int,
                              group=None) -> Optional[List[int]]:
    """One partition per rank (the reference's deployment, utils/exp_driver.py:121-123): the device
    pointer of every partition's table as seen from this process, or ``None`` when there is no
    process group whose size equals the number of partitions, or when a peer cannot be mapped
    (different host, no P2P link).  At most ``SPP_MAX_PARTS`` (16) partitions: the range partition
    book travels to the kernels in parameter space."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if num_parts > SPP_MAX_PARTS:
        raise _lib.SalientB200Error(f"at most {SPP_MAX_PARTS} feature partitions are supported (got {num_parts})")
    world = dist.get_world_size(group)
    if world != num_parts or dist.get_rank(group) != rank:
        return None
    return exchange_device_tables(local, group)


def hosted_partitions(rank: int, world: int, num_parts: int) -> List[int]:
    """Partitions resident on GPU ``rank`` when ``num_parts`` partitions are spread over ``world``
    GPUs (``num_parts`` a multiple of ``world``): a block of consecutive partitions."""
    if num_parts % world != 0:
        raise ValueError("the number of partitions must be a multiple of the number of GPUs")
    per = num_parts // world
    return list(range(rank * per, (rank + 1) * per))


def partition_pointers(rank_ptrs: List[int], offsets: List[int], world: int, pitch: int) -> List[int]:
    """Per-PARTITION table pointers from per-RANK block pointers (rank r hosts the consecutive
    partitions ``hosted_partitions(r, ...)`` in one table)."""
    P = len(offsets) - 1
    per = P // world
    out = []
    for p in range(P):
        r = p // per
        out.append(rank_ptrs[r] + (offsets[p] - offsets[r * per]) * pitch)
    return out
