"""Peer mapping of feature partitions for the P2P gather (K5).

Every rank owns one range partition of the feature matrix in its own HBM.  Once, at set-up, the
ranks exchange CUDA-IPC handles of those tables over ``torch.distributed`` (control plane only)
and map every peer's table into their own address space; from then on the gather kernel reads
peer rows with plain loads that travel over NVLink / NVSwitch -- no per-batch collective.  This
replaces the reference's three per-batch ``all_to_all`` calls
(fast_trainer/transferers.py:521,709,757).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import check

# (table data_ptr, rank, world) -> (ptrs, keep-alive)
_MAPPED: Dict[Tuple[int, int, int], List[int]] = {}


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


def exchange_partition_tables(local: torch.Tensor, rank: int, num_parts: int,
                              group=None) -> Optional[List[int]]:
    """Collective: returns the device pointer of every partition's table as seen from this
    process (own entry = local pointer), or ``None`` when there is no process group whose size
    equals the number of partitions (single-process runs must pass ``Config.partition_tables``)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size(group)
    if world != num_parts or dist.get_rank(group) != rank:
        return None
    key = (local.data_ptr(), rank, world)
    if key in _MAPPED:
        return list(_MAPPED[key])
    torch.cuda.synchronize()
    mine = export_handle(local) + (torch.cuda.current_device(),)
    gathered: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    ptrs: List[int] = []
    for p, (handle, offset, dev) in enumerate(gathered):
        if p == rank:
            ptrs.append(local.data_ptr())
        else:
            ptrs.append(import_handle(handle, offset))
    dist.barrier(group=group)  # every peer has mapped before anyone may free / move its table
    _MAPPED[key] = list(ptrs)
    return ptrs
