"""ctypes binding of ``libsalient_b200.so`` (the C ABI declared in ``include/salient_b200.h``).

The library is the product: if it is missing this module raises -- there is no CPU or PyTorch
fallback anywhere in the package.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C salient_plusplus_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, c_char_p, c_int, c_int32, c_int64, c_uint8, c_uint64,
                    c_void_p)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsalient_b200.so")
HOST_EXT_PATH = os.path.join(_PKG, "_spp_host.so")
CSRC = os.path.join(_PKG, "csrc")

ABI_VERSION = 2
SPP_MAX_PARTS = 16
SPP_MAX_HOPS = 8
SPP_MAX_FANOUT = 128
SPP_META_WORDS = 32
META_EDGES0 = 12
META_OVERFLOW = 24

EXPORTED = [
    "spp_abi_version", "spp_tune", "spp_last_error", "spp_launch_count", "spp_graph_replays", "spp_graph_captures",
    "spp_gather_rows", "spp_gather_rows_pitched", "spp_gather_partitioned", "spp_gather_by_class",
    "spp_nid2partid", "spp_nid2localnid", "spp_nid_is_local",
    "spp_cache_index_bytes", "spp_cache_build_index", "spp_nid_is_cached", "spp_nid2cachenid",
    "spp_split_scratch_words", "spp_split_by_owner",
    "spp_sampler_sizes", "spp_sample_minibatch", "spp_sample_begin", "spp_sample_hop_count",
    "spp_sample_hop_fill", "spp_sample_export_nids", "spp_debug_set_timeline",
    "spp_trace_begin", "spp_trace_end",
    "spp_batch_enqueue", "spp_batch_prepare", "spp_executor_create", "spp_executor_destroy", "spp_executor_submit",
    "spp_executor_poll", "spp_executor_wait", "spp_executor_times",
    "spp_vip_hop",
    "spp_ipc_export", "spp_ipc_import", "spp_ipc_close", "spp_enable_peer_access",
]


class FeatureMap(Structure):
    _fields_ = [("num_parts", c_int32), ("rank", c_int32),
                ("offsets", c_int64 * (SPP_MAX_PARTS + 1)),
                ("tables", c_void_p * SPP_MAX_PARTS),
                ("cache_table", c_void_p), ("cache_index", c_void_p), ("cache_index_nodes", c_int64),
                ("table_pitch", c_int64), ("cache_pitch", c_int64),
                ("local_parts", ctypes.c_uint32), ("_pad", ctypes.c_uint32)]


class Graph(Structure):
    _fields_ = [("rowptr", c_void_p), ("col", c_void_p), ("col_is_64", c_int32), ("_pad", c_int32),
                ("num_nodes", c_int64)]


class SamplerWs(Structure):
    _fields_ = [("table", c_void_p), ("table_slots", c_int64), ("n_ids", c_void_p),
                ("max_nodes", c_int64), ("tgt_start", c_void_p), ("tgt_deg", c_void_p),
                ("max_targets", c_int64), ("tile_state", c_void_p), ("tile_words", c_int64),
                ("meta", c_void_p), ("cand", c_void_p), ("cand_words", c_int64),
                ("table_direct", c_int32), ("_pad", c_int32)]


class SamplerSizes(Structure):
    _fields_ = [("max_nodes", c_int64), ("max_targets", c_int64), ("table_slots", c_int64),
                ("tile_words", c_int64), ("cand_words", c_int64), ("table_direct", c_int64),
                ("hop_targets", c_int64 * SPP_MAX_HOPS),
                ("hop_edges", c_int64 * SPP_MAX_HOPS)]


class DeviceJob(Structure):
    _fields_ = [("out_rowptr", c_void_p * SPP_MAX_HOPS), ("out_col", c_void_p * SPP_MAX_HOPS),
                ("out_col_cap", c_int64 * SPP_MAX_HOPS), ("n_id_out", c_void_p), ("x_out", c_void_p),
                ("y_out", c_void_p), ("bucket_ids", c_void_p), ("perm", c_void_p), ("seeds", c_void_p),
                ("batch_size", c_int64), ("rng_premixed", c_uint64), ("scan_epoch", c_uint64)]


class BatchJob(Structure):
    _fields_ = [("graph", Graph), ("ws", SamplerWs),
                ("seeds_host", c_void_p), ("seeds_dev", c_void_p), ("batch_size", c_int64),
                ("sizes", c_int32 * SPP_MAX_HOPS), ("n_hops", c_int32), ("replace", c_int32),
                ("rng_seed", c_uint64),
                ("out_rowptr", c_void_p * SPP_MAX_HOPS), ("out_col", c_void_p * SPP_MAX_HOPS),
                ("out_col_cap", c_int64 * SPP_MAX_HOPS), ("n_id_out", c_void_p),
                ("feature_mode", c_int32), ("do_split", c_int32), ("use_cache", c_int32), ("_pad", c_int32),
                ("table", c_void_p), ("table_pitch", c_int64), ("row_bytes", c_int64),
                ("fmap", FeatureMap), ("x_out", c_void_p), ("y_table", c_void_p), ("y_row_bytes", c_int64),
                ("y_out", c_void_p), ("bucket_ids", c_void_p), ("perm", c_void_p), ("bucket_counts", c_void_p),
                ("split_scratch", c_void_p), ("meta_host", c_void_p), ("stream", c_void_p),
                ("gather_counters", c_void_p),
                ("job_dev", c_void_p), ("job_host", c_void_p), ("seeds_stage_host", c_void_p),
                ("batch_size_cap", c_int64), ("out_col_bound", c_int64 * SPP_MAX_HOPS)]


class SalientB200Error(RuntimeError):
    """Raised for any non-zero return of the C ABI (mirrors the reference's TORCH_CHECK ->
    RuntimeError behaviour, fast_sampler/fast_sampler.cpp:572-577)."""


_lib = None


def build_library(force: bool = False, quiet: bool = True) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", CSRC, "-j", "4"], stdout=out)
    return LIB_PATH


def build_host_extension(force: bool = False, quiet: bool = True) -> str:
    """Compile ``csrc/host_session.cpp`` (the per-batch host path of a Session: pybind11 + libtorch,
    no device code) into ``_spp_host.so`` next to the package with plain g++ -- in-tree, so the built
    file travels with the source snapshot instead of living in a JIT cache."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(CSRC, "host_session.cpp")
    hdr = os.path.join(os.path.dirname(_PKG), "include", "salient_b200.h")
    if not force and os.path.exists(HOST_EXT_PATH) and \
            os.path.getmtime(HOST_EXT_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return HOST_EXT_PATH
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           "-DTORCH_EXTENSION_NAME=_spp_host", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-I" + os.path.dirname(hdr), "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(cuda_home, "include")]
    cmd += ["-I" + p for p in ce.include_paths()]
    cmd += [src, "-o", HOST_EXT_PATH + ".tmp"]
    cmd += ["-L" + p for p in ce.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if not quiet:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(HOST_EXT_PATH + ".tmp", HOST_EXT_PATH)
    return HOST_EXT_PATH


_host = None


def load_host():
    """The ``_spp_host`` module (native per-batch host path).  Missing or stale builds raise: the
    Session does not silently fall back to the interpreter path (``SPP_NATIVE_HOST=0`` selects it)."""
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_EXT_PATH):
        raise SalientB200Error(
            f"{HOST_EXT_PATH} is missing: the host-path extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`); SPP_NATIVE_HOST=0 selects the "
            "interpreter implementation of the same bookkeeping")
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension resolves its symbols)
    spec = importlib.util.spec_from_file_location("salient_plusplus_b200._spp_host", HOST_EXT_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.ABI_VERSION != ABI_VERSION or mod.BATCH_JOB_BYTES != ctypes.sizeof(BatchJob):
        raise SalientB200Error("_spp_host.so was built against a different salient_b200.h: rebuild it")
    _host = mod
    return mod


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SalientB200Error(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `make -C salient_plusplus_b200/csrc`); there is no fallback path")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, ci = c_void_p, c_int64, c_int32, c_int
    L.spp_abi_version.restype = ci
    L.spp_last_error.restype = c_char_p
    L.spp_tune.argtypes = [c_char_p, ci]
    L.spp_launch_count.restype = c_uint64
    L.spp_graph_replays.restype = c_uint64
    L.spp_graph_captures.restype = c_uint64
    L.spp_gather_rows.argtypes = [vp, i64, vp, ci, i64, vp, vp, i64, vp]
    L.spp_gather_rows_pitched.argtypes = [vp, i64, i64, vp, ci, i64, vp, vp, i64, vp]
    L.spp_gather_partitioned.argtypes = [POINTER(FeatureMap), i64, vp, ci, i64, vp, vp, vp, i64, vp, vp]
    L.spp_gather_by_class.argtypes = [POINTER(FeatureMap), i64, vp, vp, i64, ctypes.c_uint32, vp, vp, vp]
    L.spp_nid2partid.argtypes = [POINTER(i64), ci, vp, i64, vp, vp]
    L.spp_nid2localnid.argtypes = [POINTER(i64), ci, ci, vp, i64, vp, vp]
    L.spp_nid_is_local.argtypes = [POINTER(i64), ci, ci, vp, i64, vp, vp]
    L.spp_cache_index_bytes.restype = i64
    L.spp_cache_index_bytes.argtypes = [i64, i64]
    L.spp_cache_build_index.argtypes = [vp, i64, i64, vp, vp]
    L.spp_nid_is_cached.argtypes = [vp, i64, vp, i64, vp, vp]
    L.spp_nid2cachenid.argtypes = [vp, i64, vp, i64, vp, vp]
    L.spp_split_scratch_words.restype = i64
    L.spp_split_scratch_words.argtypes = [i64]
    L.spp_split_by_owner.argtypes = [POINTER(FeatureMap), ci, vp, ci, i64, vp, vp, vp, vp, vp, vp]
    L.spp_sampler_sizes.argtypes = [i64, POINTER(i32), ci, ci, i64, i64, POINTER(SamplerSizes)]
    L.spp_sample_minibatch.argtypes = [POINTER(Graph), vp, i64, POINTER(i32), ci, ci, c_uint64,
                                       POINTER(SamplerWs), POINTER(vp), POINTER(vp), POINTER(i64),
                                       vp, vp]
    L.spp_sample_begin.argtypes = [POINTER(Graph), vp, i64, POINTER(SamplerWs), vp]
    L.spp_sample_hop_count.argtypes = [POINTER(Graph), ci, i32, ci, i64, POINTER(SamplerWs), vp, vp]
    L.spp_sample_hop_fill.argtypes = [POINTER(Graph), ci, i32, ci, c_uint64, i64, i64,
                                      POINTER(SamplerWs), vp, vp, vp]
    L.spp_debug_set_timeline.restype = None
    L.spp_debug_set_timeline.argtypes = [vp]
    L.spp_sample_export_nids.argtypes = [POINTER(SamplerWs), ci, vp, ci, i64, vp]
    L.spp_trace_begin.argtypes = [i64]
    L.spp_trace_end.restype = i64
    L.spp_trace_end.argtypes = [POINTER(i32), POINTER(i32), POINTER(c_uint64), POINTER(ctypes.c_double), i64]
    L.spp_batch_enqueue.argtypes = [POINTER(BatchJob)]
    L.spp_batch_prepare.argtypes = [POINTER(BatchJob)]
    L.spp_executor_create.restype = vp
    L.spp_executor_create.argtypes = [ci]
    L.spp_executor_destroy.restype = None
    L.spp_executor_destroy.argtypes = [vp]
    L.spp_executor_submit.restype = c_uint64
    L.spp_executor_submit.argtypes = [vp, POINTER(BatchJob)]
    L.spp_executor_poll.argtypes = [vp, c_uint64]
    L.spp_executor_wait.argtypes = [vp, c_uint64]
    L.spp_executor_times.argtypes = [vp, c_uint64, POINTER(ctypes.c_double)]
    L.spp_vip_hop.argtypes = [POINTER(Graph), ctypes.c_double, ci, vp, vp, vp, vp, vp]
    L.spp_ipc_export.argtypes = [vp, POINTER(c_uint8), POINTER(i64)]
    L.spp_ipc_import.argtypes = [POINTER(c_uint8), i64, POINTER(vp)]
    L.spp_ipc_close.argtypes = [vp, i64]
    L.spp_enable_peer_access.argtypes = [ci]
    for name in EXPORTED:
        fn = getattr(L, name)  # raises AttributeError if a declared symbol is not exported
        if fn.restype is ci and name not in ("spp_abi_version",):
            pass
    if L.spp_abi_version() != ABI_VERSION:
        raise SalientB200Error("libsalient_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().spp_last_error().decode("utf-8", "replace")
        raise SalientB200Error(f"{what or 'libsalient_b200'} failed (code {rc}): {msg}")


TRACE_LABELS = ["batch_begin", "seeds_h2d", "table_clear", "seeds_init", "sample", "compact", "relabel_sort",
                "export_nid", "owner_split", "feature_gather", "label_gather", "meta_d2h", "join", "degree_count_scan",
                "sort_large_rows", "sort_rows_bitmap"]


def trace_begin(max_marks: int = 4096) -> None:
    """Arm the library's event trace (diagnostics; tools/trace_pipeline.py)."""
    check(load().spp_trace_begin(int(max_marks)), "spp_trace_begin")


def trace_end(cap: int = 4096):
    """Stop the trace; returns [(label, hop, stream handle, ms since the first mark)] in issue order."""
    lab, hop = (c_int32 * cap)(), (c_int32 * cap)()
    st, ms = (c_uint64 * cap)(), (ctypes.c_double * cap)()
    n = int(load().spp_trace_end(lab, hop, st, ms, cap))
    return [(TRACE_LABELS[lab[i]], int(hop[i]), int(st[i]), float(ms[i])) for i in range(n)]


def tune(key: str, value: int) -> None:
    """Run-time tunable of the library (see spp_tune in the header)."""
    check(load().spp_tune(key.encode(), int(value)), "spp_tune")


def launch_count() -> int:
    return int(load().spp_launch_count())
