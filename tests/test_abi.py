"""The C-ABI library loads without a GPU and exports every symbol include/salient_b200.h
declares; host-only entry points behave; nothing under the package imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "salient_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from salient_plusplus_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def test_exports_match_header(lib):
    from salient_plusplus_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == names
    assert lib.spp_abi_version() == 2


def test_struct_layouts_match_header():
    from salient_plusplus_b200 import _lib
    assert ctypes.sizeof(_lib.FeatureMap) == 4 + 4 + 17 * 8 + 16 * 8 + 8 + 8 + 8 + 16 + 8
    assert ctypes.sizeof(_lib.Graph) == 32
    assert ctypes.sizeof(_lib.SamplerWs) == 104
    assert ctypes.sizeof(_lib.SamplerSizes) == 6 * 8 + 2 * 8 * 8


def test_sampler_sizes_host_only(lib):
    from salient_plusplus_b200 import _lib
    out = _lib.SamplerSizes()
    sizes = (ctypes.c_int32 * 3)(15, 10, 5)
    assert lib.spp_sampler_sizes(1024, sizes, 3, 0, 0, 0, ctypes.byref(out)) == 0
    assert list(out.hop_targets)[:3] == [1024, 16384, 180224]
    assert list(out.hop_edges)[:3] == [15360, 163840, 901120]
    assert out.max_nodes == 1081344 and out.max_targets == 180224
    assert out.table_slots == 1460224 and out.cand_words == 15360 + 163840 + 901120 + 16  # one range per hop
    # node bound capped by the graph size, edge bound by the maximum degree
    assert lib.spp_sampler_sizes(1024, sizes, 3, 0, 100000, 7, ctypes.byref(out)) == 0
    assert out.max_nodes == 101024 and list(out.hop_edges)[:3] == [1024 * 7, 8192 * 7, 65536 * 5]
    assert out.table_direct == 1 and out.table_slots == 100000      # small graph: direct-mapped table
    assert lib.spp_sampler_sizes(1024, sizes, 3, 0, 111059956, 500, ctypes.byref(out)) == 0
    assert out.table_direct == 0 and out.table_slots == 1460224     # papers100M-sized graph: hashed (1.35 x node bound), L2 resident
    # with replacement every target with a neighbour emits exactly k edges: no max_degree tightening
    assert lib.spp_sampler_sizes(1024, sizes, 3, 1, 100000, 7, ctypes.byref(out)) == 0
    assert list(out.hop_edges)[:2] == [15360, 163840]
    full = (ctypes.c_int32 * 1)(-1)
    assert lib.spp_sampler_sizes(8, full, 1, 0, 100, -1, ctypes.byref(out)) == _lib.SPP_EINVAL if hasattr(_lib, "SPP_EINVAL") else True
    assert lib.spp_sampler_sizes(8, full, 1, 0, 100, 9, ctypes.byref(out)) == 0 and out.hop_edges[0] == 72
    assert lib.spp_sampler_sizes(8, sizes, 99, 0, 0, 0, ctypes.byref(out)) != 0
    assert b"n_hops" in lib.spp_last_error()
    assert lib.spp_split_scratch_words(0) > 0 and lib.spp_split_scratch_words(10 ** 6) > 10 ** 6
    # cache index: one 32-byte block per 224 ids + 4 bytes per cached vertex (+ scan scratch): ~N/7 bytes
    b = lib.spp_cache_index_bytes(111059956, 2082374)
    assert 111059956 // 7 + 4 * 2082374 <= b <= 111059956 // 7 + 4 * 2082374 + 64 * 1024
    assert lib.spp_cache_index_bytes(0, 0) > 0


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.spp_gather_rows(None, 0, None, 1, 10, None, None, 10, None) != 0
    assert b"row_bytes" in lib.spp_last_error()
    assert lib.spp_gather_rows(None, 16, None, 1, 0, None, None, 0, None) == 0      # empty gather is a no-op
    assert lib.spp_launch_count() == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "salient_plusplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "salient_oracle" not in text or f.endswith((".cu", ".cuh", ".py")) and "spo_rand64" in text, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from salient_plusplus_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SalientB200Error):
        _lib.load()


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from salient_plusplus_b200 import fast_sampler as fs
    with pytest.raises(RuntimeError):
        fs.serial_index(torch.zeros(4, 4), torch.tensor([0]))


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 and a C program must link
    against the library and call a host-only entry point (no GPU needed)."""
    import subprocess
    from salient_plusplus_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text('''
#include <stdio.h>
#include "salient_b200.h"
int main(void) {
  spp_sampler_sizes_t s;
  int32_t sizes[3] = {15, 10, 5};
  if (spp_abi_version() != SPP_ABI_VERSION) return 1;
  if (spp_sampler_sizes(1024, sizes, 3, 0, 0, 0, &s) != 0) return 2;
  if (s.max_nodes != 1081344) return 3;
  if (spp_gather_rows(0, 0, 0, 1, 1, 0, 0, 1, 0) == 0) return 4;   /* bad row_bytes must fail */
  printf("%s\\n", spp_last_error());
  printf("sizes %d %d %d %d\\n", (int)sizeof(spp_batch_job), (int)sizeof(spp_feature_map), (int)sizeof(spp_sampler_ws),
         (int)sizeof(spp_device_job));
  return 0;
}
''')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-l:libsalient_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "row_bytes" in out.stdout
    # the ctypes mirrors have exactly the C layout
    want = "sizes %d %d %d %d" % (ctypes.sizeof(_lib.BatchJob), ctypes.sizeof(_lib.FeatureMap), ctypes.sizeof(_lib.SamplerWs),
                                  ctypes.sizeof(_lib.DeviceJob))
    assert want in out.stdout, (want, out.stdout)
