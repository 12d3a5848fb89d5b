#!/usr/bin/env bash
# Test infrastructure, NOT product code.
#
# Builds the UNMODIFIED reference CPU sampler (the pybind11 module `fast_sampler`)
# straight from the sources where they lie under /root/reference/fast_sampler into
# oracle/_ref/ (git-ignored; it travels to the GPU box with the gpurun snapshot).
# No reference source is copied into the repository.
#
#   oracle/_ref/fast_sampler.so        - reference as is (distributed Session needs a CUDA
#                                        driver: it pins host memory unconditionally,
#                                        fast_sampler.cpp:1026,1037,... range_partition_book.cpp:188)
#   oracle/_ref/nopin/fast_sampler.so  - same sources, streamed through `sed` into a scratch
#                                        dir under /tmp with the literal `pinned_memory(true)` /
#                                        `vector_to_tensor(x, true)` switched off so the
#                                        distributed branch also runs in the GPU-less build
#                                        container. Used ONLY to generate tests/golden fixtures.
#
# Flags follow fast_sampler/setup.py:22-34 (-O3 -std=c++17 -fopenmp) except -march: the build
# container and the GPU box may have different CPUs, so -march=x86-64-v3 (AVX2/FMA/BMI2) replaces -march=native.
set -euo pipefail
REF=${REF:-/root/reference/fast_sampler}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF absent (GPU box?) - using prebuilt files in $OUT if any"; exit 0
fi
mkdir -p "$OUT" "$OUT/nopin"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC PB_INC ABI <<<"$($PY - <<'PY'
import sysconfig, torch, pybind11, os
from torch.utils import cpp_extension as ce
incs = " ".join("-I"+p for p in ce.include_paths())
print(incs.replace(" ", ","), os.path.join(os.path.dirname(torch.__file__), "lib"),
      sysconfig.get_paths()["include"], pybind11.get_include(), int(torch._C._GLIBCXX_USE_CXX11_ABI))
PY
)"
TORCH_INC=${TORCH_INC//,/ }
MARCH=${MARCH:--march=x86-64-v3}
CXXFLAGS="-O3 $MARCH -std=c++17 -fopenmp -fPIC -DNDEBUG -w -DTORCH_EXTENSION_NAME=fast_sampler \
  -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI $TORCH_INC -I$PY_INC -I$PB_INC"
LDFLAGS="-shared -fopenmp -L$TORCH_LIB -Wl,-rpath,$TORCH_LIB -ltorch -ltorch_cpu -lc10 -ltorch_python"

build_one () { # $1 = source dir, $2 = output .so
  local src="$1" out="$2" tmp; tmp="$(mktemp -d /tmp/spp_ref_obj.XXXX)"
  g++ $CXXFLAGS -I"$src" -I"$REF/parallel-hashmap" -c "$src/fast_sampler.cpp" -o "$tmp/fs.o" &
  g++ $CXXFLAGS -I"$src" -I"$REF/parallel-hashmap" -c "$src/range_partition_book.cpp" -o "$tmp/rpb.o" &
  wait
  g++ "$tmp/fs.o" "$tmp/rpb.o" $LDFLAGS -o "$out"
  rm -rf "$tmp"
}

if [ ! -f "$OUT/fast_sampler.so" ] || [ "${FORCE:-0}" = 1 ]; then
  echo "[build_ref] compiling unmodified reference -> $OUT/fast_sampler.so"
  build_one "$REF" "$OUT/fast_sampler.so"
fi
if [ ! -f "$OUT/nopin/fast_sampler.so" ] || [ "${FORCE:-0}" = 1 ]; then
  echo "[build_ref] compiling no-pin variant -> $OUT/nopin/fast_sampler.so"
  SCR="$(mktemp -d /tmp/spp_ref_nopin.XXXX)"
  for f in fast_sampler.cpp range_partition_book.cpp range_partition_book.hpp sample_cpu.hpp utils.hpp concurrentqueue.h; do
    sed -e 's/pinned_memory(true)/pinned_memory(false)/g' \
        -e 's/vector_to_tensor(\([a-z_]*\), true)/vector_to_tensor(\1, false)/g' "$REF/$f" > "$SCR/$f"
  done
  build_one "$SCR" "$OUT/nopin/fast_sampler.so"
  rm -rf "$SCR"
fi
echo "[build_ref] done"
