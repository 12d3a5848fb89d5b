// partition.cu -- K2 RangePartitionBook / Cache lookups and K3 the distributed split of a
// mini-batch's node list (fast_sampler/range_partition_book.cpp:85-195,
// fast_sampler/fast_sampler.cpp:1017-1262).  Integer, HBM/L2-bound work.
#include "common.cuh"

namespace spp {

constexpr int kEwThreads = 256;

static int ew_grid(int64_t n) {
  int64_t ctas = ceil_div(n > 0 ? n : 1, kEwThreads * 4);
  int64_t cap = (int64_t)num_sms() * 8;
  return (int)(ctas < cap ? ctas : cap);
}

static int fill_book(BookParams& b, const int64_t* offsets, int num_parts, int rank, const char* who) {
  if (!offsets) return fail(SPP_EINVAL, "%s: null partition offsets", who);
  if (num_parts < 1 || num_parts > SPP_MAX_PARTS) return fail(SPP_EINVAL, "%s: num_parts %d out of [1,%d]", who, num_parts, SPP_MAX_PARTS);
  for (int p = 0; p < num_parts; ++p)
    if (offsets[p + 1] < offsets[p]) return fail(SPP_EINVAL, "%s: partition offsets not sorted", who);
  for (int p = 0; p <= SPP_MAX_PARTS; ++p) b.off[p] = offsets[p <= num_parts ? p : num_parts];
  b.num_parts = num_parts;
  b.rank = rank;
  return 0;
}

// exact searchsorted(off[0..P], nid, right=True) - 1 : -1 below off[0], P at or above off[P]
__device__ __forceinline__ int64_t book_partid_exact(const BookParams& b, int64_t nid) {
  int cnt = 0;
#pragma unroll
  for (int q = 0; q <= SPP_MAX_PARTS; ++q)
    if (q <= b.num_parts && b.off[q] <= nid) ++cnt;
  return (int64_t)cnt - 1;
}

__global__ void k_nid2partid(const __grid_constant__ BookParams b, const int64_t* __restrict__ nids, int64_t n,
                             int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = book_partid_exact(b, nids[i]);
}

__global__ void k_nid2localnid(int64_t off, const int64_t* __restrict__ nids, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = nids[i] - off;
}

__global__ void k_nid_is_local(int64_t lo, int64_t hi, const int64_t* __restrict__ nids, int64_t n,
                               uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = nids[i];
    out[i] = (v >= lo && v < hi) ? 1 : 0;
  }
}

__global__ void k_fill_i32(int32_t* __restrict__ p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_cache_scatter(const int64_t* __restrict__ cached, int64_t n, int32_t* __restrict__ map,
                                int64_t num_nodes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = cached[i];
    if (v >= 0 && v < num_nodes) atomicMax(map + v, (int32_t)i);
  }
}

__global__ void k_nid_is_cached(const int32_t* __restrict__ map, const int64_t* __restrict__ nids, int64_t n,
                                uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(map + nids[i]) >= 0 ? 1 : 0;
}

__global__ void k_nid2cachenid(const int32_t* __restrict__ map, const int64_t* __restrict__ nids, int64_t n,
                               int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)__ldg(map + nids[i]);
}

// ------------------------------------------------------------------------------------------------
// K3 split: stable multi-way partition into P+1 classes
// ------------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 256;
constexpr int kSplitRounds = 8;
constexpr int kSplitTile = kSplitThreads * kSplitRounds;  // 2048
constexpr int kClasses = SPP_MAX_PARTS + 1;               // 17

struct SplitParams {
  BookParams book;
  const void* n_id;
  const int64_t* n_dev;
  int64_t n_max;
  const int32_t* cache_map;  // NULL when the cache is not used
  int64_t* bucket_ids;
  int64_t* perm;
  int64_t* bucket_counts;
  uint32_t* tile_hist;   // [tiles_max][kClasses] -> exclusive per-class prefix over tiles
  uint32_t* class_start; // [kClasses + 1]
  uint8_t* cls;          // [n_max]
  int64_t tiles_max;
};

__device__ __forceinline__ int64_t split_n(const SplitParams& p) {
  int64_t n = p.n_max;
  if (p.n_dev) {
    int64_t nd = *p.n_dev;
    n = nd < n ? nd : n;
  }
  return n;
}

template <typename IdxT>
__global__ void __launch_bounds__(kSplitThreads) k_split_hist(const __grid_constant__ SplitParams prm) {
  __shared__ uint32_t s_hist[kClasses];
  const int64_t n = split_n(prm);
  const IdxT* __restrict__ ids = reinterpret_cast<const IdxT*>(prm.n_id);
  const int P = prm.book.num_parts;
  for (int64_t tile = blockIdx.x; tile < prm.tiles_max; tile += gridDim.x) {
    if (threadIdx.x < kClasses) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * kSplitTile;
#pragma unroll
    for (int r = 0; r < kSplitRounds; ++r) {
      const int64_t i = base + r * kSplitThreads + threadIdx.x;
      if (i < n) {
        const int64_t nid = (int64_t)ids[i];
        int c = book_partid(prm.book, nid);
        if (c != prm.book.rank && prm.cache_map != nullptr && __ldg(prm.cache_map + nid) >= 0) c = P;
        prm.cls[i] = (uint8_t)c;
        atomicAdd(&s_hist[c], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < kClasses) prm.tile_hist[tile * kClasses + threadIdx.x] = s_hist[threadIdx.x];
    __syncthreads();
  }
}

// one CTA; warp c scans class c over the tiles
__global__ void __launch_bounds__(32 * kClasses) k_split_scan(const __grid_constant__ SplitParams prm) {
  __shared__ uint32_t s_total[kClasses];
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int64_t n = split_n(prm);
  const int64_t tiles = (n + kSplitTile - 1) / kSplitTile;
  uint32_t run = 0;
  for (int64_t t0 = 0; t0 < tiles; t0 += 32) {
    const int64_t t = t0 + lane;
    uint32_t v = t < tiles ? prm.tile_hist[t * kClasses + c] : 0u;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (t < tiles) prm.tile_hist[t * kClasses + c] = run + inc - v;
    run += __shfl_sync(kFullMask, inc, 31);
  }
  if (lane == 0) s_total[c] = run;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    const int P = prm.book.num_parts;
    for (int q = 0; q < kClasses; ++q) {
      prm.class_start[q] = acc;
      if (q <= P) prm.bucket_counts[q] = (int64_t)s_total[q];
      acc += s_total[q];
    }
    prm.bucket_counts[P + 1] = n;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(kSplitThreads) k_split_scatter(const __grid_constant__ SplitParams prm) {
  __shared__ uint32_t s_warpcnt[kSplitThreads / 32][kClasses];
  __shared__ uint32_t s_run[kClasses];
  const int64_t n = split_n(prm);
  const int64_t tiles = (n + kSplitTile - 1) / kSplitTile;
  const IdxT* __restrict__ ids = reinterpret_cast<const IdxT*>(prm.n_id);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = prm.book.num_parts;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < kClasses)
      s_run[threadIdx.x] = prm.class_start[threadIdx.x] + prm.tile_hist[tile * kClasses + threadIdx.x];
    const int64_t base = tile * kSplitTile;
    for (int r = 0; r < kSplitRounds; ++r) {
      for (int q = threadIdx.x; q < (kSplitThreads / 32) * kClasses; q += kSplitThreads) (&s_warpcnt[0][0])[q] = 0;
      __syncthreads();
      const int64_t i = base + r * kSplitThreads + threadIdx.x;
      const bool valid = i < n;
      const int c = valid ? (int)prm.cls[i] : (kClasses + lane);  // invalid lanes never match each other
      const uint32_t peers = __match_any_sync(kFullMask, c);
      const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank_in_warp == 0) s_warpcnt[warp][c] = __popc(peers);
      __syncthreads();
      if (valid) {
        uint32_t pos = s_run[c] + rank_in_warp;
        for (int w = 0; w < warp; ++w) pos += s_warpcnt[w][c];
        const int64_t nid = (int64_t)ids[i];
        prm.bucket_ids[pos] = (c == P) ? (int64_t)__ldg(prm.cache_map + nid) : nid;
        prm.perm[i] = (int64_t)pos;
      }
      __syncthreads();
      if (threadIdx.x < kClasses) {
        uint32_t add = 0;
        for (int w = 0; w < kSplitThreads / 32; ++w) add += s_warpcnt[w][threadIdx.x];
        s_run[threadIdx.x] += add;
      }
      __syncthreads();
    }
  }
}

}  // namespace spp

extern "C" {

int spp_nid2partid(const int64_t* offsets, int num_parts, const int64_t* nids, int64_t n, int64_t* out, void* stream) {
  using namespace spp;
  BookParams b;
  if (int r = fill_book(b, offsets, num_parts, 0, "spp_nid2partid")) return r;
  if (n <= 0) return 0;
  if (!nids || !out) return fail(SPP_EINVAL, "spp_nid2partid: null pointer");
  k_nid2partid<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(b, nids, n, out);
  SPP_KERNEL_CHECK("k_nid2partid");
  return 0;
}

int spp_nid2localnid(const int64_t* offsets, int num_parts, int partition_idx, const int64_t* nids, int64_t n,
                     int64_t* out, void* stream) {
  using namespace spp;
  BookParams b;
  if (int r = fill_book(b, offsets, num_parts, 0, "spp_nid2localnid")) return r;
  if (partition_idx < 0 || partition_idx > num_parts) return fail(SPP_EINVAL, "spp_nid2localnid: partition_idx %d out of range", partition_idx);
  if (n <= 0) return 0;
  if (!nids || !out) return fail(SPP_EINVAL, "spp_nid2localnid: null pointer");
  k_nid2localnid<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(offsets[partition_idx], nids, n, out);
  SPP_KERNEL_CHECK("k_nid2localnid");
  return 0;
}

int spp_nid_is_local(const int64_t* offsets, int num_parts, int rank, const int64_t* nids, int64_t n, uint8_t* out,
                     void* stream) {
  using namespace spp;
  BookParams b;
  if (int r = fill_book(b, offsets, num_parts, rank, "spp_nid_is_local")) return r;
  if (rank < 0 || rank >= num_parts) return fail(SPP_EINVAL, "spp_nid_is_local: rank %d out of range", rank);
  if (n <= 0) return 0;
  if (!nids || !out) return fail(SPP_EINVAL, "spp_nid_is_local: null pointer");
  k_nid_is_local<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(offsets[rank], offsets[rank + 1], nids, n, out);
  SPP_KERNEL_CHECK("k_nid_is_local");
  return 0;
}

int spp_cache_build_map(const int64_t* cached_vertices, int64_t n_cached, int32_t* cache_map, int64_t num_nodes,
                        void* stream) {
  using namespace spp;
  if (num_nodes < 0 || n_cached < 0 || n_cached > 0x7fffffffll) return fail(SPP_EINVAL, "spp_cache_build_map: bad sizes");
  if (num_nodes == 0) return 0;
  if (!cache_map || (n_cached > 0 && !cached_vertices)) return fail(SPP_EINVAL, "spp_cache_build_map: null pointer");
  k_fill_i32<<<ew_grid(num_nodes), kEwThreads, 0, (cudaStream_t)stream>>>(cache_map, num_nodes, -1);
  SPP_KERNEL_CHECK("k_fill_i32");
  if (n_cached > 0) {
    k_cache_scatter<<<ew_grid(n_cached), kEwThreads, 0, (cudaStream_t)stream>>>(cached_vertices, n_cached, cache_map, num_nodes);
    SPP_KERNEL_CHECK("k_cache_scatter");
  }
  return 0;
}

int spp_nid_is_cached(const int32_t* cache_map, const int64_t* nids, int64_t n, uint8_t* out, void* stream) {
  using namespace spp;
  if (n <= 0) return 0;
  if (!cache_map || !nids || !out) return fail(SPP_EINVAL, "spp_nid_is_cached: null pointer");
  k_nid_is_cached<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(cache_map, nids, n, out);
  SPP_KERNEL_CHECK("k_nid_is_cached");
  return 0;
}

int spp_nid2cachenid(const int32_t* cache_map, const int64_t* nids, int64_t n, int64_t* out, void* stream) {
  using namespace spp;
  if (n <= 0) return 0;
  if (!cache_map || !nids || !out) return fail(SPP_EINVAL, "spp_nid2cachenid: null pointer");
  k_nid2cachenid<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(cache_map, nids, n, out);
  SPP_KERNEL_CHECK("k_nid2cachenid");
  return 0;
}

int64_t spp_split_scratch_words(int64_t n_max) {
  using namespace spp;
  if (n_max < 0) n_max = 0;
  const int64_t tiles = ceil_div(n_max > 0 ? n_max : 1, kSplitTile);
  return tiles * kClasses + (kClasses + 15) + ceil_div(n_max, 4) + 4;
}

int spp_split_by_owner(const spp_feature_map* m, int use_cache, const void* n_id, int idx_is_64, int64_t n_max,
                       const int64_t* n_dev, int64_t* bucket_ids, int64_t* perm, int64_t* bucket_counts,
                       int32_t* scratch, void* stream) {
  using namespace spp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!m) return fail(SPP_EINVAL, "spp_split_by_owner: null feature map");
  SplitParams prm{};
  if (int r = fill_book(prm.book, m->offsets, m->num_parts, m->rank, "spp_split_by_owner")) return r;
  if (m->rank < 0 || m->rank >= m->num_parts) return fail(SPP_EINVAL, "spp_split_by_owner: rank %d out of range", m->rank);
  if (!bucket_counts || !scratch) return fail(SPP_EINVAL, "spp_split_by_owner: null pointer");
  if (use_cache && !m->cache_map) return fail(SPP_EINVAL, "spp_split_by_owner: use_cache without a cache_map");
  if (n_max < 0) return fail(SPP_EINVAL, "spp_split_by_owner: negative n_max");
  if (n_max > 0 && (!n_id || !bucket_ids || !perm)) return fail(SPP_EINVAL, "spp_split_by_owner: null pointer");
  if (n_max >= (1ll << 32)) return fail(SPP_EUNSUPPORTED, "spp_split_by_owner: n_max too large");
  prm.n_id = n_id;
  prm.n_dev = n_dev;
  prm.n_max = n_max;
  prm.cache_map = use_cache ? m->cache_map : nullptr;
  prm.bucket_ids = bucket_ids;
  prm.perm = perm;
  prm.bucket_counts = bucket_counts;
  prm.tiles_max = ceil_div(n_max > 0 ? n_max : 1, kSplitTile);
  prm.tile_hist = reinterpret_cast<uint32_t*>(scratch);
  prm.class_start = prm.tile_hist + prm.tiles_max * kClasses;
  prm.cls = reinterpret_cast<uint8_t*>(prm.class_start + kClasses + 15);
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(prm.tiles_max < cap ? prm.tiles_max : cap);
  if (idx_is_64) k_split_hist<int64_t><<<grid, kSplitThreads, 0, st>>>(prm);
  else k_split_hist<int32_t><<<grid, kSplitThreads, 0, st>>>(prm);
  SPP_KERNEL_CHECK("k_split_hist");
  k_split_scan<<<1, 32 * kClasses, 0, st>>>(prm);
  SPP_KERNEL_CHECK("k_split_scan");
  if (idx_is_64) k_split_scatter<int64_t><<<grid, kSplitThreads, 0, st>>>(prm);
  else k_split_scatter<int32_t><<<grid, kSplitThreads, 0, st>>>(prm);
  SPP_KERNEL_CHECK("k_split_scatter");
  return 0;
}

}  // extern "C"
