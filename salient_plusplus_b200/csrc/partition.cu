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
  b.local_mask = (rank >= 0 && rank < num_parts) ? (1u << rank) : 0u;
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

// ---- cache index construction (set-up time) ----------------------------------------------------
// index buffer: [blocks: nblocks x 32 B][rank2row: int32 x n_cached][scan scratch: uint32 x (nblocks/256 + 1)]
constexpr int kIdxThreads = 256;

__global__ void k_cache_setbits(const int64_t* __restrict__ cached, int64_t n, uint32_t* __restrict__ blocks, int64_t num_nodes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = cached[i];
    if (v >= 0 && v < num_nodes) {
      const uint32_t u = (uint32_t)v, blk = u / kCacheBlockIds, bit = u - blk * kCacheBlockIds;
      atomicOr(blocks + 8 * (size_t)blk + 1 + (bit >> 5), 1u << (bit & 31u));
    }
  }
}

// per CTA of 256 blocks: popcount of every block -> exclusive scan inside the CTA -> word 0; CTA total -> sums
__global__ void __launch_bounds__(kIdxThreads) k_cache_block_counts(uint32_t* __restrict__ blocks, int64_t nblocks,
                                                                     uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_w[kIdxThreads / 32];
  const int64_t b = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
  uint32_t cnt = 0;
  if (b < nblocks) {
#pragma unroll
    for (int q = 1; q < 8; ++q) cnt += __popc(blocks[8 * (size_t)b + q]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(cnt, lane);
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kIdxThreads / 32; ++w) {
    if (w < warp) base += s_w[w];
    total += s_w[w];
  }
  if (b < nblocks) blocks[8 * (size_t)b] = base + inc - cnt;
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of the per-CTA totals (<= a few thousand entries)
__global__ void __launch_bounds__(1024) k_cache_scan_sums(uint32_t* __restrict__ sums, int64_t n) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int64_t i0 = 0; i0 < n; i0 += 1024) {
    const int64_t i = i0 + threadIdx.x;
    const uint32_t v = i < n ? sums[i] : 0u;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t base = s_run, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) base += s_w[w];
      total += s_w[w];
    }
    if (i < n) sums[i] = base + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) s_run += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kIdxThreads) k_cache_add_base(uint32_t* __restrict__ blocks, int64_t nblocks,
                                                                 const uint32_t* __restrict__ sums) {
  const int64_t b = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
  if (b < nblocks) blocks[8 * (size_t)b] += sums[blockIdx.x];
}

// rank2row[rank(v)] = largest i with cached[i] == v (the reference's sequential overwrite keeps the
// last duplicate, range_partition_book.cpp:154-158)
__global__ void k_cache_rank2row(const int64_t* __restrict__ cached, int64_t n, CacheIndex idx, int32_t* __restrict__ rank2row) {
  const uint64_t pol = l2_policy_evict_last();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = cache_rank(idx, cached[i], pol);
    if (r >= 0) atomicMax(rank2row + r, (int32_t)i);
  }
}

__global__ void k_nid_is_cached(CacheIndex idx, const int64_t* __restrict__ nids, int64_t n, uint8_t* __restrict__ out) {
  const uint64_t pol = l2_policy_evict_last();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = cache_rank(idx, nids[i], pol) >= 0 ? 1 : 0;
}

__global__ void k_nid2cachenid(CacheIndex idx, const int64_t* __restrict__ nids, int64_t n, int64_t* __restrict__ out) {
  const uint64_t pol = l2_policy_evict_last();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)cache_lookup(idx, nids[i], pol);
}

// ------------------------------------------------------------------------------------------------
// K3 split: stable multi-way partition into P+1 classes
// ------------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 256;
constexpr int kSplitRounds = 8;
constexpr int kSplitTile = kSplitThreads * kSplitRounds;  // 2048
constexpr int kClasses = SPP_MAX_PARTS + 1;               // 17
static_assert(kSplitTile == kSplitTileRows && kClasses == kSplitClasses, "scratch layout shared with gather.cu");

struct SplitParams {
  BookParams book;
  const void* n_id;
  const int64_t* n_dev;
  int64_t n_max;
  CacheIndex cache;  // cache.nodes == 0 when the cache is not used
  int64_t* bucket_ids;
  int64_t* perm;
  int64_t* bucket_counts;
  uint32_t* tile_hist;   // [tiles_max][kClasses] -> exclusive per-class prefix over tiles
  uint32_t* class_start; // [kClasses + 1]
  int32_t* desc;         // [n_max] per-node source descriptor: p >= 0 -> partition p, < 0 -> ~cache row
  int32_t* inv;          // [n_max] inverse of perm: inv[pos] = i (the gather-by-class kernels walk buckets)
  int64_t tiles_max;
  const spp_device_job* job;  // graph replay: bucket_ids / perm come from the device job block
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
  const uint64_t pol = l2_policy_evict_last();
  for (int64_t tile = blockIdx.x; tile < prm.tiles_max; tile += gridDim.x) {
    if (threadIdx.x < kClasses) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * kSplitTile;
    // phased so that the 8 rounds' loads are independent: ids -> index blocks -> rank2row -> histogram
    // (one round at a time this kernel was bound by the latency of the two dependent probe loads)
    int64_t nid[kSplitRounds];
    int cls[kSplitRounds];
    int32_t rank[kSplitRounds];
#pragma unroll
    for (int r = 0; r < kSplitRounds; ++r) {
      const int64_t i = base + r * kSplitThreads + threadIdx.x;
      nid[r] = i < n ? (int64_t)ids[i] : -1;
    }
#pragma unroll
    for (int r = 0; r < kSplitRounds; ++r) {
      cls[r] = nid[r] >= 0 ? book_partid(prm.book, nid[r]) : -1;
      rank[r] = -1;
      // the ONE cache probe of this node: the scatter below and the fused gather read `desc`
      if (cls[r] >= 0 && !book_is_local(prm.book, cls[r]) && prm.cache.nodes > 0) rank[r] = cache_rank(prm.cache, nid[r], pol);
    }
#pragma unroll
    for (int r = 0; r < kSplitRounds; ++r)
      if (rank[r] >= 0) rank[r] = (int32_t)ld_l2hint(reinterpret_cast<const uint32_t*>(prm.cache.rank2row) + rank[r], pol);
#pragma unroll
    for (int r = 0; r < kSplitRounds; ++r) {
      const int64_t i = base + r * kSplitThreads + threadIdx.x;
      if (cls[r] >= 0) {
        int c = cls[r];
        int32_t d = c;
        if (rank[r] >= 0) {
          c = P;
          d = ~rank[r];
        }
        prm.desc[i] = d;
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
    prm.class_start[kClasses] = acc;
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
  int64_t* const bucket_ids = prm.job ? prm.job->bucket_ids : prm.bucket_ids;
  int64_t* const perm = prm.job ? prm.job->perm : prm.perm;
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
      const int32_t d = valid ? prm.desc[i] : 0;
      const int c = valid ? (d < 0 ? P : (int)d) : (kClasses + lane);  // invalid lanes never match each other
      const uint32_t peers = __match_any_sync(kFullMask, c);
      const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank_in_warp == 0) s_warpcnt[warp][c] = __popc(peers);
      __syncthreads();
      if (valid) {
        uint32_t pos = s_run[c] + rank_in_warp;
        for (int w = 0; w < warp; ++w) pos += s_warpcnt[w][c];
        bucket_ids[pos] = (d < 0) ? (int64_t)(~d) : (int64_t)ids[i];
        perm[i] = (int64_t)pos;
        prm.inv[pos] = (int32_t)i;
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

int64_t spp_cache_index_bytes(int64_t num_nodes, int64_t n_cached) {
  using namespace spp;
  if (num_nodes < 0) num_nodes = 0;
  if (n_cached < 0) n_cached = 0;
  const int64_t nblocks = cache_index_blocks(num_nodes);
  return nblocks * 32 + (n_cached + 1) * 4 + (ceil_div(nblocks > 0 ? nblocks : 1, kIdxThreads) + 1) * 4;
}

int spp_cache_build_index(const int64_t* cached_vertices, int64_t n_cached, int64_t num_nodes, void* index, void* stream) {
  using namespace spp;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_nodes < 0 || num_nodes > (1ll << 31) || n_cached < 0 || n_cached > 0x7fffffffll)
    return fail(SPP_EINVAL, "spp_cache_build_index: bad sizes");
  if (num_nodes == 0) return 0;
  if (!index || (n_cached > 0 && !cached_vertices)) return fail(SPP_EINVAL, "spp_cache_build_index: null pointer");
  const int64_t nblocks = cache_index_blocks(num_nodes);
  uint32_t* blocks = reinterpret_cast<uint32_t*>(index);
  int32_t* rank2row = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(index) + nblocks * 32);
  uint32_t* sums = reinterpret_cast<uint32_t*>(rank2row + n_cached + 1);
  SPP_CUDA(cudaMemsetAsync(blocks, 0, (size_t)nblocks * 32, st));
  SPP_CUDA(cudaMemsetAsync(rank2row, 0xFF, (size_t)(n_cached + 1) * 4, st));
  if (n_cached == 0) return 0;
  k_cache_setbits<<<ew_grid(n_cached), kEwThreads, 0, st>>>(cached_vertices, n_cached, blocks, num_nodes);
  SPP_KERNEL_CHECK("k_cache_setbits");
  const int ctas = (int)ceil_div(nblocks, kIdxThreads);
  k_cache_block_counts<<<ctas, kIdxThreads, 0, st>>>(blocks, nblocks, sums);
  SPP_KERNEL_CHECK("k_cache_block_counts");
  k_cache_scan_sums<<<1, 1024, 0, st>>>(sums, ctas);
  SPP_KERNEL_CHECK("k_cache_scan_sums");
  k_cache_add_base<<<ctas, kIdxThreads, 0, st>>>(blocks, nblocks, sums);
  SPP_KERNEL_CHECK("k_cache_add_base");
  k_cache_rank2row<<<ew_grid(n_cached), kEwThreads, 0, st>>>(cached_vertices, n_cached, make_cache_index(index, num_nodes), rank2row);
  SPP_KERNEL_CHECK("k_cache_rank2row");
  return 0;
}

int spp_nid_is_cached(const void* index, int64_t num_nodes, const int64_t* nids, int64_t n, uint8_t* out, void* stream) {
  using namespace spp;
  if (n <= 0) return 0;
  if (!index || !nids || !out) return fail(SPP_EINVAL, "spp_nid_is_cached: null pointer");
  k_nid_is_cached<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(make_cache_index(index, num_nodes), nids, n, out);
  SPP_KERNEL_CHECK("k_nid_is_cached");
  return 0;
}

int spp_nid2cachenid(const void* index, int64_t num_nodes, const int64_t* nids, int64_t n, int64_t* out, void* stream) {
  using namespace spp;
  if (n <= 0) return 0;
  if (!index || !nids || !out) return fail(SPP_EINVAL, "spp_nid2cachenid: null pointer");
  k_nid2cachenid<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(make_cache_index(index, num_nodes), nids, n, out);
  SPP_KERNEL_CHECK("k_nid2cachenid");
  return 0;
}

int64_t spp_split_scratch_words(int64_t n_max) {
  using namespace spp;
  if (n_max < 0) n_max = 0;
  const int64_t tiles = ceil_div(n_max > 0 ? n_max : 1, kSplitTile);
  return 2 * n_max + tiles * kClasses + (kClasses + 15) + 4;  // [desc n_max][inv n_max][tile_hist][class_start]
}

int spp_split_by_owner(const spp_feature_map* m, int use_cache, const void* n_id, int idx_is_64, int64_t n_max,
                       const int64_t* n_dev, int64_t* bucket_ids, int64_t* perm, int64_t* bucket_counts,
                       int32_t* scratch, void* stream) {
  return spp::split_by_owner_job(m, use_cache, n_id, idx_is_64, n_max, n_dev, bucket_ids, perm, bucket_counts, scratch,
                                 (cudaStream_t)stream, nullptr);
}

}  // extern "C"

namespace spp {

int split_by_owner_job(const spp_feature_map* m, int use_cache, const void* n_id, int idx_is_64, int64_t n_max,
                       const int64_t* n_dev, int64_t* bucket_ids, int64_t* perm, int64_t* bucket_counts, int32_t* scratch,
                       cudaStream_t st, const spp_device_job* job) {
  if (!m) return fail(SPP_EINVAL, "spp_split_by_owner: null feature map");
  SplitParams prm{};
  if (int r = fill_book(prm.book, m->offsets, m->num_parts, m->rank, "spp_split_by_owner")) return r;
  if (m->rank < 0 || m->rank >= m->num_parts) return fail(SPP_EINVAL, "spp_split_by_owner: rank %d out of range", m->rank);
  if (!bucket_counts || !scratch) return fail(SPP_EINVAL, "spp_split_by_owner: null pointer");
  if (use_cache && !m->cache_index) return fail(SPP_EINVAL, "spp_split_by_owner: use_cache without a cache index");
  if (m->local_parts) prm.book.local_mask = m->local_parts | (1u << m->rank);
  if (n_max < 0) return fail(SPP_EINVAL, "spp_split_by_owner: negative n_max");
  if (n_max > 0 && (!n_id || (!job && (!bucket_ids || !perm)))) return fail(SPP_EINVAL, "spp_split_by_owner: null pointer");
  if (n_max >= (1ll << 32)) return fail(SPP_EUNSUPPORTED, "spp_split_by_owner: n_max too large");
  prm.n_id = n_id;
  prm.n_dev = n_dev;
  prm.n_max = n_max;
  prm.cache = make_cache_index(use_cache ? m->cache_index : nullptr, m->cache_index_nodes);
  prm.bucket_ids = bucket_ids;
  prm.perm = perm;
  prm.job = job;
  prm.bucket_counts = bucket_counts;
  prm.tiles_max = ceil_div(n_max > 0 ? n_max : 1, kSplitTile);
  prm.desc = scratch;
  prm.inv = scratch + n_max;
  prm.tile_hist = reinterpret_cast<uint32_t*>(scratch + 2 * n_max);
  prm.class_start = prm.tile_hist + prm.tiles_max * kClasses;
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

}  // namespace spp
