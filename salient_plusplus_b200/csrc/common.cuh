// common.cuh -- shared host/device helpers of libsalient_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "salient_b200.h"

namespace spp {

// ---- host side -------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
void count_replay();
void count_capture();
int num_sms();

// diagnostics (runtime.cu): event mark on `st` after an operation was issued; no-op unless a
// trace is active (spp_trace_begin)
enum TraceLabel {
  kTrBatchBegin = 0, kTrSeedsH2D, kTrTableClear, kTrSeedsInit, kTrSample, kTrCompact, kTrRelabel,
  kTrExport, kTrSplit, kTrGather, kTrLabels, kTrMetaD2H, kTrJoin, kTrCount, kTrSortLarge, kTrSortBitmap
};
void trace_mark(int label, int hop, cudaStream_t st);

// Side streams of a mini-batch's main stream (runtime.cu; created on first use, one set per main
// stream): the relabel/sort kernels of the sampled hops and, optionally, the feature gather are
// forked onto them so that they stay off the sampler's dependent kernel chain.
struct AuxStreams {
  cudaStream_t relabel, gather;
  cudaEvent_t fork[SPP_MAX_HOPS], fork_gather, join_relabel, join_gather;
};
AuxStreams* aux_streams(cudaStream_t main);
// bit 0: relabel/sort kernels on the side stream; bit 1: feature + label gather on a side stream
// (SPP_FORK overrides the default)
int pipeline_flags();

// Tunables (runtime.cu): defaults come from the SPP_* environment variables of the same meaning, read
// once; spp_tune() changes them at run time (A/B tools, tests).  -1 = "not set".
struct Tunables {
  int gather_ctas_per_sm;  // SPP_GATHER_CTAS_PER_SM   0 = automatic
  int gather_bulk;         // SPP_GATHER_BULK          -1 automatic, 0 never, 1 whenever rows allow it
  int bulk_tile;           // SPP_BULK_TILE            bytes per tile of the bulk-copy gather
  int bulk_stages;         // SPP_BULK_STAGES
  int bulk_ctas_per_sm;    // SPP_BULK_CTAS_PER_SM
  int gather_split;        // SPP_GATHER_SPLIT         1: peer rows fetched by their own launch on a side stream
  int gather_tile_rows;    // SPP_GATHER_TILE_ROWS     0 = automatic (64; 128 when rows may come from peer GPUs)
  int gather_l2_hint;      // SPP_GATHER_L2HINT        1: rows of a partitioned gather stream through L2 as evict_first
                           //                          (keeps the evict_last cache index resident)
};
Tunables& tunables();

// true when `p` is a pointer spp_ipc_import returned (a peer GPU's memory mapped into this process)
bool ipc_imported(const void* p);

// sampler.cu
int sample_minibatch_impl(const spp_graph* g, const int64_t* seeds, int64_t batch_size, const int32_t* sizes,
                          int n_hops, int replace, uint64_t rng_seed, const spp_sampler_ws* ws,
                          int64_t* const* out_rowptr, int64_t* const* out_col, const int64_t* out_col_cap,
                          int64_t* n_id_out, cudaStream_t st, bool* pending, const spp_device_job* job, bool want_nid);
int sorter_attributes();
uint64_t next_scan_epoch();
int gather_attributes();
// gather.cu / partition.cu: the same entry points with the per-batch outputs taken from a device job block
int gather_rows_job(const void* table, int64_t table_pitch, int64_t row_bytes, const void* idx, int idx_is_64, int64_t n_idx,
                    const int64_t* n_idx_dev, void* out, int64_t n_out_rows, cudaStream_t st, const spp_device_job* job,
                    int job_mode);
int gather_partitioned_job(const spp_feature_map* m, int64_t row_bytes, const void* n_id, int idx_is_64, int64_t n_idx,
                           const int64_t* n_idx_dev, const int32_t* src_desc, void* out, int64_t n_out_rows,
                           int64_t* counters, cudaStream_t st, const spp_device_job* job);
int split_by_owner_job(const spp_feature_map* m, int use_cache, const void* n_id, int idx_is_64, int64_t n_max,
                       const int64_t* n_dev, int64_t* bucket_ids, int64_t* perm, int64_t* bucket_counts, int32_t* scratch,
                       cudaStream_t st, const spp_device_job* job);
bool tracing_active();
int join_relabel(cudaStream_t st, bool pending);

#define SPP_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return spp::cuda_fail(_e, #expr);  \
  } while (0)

#define SPP_KERNEL_CHECK(name)                                   \
  do {                                                           \
    spp::count_launch();                                         \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return spp::cuda_fail(_e, name);      \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Layout of spp_split_by_owner's scratch (partition.cu), shared with the gather-by-class kernel:
//   [desc: n_max][inv: n_max][tile_hist: tiles * kSplitClasses][class_start: kSplitClasses + 1 ...]
constexpr int kSplitTileRows = 2048;
constexpr int kSplitClasses = SPP_MAX_PARTS + 1;
static inline const int32_t* split_scratch_inv(const int32_t* scratch, int64_t n_max) { return scratch + n_max; }
static inline const uint32_t* split_scratch_class_start(const int32_t* scratch, int64_t n_max) {
  const int64_t tiles = ceil_div(n_max > 0 ? n_max : 1, kSplitTileRows);
  return reinterpret_cast<const uint32_t*>(scratch + 2 * n_max) + tiles * kSplitClasses;
}
int gather_by_class_job(const spp_feature_map* m, int64_t row_bytes, const int64_t* bucket_ids, const int32_t* split_scratch,
                        int64_t n_max, uint32_t class_mask, void* out, int64_t* counters, cudaStream_t st,
                        const spp_device_job* job);

// ---- device side -----------------------------------------------------------------------------
constexpr uint32_t kFullMask = 0xffffffffu;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

// Counter-based generator: specification shared with oracle/salient_oracle.c:spo_rand64.
__host__ __device__ __forceinline__ uint64_t premix_seed(uint64_t seed) {
  return mix64(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull);
}
__device__ __forceinline__ uint64_t rand64(uint64_t premixed, uint32_t hop, uint64_t target_pos,
                                           uint32_t pick) {
  uint64_t ctr = ((uint64_t)hop << 56) ^ (target_pos << 8) ^ (uint64_t)pick;
  return mix64(premixed ^ ctr);
}
// uniform integer in [0, range)
__device__ __forceinline__ uint32_t bounded(uint64_t r, uint32_t range) {
  return (uint32_t)__umul64hi(r, (uint64_t)range);
}

// 128-bit streaming loads/stores that do not allocate in L1 (gathered rows are touched once)
__device__ __forceinline__ int4 ld_nc_na(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_nc_na(const int2* p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_nc_na(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ short ld_nc_na(const short* p) {
  short r;
  asm volatile("ld.global.nc.L1::no_allocate.s16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ char ld_nc_na(const char* p) { return *p; }

__device__ __forceinline__ void st_na(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_na(int2* p, const int2& v) {
  asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_na(int* p, const int& v) {
  asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_na(short* p, const short& v) { *p = v; }
__device__ __forceinline__ void st_na(char* p, const char& v) { *p = v; }

// Flag/value handshakes between CTAs of one grid: gpu-scope relaxed accesses served by L2
// (ld.volatile / st.volatile compile to STRONG.SYS accesses, which are not needed here).
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
  return v;
}
// inclusive scan across the 32 lanes of a warp
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(kFullMask, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// ---- L2 eviction-priority hints ---------------------------------------------------------------
// The 126 MB L2 is shared by streams of feature rows that are touched once (hundreds of MB per
// mini-batch) and by small random-access structures that every mini-batch probes again (the cache
// index below, 16-35 MB).  Rows are streamed with evict_first, the index is read with evict_last,
// so the index stays resident instead of costing a DRAM line per probe.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_l2hint(const uint4* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ld_l2hint(const uint32_t* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}

// streaming row loads / stores that are also marked evict_first in L2 (policy from l2_policy_evict_first)
__device__ __forceinline__ int4 ld_stream(const int4* p, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int2 ld_stream(const int2* p, uint64_t pol) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int ld_stream(const int* p, uint64_t) { return ld_nc_na(p); }
__device__ __forceinline__ short ld_stream(const short* p, uint64_t) { return ld_nc_na(p); }
__device__ __forceinline__ char ld_stream(const char* p, uint64_t) { return ld_nc_na(p); }
__device__ __forceinline__ void st_stream(int4* p, const int4& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_stream(int2* p, const int2& v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.s32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_stream(int* p, const int& v, uint64_t) { st_na(p, v); }
__device__ __forceinline__ void st_stream(short* p, const short& v, uint64_t) { st_na(p, v); }
__device__ __forceinline__ void st_stream(char* p, const char& v, uint64_t) { st_na(p, v); }

// ---- cache index (replaces the reference's dense id -> cache-row arrays,
//      fast_sampler/range_partition_book.cpp:152-158) ------------------------------------------
// One 32-byte block (= one L2 sector) per 224 node ids: word 0 = number of cached ids in all
// earlier blocks, words 1..7 = membership bits.  A probe is ONE sector read: membership, and for a
// member its rank among the cached ids in ascending id order; rank2row[rank] is the row of the
// reference's cached_features.  N/7 bytes (papers100M: 15.9 MB, MAG240M: 34.9 MB) instead of a
// 4N-byte dense map (444 / 976 MB) whose random probes each cost a DRAM access.
constexpr uint32_t kCacheBlockIds = 224;
struct CacheIndex {
  const uint4* blocks;      // [ceil(nodes / 224)][2]
  const int32_t* rank2row;  // [number of distinct cached ids]
  int64_t nodes;            // ids >= nodes are not cached
};
__host__ __device__ __forceinline__ int64_t cache_index_blocks(int64_t nodes) {
  return (nodes + kCacheBlockIds - 1) / kCacheBlockIds;
}
__host__ __device__ __forceinline__ CacheIndex make_cache_index(const void* base, int64_t nodes) {
  CacheIndex c;
  c.blocks = reinterpret_cast<const uint4*>(base);
  c.rank2row = reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(base) + cache_index_blocks(nodes) * 32);
  c.nodes = base ? nodes : 0;
  return c;
}
// rank of `id` among the cached ids, or -1
__device__ __forceinline__ int32_t cache_rank(const CacheIndex& c, int64_t id, uint64_t pol) {
  if (id < 0 || id >= c.nodes) return -1;
  const uint32_t u = (uint32_t)id;
  const uint32_t blk = u / kCacheBlockIds, bit = u - blk * kCacheBlockIds;
  const uint4* b = c.blocks + 2 * (size_t)blk;
  const uint4 lo = ld_l2hint(b, pol), hi = ld_l2hint(b + 1, pol);
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  const uint32_t wi = (bit >> 5) + 1u, m = 1u << (bit & 31u);
  uint32_t rank = w[0];
  bool hit = false;
#pragma unroll
  for (uint32_t q = 1; q < 8; ++q) {
    if (q < wi) rank += __popc(w[q]);
    else if (q == wi) {
      hit = (w[q] & m) != 0u;
      rank += __popc(w[q] & (m - 1u));
    }
  }
  return hit ? (int32_t)rank : -1;
}
// row of cached_features holding `id`, or -1
__device__ __forceinline__ int32_t cache_lookup(const CacheIndex& c, int64_t id, uint64_t pol) {
  const int32_t r = cache_rank(c, id, pol);
  return r < 0 ? -1 : (int32_t)ld_l2hint(reinterpret_cast<const uint32_t*>(c.rank2row) + r, pol);
}

// Range partition book in kernel-parameter space (<= 17 offsets: a register-resident search)
struct BookParams {
  int64_t off[SPP_MAX_PARTS + 1];
  int num_parts;
  int rank;
  uint32_t local_mask;  // bit p: partition p is resident on this GPU (always includes `rank`)
};
__device__ __forceinline__ bool book_is_local(const BookParams& b, int p) { return (b.local_mask >> p) & 1u; }
// searchsorted(off, nid, right=True) - 1, clamped like the reference's use (ids inside [0, N))
__device__ __forceinline__ int book_partid(const BookParams& b, int64_t nid) {
  int p = 0;
#pragma unroll
  for (int q = 1; q < SPP_MAX_PARTS; ++q)
    if (q < b.num_parts && nid >= b.off[q]) p = q;
  return p;
}

}  // namespace spp
