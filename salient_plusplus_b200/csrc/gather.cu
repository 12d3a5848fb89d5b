// gather.cu -- K4 feature gather and K4+K5 fused partition-book / cache / peer-to-peer gather.
//
// HBM-bound byte movement: out[i,:] = table[idx[i],:].  One CTA owns a tile of kRows output rows:
// it first resolves the tile's source row pointers into shared memory (index load, and for the
// partitioned flavour the range-partition-book search + dense cache-map lookup + owner table
// selection), then all 256 threads stream the tile as a flat run of 16-byte chunks: the output
// side is perfectly coalesced (the tile is contiguous in `out`), the input side is coalesced
// inside each row.  kUnroll independent 128-bit loads are in flight per thread before the first
// store (L1::no_allocate on both sides: every byte is touched once).
//
// Algorithmic bytes per row: 2 * row_bytes + sizeof(index)   (DESIGN.md section 4).
#include <cstdlib>

#include "common.cuh"

namespace spp {

constexpr int kGatherThreads = 256;
constexpr int kRows = 64;   // rows per tile
constexpr int kUnroll = 8;  // independent vector loads in flight per thread

struct GatherParams {
  const char* table;     // single-table flavour
  const void* idx;       // int32 / int64 node or row ids
  const int64_t* n_dev;  // optional device-resident row count
  int64_t n_max;         // host-side bound (min(n_idx, n_out_rows))
  char* out;
  int64_t row_bytes;     // bytes copied per row (= output pitch)
  int64_t table_pitch;   // byte pitch of the source rows (>= row_bytes; padded tables)
  int64_t cache_pitch;
  uint32_t vpr;        // vectors per row
  uint32_t vpr_magic;  // ceil(2^32 / vpr): lc / vpr == umulhi(lc, magic) while lc * vpr < 2^32;
                       // 0 = use a real division (vpr == 1 or very wide rows)
  // partitioned flavour
  BookParams book;
  const char* tables[SPP_MAX_PARTS];
  const char* cache_table;
  CacheIndex cache;              // cache.nodes == 0: no cache
  const int32_t* desc;           // optional per-row source descriptors written by spp_split_by_owner
                                 // (p >= 0: partition p, < 0: ~cache row): no book search, no probe
  unsigned long long* counters;  // [3] local / cache / peer rows (optional)
  // graph replay: per-batch pointers come from the device job block (session.cu)
  const spp_device_job* job;
  int src_align16;               // every source base and pitch is a multiple of 16 bytes (bulk-copy eligibility)
  int l2_stream_hint;            // partitioned flavour: rows stream through L2 as evict_first
  int job_mode;                  // 1: out = job->x_out; 2 (labels): idx = job->seeds, out = job->y_out, n <= job->batch_size
};

struct GatherView {
  const void* idx;
  char* out;
  int64_t n;
};
__device__ __forceinline__ GatherView gather_view(const GatherParams& prm) {
  GatherView v{prm.idx, prm.out, prm.n_max};
  if (prm.n_dev != nullptr) {
    const int64_t nd = *prm.n_dev;
    v.n = nd < v.n ? nd : v.n;
  }
  if (prm.job != nullptr) {
    if (prm.job_mode == 1) {
      v.out = reinterpret_cast<char*>(prm.job->x_out);
    } else if (prm.job_mode == 2) {
      v.idx = prm.job->seeds;
      v.out = reinterpret_cast<char*>(prm.job->y_out);
      v.n = prm.job->batch_size < v.n ? prm.job->batch_size : v.n;
    }
  }
  return v;
}

// Resolution of one output row's source pointer, split so that its two dependent loads (the index
// and, for remote rows, the dense cache map) can be issued a tile ahead of their use.
template <bool kPartitioned>
struct RowResolver {
  int64_t id = 0;      // node / row id (valid when `on`)
  int p = 0;           // owner partition
  int32_t crow = -1;   // cache row (in flight until `finish`)
  bool on = false;

  __device__ __forceinline__ void begin_lookup(const GatherParams& prm, int64_t row, uint64_t pol) {
    if constexpr (kPartitioned) {
      crow = -1;
      if (on) {
        if (prm.desc != nullptr) {
          const int32_t d = __ldg(prm.desc + row);
          if (d < 0) {
            crow = ~d;
            p = -1;
          } else {
            p = d;
          }
        } else {
          p = book_partid(prm.book, id);
          if (!book_is_local(prm.book, p) && prm.cache.nodes > 0) crow = cache_lookup(prm.cache, id, pol);
        }
      }
    }
  }
  // returns the source pointer and the class (0 local, 1 cache, 2 peer, -1 none)
  __device__ __forceinline__ const char* finish(const GatherParams& prm, int& cls) const {
    cls = -1;
    if (!on) return nullptr;
    if constexpr (!kPartitioned) {
      return prm.table + id * prm.table_pitch;
    } else {
      if (crow >= 0) {
        cls = 1;
        return prm.cache_table + (int64_t)crow * prm.cache_pitch;
      }
      if (book_is_local(prm.book, p)) {
        cls = 0;
        return prm.tables[p] + (id - prm.book.off[p]) * prm.table_pitch;
      }
      cls = 2;
      return prm.tables[p] + (id - prm.book.off[p]) * prm.table_pitch;  // peer HBM over NVLink
    }
  }
};

// ROWS = rows per tile (64: 16 KB of 256-byte rows in flight per CTA; 128 for maps with peer tables, whose
// rows take an NVLink round trip: the tile must be deep enough to keep the link busy with 2 CTAs per SM)
template <typename V, bool kPartitioned, typename IdxT, int ROWS>
__global__ void __launch_bounds__(kGatherThreads, 4) k_gather(const __grid_constant__ GatherParams prm) {
  __shared__ const char* s_src[2][ROWS];
  const int tid = threadIdx.x;
  const GatherView gv = gather_view(prm);
  const int64_t n = gv.n;
  const int64_t num_tiles = (n + ROWS - 1) / ROWS;
  const IdxT* __restrict__ idx = reinterpret_cast<const IdxT*>(gv.idx);
  const uint32_t vpr = prm.vpr, magic = prm.vpr_magic;
  const bool resolver = tid < ROWS;  // warps 0 and 1
  unsigned long long cnt0 = 0, cnt1 = 0, cnt2 = 0;
  uint64_t pol = 0, pol_stream = 0;
  bool stream_hint = false;
  if constexpr (kPartitioned) {
    pol = l2_policy_evict_last();
    stream_hint = prm.l2_stream_hint != 0;
    if (stream_hint) pol_stream = l2_policy_evict_first();
  }

  auto load_id = [&](int64_t tile, RowResolver<kPartitioned>& r) {
    r.on = false;
    if (tile < num_tiles) {
      const int64_t row = tile * ROWS + tid;
      if (row < n) {
        r.id = (int64_t)idx[row];
        r.on = true;
      }
    }
  };
  auto publish = [&](const RowResolver<kPartitioned>& r, int buf) {
    int cls;
    const char* src = r.finish(prm, cls);
    s_src[buf][tid] = src;
    if constexpr (kPartitioned) {
      if (prm.counters != nullptr) {
        const uint32_t m0 = __ballot_sync(kFullMask, cls == 0);
        const uint32_t m1 = __ballot_sync(kFullMask, cls == 1);
        const uint32_t m2 = __ballot_sync(kFullMask, cls == 2);
        if ((tid & 31) == 0) {
          cnt0 += __popc(m0);
          cnt1 += __popc(m1);
          cnt2 += __popc(m2);
        }
      }
    }
  };

  // software pipeline over this CTA's tiles: ids of tile t+2 and the cache lookups of tile t+1
  // are in flight while tile t streams; one barrier per tile
  RowResolver<kPartitioned> r1, r2;
  int64_t tile = blockIdx.x;
  if (resolver) {
    RowResolver<kPartitioned> r0;
    load_id(tile, r0);
    load_id(tile + gridDim.x, r1);
    r0.begin_lookup(prm, tile * ROWS + tid, pol);
    publish(r0, 0);
  }
  __syncthreads();
  int buf = 0;
  for (; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * ROWS;
    const int rows = (int)((n - row0) < ROWS ? (n - row0) : ROWS);
    if (resolver) {
      r1.begin_lookup(prm, (tile + gridDim.x) * ROWS + tid, pol);  // tile t+1: id arrived a tile ago
      load_id(tile + 2 * (int64_t)gridDim.x, r2);  // tile t+2: id load in flight
    }
    const uint32_t chunks = (uint32_t)rows * vpr;
    V* __restrict__ dst = reinterpret_cast<V*>(gv.out + row0 * prm.row_bytes);
    const char* const* src = s_src[buf];
    for (uint32_t base = tid; base < chunks; base += kGatherThreads * kUnroll) {
      V vals[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t lc = base + u * kGatherThreads;
        if (lc < chunks) {
          const uint32_t r = magic ? __umulhi(lc, magic) : lc / vpr;
          const uint32_t v = lc - r * vpr;
          const V* q = reinterpret_cast<const V*>(src[r]) + v;
          vals[u] = stream_hint ? ld_stream(q, pol_stream) : ld_nc_na(q);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t lc = base + u * kGatherThreads;
        if (lc < chunks) {
          if (stream_hint) st_stream(dst + lc, vals[u], pol_stream);
          else st_na(dst + lc, vals[u]);
        }
      }
    }
    if (resolver) {
      publish(r1, buf ^ 1);
      r1 = r2;
    }
    __syncthreads();
    buf ^= 1;
  }
  if constexpr (kPartitioned) {
    if (prm.counters != nullptr && resolver && (tid & 31) == 0) {
      if (cnt0) atomicAdd(prm.counters + 0, cnt0);
      if (cnt1) atomicAdd(prm.counters + 1, cnt1);
      if (cnt2) atomicAdd(prm.counters + 2, cnt2);
    }
  }
}

// ================================================================================================
// Bulk-copy flavour for rows whose size and addresses are multiples of 16 bytes (128-d fp16 papers
// rows, 768-d MAG rows, fp32 arxiv rows): every row travels global -> shared as ONE asynchronous
// bulk copy (cp.async.bulk, the 1-D TMA path) and every tile leaves shared -> global as ONE bulk
// store.  No data passes through registers, a warp keeps kBulkStages - 1 tiles (~8 KB each) in
// flight with 32 threads, and the requests that reach the fabric are whole rows rather than
// 16-byte fragments -- which is what the NVLink-bound peer fetch wants.
// Each warp is an independent pipeline: lane i resolves row i of the tile and issues its copy,
// lane 0 arms the tile's mbarrier with the byte count and later issues the tile's bulk store.
// ================================================================================================
constexpr int kBulkWarps = 4;       // warps (independent pipelines) per CTA
constexpr int kBulkMaxStages = 8;   // stages per warp: S - 2 tiles being loaded, 2 being stored

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <bool kPartitioned, typename IdxT>
__global__ void __launch_bounds__(kBulkWarps * 32) k_gather_bulk(const __grid_constant__ GatherParams prm, const int tile_rows,
                                                                  const int stages) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ uint64_t s_bar[kBulkWarps][kBulkMaxStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const GatherView gv = gather_view(prm);
  const int64_t n = gv.n;
  const uint32_t row_bytes = (uint32_t)prm.row_bytes;
  // Rows whose size is not a multiple of 16 bytes (ogbn-products: 200 bytes in a 256-byte pitch) are
  // copied as round_up(row_bytes, 16) bytes -- the pitch guarantees the tail is readable -- into a
  // shared-memory tile of that stride, and the warp then stores the dense rows itself (8- or 4-byte
  // words); multiples of 16 leave as one bulk store per tile.
  const uint32_t copy_bytes = (row_bytes + 15u) & ~15u;
  const bool dense = copy_bytes == row_bytes;
  const uint32_t stage_bytes = (uint32_t)tile_rows * copy_bytes;
  unsigned char* my = s_raw + (size_t)warp * stages * stage_bytes;
  uint64_t* bar = s_bar[warp];
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const IdxT* __restrict__ idx = reinterpret_cast<const IdxT*>(gv.idx);
  const int64_t num_tiles = (n + tile_rows - 1) / tile_rows;
  const int64_t wglobal = (int64_t)blockIdx.x * kBulkWarps + warp;
  const int64_t wstride = (int64_t)gridDim.x * kBulkWarps;
  uint64_t pol = 0;
  if constexpr (kPartitioned) pol = l2_policy_evict_last();
  unsigned long long cnt0 = 0, cnt1 = 0, cnt2 = 0;

  // the id (and source descriptor) of the tile that is ISSUED next is loaded one step ahead
  RowResolver<kPartitioned> nxt;
  auto load_row = [&](int64_t tile, RowResolver<kPartitioned>& r) {
    r.on = false;
    if (tile < num_tiles && lane < tile_rows) {
      const int64_t row = tile * tile_rows + lane;
      if (row < n) {
        r.id = (int64_t)idx[row];
        r.on = true;
      }
    }
  };
  auto issue = [&](int64_t tile, int stage, RowResolver<kPartitioned>& r) {
    const int64_t row0 = tile * tile_rows;
    const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
    if (lane == 0) mbar_expect_tx(bar + stage, (uint32_t)rows * copy_bytes);
    __syncwarp();
    r.begin_lookup(prm, row0 + lane, pol);
    int cls;
    const char* src = r.finish(prm, cls);
    if (r.on) bulk_g2s(my + (size_t)stage * stage_bytes + (size_t)lane * copy_bytes, src, copy_bytes, bar + stage);
    if constexpr (kPartitioned) {
      if (prm.counters != nullptr) {
        const uint32_t m0 = __ballot_sync(kFullMask, cls == 0), m1 = __ballot_sync(kFullMask, cls == 1),
                       m2 = __ballot_sync(kFullMask, cls == 2);
        if (lane == 0) {
          cnt0 += __popc(m0);
          cnt1 += __popc(m1);
          cnt2 += __popc(m2);
        }
      }
    }
  };

  // prologue: stages - 2 tiles in flight
  int64_t t_issue = wglobal;
  load_row(t_issue, nxt);
  int issued = 0;
  for (; issued < stages - 2 && t_issue < num_tiles; ++issued) {
    RowResolver<kPartitioned> cur = nxt;
    load_row(t_issue + wstride, nxt);
    issue(t_issue, issued, cur);
    t_issue += wstride;
  }
  int k = 0, stage = 0, istage = issued % stages;  // tiles consumed; stage of tile k; stage of the next issue
  uint32_t parity = 0;
  for (int64_t tile = wglobal; tile < num_tiles; tile += wstride, ++k) {
    mbar_wait(bar + stage, parity);
    {
      const int64_t row0 = tile * tile_rows;
      const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
      if (dense) {
        if (lane == 0) {
          bulk_s2g(gv.out + row0 * prm.row_bytes, my + (size_t)stage * stage_bytes, (uint32_t)rows * row_bytes);
          bulk_commit();
        }
      } else {
        const unsigned char* tile_s = my + (size_t)stage * stage_bytes;
        char* tile_g = gv.out + row0 * prm.row_bytes;
        if ((row_bytes & 7u) == 0u && ((uintptr_t)gv.out & 7u) == 0u) {
          const uint32_t wpr = row_bytes >> 3, words = (uint32_t)rows * wpr;
          for (uint32_t e = lane; e < words; e += 32) {
            const uint32_t rr = e / wpr, o = e - rr * wpr;
            const int2 v = *reinterpret_cast<const int2*>(tile_s + (size_t)rr * copy_bytes + ((size_t)o << 3));
            st_na(reinterpret_cast<int2*>(tile_g) + e, v);
          }
        } else {
          const uint32_t wpr = row_bytes >> 2, words = (uint32_t)rows * wpr;
          for (uint32_t e = lane; e < words; e += 32) {
            const uint32_t rr = e / wpr, o = e - rr * wpr;
            const int v = *reinterpret_cast<const int*>(tile_s + (size_t)rr * copy_bytes + ((size_t)o << 2));
            st_na(reinterpret_cast<int*>(tile_g) + e, v);
          }
        }
        // generic-proxy reads of the stage are ordered before the async-proxy writes that refill it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
    }
    if (t_issue < num_tiles) {
      // the next load goes into the stage of tile k - 2: its store (two commits ago) must have finished reading
      if (lane == 0 && dense) bulk_wait_read<2>();
      __syncwarp();
      RowResolver<kPartitioned> cur = nxt;
      load_row(t_issue + wstride, nxt);
      issue(t_issue, istage, cur);
      t_issue += wstride;
      if (++istage == stages) istage = 0;
    }
    if (++stage == stages) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last stores' reads
  if constexpr (kPartitioned) {
    if (prm.counters != nullptr && lane == 0) {
      if (cnt0) atomicAdd(prm.counters + 0, cnt0);
      if (cnt1) atomicAdd(prm.counters + 1, cnt1);
      if (cnt2) atomicAdd(prm.counters + 2, cnt2);
    }
  }
}

// Does any partition table live on another GPU?  Peer tables enter a process through
// spp_ipc_import, which remembers the pointers it handed out (runtime.cu).
static bool map_has_peer_tables(const GatherParams& prm) {
  for (int p = 0; p < prm.book.num_parts; ++p)
    if (p != prm.book.rank && prm.tables[p] != nullptr && ipc_imported(prm.tables[p])) return true;
  return false;
}

// ================================================================================================
// Gather by source class.  After the owner split every node of the batch has a position `pos` in
// bucket order (partition 0 .. P-1, then cached), bucket_ids[pos] = global id (or cache row) and
// inv[pos] = its row in x.  This kernel serves ONE set of classes (a bit mask): the classes whose
// rows live in this GPU's HBM (hosted partitions + replicated cache) or the classes that live on
// peer GPUs.  The two launches run on two streams, so the NVLink-bound peer fetch (few CTAs, deep
// queues) and the HBM-bound local gather overlap instead of sharing tiles in which every CTA waits
// for its slowest, remote, rows.  Source side: dense walk over a bucket;
// destination side: one row (>= 128 contiguous bytes for every BASELINE shape) per inv[pos].
// ================================================================================================
struct ClassGatherParams {
  GatherParams g;              // tables, cache table, book (offsets), row / pitch sizes, out | job
  const int64_t* bucket_ids;   // [n] (from the job block when g.job is set)
  const int32_t* inv;          // [n]
  const uint32_t* class_start; // [kSplitClasses + 1] exclusive prefix of the bucket sizes
  uint32_t class_mask;         // bit c: this launch serves bucket c (c == num_parts: cached rows)
};

template <typename V>
__global__ void __launch_bounds__(kGatherThreads) k_gather_classes(const __grid_constant__ ClassGatherParams cp) {
  __shared__ const char* s_src[2][kRows];
  __shared__ char* s_dst[2][kRows];
  __shared__ uint32_t s_pref[kSplitClasses + 1];  // prefix of the served buckets' sizes
  __shared__ uint32_t s_first[kSplitClasses];     // first position of every bucket
  const GatherParams& prm = cp.g;
  const int tid = threadIdx.x;
  const int P = prm.book.num_parts;
  if (tid == 0) {
    uint32_t acc = 0;
    for (int c = 0; c < kSplitClasses; ++c) {
      const uint32_t lo = cp.class_start[c], hi = cp.class_start[c + 1];
      s_first[c] = lo;
      s_pref[c] = acc;
      if ((cp.class_mask >> c) & 1u) acc += hi - lo;
    }
    s_pref[kSplitClasses] = acc;
  }
  __syncthreads();
  const int64_t M = s_pref[kSplitClasses];
  const int64_t num_tiles = (M + kRows - 1) / kRows;
  char* const out = (prm.job != nullptr) ? reinterpret_cast<char*>(prm.job->x_out) : prm.out;
  const int64_t* __restrict__ bucket_ids = (prm.job != nullptr) ? prm.job->bucket_ids : cp.bucket_ids;
  const uint32_t vpr = prm.vpr, magic = prm.vpr_magic;
  const bool resolver = tid < kRows;
  unsigned long long cnt0 = 0, cnt1 = 0, cnt2 = 0;

  // resolver state of one row: bucket, id / cache row and destination row (loads issued a tile ahead)
  struct Row {
    int64_t val = 0;
    int32_t dst = 0;
    int cls = -1;
  };
  auto load_row = [&](int64_t tile, Row& r) {
    r.cls = -1;
    const int64_t m = tile * kRows + tid;
    if (tile < num_tiles && m < M) {
      int c = 0;
#pragma unroll
      for (int q = 1; q < kSplitClasses; ++q)
        if ((uint32_t)m >= s_pref[q]) c = q;
      const uint32_t pos = s_first[c] + ((uint32_t)m - s_pref[c]);
      r.val = bucket_ids[pos];
      r.dst = cp.inv[pos];
      r.cls = c;
    }
  };
  auto publish = [&](const Row& r, int buf) {
    const char* src = nullptr;
    char* dst = nullptr;
    int kind = -1;
    if (r.cls >= 0) {
      if (r.cls == P) {
        src = prm.cache_table + r.val * prm.cache_pitch;
        kind = 1;
      } else {
        src = prm.tables[r.cls] + (r.val - prm.book.off[r.cls]) * prm.table_pitch;
        kind = book_is_local(prm.book, r.cls) ? 0 : 2;
      }
      dst = out + (int64_t)r.dst * prm.row_bytes;
    }
    s_src[buf][tid] = src;
    s_dst[buf][tid] = dst;
    if (prm.counters != nullptr) {
      const uint32_t m0 = __ballot_sync(kFullMask, kind == 0), m1 = __ballot_sync(kFullMask, kind == 1),
                     m2 = __ballot_sync(kFullMask, kind == 2);
      if ((tid & 31) == 0) {
        cnt0 += __popc(m0);
        cnt1 += __popc(m1);
        cnt2 += __popc(m2);
      }
    }
  };

  Row r1;
  int64_t tile = blockIdx.x;
  if (resolver) {
    Row r0;
    load_row(tile, r0);
    load_row(tile + gridDim.x, r1);
    publish(r0, 0);
  }
  __syncthreads();
  int buf = 0;
  for (; tile < num_tiles; tile += gridDim.x) {
    const int64_t m0 = tile * kRows;
    const int rows = (int)((M - m0) < kRows ? (M - m0) : kRows);
    Row r2;
    if (resolver) load_row(tile + 2 * (int64_t)gridDim.x, r2);
    const uint32_t chunks = (uint32_t)rows * vpr;
    const char* const* src = s_src[buf];
    char* const* dst = s_dst[buf];
    for (uint32_t base = tid; base < chunks; base += kGatherThreads * kUnroll) {
      V vals[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t lc = base + u * kGatherThreads;
        if (lc < chunks) {
          const uint32_t r = magic ? __umulhi(lc, magic) : lc / vpr;
          vals[u] = ld_nc_na(reinterpret_cast<const V*>(src[r]) + (lc - r * vpr));
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t lc = base + u * kGatherThreads;
        if (lc < chunks) {
          const uint32_t r = magic ? __umulhi(lc, magic) : lc / vpr;
          st_na(reinterpret_cast<V*>(dst[r]) + (lc - r * vpr), vals[u]);
        }
      }
    }
    if (resolver) {
      publish(r1, buf ^ 1);
      r1 = r2;
    }
    __syncthreads();
    buf ^= 1;
  }
  if (prm.counters != nullptr && resolver && (tid & 31) == 0) {
    if (cnt0) atomicAdd(prm.counters + 0, cnt0);
    if (cnt1) atomicAdd(prm.counters + 1, cnt1);
    if (cnt2) atomicAdd(prm.counters + 2, cnt2);
  }
}

// opt-in shared-memory size of the bulk-copy flavour (once per device; also called before a launch
// sequence is captured into a CUDA graph)
int gather_attributes() {
  static bool done[64] = {false};
  int dev = 0;
  SPP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || done[dev]) return 0;
  SPP_CUDA(cudaFuncSetAttribute(k_gather_bulk<true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SPP_CUDA(cudaFuncSetAttribute(k_gather_bulk<true, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SPP_CUDA(cudaFuncSetAttribute(k_gather_bulk<false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SPP_CUDA(cudaFuncSetAttribute(k_gather_bulk<false, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  done[dev] = true;
  return 0;
}

// decided by the A/B runs under profiles/ (r02_ab_gather_bulk_*.txt)
static bool bulk_default_for_peers() { return false; }

template <bool kPartitioned>
static int launch_gather(GatherParams& prm, int vec_bytes, int idx_is_64, cudaStream_t st) {
  if (prm.n_max <= 0) return 0;
  prm.vpr = (uint32_t)(prm.row_bytes / vec_bytes);
  // 4 CTAs / SM x 8 loads in flight per thread saturate HBM and leave half of every SM's thread
  // slots to the latency-bound sampler kernels of the other in-flight mini-batches.  When rows
  // come from peer GPUs the kernel is NVLink bound and holds its CTAs 3-5x longer, and several
  // batches' gathers are resident at once: ONE CTA per SM with a 128-row tile (32 KB of rows in
  // flight per SM and gather) keeps the link as busy and leaves the SMs to the samplers.
  const Tunables& tn = tunables();
  int cps = tn.gather_ctas_per_sm > 0 ? tn.gather_ctas_per_sm : 0;
  bool peers = false;
  if constexpr (kPartitioned) peers = map_has_peer_tables(prm);
  const int bulk_mode = tn.gather_bulk;
  const int64_t copy_bytes = (prm.row_bytes + 15) & ~15ll;
  const bool bulk_ok = prm.row_bytes <= 8192 && (vec_bytes == 16 || (prm.src_align16 && (prm.row_bytes & 3) == 0 &&
                                                                       copy_bytes <= prm.table_pitch &&
                                                                       (!kPartitioned || prm.cache_table == nullptr || copy_bytes <= prm.cache_pitch)));
  if (bulk_ok && (bulk_mode == 1 || (bulk_mode < 0 && peers && bulk_default_for_peers()))) {
    const int tile_bytes = tn.bulk_tile > 0 ? tn.bulk_tile : 4096;
    int stages = tn.bulk_stages;
    if (stages < 3) stages = 3;
    if (stages > kBulkMaxStages) stages = kBulkMaxStages;
    const int bcps = tn.bulk_ctas_per_sm > 0 ? tn.bulk_ctas_per_sm : 2;
    int tile_rows = (int)(tile_bytes / copy_bytes);
    if (tile_rows > 32) tile_rows = 32;
    if (tile_rows < 1) tile_rows = 1;
    const size_t smem = (size_t)kBulkWarps * stages * tile_rows * copy_bytes;
    if (smem <= 200 * 1024) {
      if (int r = gather_attributes()) return r;
      const int64_t btiles = ceil_div(prm.n_max, tile_rows);
      const int64_t bmax = (int64_t)num_sms() * (cps > 0 ? cps : bcps);
      const int64_t want = ceil_div(btiles, kBulkWarps);
      const int bgrid = (int)(want < bmax ? want : bmax);
      if (idx_is_64) k_gather_bulk<kPartitioned, int64_t><<<bgrid, kBulkWarps * 32, smem, st>>>(prm, tile_rows, stages);
      else k_gather_bulk<kPartitioned, int32_t><<<bgrid, kBulkWarps * 32, smem, st>>>(prm, tile_rows, stages);
      SPP_KERNEL_CHECK("k_gather_bulk");
      return 0;
    }
  }
  if (cps == 0) {
    cps = 4;
    if (peers) cps = 1;  // round 2, with 128-row tiles: 1 / 2 CTAs per SM = 93.1 / 96.7 us per batch at 2 GPUs,
                         // 121.4 / 130.9 at 4 (profiles/r02_ab_nvlink_regime_g*.txt)
  }
  // rows per tile: deeper tiles when rows may come over NVLink (latency ~3x HBM's), see k_gather
  // (128: at 2 GPUs as fast as 256, and with ~460 k rows over 296 persistent CTAs the last round of
  // 256-row tiles is only partly filled: 28.7 k vs 30.1 k batches/s at 4 GPUs)
  int tile_rows = tn.gather_tile_rows > 0 ? tn.gather_tile_rows : (peers ? 128 : 64);
  tile_rows = tile_rows >= 256 ? 256 : tile_rows >= 128 ? 128 : 64;
  if ((uint64_t)prm.vpr * tile_rows >= (1ull << 31))
    return fail(SPP_EUNSUPPORTED, "gather: row of %lld bytes too wide", (long long)prm.row_bytes);
  {
    const bool mok = prm.vpr > 1 && (uint64_t)tile_rows * prm.vpr * prm.vpr < (1ull << 32);
    prm.vpr_magic = mok ? (uint32_t)(((1ull << 32) + prm.vpr - 1) / prm.vpr) : 0u;
  }
  const int64_t tiles = ceil_div(prm.n_max, tile_rows);
  const int64_t max_ctas = (int64_t)num_sms() * cps;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
#define SPP_GATHER_LAUNCH_R(V, R)                                                               \
  do {                                                                                          \
    if (idx_is_64)                                                                              \
      k_gather<V, kPartitioned, int64_t, R><<<grid, kGatherThreads, 0, st>>>(prm);              \
    else                                                                                        \
      k_gather<V, kPartitioned, int32_t, R><<<grid, kGatherThreads, 0, st>>>(prm);              \
  } while (0)
#define SPP_GATHER_LAUNCH(V)                                                                    \
  do {                                                                                          \
    if (tile_rows == 256) SPP_GATHER_LAUNCH_R(V, 256);                                          \
    else if (tile_rows == 128) SPP_GATHER_LAUNCH_R(V, 128);                                     \
    else SPP_GATHER_LAUNCH_R(V, 64);                                                            \
  } while (0)
  switch (vec_bytes) {
    case 16: SPP_GATHER_LAUNCH(int4); break;
    case 8: SPP_GATHER_LAUNCH(int2); break;
    case 4: SPP_GATHER_LAUNCH(int); break;
    case 2: SPP_GATHER_LAUNCH(short); break;
    default: SPP_GATHER_LAUNCH(char); break;
  }
#undef SPP_GATHER_LAUNCH
#undef SPP_GATHER_LAUNCH_R
  SPP_KERNEL_CHECK("k_gather");
  return 0;
}

static int pick_vec_bytes(int64_t row_bytes, uintptr_t align_bits) {
  for (int v = 16; v > 1; v >>= 1)
    if (row_bytes % v == 0 && (align_bits % v) == 0) return v;
  return 1;
}

}  // namespace spp

namespace spp {

int gather_rows_job(const void* table, int64_t table_pitch, int64_t row_bytes, const void* idx, int idx_is_64, int64_t n_idx,
                    const int64_t* n_idx_dev, void* out, int64_t n_out_rows, cudaStream_t st, const spp_device_job* job,
                    int job_mode) {
  if (row_bytes <= 0) return fail(SPP_EINVAL, "spp_gather_rows: row_bytes must be positive");
  if (table_pitch < row_bytes) return fail(SPP_EINVAL, "spp_gather_rows: table pitch smaller than the row");
  int64_t n = n_idx < n_out_rows ? n_idx : n_out_rows;
  if (n <= 0) return 0;
  if (!table || (!job && (!idx || !out))) return fail(SPP_EINVAL, "spp_gather_rows: null pointer");
  GatherParams prm{};
  prm.table = (const char*)table;
  prm.idx = idx;
  prm.n_dev = n_idx_dev;
  prm.n_max = n;
  prm.out = (char*)out;
  prm.row_bytes = row_bytes;
  prm.table_pitch = table_pitch;
  prm.job = job;
  prm.job_mode = job ? job_mode : 0;
  prm.src_align16 = ((((uintptr_t)table | (uintptr_t)table_pitch) & 15u) == 0u) ? 1 : 0;
  // with a job block the output address is not known to the host: torch allocations (512-byte
  // aligned) are assumed; rows that need a narrower vector because of their size still get it
  int vb = pick_vec_bytes(row_bytes, (uintptr_t)table | (job ? 0 : (uintptr_t)out) | (uintptr_t)table_pitch);
  if (job && job_mode == 2 && vb > 8) vb = 8;  // label rows live at the tail of the int64 arena: 8-byte aligned only
  return launch_gather<false>(prm, vb, idx_is_64, st);
}

int gather_partitioned_job(const spp_feature_map* m, int64_t row_bytes, const void* n_id, int idx_is_64, int64_t n_idx,
                           const int64_t* n_idx_dev, const int32_t* src_desc, void* out, int64_t n_out_rows,
                           int64_t* counters, cudaStream_t st, const spp_device_job* job) {
  if (!m) return fail(SPP_EINVAL, "spp_gather_partitioned: null feature map");
  if (m->num_parts < 1 || m->num_parts > SPP_MAX_PARTS || m->rank < 0 || m->rank >= m->num_parts)
    return fail(SPP_EINVAL, "spp_gather_partitioned: bad num_parts/rank (%d/%d)", m->num_parts, m->rank);
  if (row_bytes <= 0) return fail(SPP_EINVAL, "spp_gather_partitioned: row_bytes must be positive");
  int64_t n = n_idx < n_out_rows ? n_idx : n_out_rows;
  if (n <= 0) return 0;
  if (!n_id || (!out && !job)) return fail(SPP_EINVAL, "spp_gather_partitioned: null pointer");
  if ((m->cache_index == nullptr) != (m->cache_table == nullptr))
    return fail(SPP_EINVAL, "spp_gather_partitioned: cache_index and cache_table must be given together");
  GatherParams prm{};
  prm.idx = n_id;
  prm.n_dev = n_idx_dev;
  prm.n_max = n;
  prm.out = (char*)out;
  prm.job = job;
  prm.job_mode = job ? 1 : 0;
  prm.row_bytes = row_bytes;
  prm.table_pitch = m->table_pitch > 0 ? m->table_pitch : row_bytes;
  prm.cache_pitch = m->cache_pitch > 0 ? m->cache_pitch : row_bytes;
  if (prm.table_pitch < row_bytes || prm.cache_pitch < row_bytes)
    return fail(SPP_EINVAL, "spp_gather_partitioned: pitch smaller than the row");
  prm.book.num_parts = m->num_parts;
  prm.book.rank = m->rank;
  uintptr_t align = (job ? 0 : (uintptr_t)out) | (uintptr_t)prm.table_pitch | (uintptr_t)prm.cache_pitch;
  for (int p = 0; p <= SPP_MAX_PARTS; ++p) prm.book.off[p] = p <= m->num_parts ? m->offsets[p] : m->offsets[m->num_parts];
  for (int p = 0; p < m->num_parts; ++p) {
    if (m->offsets[p + 1] < m->offsets[p]) return fail(SPP_EINVAL, "spp_gather_partitioned: offsets not sorted");
    if (m->tables[p] == nullptr && m->offsets[p + 1] > m->offsets[p] && !(m->cache_index && p != m->rank))
      return fail(SPP_EINVAL, "spp_gather_partitioned: partition %d has no table", p);
    prm.tables[p] = (const char*)m->tables[p];
    align |= (uintptr_t)m->tables[p];
  }
  prm.cache_table = (const char*)m->cache_table;
  {
    uintptr_t sa = (uintptr_t)prm.table_pitch | (uintptr_t)m->cache_table | (m->cache_table ? (uintptr_t)prm.cache_pitch : 0);
    for (int p = 0; p < m->num_parts; ++p) sa |= (uintptr_t)m->tables[p];
    prm.src_align16 = (sa & 15u) == 0u ? 1 : 0;
  }
  prm.cache = make_cache_index(m->cache_index, m->cache_index_nodes);
  prm.l2_stream_hint = (m->cache_index != nullptr && tunables().gather_l2_hint != 0) ? 1 : 0;
  prm.desc = src_desc;
  prm.book.local_mask = (1u << m->rank) | m->local_parts;
  align |= (uintptr_t)m->cache_table;
  prm.counters = (unsigned long long*)counters;
  int vb = pick_vec_bytes(row_bytes, align);
  return launch_gather<true>(prm, vb, idx_is_64, st);
}


int gather_by_class_job(const spp_feature_map* m, int64_t row_bytes, const int64_t* bucket_ids, const int32_t* split_scratch,
                        int64_t n_max, uint32_t class_mask, void* out, int64_t* counters, cudaStream_t st,
                        const spp_device_job* job) {
  if (!m) return fail(SPP_EINVAL, "spp_gather_by_class: null feature map");
  if (m->num_parts < 1 || m->num_parts > SPP_MAX_PARTS || m->rank < 0 || m->rank >= m->num_parts)
    return fail(SPP_EINVAL, "spp_gather_by_class: bad num_parts/rank (%d/%d)", m->num_parts, m->rank);
  if (row_bytes <= 0) return fail(SPP_EINVAL, "spp_gather_by_class: row_bytes must be positive");
  if (n_max <= 0 || class_mask == 0) return 0;
  if (!split_scratch || (!job && (!bucket_ids || !out))) return fail(SPP_EINVAL, "spp_gather_by_class: null pointer");
  ClassGatherParams cp{};
  GatherParams& prm = cp.g;
  prm.out = (char*)out;
  prm.job = job;
  prm.job_mode = job ? 1 : 0;
  prm.row_bytes = row_bytes;
  prm.table_pitch = m->table_pitch > 0 ? m->table_pitch : row_bytes;
  prm.cache_pitch = m->cache_pitch > 0 ? m->cache_pitch : row_bytes;
  if (prm.table_pitch < row_bytes || prm.cache_pitch < row_bytes)
    return fail(SPP_EINVAL, "spp_gather_by_class: pitch smaller than the row");
  prm.book.num_parts = m->num_parts;
  prm.book.rank = m->rank;
  prm.book.local_mask = (1u << m->rank) | m->local_parts;
  uintptr_t align = (job ? 0 : (uintptr_t)out) | (uintptr_t)row_bytes | (uintptr_t)prm.table_pitch;
  bool peers = false;
  for (int p = 0; p <= SPP_MAX_PARTS; ++p) prm.book.off[p] = p <= m->num_parts ? m->offsets[p] : m->offsets[m->num_parts];
  for (int p = 0; p < m->num_parts; ++p) {
    prm.tables[p] = (const char*)m->tables[p];
    if ((class_mask >> p) & 1u) {
      if (m->tables[p] == nullptr && m->offsets[p + 1] > m->offsets[p])
        return fail(SPP_EINVAL, "spp_gather_by_class: partition %d has no table", p);
      align |= (uintptr_t)m->tables[p];
      peers = peers || (m->tables[p] != nullptr && ipc_imported(m->tables[p]));
    }
  }
  if (((class_mask >> m->num_parts) & 1u) && !m->cache_table) class_mask &= ~(1u << m->num_parts);  // no cache: empty bucket
  if ((class_mask >> m->num_parts) & 1u) align |= (uintptr_t)m->cache_table | (uintptr_t)prm.cache_pitch;
  if (class_mask == 0) return 0;
  prm.cache_table = (const char*)m->cache_table;
  prm.counters = (unsigned long long*)counters;
  cp.bucket_ids = bucket_ids;
  cp.inv = split_scratch_inv(split_scratch, n_max);
  cp.class_start = split_scratch_class_start(split_scratch, n_max);
  cp.class_mask = class_mask;
  const int vb = pick_vec_bytes(row_bytes, align);
  prm.vpr = (uint32_t)(row_bytes / vb);
  if ((uint64_t)prm.vpr * kRows >= (1ull << 31)) return fail(SPP_EUNSUPPORTED, "spp_gather_by_class: row too wide");
  const bool magic_ok = prm.vpr > 1 && (uint64_t)kRows * prm.vpr * prm.vpr < (1ull << 32);
  prm.vpr_magic = magic_ok ? (uint32_t)(((1ull << 32) + prm.vpr - 1) / prm.vpr) : 0u;
  const Tunables& tn = tunables();
  // NVLink-bound launches need few CTAs (their rows are all remote: every byte in flight is a peer
  // byte), HBM-bound ones the usual 4 per SM
  int cps = tn.gather_ctas_per_sm > 0 ? tn.gather_ctas_per_sm : (peers ? 2 : 4);
  const int64_t tiles = ceil_div(n_max, kRows);
  const int64_t max_ctas = (int64_t)num_sms() * cps;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  switch (vb) {
    case 16: k_gather_classes<int4><<<grid, kGatherThreads, 0, st>>>(cp); break;
    case 8: k_gather_classes<int2><<<grid, kGatherThreads, 0, st>>>(cp); break;
    case 4: k_gather_classes<int><<<grid, kGatherThreads, 0, st>>>(cp); break;
    case 2: k_gather_classes<short><<<grid, kGatherThreads, 0, st>>>(cp); break;
    default: k_gather_classes<char><<<grid, kGatherThreads, 0, st>>>(cp); break;
  }
  SPP_KERNEL_CHECK("k_gather_classes");
  return 0;
}

}  // namespace spp

extern "C" {

int spp_gather_rows(const void* table, int64_t row_bytes, const void* idx, int idx_is_64, int64_t n_idx,
                    const int64_t* n_idx_dev, void* out, int64_t n_out_rows, void* stream) {
  return spp_gather_rows_pitched(table, row_bytes, row_bytes, idx, idx_is_64, n_idx, n_idx_dev, out, n_out_rows, stream);
}

int spp_gather_rows_pitched(const void* table, int64_t table_pitch, int64_t row_bytes, const void* idx, int idx_is_64,
                            int64_t n_idx, const int64_t* n_idx_dev, void* out, int64_t n_out_rows, void* stream) {
  return spp::gather_rows_job(table, table_pitch, row_bytes, idx, idx_is_64, n_idx, n_idx_dev, out, n_out_rows,
                              (cudaStream_t)stream, nullptr, 0);
}

int spp_gather_partitioned(const spp_feature_map* m, int64_t row_bytes, const void* n_id, int idx_is_64,
                           int64_t n_idx, const int64_t* n_idx_dev, const int32_t* src_desc, void* out,
                           int64_t n_out_rows, int64_t* counters, void* stream) {
  return spp::gather_partitioned_job(m, row_bytes, n_id, idx_is_64, n_idx, n_idx_dev, src_desc, out, n_out_rows, counters,
                                     (cudaStream_t)stream, nullptr);
}

int spp_gather_by_class(const spp_feature_map* m, int64_t row_bytes, const int64_t* bucket_ids, const int32_t* split_scratch,
                        int64_t n_max, uint32_t class_mask, void* out, int64_t* counters, void* stream) {
  return spp::gather_by_class_job(m, row_bytes, bucket_ids, split_scratch, n_max, class_mask, out, counters,
                                  (cudaStream_t)stream, nullptr);
}

}  // extern "C"
