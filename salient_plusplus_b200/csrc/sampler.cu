// sampler.cu -- K1: multi-hop uniform neighbour sampling over CSR with per-hop node dedup and
// global->local relabelling.  Replaces sample_adj (fast_sampler/sample_cpu.hpp:25-143) and
// multilayer_sample (fast_sampler/fast_sampler.cpp:191-236) of the reference.
//
// Common idea of both code paths below.  Per hop h the frontier is every node discovered so far
// (T of them); a candidate edge gets a position p in the reference's sequential visiting order
// (target by target, neighbour by neighbour).  The global id of every candidate is inserted into
// an L2-resident id table and ~(T + p) is atomicMax-ed into its entry: established nodes hold
// ~local with local < T so they always win, a new node ends up holding the SMALLEST position that
// saw it.  An order-preserving compaction of the candidates that own their entry then hands out
// local ids T, T+1, ... in exactly the reference's first-discovery order, and a last pass turns
// candidate slots into local ids and sorts every row ascending (std::sort, sample_cpu.hpp:126).
//
//   General path (with replacement, fan-out > 32), 4 kernels per hop:
//     k_hop_count_scan -> k_hop_sample<mode,G,PPL> -> k_hop_compact -> k_relabel_sort_general;
//     64-bit decoupled look-back scans, data-dependent edge counts.
//   Full neighbourhood (k < 0: layer-wise inference) shares the count / compact kernels but is
//     edge parallel where rows can be hubs: k_hop_sample_edges (one thread per candidate, row by
//     binary search in the scanned out_rowptr) and, for rows longer than 32,
//     k_sort_rows_bitmap (CTA per row, shared-memory bitmap of the local ids, O(S/32 + n));
//     k_sort_large_rows (comparison network) only takes rows with repeated neighbours or
//     batches of more than 1.6 M nodes.
//   Fused path (1 <= k <= 32, without replacement: every reference configuration), 3 kernels per hop:
//     k_hop_sample_fused -> k_hop_compact_fused -> k_relabel_sort_fused; see the banner further down.
//
// Table entry: 64 bits = { key + 1 , ~local } ; empty = 0, so the table is cleared by a memset.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace spp {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 1024 items per tile
constexpr int kSampleThreads = 256;
constexpr int kMetaWork = 25;       // meta word: number of rows queued for the CTA-per-row sorters (list: tgt_deg)
constexpr int kMetaWork2 = 26;      // meta word: rows the bitmap sorter handed on (list: low words of tgt_start)
constexpr int kWarpSortCap = 1024;  // rows up to this length are sorted by one warp in shared memory
constexpr int kBlockSortSmemElems = 48 * 1024;  // int32 elements a CTA sorts in shared memory

// Two flavours behind one interface:
//   hashed  (default)  open addressing, slots = pow2 >= 1.5 * node bound: L2 resident whatever the
//                      graph size (16 MB at (15,10,5)@1024);
//   direct             slot == node id (a perfect hash), chosen by spp_sampler_sizes when the graph
//                      has at most 1.5x as many nodes as the hash table would have slots: no CAS round
//                      trip per candidate, no probing (-7 % per mini-batch on the products shape).
struct Table {
  uint32_t* w;  // interleaved {key+1, ~local}
  uint32_t slots;  // hashed: any size >= 1.25 * node bound (multiply-high range reduction, no power of two needed)
  int direct;
};

__device__ __forceinline__ uint32_t table_home(const Table& t, int32_t key) {
  return t.direct ? (uint32_t)key : __umulhi((uint32_t)key * 2654435761u, t.slots);
}

// insert-if-absent; returns the slot of `key`
__device__ __forceinline__ uint32_t table_insert(const Table& t, int32_t key) {
  const uint32_t want = (uint32_t)key + 1u;
  uint32_t slot = table_home(t, key);
  if (t.direct) {  // racing writers store the same value
    __stcg(t.w + 2 * (size_t)slot, want);
    return slot;
  }
  while (true) {  // optimistic: one L2 round trip when the slot is free or already holds the key
    const uint32_t prev = atomicCAS(t.w + 2 * (size_t)slot, 0u, want);
    if (prev == 0u || prev == want) return slot;
    slot = (slot + 1 == t.slots) ? 0u : slot + 1;
  }
}

// insert-if-absent AND raise the entry's value to `val`, normally in ONE L2 atomic (every random
// table access costs a 32-byte L2 sector whatever its width, and the sampler chain is bound by that
// sector rate): direct-mapped tables take a 64-bit atomicMax on {val : key+1} (all writers of a slot
// carry the same key, so the maximum is the maximum of val); hashed tables a 64-bit CAS on the
// empty entry -- it either inserts {val : key+1}, or returns the resident entry, and only a resident
// entry of the same key with a smaller value needs the second (32-bit) atomicMax.
__device__ __forceinline__ uint32_t table_insert_max(const Table& t, int32_t key, uint32_t val) {
  const uint32_t want = (uint32_t)key + 1u;
  unsigned long long* tab64 = reinterpret_cast<unsigned long long*>(t.w);
  const unsigned long long mine = ((unsigned long long)val << 32) | (unsigned long long)want;
  uint32_t slot = table_home(t, key);
  if (t.direct) {
    atomicMax(tab64 + slot, mine);
    return slot;
  }
  while (true) {
    const unsigned long long old = atomicCAS(tab64 + slot, 0ull, mine);
    if (old == 0ull) return slot;
    if ((uint32_t)old == want) {
      if ((uint32_t)(old >> 32) < val) atomicMax(t.w + 2 * (size_t)slot + 1, val);
      return slot;
    }
    slot = (slot + 1 == t.slots) ? 0u : slot + 1;
  }
}

__device__ __forceinline__ uint32_t table_find(const Table& t, int32_t key) {
  const uint32_t want = (uint32_t)key + 1u;
  uint32_t slot = table_home(t, key);
  while (__ldcg(t.w + 2 * (size_t)slot) != want) slot = (slot + 1 == t.slots) ? 0u : slot + 1;
  return slot;
}

// ------------------------------------------------------------------------------------------------
// seeds: n_ids[i] = seeds[i]; map[seed] = LAST position of that seed (sample_cpu.hpp:13-19)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_seeds_init(const spp_device_job* job, const int64_t* __restrict__ seeds, int64_t bs,
                                                     int32_t* __restrict__ n_ids, Table tab,
                                                     int64_t* __restrict__ meta) {
  const int tid = threadIdx.x;
  if (job != nullptr) {
    seeds = job->seeds;
    bs = job->batch_size;
  }
  for (int64_t i = tid; i < bs; i += blockDim.x) {
    const int32_t key = (int32_t)seeds[i];
    n_ids[i] = key;
    const uint32_t slot = table_insert(tab, key);
    atomicMax(tab.w + 2 * (size_t)slot + 1, (uint32_t)i + 1u);
  }
  __syncthreads();
  for (int64_t i = tid; i < bs; i += blockDim.x) {
    const uint32_t slot = table_find(tab, n_ids[i]);
    uint32_t* enc = tab.w + 2 * (size_t)slot + 1;
    if (__ldcg(enc) == (uint32_t)i + 1u) __stcg(enc, ~(uint32_t)i);
  }
  if (tid == 0) {
    meta[SPP_META_NODES(0)] = bs;
    meta[SPP_META_OVERFLOW] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// decoupled look-back scan machinery (tile_state[0] = dynamic tile counter, [1 + t] = tile t)
// state word = status << 62 | value ; status 0 invalid, 1 aggregate, 2 inclusive prefix
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kValMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* s_warp, uint64_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(kFullMask, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint64_t wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    uint64_t x = s_warp[w];
    if (w < warp) wbase += x;
    tot += x;
  }
  total = tot;
  __syncthreads();  // s_warp reusable
  return wbase + inc - v;
}

// Called by every thread of the CTA; returns the exclusive prefix of `tile`.
__device__ __forceinline__ uint64_t lookback(uint64_t* tile_state, int64_t tile, uint64_t aggregate,
                                             uint64_t* s_bcast) {
  uint64_t* st = tile_state + 1;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint64_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_volatile_u64(st + 0, (2ull << 62) | (aggregate & kValMask));
    } else {
      if (lane == 0) st_volatile_u64(st + tile, (1ull << 62) | (aggregate & kValMask));
      int64_t look = tile - 1;
      while (true) {
        const int64_t idx = look - lane;
        uint64_t s = (2ull << 62);  // virtual zero prefix in front of tile 0
        if (idx >= 0) {
          do {
            s = ld_volatile_u64(st + idx);
          } while ((s >> 62) == 0);
        }
        const uint32_t pref = __ballot_sync(kFullMask, (s >> 62) == 2);
        uint64_t v = s & kValMask;
        if (pref) {
          const int first = __ffs(pref) - 1;  // nearest tile holding an inclusive prefix
          if (lane > first) v = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
        excl += v;
        if (pref) break;
        look -= 32;
      }
      if (lane == 0) st_volatile_u64(st + tile, (2ull << 62) | ((excl + aggregate) & kValMask));
    }
    if (lane == 0) *s_bcast = excl;
  }
  __syncthreads();
  const uint64_t r = *s_bcast;
  __syncthreads();
  return r;
}

__device__ __forceinline__ int64_t next_tile(uint64_t* tile_state, int64_t* s_tile) {
  if (threadIdx.x == 0) *s_tile = (int64_t)atomicAdd((unsigned long long*)tile_state, 1ull);
  __syncthreads();
  const int64_t t = *s_tile;
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------------
// 1. kept-edge count per target + exclusive scan
// ------------------------------------------------------------------------------------------------
struct HopParams {
  const int64_t* rowptr;
  const void* col;
  int32_t* n_ids;
  int64_t* tgt_start;
  int32_t* tgt_deg;
  int64_t* meta;
  uint64_t* tile_state;
  int64_t* out_rowptr;
  int64_t* out_col;
  Table tab;
  int64_t max_targets;  // capacity of tgt_* / out_rowptr
  int64_t max_edges;    // capacity of out_col
  int64_t max_nodes;    // capacity of n_ids
  uint64_t premixed;    // RNG key
  int32_t hop;
  int32_t fanout;
  int32_t replace;
  int32_t bitmap_rows;  // != 0: rows longer than 32 go (un-relabelled) to k_sort_rows_bitmap
  // When set, the per-batch pointers / capacities / RNG key are read from this device-resident
  // block instead of the kernel parameters, so that a captured CUDA graph of the launch sequence
  // can be replayed for every mini-batch of a slot (session.cu).
  const spp_device_job* job;
};

__device__ __forceinline__ int64_t* hop_out_rowptr(const HopParams& p) { return p.job ? p.job->out_rowptr[p.hop] : p.out_rowptr; }
__device__ __forceinline__ int64_t* hop_out_col(const HopParams& p) { return p.job ? p.job->out_col[p.hop] : p.out_col; }
__device__ __forceinline__ int64_t hop_max_edges(const HopParams& p) { return p.job ? p.job->out_col_cap[p.hop] : p.max_edges; }
__device__ __forceinline__ uint64_t hop_rng_key(const HopParams& p) { return p.job ? p.job->rng_premixed : p.premixed; }
#define SPP_HOP_LOCALS                                  \
  int64_t* const o_rowptr = hop_out_rowptr(prm);        \
  int64_t* const o_col = hop_out_col(prm);              \
  const int64_t o_cap = hop_max_edges(prm);             \
  const uint64_t o_seed = hop_rng_key(prm);             \
  (void)o_rowptr; (void)o_col; (void)o_cap; (void)o_seed;

__global__ void __launch_bounds__(kScanThreads) k_hop_count_scan(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  __shared__ uint64_t s_warp[kScanThreads / 32];
  __shared__ uint64_t s_bcast;
  __shared__ int64_t s_tile;
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) {
    T = prm.max_targets;
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.meta[SPP_META_OVERFLOW] = 1;
  }
  const int64_t num_tiles = (T + kScanTile - 1) / kScanTile;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    o_rowptr[0] = 0;
    prm.meta[kMetaWork] = 0;
    prm.meta[kMetaWork2] = 0;
    if (T == 0) prm.meta[SPP_META_EDGES(prm.hop)] = 0;
  }
  const int k = prm.fanout;
  while (true) {
    const int64_t tile = next_tile(prm.tile_state, &s_tile);
    if (tile >= num_tiles) break;
    const int64_t i0 = tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t c[kScanItems];
    uint64_t sum = 0;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      const int64_t i = i0 + q;
      c[q] = 0;
      if (i < T) {
        const int32_t n = prm.n_ids[i];
        const int64_t s = __ldg(prm.rowptr + n), e = __ldg(prm.rowptr + n + 1);
        const int32_t deg = (int32_t)(e - s);
        prm.tgt_start[i] = s;
        prm.tgt_deg[i] = deg;
        uint32_t cnt;
        if (k < 0) cnt = (uint32_t)deg;
        else if (prm.replace) cnt = deg > 0 ? (uint32_t)k : 0u;
        else cnt = (uint32_t)(deg < k ? deg : k);
        c[q] = cnt;
        sum += cnt;
      }
    }
    uint64_t total;
    const uint64_t texcl = block_excl_scan(sum, s_warp, total);
    const uint64_t base = lookback(prm.tile_state, tile, total, &s_bcast);
    uint64_t run = base + texcl;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      const int64_t i = i0 + q;
      run += c[q];
      if (i < T) o_rowptr[i + 1] = (int64_t)run;
    }
    if (tile == num_tiles - 1 && threadIdx.x == 0) prm.meta[SPP_META_EDGES(prm.hop)] = (int64_t)(base + total);
  }
}

// ------------------------------------------------------------------------------------------------
// 2. neighbour choice + hash insert
// ------------------------------------------------------------------------------------------------
template <bool kCol64>
__device__ __forceinline__ int32_t load_col(const void* col, int64_t e) {
  if constexpr (kCol64) return (int32_t)__ldg(reinterpret_cast<const int64_t*>(col) + e);
  else return __ldg(reinterpret_cast<const int32_t*>(col) + e);
}

__device__ __forceinline__ void emit_candidate(const HopParams& prm, int64_t* out_col, int32_t node, uint32_t Tbase, int64_t p) {
  const uint32_t slot = table_insert_max(prm.tab, node, ~(Tbase + (uint32_t)p));
  out_col[p] = (int64_t)slot;
}

// kMode 1: with replacement; 2: without replacement (Floyd).  (Full neighbourhood is edge parallel:
// k_hop_sample_edges below.)
template <int kMode, int G, int PPL, bool kCol64>
__global__ void __launch_bounds__(kSampleThreads) k_hop_sample(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  const int64_t E = prm.meta[SPP_META_EDGES(prm.hop)];
  if (E > o_cap || (uint64_t)T + (uint64_t)E >= 0xFFFFFFF0ull) {
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.meta[SPP_META_OVERFLOW] = 1;
    return;
  }
  const uint32_t Tbase = (uint32_t)T;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);   // lane inside the group
  const int gbase = lane & ~(G - 1);
  constexpr int kGroupsPerWarp = 32 / G;
  const int64_t warp_global = ((int64_t)blockIdx.x * kSampleThreads + threadIdx.x) >> 5;
  const int64_t warps_total = ((int64_t)gridDim.x * kSampleThreads) >> 5;
  const int k = prm.fanout;

  for (int64_t i0 = warp_global * kGroupsPerWarp; i0 < T; i0 += warps_total * kGroupsPerWarp) {
    const int64_t i = i0 + (lane / G);
    const bool valid = i < T;
    int64_t start = 0, p0 = 0;
    int32_t deg = 0;
    if (valid) {
      start = prm.tgt_start[i];
      deg = prm.tgt_deg[i];
      p0 = o_rowptr[i];
    }
    if constexpr (kMode == 1) {
      const int32_t c = deg > 0 ? k : 0;
      for (int32_t j = gl; j < c; j += G) {
        const uint32_t pick = bounded(rand64(o_seed, (uint32_t)prm.hop, (uint64_t)i, (uint32_t)j), (uint32_t)deg);
        emit_candidate(prm, o_col, load_col<kCol64>(prm.col, start + pick), Tbase, p0 + j);
      }
    } else {
      const bool need = valid && deg > k;
      const int32_t basej = deg - k;
      uint32_t myr[PPL], mypick[PPL];
#pragma unroll
      for (int q = 0; q < PPL; ++q) {
        const int j = gl + G * q;
        myr[q] = 0;
        mypick[q] = 0xffffffffu;
        if (need && j < k)
          myr[q] = bounded(rand64(o_seed, (uint32_t)prm.hop, (uint64_t)i, (uint32_t)j), (uint32_t)(basej + j) + 1u);
      }
      if (__any_sync(kFullMask, need)) {
        // Floyd: step s draws t in [0, basej+s]; already chosen -> take basej+s instead.
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
#pragma unroll 1
          for (int sl = 0; sl < G; ++sl) {
            const int s = q * G + sl;
            if (s >= k) break;
            const uint32_t t = __shfl_sync(kFullMask, myr[q], gbase + sl);
            bool hit = false;
#pragma unroll
            for (int q2 = 0; q2 < PPL; ++q2) hit |= (gl + G * q2 < s) && (mypick[q2] == t);
            const uint32_t b = __ballot_sync(kFullMask, need && hit);
            const uint32_t gm = (G == 32) ? b : ((b >> gbase) & ((1u << G) - 1u));
            const uint32_t winner = gm ? (uint32_t)(basej + s) : t;
            if (gl == sl) mypick[q] = winner;
          }
        }
      }
      const int32_t c = deg < k ? deg : k;
      int32_t node[PPL];
#pragma unroll
      for (int q = 0; q < PPL; ++q) {
        const int j = gl + G * q;
        if (valid && j < c) {
          const uint32_t pick = need ? mypick[q] : (uint32_t)j;
          node[q] = load_col<kCol64>(prm.col, start + pick);
        }
      }
#pragma unroll
      for (int q = 0; q < PPL; ++q) {
        const int j = gl + G * q;
        if (valid && j < c) emit_candidate(prm, o_col, node[q], Tbase, p0 + j);
      }
    }
  }
}

// Full neighbourhood (fanout < 0), edge parallel: one thread per candidate position p.  The owning
// row is found by a binary search in out_rowptr (the exclusive scan k_hop_count_scan just wrote),
// so a hub with 10^5 neighbours is spread over 10^5 threads instead of looping in one warp, and
// consecutive lanes read consecutive col entries.
template <bool kCol64>
__global__ void __launch_bounds__(kSampleThreads) k_hop_sample_edges(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  const int64_t E = prm.meta[SPP_META_EDGES(prm.hop)];
  if (E > o_cap || (uint64_t)T + (uint64_t)E >= 0xFFFFFFF0ull) {
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.meta[SPP_META_OVERFLOW] = 1;
    return;
  }
  const uint32_t Tbase = (uint32_t)T;
  const int64_t stride = (int64_t)gridDim.x * kSampleThreads;
  for (int64_t p = (int64_t)blockIdx.x * kSampleThreads + threadIdx.x; p < E; p += stride) {
    int64_t lo = 0, hi = T;  // out_rowptr[lo] <= p < out_rowptr[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(o_rowptr + mid) <= p) lo = mid;
      else hi = mid;
    }
    const int64_t j = p - __ldg(o_rowptr + lo);
    emit_candidate(prm, o_col, load_col<kCol64>(prm.col, prm.tgt_start[lo] + j), Tbase, p);
  }
}

// ------------------------------------------------------------------------------------------------
// 3. order-preserving compaction of first discoverers -> new local ids
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) k_hop_compact(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  __shared__ uint64_t s_warp[kScanThreads / 32];
  __shared__ uint64_t s_bcast;
  __shared__ int64_t s_tile;
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  int64_t E = prm.meta[SPP_META_EDGES(prm.hop)];
  if (E > o_cap || (uint64_t)T + (uint64_t)E >= 0xFFFFFFF0ull) E = 0;  // overflow already flagged
  const uint32_t Tbase = (uint32_t)T;
  const int64_t num_tiles = (E + kScanTile - 1) / kScanTile;
  if (E == 0 && blockIdx.x == 0 && threadIdx.x == 0) prm.meta[SPP_META_NODES(prm.hop + 1)] = T;
  while (true) {
    const int64_t tile = next_tile(prm.tile_state, &s_tile);
    if (tile >= num_tiles) break;
    const int64_t p0 = tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t slot[kScanItems];
    uint64_t ent[kScanItems];
    uint32_t flags = 0, cnt = 0;
    const uint64_t* tab64 = reinterpret_cast<const uint64_t*>(prm.tab.w);
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) slot[q] = (p0 + q < E) ? (uint32_t)o_col[p0 + q] : 0u;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) ent[q] = __ldcg(tab64 + slot[q]);
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      const int64_t p = p0 + q;
      if (p < E && (uint32_t)(ent[q] >> 32) == ~(Tbase + (uint32_t)p)) {
        flags |= 1u << q;
        ++cnt;
      }
    }
    uint64_t total;
    const uint64_t texcl = block_excl_scan(cnt, s_warp, total);
    const uint64_t base = lookback(prm.tile_state, tile, total, &s_bcast);
    uint64_t r = base + texcl;
#pragma unroll
    for (int q = 0; q < kScanItems; ++q) {
      if (flags & (1u << q)) {
        const int64_t L = T + (int64_t)r;
        if (L < prm.max_nodes) {
          prm.n_ids[L] = (int32_t)((uint32_t)ent[q] - 1u);
          __stcg(prm.tab.w + 2 * (size_t)slot[q] + 1, ~(uint32_t)L);
        }
        ++r;
      }
    }
    if (tile == num_tiles - 1 && threadIdx.x == 0) {
      int64_t S = T + (int64_t)(base + total);
      if (S > prm.max_nodes) {
        prm.meta[SPP_META_OVERFLOW] = 1;
        S = prm.max_nodes;
      }
      prm.meta[SPP_META_NODES(prm.hop + 1)] = S;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 4. relabel + per-row ascending sort
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ int32_t group_bitonic_sort(int32_t v, int gl) {
#pragma unroll
  for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
    for (int d = k >> 1; d > 0; d >>= 1) {
      const int32_t o = __shfl_xor_sync(kFullMask, v, d);
      const bool lower = (gl & d) == 0;
      const bool asc = (gl & k) == 0 || k == G;
      const bool take_min = lower == asc;
      v = take_min ? (v < o ? v : o) : (v > o ? v : o);
    }
  }
  return v;
}

// every row has at most G entries (sampled hops with fanout <= 32)
template <int G>
__global__ void __launch_bounds__(kSampleThreads) k_relabel_sort_small(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  if (prm.meta[SPP_META_OVERFLOW]) return;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  constexpr int kGroupsPerWarp = 32 / G;
  const int64_t warp_global = ((int64_t)blockIdx.x * kSampleThreads + threadIdx.x) >> 5;
  const int64_t warps_total = ((int64_t)gridDim.x * kSampleThreads) >> 5;
  for (int64_t i0 = warp_global * kGroupsPerWarp; i0 < T; i0 += warps_total * kGroupsPerWarp) {
    const int64_t i = i0 + (lane / G);
    int64_t p0 = 0;
    int n = 0;
    if (i < T) {
      p0 = o_rowptr[i];
      n = (int)(o_rowptr[i + 1] - p0);
    }
    int32_t v = 0x7fffffff;
    if (gl < n) {
      const uint32_t slot = (uint32_t)o_col[p0 + gl];
      v = (int32_t)~__ldcg(prm.tab.w + 2 * (size_t)slot + 1);
    }
    v = group_bitonic_sort<G>(v, gl);
    if (gl < n) o_col[p0 + gl] = (int64_t)v;
  }
}

// compare-exchange network that sorts n (arbitrary) elements ascending: bitonic sort with the
// "flip" first stage so every comparator is ascending and the virtual +inf padding never moves
template <typename T, typename SyncF>
__device__ __forceinline__ void network_sort(T* a, int64_t n, int tid, int nthreads, SyncF sync) {
  for (int64_t k = 2; (k >> 1) < n; k <<= 1) {
    for (int64_t i = tid; i < n; i += nthreads) {
      const int64_t j = i ^ (k - 1);
      if (j > i && j < n) {
        const T x = a[i], y = a[j];
        if (y < x) { a[i] = y; a[j] = x; }
      }
    }
    sync();
    for (int64_t d = k >> 2; d > 0; d >>= 1) {
      for (int64_t i = tid; i < n; i += nthreads) {
        const int64_t j = i ^ d;
        if (j > i && j < n) {
          const T x = a[i], y = a[j];
          if (y < x) { a[i] = y; a[j] = x; }
        }
      }
      sync();
    }
  }
}

// general rows: <= 32 in registers, <= kWarpSortCap in the warp's shared-memory slice, longer rows
// are relabelled in place and queued (row index into tgt_deg, count in meta[kMetaWork])
constexpr int kGeneralWarps = 4;
__global__ void __launch_bounds__(kGeneralWarps * 32) k_relabel_sort_general(const __grid_constant__ HopParams prm) {
  SPP_HOP_LOCALS
  __shared__ int32_t s_buf[kGeneralWarps][kWarpSortCap];
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  if (prm.meta[SPP_META_OVERFLOW]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * kGeneralWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kGeneralWarps;
  int32_t* buf = s_buf[warp];
  for (int64_t i = warp_global; i < T; i += warps_total) {
    const int64_t p0 = o_rowptr[i];
    const int64_t n = o_rowptr[i + 1] - p0;
    if (n <= 32) {
      int32_t v = 0x7fffffff;
      if (lane < n) {
        const uint32_t slot = (uint32_t)o_col[p0 + lane];
        v = (int32_t)~__ldcg(prm.tab.w + 2 * (size_t)slot + 1);
      }
      v = group_bitonic_sort<32>(v, lane);
      if (lane < n) o_col[p0 + lane] = (int64_t)v;
    } else if (prm.bitmap_rows) {  // k_sort_rows_bitmap relabels and sorts the row with a whole CTA
      if (lane == 0) {
        const unsigned long long w = atomicAdd((unsigned long long*)(prm.meta + kMetaWork), 1ull);
        prm.tgt_deg[w] = (int32_t)i;
      }
    } else if (n <= kWarpSortCap) {
      for (int64_t j = lane; j < n; j += 32) {
        const uint32_t slot = (uint32_t)o_col[p0 + j];
        buf[j] = (int32_t)~__ldcg(prm.tab.w + 2 * (size_t)slot + 1);
      }
      __syncwarp();
      network_sort(buf, n, lane, 32, [] { __syncwarp(); });
      for (int64_t j = lane; j < n; j += 32) o_col[p0 + j] = (int64_t)buf[j];
      __syncwarp();
    } else {
      for (int64_t j = lane; j < n; j += 32) {
        const uint32_t slot = (uint32_t)o_col[p0 + j];
        o_col[p0 + j] = (int64_t)(int32_t)~__ldcg(prm.tab.w + 2 * (size_t)slot + 1);
      }
      if (lane == 0) {
        const unsigned long long w = atomicAdd((unsigned long long*)(prm.meta + kMetaWork), 1ull);
        prm.tgt_deg[w] = (int32_t)i;  // tgt_deg is dead after k_hop_sample: reuse as the work list
      }
    }
  }
}

// Rows of a full-neighbourhood hop longer than 32 entries, one CTA per row: the row's local ids
// are a (nearly always duplicate-free) subset of [0, S), so they are sorted by setting their bits
// in a shared-memory bitmap and enumerating the set bits in order -- O(S/32 + n) per row instead
// of the O(n log^2 n) of a comparison network, which matters for hub rows of 10^4..10^5 entries.
// `cap_bits` = bitmap capacity of this launch (two launches: a small bitmap with many CTAs per SM
// and a large one); the launch whose range [min_bits, cap_bits) contains S does the work.  A row
// with a repeated neighbour (multigraph) is relabelled in place and handed on to
// k_sort_large_rows through the second work list, like every row when S exceeds both bitmaps.
__global__ void __launch_bounds__(256) k_sort_rows_bitmap(const __grid_constant__ HopParams prm, const int64_t min_bits,
                                                          const int64_t cap_bits, const int last) {
  SPP_HOP_LOCALS
  extern __shared__ uint32_t s_bits[];
  __shared__ uint32_t s_wsum[8];
  if (prm.meta[SPP_META_OVERFLOW]) return;
  const int64_t S = prm.meta[SPP_META_NODES(prm.hop + 1)];
  const bool mine = S > min_bits && S <= cap_bits;
  if (!mine && !(last && S > cap_bits)) return;
  const int64_t work = prm.meta[kMetaWork];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* list2 = reinterpret_cast<int32_t*>(prm.tgt_start);  // dead after the sampling kernel
  const int words = (int)((S + 31) >> 5);
  const int per = (words + 255) / 256;  // bitmap words per thread in the enumeration
  for (int64_t w = blockIdx.x; w < work; w += gridDim.x) {
    const int64_t i = prm.tgt_deg[w];
    const int64_t p0 = o_rowptr[i];
    const int64_t n = o_rowptr[i + 1] - p0;
    int64_t* row = o_col + p0;
    bool dup = !mine;
    if (mine)
      for (int t = tid; t < words; t += 256) s_bits[t] = 0u;
    __syncthreads();
    for (int64_t j = tid; j < n; j += 256) {
      const uint32_t slot = (uint32_t)row[j];
      const uint32_t l = ~__ldcg(prm.tab.w + 2 * (size_t)slot + 1);
      row[j] = (int64_t)l;  // relabelled in place (what the comparison sorter expects)
      if (mine) {
        const uint32_t bit = 1u << (l & 31);
        if (atomicOr(&s_bits[l >> 5], bit) & bit) dup = true;
      }
    }
    if (__syncthreads_or(dup)) {
      if (tid == 0) {
        const unsigned long long q = atomicAdd((unsigned long long*)(prm.meta + kMetaWork2), 1ull);
        list2[q] = (int32_t)i;
      }
      continue;
    }
    // enumerate the set bits in order: thread t owns words [t * per, (t + 1) * per)
    const int w0 = tid * per, w1 = (w0 + per < words) ? w0 + per : words;
    uint32_t cnt = 0;
    for (int t = w0; t < w1; ++t) cnt += __popc(s_bits[t]);
    uint32_t inc = warp_incl_scan(cnt, lane);
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < warp) base += s_wsum[q];
    uint32_t o = base + inc - cnt;
    for (int t = w0; t < w1; ++t) {
      uint32_t m = s_bits[t];
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        row[o++] = (int64_t)(((uint32_t)t << 5) + (uint32_t)b);
      }
    }
    __syncthreads();  // s_bits / s_wsum are reused by the next row
  }
}

// one CTA per queued long row (comparison network).  second_list: rows handed on by
// k_sort_rows_bitmap (already relabelled), otherwise the rows queued by k_relabel_sort_general
__global__ void __launch_bounds__(256) k_sort_large_rows(const __grid_constant__ HopParams prm, const int second_list) {
  SPP_HOP_LOCALS
  extern __shared__ int32_t s_big[];
  const int64_t work = prm.meta[second_list ? kMetaWork2 : kMetaWork];
  const int32_t* list = second_list ? reinterpret_cast<const int32_t*>(prm.tgt_start) : prm.tgt_deg;
  for (int64_t w = blockIdx.x; w < work; w += gridDim.x) {
    const int64_t i = list[w];
    const int64_t p0 = o_rowptr[i];
    const int64_t n = o_rowptr[i + 1] - p0;
    int64_t* row = o_col + p0;
    if (n <= kBlockSortSmemElems) {
      for (int64_t j = threadIdx.x; j < n; j += blockDim.x) s_big[j] = (int32_t)row[j];
      __syncthreads();
      network_sort(s_big, n, threadIdx.x, blockDim.x, [] { __syncthreads(); });
      for (int64_t j = threadIdx.x; j < n; j += blockDim.x) row[j] = (int64_t)s_big[j];
      __syncthreads();
    } else {
      __syncthreads();
      network_sort(row, n, threadIdx.x, blockDim.x, [] { __syncthreads(); });
    }
  }
}

template <typename OutT>
__global__ void k_export_nids(const spp_device_job* job, const int32_t* __restrict__ n_ids, const int64_t* __restrict__ meta,
                              int word, int64_t max_nodes, OutT* __restrict__ out) {
  if (job != nullptr) out = reinterpret_cast<OutT*>(job->n_id_out);
  int64_t n = meta[word];
  if (n > max_nodes) n = max_nodes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (OutT)n_ids[i];
}

// ================================================================================================
// Fused path for sampled hops with 1 <= fanout <= 32 (every reference configuration):
//   A. k_hop_sample_fused   reads rowptr itself (no separate degree pass), candidate (i, j) lives at
//                           the fixed-stride virtual position v = i * k + j (ordering key T + v),
//                           slot of each candidate -> cand[v] (kInvalidCand for j >= kept count);
//                           optimistic CAS insert: one L2 round trip per candidate.
//   B. k_hop_compact_fused  one pass over the virtual positions: kept-edge count and first-discoverer
//                           count scanned together -> out_rowptr, n_ids, new local ids.  The scan is
//                           "aggregate only": a tile publishes its own totals at once and sums the
//                           totals of all earlier tiles in parallel, so there is no serial prefix
//                           chain (the tile counts here are <= ~1000).
//   C. k_relabel_sort_fused slot -> local id, register bitonic sort per row, compacted int64 row.
// ================================================================================================
constexpr uint32_t kInvalidCand = 0xFFFFFFFFu;
// The compaction pass rewrites cand[v] to (local id | kResolvedCand) for every candidate whose local id
// it already knows -- first discoverers (it assigns the id) and nodes of earlier hops (the entry holds
// ~local with local < T) -- so the relabel pass only goes back to the table for the repeats of nodes
// that are new in this hop (slots are < 2^31, local ids far below: the flag bit is free).
constexpr uint32_t kResolvedCand = 0x80000000u;
constexpr int kFusedItems = 8;
constexpr int kFusedTile = kScanThreads * kFusedItems;  // 2048 virtual positions per tile

struct FusedParams {
  HopParams h;
  uint32_t* cand;    // uint32[T * k] slot of every virtual candidate
  int64_t cand_cap;  // capacity of cand (elements)
  uint32_t epoch;    // tag of this launch's tile aggregates
  unsigned long long* timeline;  // diagnostics (spp_debug_set_timeline): 8 globaltimer stamps per tile
};

// Lane groups of exactly k lanes (32 / k targets per warp) and a three-stage software pipeline
// per group: [n_ids -> rowptr of target t+2 / t+1] | [picks + col load of target t+1] | [table CAS +
// atomicMax of target t], so the random col read and the L2 atomics of consecutive targets overlap.
template <bool kCol64>
__global__ void __launch_bounds__(kSampleThreads) k_hop_sample_fused(const __grid_constant__ FusedParams fp) {
  const HopParams& prm = fp.h;
  SPP_HOP_LOCALS
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  const int k = prm.fanout;
  if ((uint64_t)T * (uint64_t)k > (uint64_t)fp.cand_cap || (uint64_t)T * (uint64_t)(k + 1) >= 0xFFFFFFF0ull) {
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.meta[SPP_META_OVERFLOW] = 1;
    return;
  }
  if (blockIdx.x == 0) {
    // ticket counter of k_hop_compact_fused, and every aggregate word that kernel will poll: it runs
    // after this one in stream order, so no stale aggregate of an earlier launch (whatever its epoch
    // tag, even after the 2^30 tags have wrapped) can ever be taken for a current one
    if (threadIdx.x == 0) prm.tile_state[0] = 0;
    const int64_t tiles = ((int64_t)T * k + kFusedTile - 1) / kFusedTile;
    for (int64_t t = threadIdx.x; t < tiles; t += kSampleThreads) prm.tile_state[2 + t] = 0;
  }
  const uint32_t Tbase = (uint32_t)T;
  const int lane = threadIdx.x & 31;
  const int gpw = 32 / k;                 // groups (targets) per warp
  const int g = lane / k;                 // my group
  const bool lane_on = g < gpw;           // tail lanes of the warp idle when 32 % k != 0
  const int gl = lane - g * k;            // lane inside the group
  const int gbase = lane_on ? g * k : 0;
  const uint32_t kmask = k == 32 ? 0xffffffffu : ((1u << k) - 1u);
  const int64_t warp_global = ((int64_t)blockIdx.x * kSampleThreads + threadIdx.x) >> 5;
  const int64_t stride = (((int64_t)gridDim.x * kSampleThreads) >> 5) * gpw;

  int64_t i0 = warp_global * gpw;  // warp-uniform index of the warp's first target this round
  // stage A registers: rowptr pair of target (i0 + g), node id of target (i0 + stride + g)
  int64_t nstart = 0, nend = 0;
  int32_t n_next = 0;
  {
    const int64_t i = i0 + g;
    if (lane_on && i < T) {
      const int32_t n = prm.n_ids[i];
      nstart = __ldg(prm.rowptr + n);
      nend = __ldg(prm.rowptr + n + 1);
    }
    if (lane_on && i + stride < T) n_next = prm.n_ids[i + stride];
  }
  // stage C registers (candidate whose col value is in flight / has landed)
  bool c_valid = false;
  int32_t c_node = 0;
  uint32_t c_v = 0;

  while (true) {
    const bool have = i0 < T;  // warp-uniform
    bool b_valid = false;
    int32_t b_node = 0;
    uint32_t b_v = 0;
    if (have) {
      // ---- stage B: picks of target i0 + g, col load issued (not consumed here) ----------------
      const int64_t i = i0 + g;
      const bool valid = lane_on && i < T;
      const int64_t start = nstart;
      const int32_t deg = (int32_t)(nend - nstart);
      const bool need = valid && deg > k;
      const int32_t basej = deg - k;
      uint32_t myr = 0, mypick = 0xffffffffu;
      if (need) myr = bounded(rand64(o_seed, (uint32_t)prm.hop, (uint64_t)i, (uint32_t)gl), (uint32_t)(basej + gl) + 1u);
      if (__any_sync(kFullMask, need)) {
#pragma unroll 1
        for (int s = 0; s < k; ++s) {  // Floyd: t uniform on [0, basej+s]; taken already -> basej+s
          const uint32_t t = __shfl_sync(kFullMask, myr, gbase + s);
          const uint32_t b = __ballot_sync(kFullMask, need && gl < s && mypick == t);
          if (gl == s) mypick = ((b >> gbase) & kmask) ? (uint32_t)(basej + s) : t;
        }
      }
      if (valid) {
        const int32_t c = deg < k ? deg : k;
        b_v = (uint32_t)(i * k + gl);
        if (gl < c) {
          const uint32_t pick = need ? mypick : (uint32_t)gl;
          b_node = load_col<kCol64>(prm.col, start + pick);
          b_valid = true;
        } else {
          fp.cand[b_v] = kInvalidCand;
        }
      }
      // ---- stage A: rowptr of the next target, node id of the one after -----------------------
      const int64_t in = i + stride;
      if (lane_on && in < T) {
        nstart = __ldg(prm.rowptr + n_next);
        nend = __ldg(prm.rowptr + n_next + 1);
      }
      if (lane_on && in + stride < T) n_next = prm.n_ids[in + stride];
    }
    // ---- stage C: insert the previous round's candidate ------------------------------------------
    if (c_valid) {
      fp.cand[c_v] = table_insert_max(prm.tab, c_node, ~(Tbase + c_v));
    }
    if (!have) break;
    c_valid = b_valid;
    c_node = b_node;
    c_v = b_v;
    i0 += stride;
  }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SPP_STAMP(i)                                                                         \
  do {                                                                                       \
    if (fp.timeline != nullptr && threadIdx.x == 0) fp.timeline[tile * 8 + (i)] = globaltimer_ns(); \
  } while (0)

// tag of this launch's tile aggregates: a kernel parameter for plain launches; under graph replay
// (where parameters are frozen at capture time) it comes from the device job block, which the host
// advances for every mini-batch
__device__ __forceinline__ uint32_t scan_epoch(const FusedParams& fp) {
  if (fp.h.job == nullptr) return fp.epoch;
  return (uint32_t)((fp.h.job->scan_epoch + (uint64_t)fp.h.hop) % 0x3FFFFFFFull) + 1u;
}

__global__ void __launch_bounds__(kScanThreads) k_hop_compact_fused(const __grid_constant__ FusedParams fp) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_red[kScanThreads / 32];
  __shared__ unsigned long long s_base;
  __shared__ int64_t s_tile;
  const HopParams& prm = fp.h;
  SPP_HOP_LOCALS
  const uint32_t epoch = scan_epoch(fp);
  // first ticket: issued at once, in flight together with the meta loads.  The ticket counter
  // (tile_state[0]) was zeroed by this hop's k_hop_sample_fused, which precedes us in stream order.
  if (threadIdx.x == 0) s_tile = (int64_t)atomicAdd((unsigned long long*)prm.tile_state, 1ull);
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  const int k = prm.fanout;
  int64_t V = T * k;
  if (V > fp.cand_cap || (uint64_t)T * (uint64_t)(k + 1) >= 0xFFFFFFF0ull) V = 0;  // overflow flagged by the sampler
  const uint32_t Tbase = (uint32_t)T;
  const int64_t num_tiles = (V + kFusedTile - 1) / kFusedTile;
  unsigned long long* ctr = (unsigned long long*)prm.tile_state;        // [0] ticket counter
  uint64_t* agg = prm.tile_state + 2;                                   // [2 + t] epoch | kept | new
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    o_rowptr[0] = 0;
    if (V == 0) {
      prm.meta[SPP_META_EDGES(prm.hop)] = 0;
      prm.meta[SPP_META_NODES(prm.hop + 1)] = T;
    }
  }
  // rows without a virtual position (V == 0 because of an overflow) are left untouched
  while (true) {
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= num_tiles) break;
    SPP_STAMP(0);
    const int64_t v0 = tile * kFusedTile + (int64_t)threadIdx.x * kFusedItems;
    uint32_t slot[kFusedItems];
    uint64_t ent[kFusedItems];
    uint32_t keptm = 0, newm = 0;
    if (v0 + kFusedItems <= V) {  // 32-byte aligned run of candidates: two 128-bit loads
      const uint4 a = __ldcg(reinterpret_cast<const uint4*>(fp.cand + v0));
      const uint4 b = __ldcg(reinterpret_cast<const uint4*>(fp.cand + v0) + 1);
      slot[0] = a.x; slot[1] = a.y; slot[2] = a.z; slot[3] = a.w;
      slot[4] = b.x; slot[5] = b.y; slot[6] = b.z; slot[7] = b.w;
    } else {
#pragma unroll
      for (int q = 0; q < kFusedItems; ++q) slot[q] = (v0 + q < V) ? fp.cand[v0 + q] : kInvalidCand;
    }
    // all table entries {key+1, ~local} of the thread's candidates as independent 64-bit loads
    // (the key is needed later for n_ids; loading it here keeps it off the dependent path)
    const uint64_t* tab64 = reinterpret_cast<const uint64_t*>(prm.tab.w);
#pragma unroll
    for (int q = 0; q < kFusedItems; ++q) ent[q] = __ldcg(tab64 + (slot[q] != kInvalidCand ? slot[q] : 0u));
#pragma unroll
    for (int q = 0; q < kFusedItems; ++q) {
      if (slot[q] != kInvalidCand) {
        keptm |= 1u << q;
        if ((uint32_t)(ent[q] >> 32) == ~(Tbase + (uint32_t)(v0 + q))) newm |= 1u << q;
      }
    }
    SPP_STAMP(1);
    // packed (kept << 16 | new) block scan: per-tile sums are <= 2048 each
    const uint32_t mine = ((uint32_t)__popc(keptm) << 16) | (uint32_t)__popc(newm);
    uint32_t inc = warp_incl_scan(mine, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
      const uint32_t x = s_warp[w];
      if (w < warp) wbase += x;
      total += x;
    }
    const uint32_t texcl = wbase + inc - mine;
    if (threadIdx.x == 0) st_volatile_u64(agg + tile, ((uint64_t)epoch << 32) | (uint64_t)total);
    SPP_STAMP(2);
    // Sum of the aggregates of every earlier tile (each is published independently of its owner's
    // own wait, so there is no serial chain).  Every thread reads its share of the flags at once
    // (one L2 round trip for the whole CTA); only flags that are not there yet are polled again,
    // with a back-off, so waiting CTAs do not flood L2 while earlier tiles still load their
    // table entries (unthrottled polling by 256 threads per CTA tripled the kernel time).
    {
      unsigned long long acc = 0;  // kept in the high 32 bits, new in the low 32 bits
      for (int64_t t = threadIdx.x; t < tile; t += kScanThreads) {
        uint64_t w = ld_volatile_u64(agg + t);
        while ((uint32_t)(w >> 32) != epoch) {
          __nanosleep(64);
          w = ld_volatile_u64(agg + t);
        }
        acc += ((unsigned long long)((uint32_t)w >> 16) << 32) | (unsigned long long)((uint32_t)w & 0xffffu);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFullMask, acc, d);
      if (lane == 0) s_red[warp] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = 0;
#pragma unroll
      for (int w = 0; w < kScanThreads / 32; ++w) b += s_red[w];
      s_base = b;
    }
    __syncthreads();
    const unsigned long long base = s_base;
    SPP_STAMP(3);
    uint64_t kept_run = (base >> 32) + (texcl >> 16);
    uint64_t new_run = (base & 0xffffffffull) + (texcl & 0xffffu);
    // row bookkeeping: v = i * k + r
    int64_t row = (int64_t)((uint32_t)v0 / (uint32_t)k);  // V < 2^32 (checked above)
    int r = (int)(v0 - row * k);
#pragma unroll
    for (int q = 0; q < kFusedItems; ++q) {
      if (v0 + q < V) {
        if (keptm & (1u << q)) {
          ++kept_run;
          const uint32_t local = ~(uint32_t)(ent[q] >> 32);
          if (local < Tbase) slot[q] = local | kResolvedCand;  // node of an earlier hop
        }
        if (newm & (1u << q)) {
          const int64_t L = T + (int64_t)new_run;
          if (L < prm.max_nodes) {
            prm.n_ids[L] = (int32_t)((uint32_t)ent[q] - 1u);
            __stcg(prm.tab.w + 2 * (size_t)slot[q] + 1, ~(uint32_t)L);
            slot[q] = (uint32_t)L | kResolvedCand;             // first discoverer: id assigned here
          }
          ++new_run;
        }
        if (r == k - 1) o_rowptr[row + 1] = (int64_t)kept_run;
        if (++r == k) {
          r = 0;
          ++row;
        }
      }
    }
    // candidates with their resolution back to cand (same 32-byte run this thread loaded)
    if (v0 + kFusedItems <= V) {
      uint4* cw = reinterpret_cast<uint4*>(fp.cand + v0);
      cw[0] = make_uint4(slot[0], slot[1], slot[2], slot[3]);
      cw[1] = make_uint4(slot[4], slot[5], slot[6], slot[7]);
    } else {
#pragma unroll
      for (int q = 0; q < kFusedItems; ++q)
        if (v0 + q < V) fp.cand[v0 + q] = slot[q];
    }
    if (tile == num_tiles - 1 && threadIdx.x == 0) {
      const uint64_t kept_total = (base >> 32) + (total >> 16);
      int64_t S = T + (int64_t)((base & 0xffffffffull) + (total & 0xffffu));
      if (S > prm.max_nodes) {
        prm.meta[SPP_META_OVERFLOW] = 1;
        S = prm.max_nodes;
      }
      if ((int64_t)kept_total > o_cap) prm.meta[SPP_META_OVERFLOW] = 1;  // out_col too small
      prm.meta[SPP_META_EDGES(prm.hop)] = (int64_t)kept_total;
      prm.meta[SPP_META_NODES(prm.hop + 1)] = S;
    }
    SPP_STAMP(4);
    if (num_tiles <= (int64_t)gridDim.x) break;  // every tile has its own CTA: no second ticket
    __syncthreads();                               // s_tile / s_base consumed
    if (threadIdx.x == 0) s_tile = (int64_t)atomicAdd(ctr, 1ull);
  }
}

template <int G>
__global__ void __launch_bounds__(kSampleThreads) k_relabel_sort_fused(const __grid_constant__ FusedParams fp) {
  const HopParams& prm = fp.h;
  SPP_HOP_LOCALS
  int64_t T = prm.meta[SPP_META_NODES(prm.hop)];
  if (T > prm.max_targets) T = prm.max_targets;
  if (prm.meta[SPP_META_OVERFLOW]) return;
  const int k = prm.fanout;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  constexpr int kGroupsPerWarp = 32 / G;
  const int64_t warp_global = ((int64_t)blockIdx.x * kSampleThreads + threadIdx.x) >> 5;
  const int64_t warps_total = ((int64_t)gridDim.x * kSampleThreads) >> 5;
  for (int64_t i0 = warp_global * kGroupsPerWarp; i0 < T; i0 += warps_total * kGroupsPerWarp) {
    const int64_t i = i0 + (lane / G);
    int64_t p0 = 0;
    int n = 0;
    if (i < T) {
      p0 = o_rowptr[i];
      n = (int)(o_rowptr[i + 1] - p0);
    }
    int32_t v = 0x7fffffff;
    if (gl < n) {
      const uint32_t c = fp.cand[i * k + gl];
      v = (c & kResolvedCand) ? (int32_t)(c & ~kResolvedCand) : (int32_t)~__ldcg(prm.tab.w + 2 * (size_t)c + 1);
    }
    v = group_bitonic_sort<G>(v, gl);
    if (gl < n) o_col[p0 + gl] = (int64_t)v;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int check_ws(const spp_sampler_ws* ws) {
  if (!ws) return fail(SPP_EINVAL, "sampler: null workspace");
  if (!ws->table || !ws->n_ids || !ws->tgt_start || !ws->tgt_deg || !ws->tile_state || !ws->meta)
    return fail(SPP_EINVAL, "sampler: workspace has null members");
  if (ws->table_direct) {
    if (ws->table_slots < 1 || ws->table_slots > (1ll << 31))
      return fail(SPP_EINVAL, "sampler: direct table needs 1 <= table_slots <= 2^31");
    return 0;  // table_slots >= num_nodes is checked where the graph is known
  }
  if (ws->table_slots < 2 || ws->table_slots > (1ll << 31))
    return fail(SPP_EINVAL, "sampler: table_slots must lie in [2, 2^31]");
  if (ws->table_slots * 4 < 5 * ws->max_nodes)
    return fail(SPP_ECAPACITY, "sampler: table_slots (%lld) < 1.25 * max_nodes (%lld)", (long long)ws->table_slots,
                (long long)ws->max_nodes);
  return 0;
}

static Table make_table(const spp_sampler_ws* ws) {
  Table t;
  t.w = reinterpret_cast<uint32_t*>(ws->table);
  t.slots = (uint32_t)(ws->table_slots >= (1ll << 32) ? 0xFFFFFFFFll : ws->table_slots);
  t.direct = ws->table_direct != 0;
  return t;
}

static HopParams make_params(const spp_graph* g, const spp_sampler_ws* ws, int hop, int32_t fanout, int replace,
                             uint64_t rng_seed, int64_t max_targets, int64_t max_edges, int64_t* out_rowptr,
                             int64_t* out_col, const spp_device_job* job = nullptr) {
  HopParams p{};
  p.job = job;
  p.rowptr = g->rowptr;
  p.col = g->col;
  p.n_ids = ws->n_ids;
  p.tgt_start = ws->tgt_start;
  p.tgt_deg = ws->tgt_deg;
  p.meta = ws->meta;
  p.tile_state = ws->tile_state;
  p.out_rowptr = out_rowptr;
  p.out_col = out_col;
  p.tab = make_table(ws);
  p.max_targets = max_targets < ws->max_targets ? max_targets : ws->max_targets;
  p.max_edges = max_edges;
  p.max_nodes = ws->max_nodes;
  p.premixed = premix_seed(rng_seed);
  p.hop = hop;
  p.fanout = fanout;
  p.replace = replace;
  return p;
}

static int scan_grid(int64_t bound_items) {
  int64_t tiles = ceil_div(bound_items > 0 ? bound_items : 1, kScanTile);
  int64_t cap = (int64_t)num_sms() * 4;
  return (int)(tiles < cap ? tiles : cap);
}

// The general (look-back) path owns tile_state[2..] and clears what it uses before every scan;
// words [0] and [1] belong to the fused path (self-resetting tile counter / finished-CTA count).
constexpr int kGeneralTileOffset = 2;

static int reset_tiles(const spp_sampler_ws* ws, int64_t bound_items, cudaStream_t st) {
  int64_t words = 1 + ceil_div(bound_items > 0 ? bound_items : 1, kScanTile);
  if (words + kGeneralTileOffset > ws->tile_words)
    return fail(SPP_ECAPACITY, "sampler: tile_state too small (%lld words needed, %lld given)",
                (long long)(words + kGeneralTileOffset), (long long)ws->tile_words);
  SPP_CUDA(cudaMemsetAsync(ws->tile_state + kGeneralTileOffset, 0, (size_t)words * sizeof(uint64_t), st));
  return 0;
}

static int launch_begin(const spp_graph* g, const int64_t* seeds, int64_t bs, const spp_sampler_ws* ws,
                        cudaStream_t st, const spp_device_job* job = nullptr) {
  if (int r = check_ws(ws)) return r;
  if (!g || !g->rowptr) return fail(SPP_EINVAL, "sampler: null graph");  // col may be NULL when nnz == 0
  if (ws->table_direct && ws->table_slots < g->num_nodes)
    return fail(SPP_ECAPACITY, "sampler: direct table has %lld slots for %lld nodes", (long long)ws->table_slots,
                (long long)g->num_nodes);
  if (bs < 0 || bs > ws->max_nodes) return fail(SPP_ECAPACITY, "sampler: batch_size %lld exceeds max_nodes %lld",
                                                 (long long)bs, (long long)ws->max_nodes);
  if (bs > 0 && !seeds && !job) return fail(SPP_EINVAL, "sampler: null seeds");
  SPP_CUDA(cudaMemsetAsync(ws->table, 0, (size_t)ws->table_slots * sizeof(uint64_t), st));
  trace_mark(kTrTableClear, 0, st);
  k_seeds_init<<<1, 1024, 0, st>>>(job, seeds, bs, ws->n_ids, make_table(ws), ws->meta);
  SPP_KERNEL_CHECK("k_seeds_init");
  trace_mark(kTrSeedsInit, 0, st);
  return 0;
}

static int launch_count(const spp_graph* g, int hop, int32_t fanout, int replace, int64_t max_targets,
                        const spp_sampler_ws* ws, int64_t* out_rowptr, cudaStream_t st, const spp_device_job* job = nullptr) {
  if (hop < 0 || hop >= SPP_MAX_HOPS) return fail(SPP_EINVAL, "sampler: hop %d out of range", hop);
  if (!out_rowptr && !job) return fail(SPP_EINVAL, "sampler: null out_rowptr");
  if (int r = reset_tiles(ws, max_targets, st)) return r;
  HopParams p = make_params(g, ws, hop, fanout, replace, 0, max_targets, 0, out_rowptr, nullptr, job);
  p.tile_state += kGeneralTileOffset;
  k_hop_count_scan<<<scan_grid(p.max_targets), kScanThreads, 0, st>>>(p);
  SPP_KERNEL_CHECK("k_hop_count_scan");
  trace_mark(kTrCount, hop, st);
  return 0;
}

template <int kMode, int G, int PPL>
static void launch_sample_kernel(const HopParams& p, bool col64, int grid, cudaStream_t st) {
  if (col64) k_hop_sample<kMode, G, PPL, true><<<grid, kSampleThreads, 0, st>>>(p);
  else k_hop_sample<kMode, G, PPL, false><<<grid, kSampleThreads, 0, st>>>(p);
}

// opt-in shared-memory sizes of the CTA-per-row sorters (once per device; also called before a
// launch sequence is captured into a CUDA graph)
int sorter_attributes() {
  static bool done[64] = {false};
  int dev = 0;
  SPP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || done[dev]) return 0;
  SPP_CUDA(cudaFuncSetAttribute(k_sort_large_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((size_t)kBlockSortSmemElems * sizeof(int32_t))));
  SPP_CUDA(cudaFuncSetAttribute(k_sort_rows_bitmap, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  done[dev] = true;
  return 0;
}

static int launch_fill(const spp_graph* g, int hop, int32_t fanout, int replace, uint64_t rng_seed,
                       int64_t max_targets, int64_t max_edges, const spp_sampler_ws* ws, const int64_t* out_rowptr,
                       int64_t* out_col, cudaStream_t st, const spp_device_job* job = nullptr) {
  if (hop < 0 || hop >= SPP_MAX_HOPS) return fail(SPP_EINVAL, "sampler: hop %d out of range", hop);
  if (!job && (!out_rowptr || (!out_col && max_edges > 0))) return fail(SPP_EINVAL, "sampler: null output");
  if (fanout >= 0 && !replace && fanout > SPP_MAX_FANOUT)
    return fail(SPP_EUNSUPPORTED, "sampler: fanout %d > SPP_MAX_FANOUT (%d)", fanout, SPP_MAX_FANOUT);
  HopParams p = make_params(g, ws, hop, fanout, replace, rng_seed, max_targets, max_edges,
                            const_cast<int64_t*>(out_rowptr), out_col, job);
  p.tile_state += kGeneralTileOffset;
  const bool c64 = g->col_is_64 != 0;
  const int sms = num_sms();
  // groups needed ~ targets; persistent grid, 8 CTAs / SM at most
  auto grid_for = [&](int G) {
    int64_t warps = ceil_div(p.max_targets > 0 ? p.max_targets : 1, 32 / G);
    int64_t ctas = ceil_div(warps, kSampleThreads / 32);
    int64_t cap = (int64_t)sms * 8;
    return (int)(ctas < cap ? ctas : cap);
  };
  // full neighbourhood: rows longer than 32 are sorted by the bitmap kernels when the node bound
  // fits the large bitmap (the device picks the launch by the actual node count)
  constexpr int64_t kBitsSmall = 128 * 1024, kBitsLarge = 200 * 1024 * 8;
  const bool bitmap = fanout < 0;
  p.bitmap_rows = bitmap ? 1 : 0;
  if (fanout < 0) {
    int64_t ctas = ceil_div(max_edges > 0 ? max_edges : 1, kSampleThreads);
    const int64_t cap = (int64_t)sms * 8;
    const int egrid = (int)(ctas < cap ? ctas : cap);
    if (c64) k_hop_sample_edges<true><<<egrid, kSampleThreads, 0, st>>>(p);
    else k_hop_sample_edges<false><<<egrid, kSampleThreads, 0, st>>>(p);
  } else if (replace) {
    if (fanout <= 8) launch_sample_kernel<1, 8, 1>(p, c64, grid_for(8), st);
    else if (fanout <= 16) launch_sample_kernel<1, 16, 1>(p, c64, grid_for(16), st);
    else launch_sample_kernel<1, 32, 1>(p, c64, grid_for(32), st);
  } else {
    if (fanout <= 4) launch_sample_kernel<2, 4, 1>(p, c64, grid_for(4), st);
    else if (fanout <= 8) launch_sample_kernel<2, 8, 1>(p, c64, grid_for(8), st);
    else if (fanout <= 16) launch_sample_kernel<2, 16, 1>(p, c64, grid_for(16), st);
    else if (fanout <= 32) launch_sample_kernel<2, 32, 1>(p, c64, grid_for(32), st);
    else if (fanout <= 64) launch_sample_kernel<2, 32, 2>(p, c64, grid_for(32), st);
    else launch_sample_kernel<2, 32, 4>(p, c64, grid_for(32), st);
  }
  SPP_KERNEL_CHECK("k_hop_sample");
  trace_mark(kTrSample, hop, st);

  if (int r = reset_tiles(ws, max_edges, st)) return r;
  k_hop_compact<<<scan_grid(max_edges), kScanThreads, 0, st>>>(p);
  SPP_KERNEL_CHECK("k_hop_compact");
  trace_mark(kTrCompact, hop, st);

  const bool small_rows = fanout >= 0 && fanout <= 32;
  if (small_rows) {
    const int G = fanout <= 4 ? 4 : fanout <= 8 ? 8 : fanout <= 16 ? 16 : 32;
    const int grid = grid_for(G);
    switch (G) {
      case 4: k_relabel_sort_small<4><<<grid, kSampleThreads, 0, st>>>(p); break;
      case 8: k_relabel_sort_small<8><<<grid, kSampleThreads, 0, st>>>(p); break;
      case 16: k_relabel_sort_small<16><<<grid, kSampleThreads, 0, st>>>(p); break;
      default: k_relabel_sort_small<32><<<grid, kSampleThreads, 0, st>>>(p); break;
    }
    SPP_KERNEL_CHECK("k_relabel_sort_small");
    trace_mark(kTrRelabel, hop, st);
  } else {
    int64_t ctas = ceil_div(p.max_targets > 0 ? p.max_targets : 1, kGeneralWarps);
    int64_t cap = (int64_t)sms * 8;
    k_relabel_sort_general<<<(int)(ctas < cap ? ctas : cap), kGeneralWarps * 32, 0, st>>>(p);
    SPP_KERNEL_CHECK("k_relabel_sort_general");
    trace_mark(kTrRelabel, hop, st);
    if (fanout < 0 || fanout > kWarpSortCap) {
      const size_t smem = (size_t)kBlockSortSmemElems * sizeof(int32_t);
      if (int r = sorter_attributes()) return r;
      if (bitmap) {
        // small bitmap (16 KB: many CTAs per SM) for the usual batch, large one (200 KB) behind it;
        // whichever launch does not match the node count returns at once
        const int64_t rows_cap = p.max_targets < (int64_t)sms * 8 ? p.max_targets : (int64_t)sms * 8;
        const bool small_only = p.max_targets + max_edges <= kBitsSmall;  // host bound on the node count
        k_sort_rows_bitmap<<<(int)(rows_cap > 0 ? rows_cap : 1), 256, kBitsSmall / 8, st>>>(p, -1, kBitsSmall,
                                                                                            small_only ? 1 : 0);
        SPP_KERNEL_CHECK("k_sort_rows_bitmap");
        if (!small_only) {
          k_sort_rows_bitmap<<<sms, 256, kBitsLarge / 8, st>>>(p, kBitsSmall, kBitsLarge, 1);
          SPP_KERNEL_CHECK("k_sort_rows_bitmap");
        }
        trace_mark(kTrSortBitmap, hop, st);
      }
      k_sort_large_rows<<<sms, 256, smem, st>>>(p, bitmap ? 1 : 0);
      SPP_KERNEL_CHECK("k_sort_large_rows");
      trace_mark(kTrSortLarge, hop, st);
    }
  }
  return 0;
}

// ONE process-wide counter feeds the tags of plain launches (one value per hop launch) and of
// replayed jobs (SPP_MAX_HOPS consecutive values per mini-batch), so a tag cannot repeat on any
// workspace before the counter has advanced by 2^30 - 1
static std::atomic<uint64_t> g_epoch{1};
uint64_t next_scan_epoch() { return g_epoch.fetch_add(SPP_MAX_HOPS, std::memory_order_relaxed); }
static std::atomic<unsigned long long*> g_timeline{nullptr};

static bool fused_ok(int32_t fanout, int replace, const spp_sampler_ws* ws) {
  return fanout >= 1 && fanout <= 32 && !replace && ws->cand != nullptr;
}

// one sampled hop through the fused path (no host synchronisation, no memset)
// `cand_off`: this hop's first word inside ws->cand (every hop of a mini-batch has its own range,
// so a hop's relabel/sort kernel may still be reading its candidates while the next hop samples).
// `relabel_st`: stream of the relabel/sort kernel; when it differs from `st` the kernel is forked
// off behind `fork_ev` and the caller joins later (see sample_minibatch_impl).
static int launch_hop_fused(const spp_graph* g, int hop, int32_t fanout, uint64_t rng_seed, int64_t max_targets,
                            int64_t max_edges, const spp_sampler_ws* ws, int64_t* out_rowptr, int64_t* out_col,
                            cudaStream_t st, int64_t cand_off = 0, cudaStream_t relabel_st = nullptr,
                            cudaEvent_t fork_ev = nullptr, const spp_device_job* job = nullptr) {
  if (hop < 0 || hop >= SPP_MAX_HOPS) return fail(SPP_EINVAL, "sampler: hop %d out of range", hop);
  if (!job && (!out_rowptr || (!out_col && max_edges > 0))) return fail(SPP_EINVAL, "sampler: null output");
  FusedParams fp{};
  fp.h = make_params(g, ws, hop, fanout, 0, rng_seed, max_targets, max_edges, out_rowptr, out_col, job);
  const int64_t vmax = fp.h.max_targets * (int64_t)fanout;
  if (cand_off < 0 || (cand_off & 7) || cand_off + vmax > ws->cand_words)
    return fail(SPP_ECAPACITY, "sampler: cand buffer too small (%lld needed, %lld given)",
                (long long)(cand_off + vmax), (long long)ws->cand_words);
  const int64_t tiles = ceil_div(vmax > 0 ? vmax : 1, kFusedTile);
  if (2 + tiles > ws->tile_words)
    return fail(SPP_ECAPACITY, "sampler: tile_state too small (%lld words needed)", (long long)(2 + tiles));
  fp.cand = reinterpret_cast<uint32_t*>(ws->cand) + cand_off;
  fp.cand_cap = ws->cand_words - cand_off;
  fp.timeline = g_timeline.load(std::memory_order_relaxed);
  // epochs stay in [1, 2^30): they can never equal the high word of a look-back state
  // (status << 30) left behind in the shared aggregate area by the general path
  fp.epoch = (uint32_t)(g_epoch.fetch_add(1, std::memory_order_relaxed) % 0x3FFFFFFFull) + 1u;
  const bool c64 = g->col_is_64 != 0;
  const int sms = num_sms();
  const int G = fanout <= 4 ? 4 : fanout <= 8 ? 8 : fanout <= 16 ? 16 : 32;  // relabel/sort groups (power of two)
  static int s_cps = 0, r_cps = 0;  // CTAs per SM of the sampling / relabel kernels (tunable for experiments)
  if (s_cps == 0) {
    const char* e = getenv("SPP_SAMPLE_CTAS_PER_SM");
    s_cps = (e && atoi(e) > 0) ? atoi(e) : 8;
    e = getenv("SPP_RELABEL_CTAS_PER_SM");
    r_cps = (e && atoi(e) > 0) ? atoi(e) : 8;
  }
  int64_t cap = (int64_t)sms * s_cps;
  {
    const int gpw = 32 / fanout;
    int64_t warps = ceil_div(fp.h.max_targets > 0 ? fp.h.max_targets : 1, gpw);
    int64_t ctas = ceil_div(warps, kSampleThreads / 32);
    const int sgrid = (int)(ctas < cap ? ctas : cap);
    if (c64) k_hop_sample_fused<true><<<sgrid, kSampleThreads, 0, st>>>(fp);
    else k_hop_sample_fused<false><<<sgrid, kSampleThreads, 0, st>>>(fp);
  }
  cap = (int64_t)sms * r_cps;
  int64_t warps = ceil_div(fp.h.max_targets > 0 ? fp.h.max_targets : 1, 32 / G);
  int64_t ctas = ceil_div(warps, kSampleThreads / 32);
  const int grid = (int)(ctas < cap ? ctas : cap);
  SPP_KERNEL_CHECK("k_hop_sample_fused");
  trace_mark(kTrSample, hop, st);
  const int64_t scap = (int64_t)sms * 6;
  k_hop_compact_fused<<<(int)(tiles < scap ? tiles : scap), kScanThreads, 0, st>>>(fp);
  SPP_KERNEL_CHECK("k_hop_compact_fused");
  trace_mark(kTrCompact, hop, st);
  cudaStream_t rst = st;
  if (relabel_st != nullptr && relabel_st != st) {
    SPP_CUDA(cudaEventRecord(fork_ev, st));
    SPP_CUDA(cudaStreamWaitEvent(relabel_st, fork_ev, 0));
    rst = relabel_st;
  }
  switch (G) {
    case 4: k_relabel_sort_fused<4><<<grid, kSampleThreads, 0, rst>>>(fp); break;
    case 8: k_relabel_sort_fused<8><<<grid, kSampleThreads, 0, rst>>>(fp); break;
    case 16: k_relabel_sort_fused<16><<<grid, kSampleThreads, 0, rst>>>(fp); break;
    default: k_relabel_sort_fused<32><<<grid, kSampleThreads, 0, rst>>>(fp); break;
  }
  SPP_KERNEL_CHECK("k_relabel_sort_fused");
  trace_mark(kTrRelabel, hop, rst);
  return 0;
}

static int launch_export(const spp_sampler_ws* ws, int word, void* out, int out_is_64, int64_t max_nodes,
                         cudaStream_t st, const spp_device_job* job = nullptr) {
  if ((!out && !job) || max_nodes <= 0) return 0;
  int64_t cap = max_nodes < ws->max_nodes ? max_nodes : ws->max_nodes;
  int64_t ctas = ceil_div(cap, 256);
  int64_t lim = (int64_t)num_sms() * 8;
  int grid = (int)(ctas < lim ? ctas : lim);
  if (out_is_64) k_export_nids<int64_t><<<grid, 256, 0, st>>>(job, ws->n_ids, ws->meta, word, cap, (int64_t*)out);
  else k_export_nids<int32_t><<<grid, 256, 0, st>>>(job, ws->n_ids, ws->meta, word, cap, (int32_t*)out);
  SPP_KERNEL_CHECK("k_export_nids");
  return 0;
}

// All hops of one mini-batch.  With the relabel fork on (pipeline_flags() & 1) the relabel/sort
// kernel of every fused hop runs on the main stream's side stream: it only reads table entries
// that its own hop's compaction finalised (the next hop's sampler never lowers the value of an
// established entry and only adds new slots) and its own range of `cand`, and nothing on the
// critical path (next hop, owner split, feature gather) reads out_col.  *pending tells the caller
// that join_relabel() must be called on `st` before the batch is complete.
int sample_minibatch_impl(const spp_graph* g, const int64_t* seeds, int64_t batch_size, const int32_t* sizes,
                          int n_hops, int replace, uint64_t rng_seed, const spp_sampler_ws* ws,
                          int64_t* const* out_rowptr, int64_t* const* out_col, const int64_t* out_col_cap,
                          int64_t* n_id_out, cudaStream_t st, bool* pending, const spp_device_job* job, bool want_nid) {
  // `job` != NULL: per-batch pointers, capacities, seeds and the RNG key are read by the kernels from
  // that device block (graph replay); out_col_cap[] then only bounds grids and scratch ranges
  *pending = false;
  if (n_hops < 0 || n_hops > SPP_MAX_HOPS) return fail(SPP_EINVAL, "spp_sample_minibatch: n_hops out of range");
  if (n_hops > 0 && (!sizes || !out_rowptr || !out_col || !out_col_cap))
    return fail(SPP_EINVAL, "spp_sample_minibatch: null argument");
  if (int r = launch_begin(g, seeds, batch_size, ws, st, job)) return r;
  // one range of ws->cand per fused hop when they all fit (workspaces sized by spp_sampler_sizes
  // do); otherwise every hop reuses the start of the buffer and nothing is forked
  int64_t need = 0, T = batch_size;
  for (int h = 0; h < n_hops; ++h) {
    const int64_t Tb = T < ws->max_targets ? T : ws->max_targets;
    if (fused_ok(sizes[h], replace, ws)) need += (Tb * (int64_t)sizes[h] + 7) & ~7ll;
    const int64_t next = T + out_col_cap[h];
    T = next < ws->max_nodes ? next : ws->max_nodes;
  }
  const bool ranges = ws->cand != nullptr && need <= ws->cand_words;
  AuxStreams* aux = (ranges && !job && (pipeline_flags() & 1)) ? aux_streams(st) : nullptr;
  // host-side frontier bounds (the device clamps to them and raises SPP_META_OVERFLOW)
  int64_t cand_off = 0;
  T = batch_size;
  for (int h = 0; h < n_hops; ++h) {
    int64_t Tb = T < ws->max_targets ? T : ws->max_targets;
    if (fused_ok(sizes[h], replace, ws)) {
      if (int r = launch_hop_fused(g, h, sizes[h], rng_seed, Tb, out_col_cap[h], ws, out_rowptr[h], out_col[h], st,
                                   cand_off, aux ? aux->relabel : nullptr, aux ? aux->fork[h] : nullptr, job))
        return r;
      if (aux) *pending = true;
      if (ranges) cand_off += (Tb * (int64_t)sizes[h] + 7) & ~7ll;
    } else {
      // the general path rewrites table entries wholesale: earlier relabels must have finished
      if (int r = join_relabel(st, *pending)) return r;
      *pending = false;
      if (int r = launch_count(g, h, sizes[h], replace, Tb, ws, out_rowptr[h], st, job)) return r;
      if (int r = launch_fill(g, h, sizes[h], replace, rng_seed, Tb, out_col_cap[h], ws, out_rowptr[h], out_col[h], st, job))
        return r;
    }
    int64_t next = T + out_col_cap[h];
    T = next < ws->max_nodes ? next : ws->max_nodes;
  }
  if (n_id_out || (job && want_nid)) {
    if (int r = launch_export(ws, SPP_META_NODES(n_hops), n_id_out, 1, ws->max_nodes, st, job)) return r;
    trace_mark(kTrExport, 0, st);
  }
  return 0;
}

int join_relabel(cudaStream_t st, bool pending) {
  if (!pending) return 0;
  AuxStreams* aux = aux_streams(st);
  if (!aux) return fail(SPP_EINVAL, "sampler: side stream missing at join");
  SPP_CUDA(cudaEventRecord(aux->join_relabel, aux->relabel));
  SPP_CUDA(cudaStreamWaitEvent(st, aux->join_relabel, 0));
  trace_mark(kTrJoin, 0, st);
  return 0;
}

}  // namespace spp

extern "C" {

/* diagnostics: device buffer (8 uint64 per compaction tile) receiving %globaltimer stamps of
 * k_hop_compact_fused's phases; NULL switches it off */
void spp_debug_set_timeline(void* dev_ptr) { spp::g_timeline.store((unsigned long long*)dev_ptr); }

int spp_sampler_sizes(int64_t batch_size, const int32_t* sizes, int n_hops, int replace, int64_t num_nodes,
                      int64_t max_degree, spp_sampler_sizes_t* out) {
  using namespace spp;
  if (!out || (n_hops > 0 && !sizes)) return fail(SPP_EINVAL, "spp_sampler_sizes: null argument");
  if (n_hops < 0 || n_hops > SPP_MAX_HOPS) return fail(SPP_EINVAL, "spp_sampler_sizes: n_hops %d out of range", n_hops);
  if (batch_size < 0) return fail(SPP_EINVAL, "spp_sampler_sizes: negative batch size");
  int64_t T = batch_size, maxE = 0, maxT = batch_size;
  const int64_t node_cap = num_nodes > 0 ? batch_size + num_nodes : INT64_MAX;
  for (int h = 0; h < n_hops; ++h) {
    out->hop_targets[h] = T;
    maxT = T;
    int64_t k = sizes[h];
    int64_t E;
    if (k >= 0) {
      // without replacement a target keeps min(k, deg) edges; with replacement always k
      int64_t per = (!replace && max_degree > 0 && max_degree < k) ? max_degree : k;
      E = T * per;
    } else if (max_degree >= 0) {
      E = T * max_degree;
    } else {
      return fail(SPP_EINVAL, "spp_sampler_sizes: full-neighbourhood hop needs max_degree");
    }
    out->hop_edges[h] = E;
    if (E > maxE) maxE = E;
    int64_t next = T + E;
    T = next < node_cap ? next : node_cap;
  }
  for (int h = n_hops; h < SPP_MAX_HOPS; ++h) out->hop_targets[h] = out->hop_edges[h] = 0;
  out->max_nodes = T > 0 ? T : 1;
  out->max_targets = maxT > 0 ? maxT : 1;
  int64_t pow2 = 1024;
  while (2 * pow2 < 3 * out->max_nodes) pow2 <<= 1;  // round 1's table size: still the yardstick for "direct or hashed"
  if (pow2 > (1ll << 31)) return fail(SPP_EUNSUPPORTED, "spp_sampler_sizes: node bound %lld too large", (long long)T);
  // hashed table: 1.35 x the node bound, rounded up to 1024 slots (any size works: the home slot is a
  // multiply-high range reduction).  Load factor <= 0.74 at the bound, ~0.3 for typical batches; a third
  // less to clear per batch and to keep L2 resident than the next power of two (11.7 vs 16.8 MB at
  // (15,10,5) @ 1024).
  int64_t slots = ((out->max_nodes * 27 + 19) / 20 + 1023) & ~1023ll;
  if (slots < 1024) slots = 1024;
  out->table_slots = slots;
  out->table_direct = 0;
  const char* force = getenv("SPP_TABLE");  // "hash" / "direct": override the choice (tests, experiments)
  const bool want_direct = force && force[0] == 'd' ? true : force && force[0] == 'h' ? false : 2 * num_nodes <= 3 * pow2;
  if (num_nodes > 0 && num_nodes <= (1ll << 31) && want_direct) {  // direct map no bigger than 1.5x the hash table
    out->table_direct = 1;
    out->table_slots = num_nodes;
  }
  int64_t items = maxE > maxT ? maxE : maxT;
  out->tile_words = 2 + ceil_div(items > 0 ? items : 1, kScanTile) + 30;
  int64_t cw = 0;  // fused path: virtual candidates of every sampled hop (one range per hop)
  for (int h = 0; h < n_hops; ++h)
    if (sizes[h] >= 1 && sizes[h] <= 32) cw += (out->hop_targets[h] * (int64_t)sizes[h] + 7) & ~7ll;
  out->cand_words = cw + 16;
  return 0;
}

int spp_sample_begin(const spp_graph* g, const int64_t* seeds, int64_t batch_size, const spp_sampler_ws* ws,
                     void* stream) {
  return spp::launch_begin(g, seeds, batch_size, ws, (cudaStream_t)stream);
}

int spp_sample_hop_count(const spp_graph* g, int hop, int32_t fanout, int replace, int64_t max_targets,
                         const spp_sampler_ws* ws, int64_t* out_rowptr, void* stream) {
  if (int r = spp::check_ws(ws)) return r;
  return spp::launch_count(g, hop, fanout, replace, max_targets, ws, out_rowptr, (cudaStream_t)stream);
}

int spp_sample_hop_fill(const spp_graph* g, int hop, int32_t fanout, int replace, uint64_t rng_seed,
                        int64_t max_targets, int64_t max_edges, const spp_sampler_ws* ws, const int64_t* out_rowptr,
                        int64_t* out_col, void* stream) {
  if (int r = spp::check_ws(ws)) return r;
  return spp::launch_fill(g, hop, fanout, replace, rng_seed, max_targets, max_edges, ws, out_rowptr, out_col,
                          (cudaStream_t)stream);
}

int spp_sample_export_nids(const spp_sampler_ws* ws, int hop, void* n_id_out, int out_is_64, int64_t max_nodes,
                           void* stream) {
  if (int r = spp::check_ws(ws)) return r;
  if (hop < 0 || hop > SPP_MAX_HOPS) return spp::fail(SPP_EINVAL, "spp_sample_export_nids: hop out of range");
  return spp::launch_export(ws, SPP_META_NODES(hop), n_id_out, out_is_64, max_nodes, (cudaStream_t)stream);
}

int spp_sample_minibatch(const spp_graph* g, const int64_t* seeds, int64_t batch_size, const int32_t* sizes,
                         int n_hops, int replace, uint64_t rng_seed, const spp_sampler_ws* ws,
                         int64_t* const* out_rowptr, int64_t* const* out_col, const int64_t* out_col_cap,
                         int64_t* n_id_out, void* stream) {
  bool pending = false;
  if (int r = spp::sample_minibatch_impl(g, seeds, batch_size, sizes, n_hops, replace, rng_seed, ws, out_rowptr, out_col,
                                         out_col_cap, n_id_out, (cudaStream_t)stream, &pending, nullptr, false))
    return r;
  return spp::join_relabel((cudaStream_t)stream, pending);
}

}  // extern "C"
