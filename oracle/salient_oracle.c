/*
 * oracle/salient_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the SALIENT++ mini-batch generation
 * algorithm (the reference's `fast_sampler` C++ module).  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (salient_plusplus_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * compiled, unmodified reference (oracle/_ref/fast_sampler.so, built by oracle/build_ref.sh) on
 * seeded graphs, including the stochastic path (same std::mt19937 stream, same biased
 * Floyd variant), and tests/golden/ holds fixtures generated from that reference.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 *
 * Two RNG modes for without-replacement / with-replacement sampling:
 *   SPO_RNG_REFERENCE (0): std::mt19937 seeded per batch with stop*17+5
 *       (fast_sampler/fast_sampler.cpp:994), `gen() % j` Floyd variant exactly as
 *       fast_sampler/sample_cpu.hpp:97-110 (known to be non-uniform).
 *   SPO_RNG_COUNTER (1): the counter-based generator the CUDA kernels use (splitmix64
 *       finaliser keyed by (seed, hop, target position, pick)), with the *correct* Floyd
 *       step t = U[0, j].  This mode exists so the stochastic GPU path can also be checked
 *       bit-for-bit; it is a specification of the new kernel, not of the reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SPO_RNG_REFERENCE 0
#define SPO_RNG_COUNTER 1

/* ------------------------------------------------------------------------------------------
 * std::mt19937 (32-bit Mersenne Twister, Matsumoto & Nishimura 1998) as libstdc++ implements
 * it: seed(s) is the Knuth initialisation with multiplier 1812433253, operator() tempering
 * (u=11, s=7,b=0x9D2C5680, t=15,c=0xEFC60000, l=18).  Used at fast_sampler/sample_cpu.hpp:11.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t mt[624];
  int idx;
} spo_mt19937;

void spo_mt_seed(spo_mt19937* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; ++i)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

uint32_t spo_mt_next(spo_mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* ------------------------------------------------------------------------------------------
 * Counter-based generator shared (by specification) with csrc/sampler.cu:spp_rand64.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t spo_mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

uint64_t spo_rand64(uint64_t seed, uint32_t hop, uint64_t target_pos, uint32_t pick) {
  uint64_t ctr = ((uint64_t)hop << 56) ^ (target_pos << 8) ^ (uint64_t)pick;
  return spo_mix64(spo_mix64(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) ^ ctr);
}

/* uniform integer in [0, range) from 64 random bits (multiply-high reduction) */
static inline uint32_t spo_bounded(uint64_t r, uint32_t range) {
  return (uint32_t)(((unsigned __int128)r * (unsigned __int128)range) >> 64);
}

/* ------------------------------------------------------------------------------------------
 * Global id -> local index map.  The reference uses phmap::flat_hash_map<int32,int32>
 * (fast_sampler/sample_cpu.hpp:13-19,27,54); only map *semantics* matter for the result
 * (insert-if-absent, lookup), so a small open-addressing table is used here.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t* keys;
  int32_t* vals;
  uint64_t cap; /* power of two */
  uint64_t size;
} spo_map;

static void spo_map_init(spo_map* m, uint64_t cap) {
  m->cap = 16;
  while (m->cap < cap) m->cap <<= 1;
  m->keys = (int32_t*)malloc(m->cap * sizeof(int32_t));
  m->vals = (int32_t*)malloc(m->cap * sizeof(int32_t));
  memset(m->keys, 0xff, m->cap * sizeof(int32_t)); /* -1 = empty (node ids are >= 0) */
  m->size = 0;
}

static void spo_map_free(spo_map* m) {
  free(m->keys);
  free(m->vals);
}

static inline uint64_t spo_hash32(int32_t k) { return spo_mix64((uint64_t)(uint32_t)k + 1); }

static void spo_map_grow(spo_map* m);

/* returns pointer to value slot; *inserted tells whether key was absent */
static int32_t* spo_map_insert(spo_map* m, int32_t key, int32_t val, int* inserted) {
  if ((m->size + 1) * 2 > m->cap) spo_map_grow(m);
  uint64_t h = spo_hash32(key) & (m->cap - 1);
  while (m->keys[h] != -1) {
    if (m->keys[h] == key) {
      *inserted = 0;
      return &m->vals[h];
    }
    h = (h + 1) & (m->cap - 1);
  }
  m->keys[h] = key;
  m->vals[h] = val;
  m->size++;
  *inserted = 1;
  return &m->vals[h];
}

static void spo_map_grow(spo_map* m) {
  spo_map n;
  spo_map_init(&n, m->cap * 2);
  for (uint64_t i = 0; i < m->cap; ++i)
    if (m->keys[i] != -1) {
      int ins;
      spo_map_insert(&n, m->keys[i], m->vals[i], &ins);
    }
  spo_map_free(m);
  *m = n;
}

/* ------------------------------------------------------------------------------------------
 * Sampler state for one mini-batch: n_ids vector, id map, RNG, the per-hop adjacencies.
 * Mirrors the locals of multilayer_sample (fast_sampler/fast_sampler.cpp:191-227).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t* rowptr; /* [T+1] */
  int64_t* col;    /* [E] local ids, ascending in each row */
  int64_t T, E, S; /* targets, edges, |n_id| after the hop */
} spo_adj;

typedef struct {
  int32_t* n_ids;
  int64_t n, n_cap;
  spo_map map;
  spo_mt19937 gen;
  uint64_t seed; /* counter-mode key */
  int rng_mode;
  spo_adj* adjs;
  int n_adjs, adj_cap;
} spo_state;

static void spo_push_nid(spo_state* s, int32_t v) {
  if (s->n == s->n_cap) {
    s->n_cap = s->n_cap ? s->n_cap * 2 : 1024;
    s->n_ids = (int32_t*)realloc(s->n_ids, s->n_cap * sizeof(int32_t));
  }
  s->n_ids[s->n++] = v;
}

/* fast_sampler.cpp:196-202: narrow seeds to int32 and build the initial map with
 * n_id_map[n_ids[i]] = i (sample_cpu.hpp:13-19) -- for a duplicated seed the LAST position
 * wins, because operator[] overwrites.
 * `mt_seed` follows fast_sampler.cpp:994 (gen.seed(range.second*17+5)) when called for a
 * Session batch. */
spo_state* spo_state_new(const int64_t* seeds, int64_t n, int rng_mode, uint64_t rng_seed) {
  spo_state* s = (spo_state*)calloc(1, sizeof(spo_state));
  spo_map_init(&s->map, (uint64_t)(n * 4 + 16));
  for (int64_t i = 0; i < n; ++i) {
    int32_t v = (int32_t)seeds[i];
    spo_push_nid(s, v);
    int ins;
    int32_t* slot = spo_map_insert(&s->map, v, (int32_t)i, &ins);
    *slot = (int32_t)i;
  }
  s->rng_mode = rng_mode;
  s->seed = rng_seed;
  spo_mt_seed(&s->gen, (uint32_t)rng_seed);
  return s;
}

void spo_state_free(spo_state* s) {
  for (int i = 0; i < s->n_adjs; ++i) {
    free(s->adjs[i].rowptr);
    free(s->adjs[i].col);
  }
  free(s->adjs);
  free(s->n_ids);
  spo_map_free(&s->map);
  free(s);
}

static int spo_cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* One hop == sample_adj (fast_sampler/sample_cpu.hpp:25-143).
 *   num_neighbors < 0 : full neighbourhood, col order        (:67-73)
 *   replace           : k draws gen() % deg, duplicates kept  (:74-82)
 *   otherwise         : deg <= k -> all, else Floyd variant   (:83-113)
 * n_ids grows in first-discovery order (:54-57); each output row holds the local ids of the
 * chosen neighbours sorted ascending (:123-139); e_id is empty (:120).
 * col64 != NULL: int64 column array (the reference's layout); else col32 is read.
 * Returns E. */
int64_t spo_hop(spo_state* s, const int64_t* rowptr, const int64_t* col64, const int32_t* col32,
                int32_t num_neighbors, int replace) {
  const int64_t T = s->n; /* idx_size: frontier = every node discovered so far (:30) */
  const uint32_t hop = (uint32_t)s->n_adjs;
  int64_t* out_rowptr = (int64_t*)malloc((T + 1) * sizeof(int64_t));
  int64_t ecap = 1024, E = 0;
  int32_t* cols = (int32_t*)malloc(ecap * sizeof(int32_t));
  int32_t* perm = (int32_t*)malloc(((num_neighbors > 0 ? num_neighbors : 0) + 1) * sizeof(int32_t));
  out_rowptr[0] = 0;

  for (int64_t i = 0; i < T; ++i) {
    const int32_t n = s->n_ids[i];
    const int64_t row_start = rowptr[n], row_end = rowptr[n + 1];
    const int32_t deg = (int32_t)(row_end - row_start); /* :48,62 (narrowed to int32) */
    int64_t row_begin_E = E;

#define SPO_ADD_NEIGHBOR(p)                                                         \
  do {                                                                              \
    const int64_t e_ = row_start + (p);                                             \
    const int32_t c_ = col64 ? (int32_t)col64[e_] : col32[e_];                      \
    int ins_;                                                                       \
    int32_t* slot_ = spo_map_insert(&s->map, c_, (int32_t)s->n, &ins_);             \
    if (ins_) spo_push_nid(s, c_);                                                  \
    if (E == ecap) {                                                                \
      ecap *= 2;                                                                    \
      cols = (int32_t*)realloc(cols, ecap * sizeof(int32_t));                       \
    }                                                                               \
    cols[E++] = *slot_;                                                             \
  } while (0)

    if (num_neighbors < 0) {
      for (int32_t j = 0; j < deg; ++j) SPO_ADD_NEIGHBOR(j);
    } else if (replace) {
      if (deg > 0) {
        for (int32_t j = 0; j < num_neighbors; ++j) {
          int32_t p;
          if (s->rng_mode == SPO_RNG_REFERENCE)
            p = (int32_t)((uint64_t)spo_mt_next(&s->gen) % (uint64_t)(int64_t)deg);
          else
            p = (int32_t)spo_bounded(spo_rand64(s->seed, hop, (uint64_t)i, (uint32_t)j), (uint32_t)deg);
          SPO_ADD_NEIGHBOR(p);
        }
      }
    } else {
      if (deg <= num_neighbors) {
        for (int32_t j = 0; j < deg; ++j) SPO_ADD_NEIGHBOR(j);
      } else {
        int32_t np = 0;
        for (int32_t j = deg - num_neighbors; j < deg; ++j) {
          int32_t option;
          if (s->rng_mode == SPO_RNG_REFERENCE) /* sample_cpu.hpp:99 -- `% j`, sic */
            option = (int32_t)((uint64_t)spo_mt_next(&s->gen) % (uint64_t)(int64_t)j);
          else /* Floyd: uniform on [0, j] */
            option = (int32_t)spo_bounded(
                spo_rand64(s->seed, hop, (uint64_t)i, (uint32_t)(j - (deg - num_neighbors))),
                (uint32_t)j + 1u);
          int found = 0;
          for (int32_t q = 0; q < np; ++q)
            if (perm[q] == option) {
              found = 1;
              break;
            }
          int32_t winner = found ? j : option;
          perm[np++] = winner;
          SPO_ADD_NEIGHBOR(winner);
        }
      }
    }
#undef SPO_ADD_NEIGHBOR
    qsort(cols + row_begin_E, (size_t)(E - row_begin_E), sizeof(int32_t), spo_cmp_i32); /* :126 */
    out_rowptr[i + 1] = E;
  }
  free(perm);

  if (s->n_adjs == s->adj_cap) {
    s->adj_cap = s->adj_cap ? s->adj_cap * 2 : 4;
    s->adjs = (spo_adj*)realloc(s->adjs, s->adj_cap * sizeof(spo_adj));
  }
  spo_adj* a = &s->adjs[s->n_adjs++];
  a->rowptr = out_rowptr;
  a->col = (int64_t*)malloc((E > 0 ? E : 1) * sizeof(int64_t));
  for (int64_t e = 0; e < E; ++e) a->col[e] = cols[e];
  a->T = T;
  a->E = E;
  a->S = s->n;
  free(cols);
  return E;
}

int64_t spo_state_num_nodes(const spo_state* s) { return s->n; }
int spo_state_num_adjs(const spo_state* s) { return s->n_adjs; }
void spo_state_adj_sizes(const spo_state* s, int i, int64_t* T, int64_t* E, int64_t* S) {
  *T = s->adjs[i].T;
  *E = s->adjs[i].E;
  *S = s->adjs[i].S;
}
void spo_state_copy_adj(const spo_state* s, int i, int64_t* rowptr, int64_t* col) {
  memcpy(rowptr, s->adjs[i].rowptr, (size_t)(s->adjs[i].T + 1) * sizeof(int64_t));
  memcpy(col, s->adjs[i].col, (size_t)s->adjs[i].E * sizeof(int64_t));
}
/* fast_sampler.cpp:219-222: widen n_ids back to int64 */
void spo_state_copy_nids(const spo_state* s, int64_t* out) {
  for (int64_t i = 0; i < s->n; ++i) out[i] = (int64_t)s->n_ids[i];
}

/* ------------------------------------------------------------------------------------------
 * serial_index (fast_sampler/fast_sampler.cpp:238-259): out[i,:] = in[idx[i],:] for
 * i < min(len(idx), n); rows beyond len(idx) are left untouched (uninitialised in the
 * reference).  Byte copy, dtype-agnostic.
 * ---------------------------------------------------------------------------------------- */
void spo_serial_index(const uint8_t* in, int64_t row_bytes, const int64_t* idx, int64_t n_idx,
                      int64_t n, uint8_t* out) {
  int64_t m = n_idx < n ? n_idx : n;
  for (int64_t i = 0; i < m; ++i)
    memcpy(out + i * row_bytes, in + idx[i] * row_bytes, (size_t)row_bytes);
}

/* Timed entry for bench.py's cpu_baseline (kind "port"): sample one batch with the reference
 * RNG mode and gather its feature rows; returns N_b.  Not used by the product. */
int64_t spo_minibatch(const int64_t* rowptr, const int64_t* col64, const int64_t* seeds, int64_t bs,
                      const int32_t* sizes, int n_hops, uint64_t mt_seed, const uint8_t* x,
                      int64_t row_bytes, uint8_t* x_out, int64_t x_out_rows) {
  spo_state* s = spo_state_new(seeds, bs, SPO_RNG_REFERENCE, mt_seed);
  for (int h = 0; h < n_hops; ++h) spo_hop(s, rowptr, col64, NULL, sizes[h], 0);
  int64_t nb = s->n;
  if (x && x_out) {
    int64_t m = nb < x_out_rows ? nb : x_out_rows;
    for (int64_t i = 0; i < m; ++i)
      memcpy(x_out + i * row_bytes, x + (int64_t)s->n_ids[i] * row_bytes, (size_t)row_bytes);
  }
  spo_state_free(s);
  return nb;
}
