/*
 * salient_b200.h -- C ABI of libsalient_b200.so: the B200 (sm_100a) mini-batch generation
 * path of SALIENT++ (neighbour sampling -> dedup/relabel -> partition-book translation ->
 * VIP-cache split -> feature gather -> peer-to-peer miss fetch).
 *
 * This is the drop-in boundary.  The reference exposes the same operations as a pybind11
 * module with torch::Tensor arguments (fast_sampler/fast_sampler.cpp:1280-1396); here every
 * entry point takes plain device pointers, sizes and a CUDA stream, so any host language can
 * bind it (INTEGRATION.md shows the ctypes and the pybind/libtorch stubs).  There is no CPU
 * fallback anywhere behind this interface.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all functions are asynchronous with respect to the host unless stated otherwise;
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or an SPP_E* code (<0);
 *     spp_last_error() returns a message for the calling thread;
 *   - node ids are < 2^31 (the reference narrows to int32 too, fast_sampler.cpp:196-199).
 */
#ifndef SALIENT_B200_H_
#define SALIENT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPP_ABI_VERSION 2
#define SPP_MAX_PARTS 16   /* partitions in a RangePartitionBook                               */
#define SPP_MAX_HOPS 8     /* hops per mini-batch                                              */
#define SPP_MAX_FANOUT 128 /* without-replacement fanout per hop (full neighbourhood: no cap)  */

#define SPP_EINVAL (-1)      /* bad argument                                                   */
#define SPP_ECAPACITY (-2)   /* a caller-provided buffer is too small (host-detectable cases)  */
#define SPP_EUNSUPPORTED (-3)

/* meta block written by the sampler (int64 each, device memory, SPP_META_WORDS long) */
#define SPP_META_WORDS 32
#define SPP_META_NODES(h) (h)       /* [0..L]: |n_id| before hop h; [L] = final N_b (h = hop)   */
#define SPP_META_EDGES(h) (12 + (h)) /* [12..12+L): edges kept by hop h                         */
#define SPP_META_OVERFLOW 24        /* !=0: a buffer bound was exceeded on the device           */

int spp_abi_version(void);
/* run-time tunables (A/B tools, tests); defaults come from the SPP_* environment variables:
 *   "gather_ctas_per_sm" (0 = automatic), "gather_bulk" (-1 automatic, 0 never, 1 whenever the rows are
 *   multiples of 16 bytes: bulk-copy flavour of the gather), "bulk_tile" (bytes), "bulk_stages",
 *   "bulk_ctas_per_sm", "gather_split" (1: a batch's peer rows are fetched by their own launch on a
 *   side stream, see spp_gather_by_class; 0 (default): one fused launch), "gather_tile_rows" (0 automatic) */
int spp_tune(const char* key, int value);
const char* spp_last_error(void);
/* number of kernels this library has launched in this process (bench.py `gpu_launches`) */
uint64_t spp_launch_count(void);
/* number of mini-batches issued as ONE CUDA-graph launch (spp_batch_job.job_dev set; SPP_GRAPH=0 disables) */
uint64_t spp_graph_replays(void);
/* number of graph captures so far (one per slot and job shape; a growing count during steady state
 * means jobs keep changing their static part) */
uint64_t spp_graph_captures(void);

/* ------------------------------------------------------------------------------------------
 * K4 -- feature gather.  Replaces serial_index (fast_sampler/fast_sampler.cpp:238-279) and,
 * on the GPU side of the reference, features_gpu[idx] (fast_trainer/transferers.py:649-653).
 *   out[i, :] = table[idx[i], :]   for i < min(n_idx, n_out_rows)       (byte copy, any dtype)
 * idx: int64 (idx_is_64 != 0) or int32.  If n_idx_dev != NULL the row count is read from
 * device memory (*n_idx_dev, int64) and n_idx is only the upper bound used to size the grid.
 * ---------------------------------------------------------------------------------------- */
int spp_gather_rows(const void* table, int64_t row_bytes, const void* idx, int idx_is_64,
                    int64_t n_idx, const int64_t* n_idx_dev, void* out, int64_t n_out_rows,
                    void* stream);
/* Same, for a source table whose rows are `table_pitch` bytes apart (>= row_bytes).  A resident
 * copy with a 256-byte pitch keeps rows whose size is not a multiple of 128 bytes (ogbn-products:
 * 200 bytes) from straddling three DRAM lines; the output stays dense ([n, row_bytes]). */
int spp_gather_rows_pitched(const void* table, int64_t table_pitch, int64_t row_bytes,
                            const void* idx, int idx_is_64, int64_t n_idx,
                            const int64_t* n_idx_dev, void* out, int64_t n_out_rows, void* stream);

/* RangePartitionBook + feature placement used by the partitioned gather and the split.
 * Mirrors RangePartitionBook{rank, world_size, partition_offsets}
 * (fast_sampler/range_partition_book.hpp:31-57) and Cache (:60-91). */
typedef struct spp_feature_map {
  int32_t num_parts;                    /* world_size                                          */
  int32_t rank;                         /* partition owned by this GPU                         */
  int64_t offsets[SPP_MAX_PARTS + 1];   /* partition_offsets, offsets[0] = 0 ... [P] = N       */
  const void* tables[SPP_MAX_PARTS];    /* tables[p]: rows of partition p (local HBM or a peer */
                                        /* GPU's HBM mapped through CUDA IPC); may be NULL for */
                                        /* partitions this GPU cannot reach                    */
  const void* cache_table;              /* cached_features [C, F] or NULL                      */
  const void* cache_index;              /* membership + rank index over the node ids, built by */
                                        /* spp_cache_build_index (replaces the reference's     */
                                        /* dense arrays, range_partition_book.cpp:152-158);    */
                                        /* NULL = no cache                                     */
  int64_t cache_index_nodes;            /* num_nodes the index was built for                   */
  int64_t table_pitch;                  /* byte pitch of tables[*] rows; 0 = dense (row_bytes) */
  int64_t cache_pitch;                  /* byte pitch of cache_table rows; 0 = dense           */
  uint32_t local_parts;                 /* bit p set: partition p is resident on THIS GPU as   */
                                        /* well (a GPU hosting several partitions when there   */
                                        /* are fewer GPUs than partitions); bit `rank` implied */
  uint32_t _pad;
} spp_feature_map;

/* K4+K5 -- fused partition-book translate + cache lookup + local/cached gather + P2P miss
 * fetch, written ONCE in MFG (n_id) order.  Replaces stages slicing1..4, the three
 * all_to_alls and combine_features of fast_trainer/transferers.py:462-766:
 *   p = nid2partid(n_id[i]);  row = p == rank ? tables[rank][n_id[i]-off[rank]]
 *                                  : cached(n_id[i]) ? cache_table[nid2cachenid(n_id[i])]
 *                                  : tables[p][n_id[i]-off[p]]           (peer HBM over NVLink)
 * src_desc (optional, int32[n]): the per-node source descriptors spp_split_by_owner left in
 *   scratch[0..n) for the SAME n_id list (p >= 0: partition p, < 0: ~cache row); the kernel then
 *   reads them sequentially instead of searching the book and probing the cache index.
 * counters (optional, int64[3]): rows served local / cache / peer are ADDED to it. */
int spp_gather_partitioned(const spp_feature_map* map_host, int64_t row_bytes, const void* n_id,
                           int idx_is_64, int64_t n_idx, const int64_t* n_idx_dev,
                           const int32_t* src_desc, void* out, int64_t n_out_rows,
                           int64_t* counters, void* stream);

/* K4 / K5 by source class.  After spp_split_by_owner every node has a position in bucket order;
 * this gathers the rows of the buckets selected by class_mask (bit p: partition p, bit num_parts:
 * cached rows) -- source side a dense walk over the buckets (bucket_ids + the split's scratch),
 * destination side row inv[pos] of `out`.  A Session issues two of these on two streams, one for the
 * buckets resident in local HBM and one for the buckets on peer GPUs, so that the NVLink-bound miss
 * fetch overlaps the HBM-bound gather; together they write exactly what spp_gather_partitioned
 * writes.  n_max is the n_max the split was called with. */
int spp_gather_by_class(const spp_feature_map* map_host, int64_t row_bytes, const int64_t* bucket_ids,
                        const int32_t* split_scratch, int64_t n_max, uint32_t class_mask, void* out,
                        int64_t* counters, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2 -- RangePartitionBook kernels (fast_sampler/range_partition_book.cpp:89-107).
 * ---------------------------------------------------------------------------------------- */
/* out[i] = searchsorted(offsets, nids[i], right=True) - 1 */
int spp_nid2partid(const int64_t* offsets_host, int num_parts, const int64_t* nids, int64_t n,
                   int64_t* out, void* stream);
/* out[i] = nids[i] - offsets[partition_idx] */
int spp_nid2localnid(const int64_t* offsets_host, int num_parts, int partition_idx,
                     const int64_t* nids, int64_t n, int64_t* out, void* stream);
/* out[i] = offsets[rank] <= nids[i] < offsets[rank+1]   (bool bytes) */
int spp_nid_is_local(const int64_t* offsets_host, int num_parts, int rank, const int64_t* nids,
                     int64_t n, uint8_t* out, void* stream);

/* Cache (fast_sampler/range_partition_book.cpp:116-195).  The reference keeps two dense host
 * arrays indexed by node id (bool cached[200M], int32 row[200M]); a random probe of a dense map
 * that size costs a DRAM access, three times per remote node and mini-batch.  Here the lookup
 * structure is an L2-resident index: one 32-byte block per 224 ids holding the membership bits
 * and the number of cached ids in front of the block, plus rank -> cache row (int32 per cached
 * vertex).  index must hold spp_cache_index_bytes(num_nodes, n_cached) bytes (256-byte aligned).
 * A vertex listed twice maps to its LAST position, like the reference's sequential overwrite
 * (:154-158); ids outside [0, num_nodes) are ignored. */
int64_t spp_cache_index_bytes(int64_t num_nodes, int64_t n_cached);
int spp_cache_build_index(const int64_t* cached_vertices, int64_t n_cached, int64_t num_nodes,
                          void* index, void* stream);
/* out[i] = nids[i] is cached   (bool bytes) */
int spp_nid_is_cached(const void* index, int64_t num_nodes, const int64_t* nids, int64_t n,
                      uint8_t* out, void* stream);
/* out[i] = cache row of nids[i], -1 if not cached   (int64) */
int spp_nid2cachenid(const void* index, int64_t num_nodes, const int64_t* nids, int64_t n,
                     int64_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3 -- distributed split of a mini-batch's node list (fast_sampler/fast_sampler.cpp:1017-1262).
 * Stable split of n_id[0..n) into P+1 buckets -- bucket p < P: nodes fetched from partition p
 * (own rank: every local node; others: remote and NOT cached), bucket P: remote cached nodes --
 * each bucket in n_id order, concatenated in bucket order:
 *   bucket_ids[pos]  = global id            (buckets 0..P-1, = partition_nids)
 *                    = cache row index      (bucket P,       = cached_nids)
 *   perm[i]          = pos of n_id[i]       (= perm_partition_to_mfg: cat(...)[perm] == n_id)
 *   bucket_counts[b] = size of bucket b     (int64[P+1]); bucket_counts[P+1] = n
 * use_cache == 0 reproduces the no-cache branch (:1031-1107).
 * n_dev (optional): take n from device memory (upper bound n_max sizes the grid).
 * scratch: int32[spp_split_scratch_words(n_max)] device words; on return scratch[0..n) holds the
 *   per-node source descriptor (p >= 0: fetched from partition p, < 0: ~cache row) that
 *   spp_gather_partitioned accepts as src_desc -- the cache index is probed once per node.
 * ---------------------------------------------------------------------------------------- */
int64_t spp_split_scratch_words(int64_t n_max);
int spp_split_by_owner(const spp_feature_map* map_host, int use_cache, const void* n_id,
                       int idx_is_64, int64_t n_max, const int64_t* n_dev, int64_t* bucket_ids,
                       int64_t* perm, int64_t* bucket_counts, int32_t* scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1 -- multi-hop neighbour sampling with per-hop dedup and global->local relabelling.
 * Replaces sample_adj (fast_sampler/sample_cpu.hpp:25-143) and multilayer_sample
 * (fast_sampler/fast_sampler.cpp:191-236).
 * ---------------------------------------------------------------------------------------- */
typedef struct spp_graph {
  const int64_t* rowptr; /* int64[num_nodes + 1]                                              */
  const void* col;       /* int32 or int64 [nnz]                                              */
  int32_t col_is_64;
  int32_t _pad;
  int64_t num_nodes;
} spp_graph;

/* Per-stream scratch of the sampler (caller allocates; sizes from spp_sampler_sizes). */
typedef struct spp_sampler_ws {
  uint64_t* table;       /* hash table, table_slots entries of {key+1, ~local}                 */
  int64_t table_slots;   /* >= 1.25 * max_nodes, any value (spp_sampler_sizes gives 1.35x)     */
  int32_t* n_ids;        /* int32[max_nodes]   global ids in first-discovery order             */
  int64_t max_nodes;
  int64_t* tgt_start;    /* int64[max_targets] rowptr[n] of each frontier node                  */
  int32_t* tgt_deg;      /* int32[max_targets]                                                  */
  int64_t max_targets;   /* frontier bound of the last hop                                     */
  uint64_t* tile_state;  /* uint64[tile_words] scan state; MUST be zero when first used         */
  int64_t tile_words;
  int64_t* meta;         /* int64[SPP_META_WORDS]                                               */
  int32_t* cand;         /* int32[cand_words] candidate slots of the fused sampled-hop path     */
                         /* (NULL: always use the general path)                                  */
  int64_t cand_words;
  int32_t table_direct;  /* != 0: slot == node id, table_slots >= num_nodes (need not be pow2)   */
  int32_t _pad;
} spp_sampler_ws;

typedef struct spp_sampler_sizes_t {
  int64_t max_nodes;        /* bound on |n_id|                                                 */
  int64_t max_targets;      /* bound on the last hop's frontier                                */
  int64_t table_slots;
  int64_t tile_words;
  int64_t cand_words;
  int64_t table_direct;     /* 1: use a direct-mapped table of table_slots = num_nodes entries  */
  int64_t hop_targets[SPP_MAX_HOPS]; /* bound T_h                                              */
  int64_t hop_edges[SPP_MAX_HOPS];   /* bound E_h (-1: data dependent, full neighbourhood)     */
} spp_sampler_sizes_t;

/* Upper bounds for a batch of `batch_size` seeds and fanouts `sizes` (negative = all
 * neighbours; replace != 0: k draws with replacement, i.e. k edges even when deg < k).  For full-neighbourhood hops the bounds need the graph: pass max_degree
 * (any upper bound on the degree) and num_nodes; edges bound = targets * max_degree. */
int spp_sampler_sizes(int64_t batch_size, const int32_t* sizes_host, int n_hops, int replace,
                      int64_t num_nodes, int64_t max_degree, spp_sampler_sizes_t* out_host);

/* One mini-batch, all hops, no host synchronisation:
 *   seeds        int64[batch_size]  (global ids; duplicates allowed, last position wins in the
 *                                    id map exactly like sample_cpu.hpp:13-19)
 *   sizes_host   fanout per hop, seed side first (fast_sampler.cpp:207); <0 = all neighbours
 *   out_rowptr[h] int64[hop_targets[h] + 1], out_col[h] int64[hop_edges[h]]  (hop order, NOT
 *                reversed: the host reverses like fast_sampler.cpp:224)
 *   out_col_cap  capacity (elements) of each out_col[h]
 *   n_id_out     int64[max_nodes] or NULL: widened n_id (fast_sampler.cpp:219-222)
 *   rng_seed     key of the counter-based generator; a Session uses stop*17+5 like
 *                fast_sampler.cpp:994
 * Sampling rule per target (sample_cpu.hpp:67-113): deg <= k or k < 0: every neighbour in col
 * order; otherwise Floyd's algorithm with t uniform on [0, j] (the reference's `% j` is a
 * defect), neighbours emitted in Floyd order.  Local ids are assigned in
 * sequential first-discovery order and each output row is sorted ascending.
 * ws->meta receives the node / edge counts; SPP_META_OVERFLOW is set if a bound was hit. */
int spp_sample_minibatch(const spp_graph* graph_host, const int64_t* seeds, int64_t batch_size,
                         const int32_t* sizes_host, int n_hops, int replace, uint64_t rng_seed,
                         const spp_sampler_ws* ws_host, int64_t* const* out_rowptr_host,
                         int64_t* const* out_col_host, const int64_t* out_col_cap_host,
                         int64_t* n_id_out, void* stream);

/* Single-hop building blocks (used by sample_adj and by full-neighbourhood hops whose edge
 * count must be known before out_col can be allocated):
 *   begin : clear table, insert seeds                       -> meta[NODES(0)] = batch_size
 *   count : degrees + exclusive scan -> out_rowptr, meta[EDGES(hop)]        (then the host may
 *           read meta to size out_col)
 *   fill  : sample + insert + compact + relabel/sort -> out_col, meta[NODES(hop+1)] */
int spp_sample_begin(const spp_graph* graph_host, const int64_t* seeds, int64_t batch_size,
                     const spp_sampler_ws* ws_host, void* stream);
int spp_sample_hop_count(const spp_graph* graph_host, int hop, int32_t fanout, int replace,
                         int64_t max_targets, const spp_sampler_ws* ws_host, int64_t* out_rowptr,
                         void* stream);
int spp_sample_hop_fill(const spp_graph* graph_host, int hop, int32_t fanout, int replace,
                        uint64_t rng_seed, int64_t max_targets, int64_t max_edges,
                        const spp_sampler_ws* ws_host, const int64_t* out_rowptr, int64_t* out_col,
                        void* stream);
/* diagnostics: device buffer (8 uint64 per compaction tile) receiving %globaltimer stamps of the
 * fused compaction kernel's phases (tools/compact_timeline.py); NULL switches it off */
void spp_debug_set_timeline(void* dev_ptr);

/* diagnostics: event trace of the launch sequence.  begin(max_marks) arms it (0 disarms); every
 * kernel / copy issued by the library is then followed by a timing event on its stream.  end()
 * synchronises the device and returns the marks in issue order: label (0 batch begin, 1 seeds
 * H2D, 2 table clear, 3 seeds init, 4 sample, 5 compact, 6 relabel+sort, 7 export n_id, 8 owner
 * split, 9 feature gather, 10 label gather, 11 meta D2H, 12 join, 13 degree count + scan, 14 large-row
 * comparison sort, 15 bitmap row sort), hop, stream handle and
 * milliseconds since the first mark (tools/trace_pipeline.py). */
int spp_trace_begin(int64_t max_marks);
int64_t spp_trace_end(int32_t* labels_host, int32_t* hops_host, uint64_t* streams_host,
                      double* ms_host, int64_t cap);

/* n_id_out[i] = (int64) ws->n_ids[i], i < meta[NODES(hop)]  (or int32 copy if out_is_64 == 0) */
int spp_sample_export_nids(const spp_sampler_ws* ws_host, int hop, void* n_id_out, int out_is_64,
                           int64_t max_nodes, void* stream);


/* ------------------------------------------------------------------------------------------
 * One mini-batch end to end, and the native enqueue executor.
 * Replaces the body of fast_sampler_thread (fast_sampler/fast_sampler.cpp:963-1274): sample ->
 * (distributed split) -> feature gather -> label gather, plus the H2D copy of the seeds and the
 * D2H copy of the size block.  The reference runs this on a pool of CPU worker threads; here the
 * work is ~16 asynchronous CUDA calls, issued either on the caller's thread
 * (spp_batch_enqueue) or by a per-device executor thread (spp_executor_*), so that a Python
 * consumer never spends its own time inside the CUDA driver.
 * ---------------------------------------------------------------------------------------- */
/* Per-batch part of a job, resident in DEVICE memory (one block per in-flight slot).  When a job
 * carries `job_dev`, every kernel of the launch sequence reads the batch's output pointers,
 * capacities, seeds and RNG key from this block instead of from its launch parameters, so the
 * whole sequence is captured ONCE per slot into a CUDA graph and replayed for every mini-batch:
 * per batch the host writes the pinned copy `job_host` (+ the seeds into `seeds_stage_host`) and
 * launches the graph, whose first nodes copy both to the device. */
typedef struct spp_device_job {
  int64_t* out_rowptr[SPP_MAX_HOPS];
  int64_t* out_col[SPP_MAX_HOPS];
  int64_t out_col_cap[SPP_MAX_HOPS];
  int64_t* n_id_out;
  void* x_out;
  void* y_out;
  int64_t* bucket_ids;
  int64_t* perm;
  const int64_t* seeds;
  int64_t batch_size;
  uint64_t rng_premixed;
  uint64_t scan_epoch;             /* tag of this batch's scan aggregates (hop h uses epoch + h)  */
} spp_device_job;

typedef struct spp_batch_job {
  spp_graph graph;
  spp_sampler_ws ws;
  const int64_t* seeds_host;       /* host seeds of this batch or NULL (device seeds).  Graph    */
                                   /* replay copies them into seeds_stage_host on the CPU; plain */
                                   /* launches hand them to cudaMemcpyAsync (pin them, or the    */
                                   /* copy is staged synchronously by the driver)                */
  int64_t* seeds_dev;              /* device seeds (destination of the copy, or the input)      */
  int64_t batch_size;
  int32_t sizes[SPP_MAX_HOPS];
  int32_t n_hops;
  int32_t replace;
  uint64_t rng_seed;
  int64_t* out_rowptr[SPP_MAX_HOPS];
  int64_t* out_col[SPP_MAX_HOPS];
  int64_t out_col_cap[SPP_MAX_HOPS];
  int64_t* n_id_out;               /* int64[ws.max_nodes] or NULL                               */
  /* features: 0 = none, 1 = single (pitched) table, 2 = partitioned map (K4+K5) */
  int32_t feature_mode;
  int32_t do_split;                /* != 0: spp_split_by_owner with `fmap`                      */
  int32_t use_cache;
  int32_t _pad;
  const void* table;
  int64_t table_pitch;
  int64_t row_bytes;
  spp_feature_map fmap;
  void* x_out;                     /* [ws.max_nodes, row_bytes]                                 */
  const void* y_table;             /* label rows (or NULL)                                      */
  int64_t y_row_bytes;
  void* y_out;                     /* [batch_size, y_row_bytes]                                 */
  int64_t* bucket_ids;             /* split outputs (do_split)                                  */
  int64_t* perm;
  int64_t* bucket_counts;          /* device int64[SPP_MAX_PARTS + 2]                           */
  int32_t* split_scratch;
  int64_t* meta_host;              /* pinned int64[SPP_META_WORDS + SPP_MAX_PARTS + 2] or NULL  */
  void* stream;
  int64_t* gather_counters;        /* optional device int64[3]: rows served local / cache / peer */
  /* graph replay (all optional; job_dev == NULL: plain stream launches) */
  spp_device_job* job_dev;         /* device block of this slot                                  */
  spp_device_job* job_host;        /* pinned host staging copy of it                             */
  int64_t* seeds_stage_host;       /* pinned int64[batch_size_cap]: seeds_host is copied here    */
  int64_t batch_size_cap;          /* largest batch_size this slot will see (grids are sized by  */
                                   /* it); 0 = batch_size                                        */
  int64_t out_col_bound[SPP_MAX_HOPS]; /* static upper bound of out_col_cap[h] (grid sizing);   */
                                   /* 0 = out_col_cap[h]                                         */
} spp_batch_job;

/* issue every call of the job on the calling thread (asynchronous w.r.t. the GPU) */
int spp_batch_enqueue(const spp_batch_job* job);
/* graph replay only: capture, instantiate and upload the slot's graph now (set-up time) instead of
 * with its first batch; needs the static fields of the job (the per-batch pointers may be dummies,
 * only whether n_id_out / seeds_host are NULL matters).  No-op when job_dev is not set. */
int spp_batch_prepare(const spp_batch_job* job);

/* executor: one worker thread bound to `device`; jobs are issued in submission order */
void* spp_executor_create(int device);
void spp_executor_destroy(void* executor);
/* copies *job, returns a ticket (> 0), or 0 on error */
uint64_t spp_executor_submit(void* executor, const spp_batch_job* job);
/* 1: the job's GPU work has completed, 0: not yet, < 0 / cudaError: failed (spp_last_error) */
int spp_executor_poll(void* executor, uint64_t ticket);
/* blocks until the job's GPU work has completed; 0 or an error code */
int spp_executor_wait(void* executor, uint64_t ticket);
/* diagnostics: steady-clock seconds at which the job was submitted, started and finished issuing */
int spp_executor_times(void* executor, uint64_t ticket, double* out3_host);

/* ------------------------------------------------------------------------------------------
 * VIP (vertex inclusion probability) propagation -- set-up time, decides the replicated cache
 * (driver/drivers/ddp.py:134-239, caching/vip.py:123-180).  One hop, fp64:
 *   t(u)     = min(1, fanout/deg(u)) * p_in[u]
 *   p_out[v] = 1 - exp(-sum_{u in N(v)} t(u))              exact == 0  (driver, ddp.py:219-224)
 *            = 1 - exp( sum_{u in N(v)} log(1 - t(u)))     exact != 0  (caching/vip.py:166-172)
 *   not_total[v] *= 1 - p_out[v]   (if not_total != NULL)
 * scratch: double[num_nodes].  p_in and p_out must not alias.
 * ---------------------------------------------------------------------------------------- */
int spp_vip_hop(const spp_graph* graph_host, double fanout, int exact, const double* p_in,
                double* p_out, double* not_total, double* scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Peer mapping (CUDA IPC) for the P2P gather.  Host-synchronous.
 *   export: handle_host receives 64 bytes, *offset_host the offset of `ptr` inside its
 *           allocation; import (in another process) returns a device pointer valid there.
 * ---------------------------------------------------------------------------------------- */
int spp_ipc_export(const void* ptr, uint8_t* handle_host, int64_t* offset_host);
int spp_ipc_import(const uint8_t* handle_host, int64_t offset, void** ptr_host);
int spp_ipc_close(void* ptr, int64_t offset);
int spp_enable_peer_access(int peer_device);

#ifdef __cplusplus
}
#endif
#endif /* SALIENT_B200_H_ */
