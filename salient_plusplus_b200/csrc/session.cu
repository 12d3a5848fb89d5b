// session.cu -- one mini-batch end to end (spp_batch_enqueue) and the native enqueue executor.
//
// The reference produces mini-batches on a pool of CPU worker threads
// (fast_sampler/fast_sampler.cpp:368-513, 963-1274).  Here a mini-batch is ~16 asynchronous CUDA
// calls; the executor is a single native thread per device that issues them in submission order
// and records a completion event, so the (Python) consumer thread only allocates outputs, posts a
// job descriptor and later waits on the ticket.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

#include <map>

namespace spp {

// buckets (bit p: partition p) whose rows do not live on this GPU
static uint32_t remote_class_mask(const spp_feature_map* m) {
  const uint32_t local = (1u << m->rank) | m->local_parts;
  uint32_t mask = 0;
  for (int p = 0; p < m->num_parts; ++p)
    if (!((local >> p) & 1u) && m->offsets[p + 1] > m->offsets[p]) mask |= 1u << p;
  return mask;
}

// The launch sequence of one mini-batch.  `job` == NULL: plain stream launches with the per-batch
// pointers in the kernel parameters.  `job` != NULL (graph capture): every kernel reads them from
// the device job block; bounds used for grid sizing come from the static fields of `j`.
static int issue_sequence(const spp_batch_job* j, cudaStream_t st, const spp_device_job* job) {
  const bool replay = job != nullptr;
  const int64_t bs = (replay && j->batch_size_cap > 0) ? j->batch_size_cap : j->batch_size;
  trace_mark(kTrBatchBegin, 0, st);
  if (replay) {
    SPP_CUDA(cudaMemcpyAsync(j->job_dev, j->job_host, sizeof(spp_device_job), cudaMemcpyHostToDevice, st));
    if (j->seeds_host && bs > 0)
      SPP_CUDA(cudaMemcpyAsync(j->seeds_dev, j->seeds_stage_host, (size_t)bs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  } else if (j->seeds_host && bs > 0) {
    if (!j->seeds_dev) return fail(SPP_EINVAL, "spp_batch_enqueue: seeds_dev missing");
    SPP_CUDA(cudaMemcpyAsync(j->seeds_dev, j->seeds_host, (size_t)bs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    trace_mark(kTrSeedsH2D, 0, st);
  }
  int64_t caps[SPP_MAX_HOPS];
  for (int h = 0; h < SPP_MAX_HOPS; ++h)
    caps[h] = (replay && j->out_col_bound[h] > 0) ? j->out_col_bound[h] : j->out_col_cap[h];
  // the relabel/sort kernels may be left running on the side stream: nothing below reads out_col,
  // the join happens before the size block goes back to the host
  bool relabel_pending = false;
  if (int r = sample_minibatch_impl(&j->graph, j->seeds_dev, bs, j->sizes, j->n_hops, j->replace, j->rng_seed, &j->ws,
                                    j->out_rowptr, j->out_col, caps, j->n_id_out, st, &relabel_pending, job,
                                    j->n_id_out != nullptr))
    return r;
  const int64_t* n_dev = j->ws.meta + SPP_META_NODES(j->n_hops);
  cudaStream_t gst = st;
  AuxStreams* aux = nullptr;
  // an error after the fork must still join the side stream back into `st` (otherwise work queued on
  // it would be left dangling behind the caller's synchronisation point)
  auto bail = [&](int r) {
    if (gst != st) {
      cudaEventRecord(aux->join_gather, gst);
      cudaStreamWaitEvent(st, aux->join_gather, 0);
    }
    return r;
  };
  if (j->do_split) {
    if (int r = split_by_owner_job(&j->fmap, j->use_cache, j->ws.n_ids, 0, j->ws.max_nodes, n_dev, j->bucket_ids, j->perm,
                                   j->bucket_counts, j->split_scratch, st, job))
      return bail(r);
    trace_mark(kTrSplit, 0, st);
  }
  // optional (SPP_FORK bit 1): the feature + label gather on its own stream -- forked AFTER the owner
  // split, whose source descriptors the gather reads
  if (!replay && (pipeline_flags() & 2) && (j->feature_mode || (j->y_table && bs > 0))) {
    aux = aux_streams(st);
    if (aux) {
      SPP_CUDA(cudaEventRecord(aux->fork_gather, st));
      SPP_CUDA(cudaStreamWaitEvent(aux->gather, aux->fork_gather, 0));
      gst = aux->gather;
    }
  }
  if (j->feature_mode == 1) {
    if (int r = gather_rows_job(j->table, j->table_pitch, j->row_bytes, j->ws.n_ids, 0, j->ws.max_nodes, n_dev, j->x_out,
                                j->ws.max_nodes, gst, job, 1))
      return bail(r);
    trace_mark(kTrGather, 0, gst);
  } else if (j->feature_mode == 2 && j->do_split && tunables().gather_split != 0 && remote_class_mask(&j->fmap) != 0) {
    // Rows of other GPUs' partitions are fetched by their own launch on a side stream, bucket by
    // bucket (the owner split just built the buckets), while this stream gathers the rows that live
    // in local HBM (hosted partitions + replicated cache): the NVLink-bound fetch and the HBM-bound
    // gather overlap instead of sharing tiles.
    const uint32_t peer_mask = remote_class_mask(&j->fmap);
    const uint32_t local_mask = ((1u << (j->fmap.num_parts + 1)) - 1u) & ~peer_mask;
    AuxStreams* ax = aux_streams(st);
    if (!ax) return bail(fail(SPP_EINVAL, "spp_batch_enqueue: side stream missing"));
    SPP_CUDA(cudaEventRecord(ax->fork_gather, st));
    SPP_CUDA(cudaStreamWaitEvent(ax->gather, ax->fork_gather, 0));
    int r = gather_by_class_job(&j->fmap, j->row_bytes, j->bucket_ids, j->split_scratch, j->ws.max_nodes, peer_mask, j->x_out,
                                j->gather_counters, ax->gather, job);
    trace_mark(kTrGather, 1, ax->gather);
    // join unconditionally (a capture must not end with an un-joined stream)
    cudaError_t je = cudaEventRecord(ax->join_gather, ax->gather);
    if (r == 0)
      r = gather_by_class_job(&j->fmap, j->row_bytes, j->bucket_ids, j->split_scratch, j->ws.max_nodes, local_mask, j->x_out,
                              j->gather_counters, st, job);
    trace_mark(kTrGather, 0, st);
    if (je == cudaSuccess) je = cudaStreamWaitEvent(st, ax->join_gather, 0);
    if (r) return bail(r);
    SPP_CUDA(je);
  } else if (j->feature_mode == 2) {
    // the owner split (when it ran) left one source descriptor per node in its scratch: the gather
    // reads them sequentially instead of probing the cache index a second time
    if (int r = gather_partitioned_job(&j->fmap, j->row_bytes, j->ws.n_ids, 0, j->ws.max_nodes, n_dev,
                                       j->do_split ? j->split_scratch : nullptr, j->x_out, j->ws.max_nodes,
                                       j->gather_counters, gst, job))
      return bail(r);
    trace_mark(kTrGather, 0, gst);
  }
  if (j->y_table && bs > 0) {
    if (int r = gather_rows_job(j->y_table, j->y_row_bytes, j->y_row_bytes, j->seeds_dev, 1, bs, nullptr, j->y_out, bs, gst,
                                job, 2))
      return bail(r);
    trace_mark(kTrLabels, 0, gst);
  }
  if (gst != st) {
    SPP_CUDA(cudaEventRecord(aux->join_gather, gst));
    SPP_CUDA(cudaStreamWaitEvent(st, aux->join_gather, 0));
  }
  if (int r = join_relabel(st, relabel_pending)) return r;
  if (j->meta_host) {
    SPP_CUDA(cudaMemcpyAsync(j->meta_host, j->ws.meta, SPP_META_WORDS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (j->do_split)
      SPP_CUDA(cudaMemcpyAsync(j->meta_host + SPP_META_WORDS, j->bucket_counts, (SPP_MAX_PARTS + 2) * sizeof(int64_t),
                               cudaMemcpyDeviceToHost, st));
    trace_mark(kTrMetaD2H, 0, st);
  }
  return 0;
}

// ---- graph replay ------------------------------------------------------------------------------
// One instantiated graph per stream (= per in-flight slot), valid for every job whose static part
// (everything but the per-batch fields that travel through the device job block) is unchanged.
struct GraphEntry {
  spp_batch_job key;        // static part of the captured job
  cudaGraphExec_t exec = nullptr;
  uint64_t kernels = 0;     // kernel launches inside the graph (launch accounting)
  bool broken = false;      // capture failed once on this stream: stay on plain launches
};
static std::mutex g_graph_mu;
static std::map<cudaStream_t, GraphEntry> g_graphs;

static int graph_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPP_GRAPH");
    v = (e && *e) ? atoi(e) : 1;
  }
  return v;
}

static void static_key(const spp_batch_job* j, spp_batch_job* k) {
  memcpy(k, j, sizeof(*k));
  k->seeds_host = j->seeds_host ? (const int64_t*)1 : nullptr;  // only whether seeds are staged from the host
  if (j->seeds_host == nullptr) k->seeds_dev = nullptr;          // device seeds: pointer travels in the job block
  k->batch_size = 0;
  k->rng_seed = 0;
  for (int h = 0; h < SPP_MAX_HOPS; ++h) {
    k->out_rowptr[h] = nullptr;
    k->out_col[h] = nullptr;
    k->out_col_cap[h] = j->out_col_bound[h] > 0 ? 0 : j->out_col_cap[h];
  }
  k->n_id_out = j->n_id_out ? (int64_t*)1 : nullptr;
  k->x_out = nullptr;
  k->y_out = nullptr;
  k->bucket_ids = nullptr;
  k->perm = nullptr;
  if (k->batch_size_cap <= 0) k->batch_size_cap = j->batch_size;
}

static void fill_device_job(const spp_batch_job* j, spp_device_job* d) {
  for (int h = 0; h < SPP_MAX_HOPS; ++h) {
    d->out_rowptr[h] = j->out_rowptr[h];
    d->out_col[h] = j->out_col[h];
    d->out_col_cap[h] = j->out_col_cap[h];
  }
  d->n_id_out = j->n_id_out;
  d->x_out = j->x_out;
  d->y_out = j->y_out;
  d->bucket_ids = j->bucket_ids;
  d->perm = j->perm;
  d->seeds = j->seeds_dev;
  d->batch_size = j->batch_size;
  d->rng_premixed = premix_seed(j->rng_seed);
  d->scan_epoch = next_scan_epoch();
}

// launch == false: only make sure the slot's graph exists (captured, instantiated, uploaded)
static int enqueue_replay(const spp_batch_job* j, cudaStream_t st, bool* done, bool launch = true) {
  *done = false;
  spp_batch_job key;
  static_key(j, &key);
  GraphEntry* e;
  {
    std::lock_guard<std::mutex> lk(g_graph_mu);
    e = &g_graphs[st];
  }
  if (e->broken) return 0;
  bool hit = e->exec != nullptr && j->batch_size <= e->key.batch_size_cap;
  if (hit) {
    const int64_t cap_have = e->key.batch_size_cap;
    key.batch_size_cap = cap_have;  // a smaller batch replays the graph captured for the larger cap
    hit = memcmp(&e->key, &key, sizeof(key)) == 0;
    if (!hit) static_key(j, &key);
  }
  if (!hit) {
    if (e->exec) {
      cudaGraphExecDestroy(e->exec);
      e->exec = nullptr;
    }
    if (int r = sorter_attributes()) return r;
    if (int r = gather_attributes()) return r;
    aux_streams(st);  // side streams are created before the capture begins
    spp_batch_job cap;
    memcpy(&cap, j, sizeof(cap));
    cap.batch_size_cap = key.batch_size_cap;
    const uint64_t k0 = spp_launch_count();
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (ce == cudaSuccess) {
      rc = issue_sequence(&cap, st, j->job_dev);
      ce = cudaStreamEndCapture(st, &graph);
    }
    const uint64_t kernels = spp_launch_count() - k0;
    count_launch(-(int)kernels);  // capture is not execution
    if (ce == cudaSuccess && rc == 0 && graph) ce = cudaGraphInstantiate(&e->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != 0 || !e->exec) {
      cudaGetLastError();
      e->exec = nullptr;
      e->broken = true;  // fall back to plain launches on this stream (reported by spp_graph_replays() staying 0)
      return 0;
    }
    memcpy(&e->key, &key, sizeof(key));
    e->kernels = kernels;
    count_capture();
    cudaGraphUpload(e->exec, st);  // the first launch does not pay for the upload
    cudaGetLastError();
  }
  if (!launch) {
    *done = true;
    return 0;
  }
  fill_device_job(j, j->job_host);
  if (j->seeds_host && j->batch_size > 0) memcpy(j->seeds_stage_host, j->seeds_host, (size_t)j->batch_size * sizeof(int64_t));
  SPP_CUDA(cudaGraphLaunch(e->exec, st));
  count_launch((int)e->kernels);
  count_replay();
  *done = true;
  return 0;
}

}  // namespace spp

static bool replay_eligible(const spp_batch_job* j) {
  return j->job_dev && j->job_host && (!j->seeds_host || j->seeds_stage_host) && spp::graph_mode_enabled() &&
         !spp::tracing_active() && spp::pipeline_flags() == 0;
}

// Capture + instantiate + upload the slot's graph ahead of its first batch (a Session does this at
// set-up, so no mini-batch pays the ~1 ms of a capture).  Only the static part of the job matters.
extern "C" int spp_batch_prepare(const spp_batch_job* j) {
  using namespace spp;
  if (!j) return fail(SPP_EINVAL, "spp_batch_prepare: null job");
  if (j->n_hops < 0 || j->n_hops > SPP_MAX_HOPS) return fail(SPP_EINVAL, "spp_batch_prepare: n_hops out of range");
  if (!replay_eligible(j)) return 0;
  bool done = false;
  return enqueue_replay(j, (cudaStream_t)j->stream, &done, false);
}

extern "C" int spp_batch_enqueue(const spp_batch_job* j) {
  using namespace spp;
  if (!j) return fail(SPP_EINVAL, "spp_batch_enqueue: null job");
  cudaStream_t st = (cudaStream_t)j->stream;
  if (j->n_hops < 0 || j->n_hops > SPP_MAX_HOPS) return fail(SPP_EINVAL, "spp_batch_enqueue: n_hops out of range");
  if (replay_eligible(j)) {
    bool done = false;
    if (int r = enqueue_replay(j, st, &done)) return r;
    if (done) return 0;
  }
  return issue_sequence(j, st, nullptr);
}

namespace spp {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Tickets whose state is kept (ring indexed by ticket number).  A Session keeps <= 8 in flight, but
// several Sessions share a device's executor -- a training iterator can sit on its pending batches
// while an evaluation pass consumes thousands through the same executor -- so the ring is sized for
// 65 536 newer tickets before an unconsumed one expires; its events are created on first use.
constexpr int kRing = 1 << 16;
constexpr int kSpinIters = 4000;  // ~100 us of pause instructions before the worker sleeps

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#endif
}

struct Executor {
  int device;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<std::pair<uint64_t, spp_batch_job>> queue;
  std::atomic<uint64_t> submitted{0};  // jobs ever queued (the worker spins on it before it sleeps)
  bool stop = false;
  uint64_t next_ticket = 1;
  uint64_t issued = 0;  // every ticket <= issued has had its CUDA calls issued
  cudaEvent_t events[kRing];
  double t_submit[kRing], t_begin[kRing], t_end[kRing];  // steady-clock seconds (diagnostics)
  int status[kRing];
  std::string errors[kRing];

  explicit Executor(int dev) : device(dev) {
    for (int i = 0; i < kRing; ++i) {
      events[i] = nullptr;
      status[i] = 0;
    }
    worker = std::thread([this] { run(); });
  }

  void run() {
    cudaSetDevice(device);
    uint64_t taken = 0;
    while (true) {
      std::pair<uint64_t, spp_batch_job> item;
      // A Session hands over a job every 30-100 us.  Sleeping on the condition variable between two
      // jobs costs the submitting (consumer) thread a futex wake and the job a wake-up latency of
      // several microseconds, so the worker first spins for about that long (~100 us) and only then
      // sleeps; between Sessions it sleeps.
      for (int spin = 0; spin < kSpinIters && submitted.load(std::memory_order_acquire) == taken; ++spin) cpu_relax();
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [this] { return stop || !queue.empty(); });
        if (queue.empty()) break;  // stop requested and drained
        item = queue.front();
        queue.pop_front();
        ++taken;
      }
      const int slot = (int)(item.first % kRing);
      t_begin[slot] = now_s();
      int rc = spp_batch_enqueue(&item.second);
      std::string err;
      if (rc != 0) err = spp_last_error();
      cudaError_t e = cudaSuccess;
      if (!events[slot]) e = cudaEventCreateWithFlags(&events[slot], cudaEventDisableTiming);  // first lap of the ring
      if (e == cudaSuccess) e = cudaEventRecord(events[slot], (cudaStream_t)item.second.stream);
      t_end[slot] = now_s();
      if (rc == 0 && e != cudaSuccess) {
        rc = (int)e;
        err = cudaGetErrorString(e);
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        status[slot] = rc;
        errors[slot] = err;
        issued = item.first;
      }
      cv_done.notify_all();
    }
    for (int i = 0; i < kRing; ++i)
      if (events[i]) cudaEventDestroy(events[i]);
  }

  // waits until `ticket` has been issued; returns its issue status
  int wait_issued(uint64_t ticket, bool block, bool& ready) {
    std::unique_lock<std::mutex> lk(mu);
    if (ticket == 0 || ticket >= next_ticket) return fail(SPP_EINVAL, "executor: unknown ticket %llu", (unsigned long long)ticket);
    if (ticket + kRing <= next_ticket)  // its ring slot may have been reused
      return fail(SPP_EINVAL, "executor: ticket %llu expired", (unsigned long long)ticket);
    if (block) cv_done.wait(lk, [&] { return issued >= ticket; });
    ready = issued >= ticket;
    if (!ready) return 0;
    const int slot = (int)(ticket % kRing);
    if (status[slot] != 0) return fail(status[slot], "executor job failed: %s", errors[slot].c_str());
    return 0;
  }
};

}  // namespace spp

extern "C" {

void* spp_executor_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    spp::fail(SPP_EINVAL, "spp_executor_create: bad device %d", device);
    return nullptr;
  }
  return new spp::Executor(device);
}

void spp_executor_destroy(void* executor) {
  auto* ex = static_cast<spp::Executor*>(executor);
  if (!ex) return;
  {
    std::lock_guard<std::mutex> lk(ex->mu);
    ex->stop = true;
  }
  ex->cv_work.notify_all();
  if (ex->worker.joinable()) ex->worker.join();
  delete ex;
}

uint64_t spp_executor_submit(void* executor, const spp_batch_job* job) {
  auto* ex = static_cast<spp::Executor*>(executor);
  if (!ex || !job) {
    spp::fail(SPP_EINVAL, "spp_executor_submit: null argument");
    return 0;
  }
  uint64_t t;
  {
    std::lock_guard<std::mutex> lk(ex->mu);
    if (ex->next_ticket - 1 - ex->issued >= (uint64_t)spp::kRing - 1) {
      spp::fail(SPP_ECAPACITY, "spp_executor_submit: too many jobs in flight");
      return 0;
    }
    t = ex->next_ticket++;
    ex->t_submit[t % spp::kRing] = spp::now_s();
    ex->queue.emplace_back(t, *job);
    ex->submitted.fetch_add(1, std::memory_order_release);
  }
  ex->cv_work.notify_one();  // no system call unless the worker sleeps
  return t;
}

int spp_executor_poll(void* executor, uint64_t ticket) {
  auto* ex = static_cast<spp::Executor*>(executor);
  if (!ex) return spp::fail(SPP_EINVAL, "spp_executor_poll: null executor");
  bool ready = false;
  if (int r = ex->wait_issued(ticket, false, ready)) return r < 0 ? r : -r;
  if (!ready) return 0;
  cudaError_t e = cudaEventQuery(ex->events[ticket % spp::kRing]);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) {
    cudaGetLastError();
    return 0;
  }
  int rc = spp::cuda_fail(e, "cudaEventQuery");
  return rc > 0 ? -rc : rc;
}

/* diagnostics: steady-clock seconds at which `ticket` was submitted / started / finished issuing */
int spp_executor_times(void* executor, uint64_t ticket, double* out3) {
  auto* ex = static_cast<spp::Executor*>(executor);
  if (!ex || !out3) return spp::fail(SPP_EINVAL, "spp_executor_times: null argument");
  std::lock_guard<std::mutex> lk(ex->mu);
  const int slot = (int)(ticket % spp::kRing);
  out3[0] = ex->t_submit[slot];
  out3[1] = ex->t_begin[slot];
  out3[2] = ex->t_end[slot];
  return 0;
}

int spp_executor_wait(void* executor, uint64_t ticket) {
  auto* ex = static_cast<spp::Executor*>(executor);
  if (!ex) return spp::fail(SPP_EINVAL, "spp_executor_wait: null executor");
  bool ready = false;
  if (int r = ex->wait_issued(ticket, true, ready)) return r;
  SPP_CUDA(cudaEventSynchronize(ex->events[ticket % spp::kRing]));
  return 0;
}

}  // extern "C"
