// session.cu -- one mini-batch end to end (spp_batch_enqueue) and the native enqueue executor.
//
// The reference produces mini-batches on a pool of CPU worker threads
// (fast_sampler/fast_sampler.cpp:368-513, 963-1274).  Here a mini-batch is ~16 asynchronous CUDA
// calls; the executor is a single native thread per device that issues them in submission order
// and records a completion event, so the (Python) consumer thread only allocates outputs, posts a
// job descriptor and later waits on the ticket.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int spp_batch_enqueue(const spp_batch_job* j) {
  using namespace spp;
  if (!j) return fail(SPP_EINVAL, "spp_batch_enqueue: null job");
  cudaStream_t st = (cudaStream_t)j->stream;
  if (j->n_hops < 0 || j->n_hops > SPP_MAX_HOPS) return fail(SPP_EINVAL, "spp_batch_enqueue: n_hops out of range");
  const int64_t bs = j->batch_size;
  trace_mark(kTrBatchBegin, 0, st);
  if (j->seeds_host && bs > 0) {
    if (!j->seeds_dev) return fail(SPP_EINVAL, "spp_batch_enqueue: seeds_dev missing");
    SPP_CUDA(cudaMemcpyAsync(j->seeds_dev, j->seeds_host, (size_t)bs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    trace_mark(kTrSeedsH2D, 0, st);
  }
  // the relabel/sort kernels may be left running on the side stream: nothing below reads out_col,
  // the join happens before the size block goes back to the host
  bool relabel_pending = false;
  if (int r = sample_minibatch_impl(&j->graph, j->seeds_dev, bs, j->sizes, j->n_hops, j->replace, j->rng_seed, &j->ws,
                                    j->out_rowptr, j->out_col, j->out_col_cap, j->n_id_out, st, &relabel_pending))
    return r;
  const int64_t* n_dev = j->ws.meta + SPP_META_NODES(j->n_hops);
  // optional: the feature + label gather on its own (lower-priority) stream
  cudaStream_t gst = st;
  AuxStreams* aux = nullptr;
  if ((pipeline_flags() & 2) && (j->feature_mode || (j->y_table && bs > 0))) {
    aux = aux_streams(st);
    if (aux) {
      SPP_CUDA(cudaEventRecord(aux->fork_gather, st));
      SPP_CUDA(cudaStreamWaitEvent(aux->gather, aux->fork_gather, 0));
      gst = aux->gather;
    }
  }
  if (j->do_split) {
    if (int r = spp_split_by_owner(&j->fmap, j->use_cache, j->ws.n_ids, 0, j->ws.max_nodes, n_dev, j->bucket_ids,
                                   j->perm, j->bucket_counts, j->split_scratch, st))
      return r;
    trace_mark(kTrSplit, 0, st);
  }
  if (j->feature_mode == 1) {
    if (int r = spp_gather_rows_pitched(j->table, j->table_pitch, j->row_bytes, j->ws.n_ids, 0, j->ws.max_nodes, n_dev,
                                        j->x_out, j->ws.max_nodes, gst))
      return r;
    trace_mark(kTrGather, 0, gst);
  } else if (j->feature_mode == 2) {
    // the owner split (when it ran) left one source descriptor per node in its scratch: the gather
    // reads them sequentially instead of probing the cache index a second time
    if (int r = spp_gather_partitioned(&j->fmap, j->row_bytes, j->ws.n_ids, 0, j->ws.max_nodes, n_dev,
                                       j->do_split ? j->split_scratch : nullptr, j->x_out, j->ws.max_nodes,
                                       j->gather_counters, gst))
      return r;
    trace_mark(kTrGather, 0, gst);
  }
  if (j->y_table && bs > 0) {
    if (int r = spp_gather_rows(j->y_table, j->y_row_bytes, j->seeds_dev, 1, bs, nullptr, j->y_out, bs, gst)) return r;
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

namespace spp {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr int kRing = 256;  // tickets in flight (a Session keeps <= 8)

struct Executor {
  int device;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<std::pair<uint64_t, spp_batch_job>> queue;
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
    for (int i = 0; i < kRing; ++i) cudaEventCreateWithFlags(&events[i], cudaEventDisableTiming);
    while (true) {
      std::pair<uint64_t, spp_batch_job> item;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [this] { return stop || !queue.empty(); });
        if (queue.empty()) break;  // stop requested and drained
        item = queue.front();
        queue.pop_front();
      }
      const int slot = (int)(item.first % kRing);
      t_begin[slot] = now_s();
      int rc = spp_batch_enqueue(&item.second);
      std::string err;
      if (rc != 0) err = spp_last_error();
      cudaError_t e = cudaEventRecord(events[slot], (cudaStream_t)item.second.stream);
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
  }
  ex->cv_work.notify_one();
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
