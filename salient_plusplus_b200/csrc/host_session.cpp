// host_session.cpp -- the per-batch host path of a Session in native code (module `_spp_host`).
//
// The reference's consumer pops finished batches from a C++ queue (fast_sampler/fast_sampler.cpp:
// 777-828: try_get_batch / blocking_get_batch[_distributed]) that a pool of C++ worker threads fills
// (fast_sampler_thread, :963-1274); nothing per batch happens in Python there.  On a B200 a mini-batch
// of the small BASELINE shapes (ogbn-arxiv-shaped, layer-wise inference) takes 30-35 us of GPU time,
// less than the ~50 us the Python thread needed to allocate a batch's outputs, fill the job
// descriptor, submit it and later cut the exact-size views -- so the public API was bound by the
// interpreter, not by the GPU.  This file is that bookkeeping as one native call per batch:
//
//   fill()          for every free slot: allocate the batch's outputs from PyTorch's caching allocator
//                   on the slot's stream -- ONE block per batch holding the int64 arena, x and y at
//                   their upper bounds (a block handed to the consumer costs a record_stream, and when
//                   it is freed an event record + queries inside the allocator: per block, not per
//                   byte) -- write the per-batch fields of the slot's spp_batch_job and hand it to the
//                   executor thread of libsalient_b200.so (spp_executor_submit);
//   get(blocking)   poll / wait for the oldest in-flight batch (in idx_range order, like
//                   fast_sampler.cpp:672-712), read its pinned size block, cut the exact-size views
//                   (rowptr / col per hop, n_id, partition buckets, cached ids, perm), re-arm the
//                   slot with the next batch and return the pieces as one tuple.
//
// It is host logic only: no kernel, no CUDA call of its own (allocation goes through at::empty, the
// launch sequence through the C ABI of include/salient_b200.h, whose entry points are injected as
// function pointers so the CPU test suite can drive this file with a recording stand-in).
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <chrono>
#include <deque>
#include <optional>
#include <string>
#include <vector>

#include "salient_b200.h"

namespace py = pybind11;

namespace {

using submit_fn = uint64_t (*)(void*, const spp_batch_job*);
using poll_fn = int (*)(void*, uint64_t);
using wait_fn = int (*)(void*, uint64_t);
using error_fn = const char* (*)();

struct Slot {
  spp_batch_job* job = nullptr;   // the slot's descriptor (owned by the Python _Slot, static part filled there)
  const int64_t* meta_host = nullptr;  // pinned size block the batch's last copy writes
  int64_t* seeds_dev = nullptr;   // device staging buffer of host seeds
  std::optional<c10::Stream> stream;
  uint64_t ticket = 0;
  at::Tensor block;               // the batch's single allocation (bytes)
  at::Tensor arena, x, y;         // typed windows of it: int64 structure arena, features, labels
  int64_t start = 0, stop = 0;
};

// `sizes` elements (row-major, at most four dimensions) of `owner`'s storage as a tensor of its own,
// starting `element_offset` elements of type `dtype` into the storage.  Built the way ATen's own
// alias_with_sizes_and_strides builds a view -- a TensorImpl on the shared Storage -- instead of through
// narrow() / view() and the dispatcher: a batch is cut into 8-20 such windows and each dispatched view
// op costs microseconds of host time, which is the whole budget on the small graph shapes.
at::Tensor storage_window(const at::Tensor& owner, caffe2::TypeMeta dtype, int64_t element_offset, c10::IntArrayRef sizes) {
  auto t = at::detail::make_tensor<c10::TensorImpl>(c10::TensorImpl::VIEW, c10::Storage(owner.storage()), owner.key_set(), dtype);
  auto* impl = t.unsafeGetTensorImpl();
  impl->set_storage_offset(element_offset);
  int64_t strides[4] = {1, 1, 1, 1};
  for (int d = (int)sizes.size() - 2; d >= 0; --d) strides[d] = strides[d + 1] * std::max<int64_t>(sizes[d + 1], 1);
  impl->set_sizes_and_strides(sizes, c10::IntArrayRef(strides, sizes.size()));
  return t;
}
// a window of the view `base`, `offset` of its elements in
at::Tensor window(const at::Tensor& base, int64_t offset, c10::IntArrayRef sizes) {
  return storage_window(base, base.dtype(), base.storage_offset() + offset, sizes);
}
at::Tensor window(const at::Tensor& base, int64_t offset, int64_t rows) {
  return window(base, offset, c10::IntArrayRef(&rows, 1));
}
at::Tensor window(const at::Tensor& base, int64_t offset, int64_t rows, int64_t cols) {
  const int64_t sizes[2] = {rows, cols};
  return window(base, offset, c10::IntArrayRef(sizes, 2));
}
// a window of another element type, `byte_offset` bytes into the byte block `block`
at::Tensor typed_window(const at::Tensor& block, int64_t byte_offset, at::ScalarType dtype, c10::IntArrayRef sizes) {
  return storage_window(block, c10::scalarTypeToTypeMeta(dtype),
                        (block.storage_offset() + byte_offset) / (int64_t)c10::elementSize(dtype), sizes);
}

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

template <typename T>
T item(const py::dict& d, const char* key) {
  if (!d.contains(key)) throw std::invalid_argument(std::string("HostSession: missing field '") + key + "'");
  return d[key].cast<T>();
}

template <typename F>
F fn_ptr(const py::dict& d, const char* key) {
  return reinterpret_cast<F>(item<uintptr_t>(d, key));
}

class HostSession {
 public:
  explicit HostSession(const py::dict& spec)
      : device_(item<c10::Device>(spec, "device")),
        n_hops_(item<int>(spec, "n_hops")),
        parts_(item<int>(spec, "num_parts")),
        hop_off_(item<std::vector<std::pair<int64_t, int64_t>>>(spec, "hop_offsets")),
        nid_off_(item<int64_t>(spec, "nid_offset")),
        y_off_(item<int64_t>(spec, "y_offset")),
        arena_words_(item<int64_t>(spec, "arena_words")),
        node_bound_(item<int64_t>(spec, "node_bound")),
        has_x_(item<bool>(spec, "has_x")),
        feat_dim_(item<int64_t>(spec, "feat_dim")),
        feat_dtype_(item<at::ScalarType>(spec, "feat_dtype")),
        has_y_(item<bool>(spec, "has_y")),
        y_in_arena_(item<bool>(spec, "y_in_arena")),
        y_cols_(item<int64_t>(spec, "y_cols")),
        y_dtype_(item<at::ScalarType>(spec, "y_dtype")),
        ranges_(item<std::vector<std::pair<int64_t, int64_t>>>(spec, "ranges")),
        batch_edges_(item<std::vector<int64_t>>(spec, "batch_edges")),
        idx_host_(reinterpret_cast<const int64_t*>(item<uintptr_t>(spec, "idx_host_ptr"))),
        idx_dev_(reinterpret_cast<int64_t*>(item<uintptr_t>(spec, "idx_dev_ptr"))),
        executor_(reinterpret_cast<void*>(item<uintptr_t>(spec, "executor"))),
        submit_(fn_ptr<submit_fn>(spec, "submit_fn")),
        poll_(fn_ptr<poll_fn>(spec, "poll_fn")),
        wait_(fn_ptr<wait_fn>(spec, "wait_fn")),
        last_error_(fn_ptr<error_fn>(spec, "last_error_fn")),
        e_id_(item<at::Tensor>(spec, "e_id")),
        error_cls_(spec["error_cls"]) {
    if (n_hops_ < 1 || n_hops_ > SPP_MAX_HOPS || (int)hop_off_.size() != n_hops_)
      throw std::invalid_argument("HostSession: hop_offsets must have one (rowptr, col) pair per hop");
    if (parts_ > SPP_MAX_PARTS) throw std::invalid_argument("HostSession: too many partitions");
    if (!batch_edges_.empty() && batch_edges_.size() != ranges_.size())
      throw std::invalid_argument("HostSession: batch_edges must have one entry per batch");
    if (!submit_ || !poll_ || !wait_ || !last_error_ || !executor_)
      throw std::invalid_argument("HostSession: executor entry points are required");
    if ((idx_host_ == nullptr) == (idx_dev_ == nullptr) && !ranges_.empty())
      throw std::invalid_argument("HostSession: exactly one of idx_host_ptr / idx_dev_ptr must be set");
    for (const auto& r : ranges_)
      if (r.first < 0 || r.second < r.first) throw std::invalid_argument("HostSession: bad batch range");
    // slots: (job address, pinned size-block address, seeds staging address, stream_id, device_index, device_type)
    for (const auto& h : item<py::list>(spec, "slots")) {
      auto t = h.cast<py::tuple>();
      if (t.size() != 6) throw std::invalid_argument("HostSession: a slot is a 6-tuple");
      slots_.emplace_back();
      Slot& s = slots_.back();
      s.job = reinterpret_cast<spp_batch_job*>(t[0].cast<uintptr_t>());
      s.meta_host = reinterpret_cast<const int64_t*>(t[1].cast<uintptr_t>());
      s.seeds_dev = reinterpret_cast<int64_t*>(t[2].cast<uintptr_t>());
      if (!s.job || !s.meta_host) throw std::invalid_argument("HostSession: null slot pointer");
      if (device_.is_cuda())
        s.stream = c10::Stream::unpack3(t[3].cast<int64_t>(), (c10::DeviceIndex)t[4].cast<int64_t>(),
                                        (c10::DeviceType)t[5].cast<int64_t>());
    }
    if (slots_.empty()) throw std::invalid_argument("HostSession: no slots");
    for (auto& s : slots_) free_.push_back(&s);
    total_ = (int64_t)ranges_.size();
    // byte layout of a batch's block: [arena int64 | pad | x | pad | y (labels that are not int64)]
    int64_t max_bs = 0;
    for (const auto& r : ranges_) max_bs = std::max(max_bs, r.second - r.first);
    y_separate_ = has_y_ && !y_in_arena_;
    x_byte_off_ = round_up(arena_words_ * 8, 256);
    const int64_t x_bytes = has_x_ ? node_bound_ * feat_dim_ * (int64_t)c10::elementSize(feat_dtype_) : 0;
    y_byte_off_ = round_up(x_byte_off_ + x_bytes, 256);
    block_bytes_ = y_byte_off_ + (y_separate_ ? max_bs * y_cols_ * (int64_t)c10::elementSize(y_dtype_) : 0);
    if (block_bytes_ < 8) block_bytes_ = 8;
  }

  // enqueue batches while a slot is free (Session._enqueue)
  void fill() {
    while (!free_.empty() && next_ < total_) enqueue();
  }

  // oldest in-flight batch, or None (not ready and !blocking, or every batch consumed)
  py::object get(bool blocking) {
    if (consumed_ == total_ || pending_.empty()) return py::none();
    Slot& s = *pending_.front();
    int r = poll_(executor_, s.ticket);
    if (r < 0) fail("spp_executor_poll", r);
    if (r != 1) {
      if (!blocking) return py::none();
      const auto t0 = std::chrono::steady_clock::now();
      int rc;
      {
        py::gil_scoped_release release;  // the consumer waits on the GPU, not on the interpreter
        rc = wait_(executor_, s.ticket);
      }
      blocked_us_ += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
      ++blocked_occasions_;
      if (rc != 0) fail("spp_executor_wait", rc);
    }
    pending_.pop_front();
    py::object out = finalize(s);
    ++consumed_;
    free_.push_back(&s);
    fill();
    return out;
  }

  // batches consumed + in-flight batches whose GPU work has completed
  int64_t complete_count() {
    int64_t n = consumed_;
    for (Slot* s : pending_) n += poll_(executor_, s->ticket) == 1 ? 1 : 0;
    return n;
  }

  // wait for abandoned in-flight work and drop its outputs (Session dropped before its last batch)
  void release() {
    std::deque<Slot*> pend;
    pend.swap(pending_);
    if (!pend.empty()) {
      py::gil_scoped_release release;
      for (Slot* s : pend) wait_(executor_, s->ticket);
    }
    for (Slot* s : pend) drop(*s);
    free_.clear();
    next_ = total_;  // nothing can be enqueued any more
  }

  int64_t consumed() const { return consumed_; }
  int64_t issued() const { return next_; }
  int64_t total() const { return total_; }
  int64_t blocked_us() const { return blocked_us_; }
  int64_t blocked_occasions() const { return blocked_occasions_; }
  int64_t in_flight() const { return (int64_t)pending_.size(); }
  int64_t block_bytes() const { return block_bytes_; }

 private:
  [[noreturn]] void fail(const char* what, int rc) {
    const char* m = last_error_();
    std::string msg = std::string(what) + " failed (code " + std::to_string(rc) + "): " + (m ? m : "");
    PyErr_SetString(error_cls_.ptr(), msg.c_str());
    throw py::error_already_set();
  }

  static void drop(Slot& s) {
    s.block = at::Tensor();
    s.arena = at::Tensor();
    s.x = at::Tensor();
    s.y = at::Tensor();
    s.ticket = 0;
  }

  void enqueue() {
    Slot& s = *free_.front();
    const int64_t b = next_;
    const int64_t start = ranges_[b].first, stop = ranges_[b].second, bs = stop - start;
    spp_batch_job& j = *s.job;
    {
      // the block belongs to the slot's stream (caching-allocator ownership)
      c10::cuda::OptionalCUDAStreamGuard guard;
      if (s.stream) guard.reset_stream(c10::cuda::CUDAStream(*s.stream));
      s.block = at::empty({block_bytes_}, at::TensorOptions().device(device_).dtype(at::kByte));
    }
    // exact-size views are cut in finalize(); these are the upper-bound windows the kernels write
    s.arena = typed_window(s.block, 0, at::kLong, {arena_words_});
    s.x = typed_window(s.block, x_byte_off_, feat_dtype_, {has_x_ ? node_bound_ : 0, feat_dim_});
    if (y_separate_)
      s.y = typed_window(s.block, y_byte_off_, y_dtype_, {bs, y_cols_});
    else if (has_y_)  // int64 labels live at the tail of the arena
      s.y = window(s.arena, y_off_, bs, y_cols_);
    else
      s.y = at::Tensor();
    int64_t* base = s.arena.data_ptr<int64_t>();
    for (int h = 0; h < n_hops_; ++h) {
      j.out_rowptr[h] = base + hop_off_[h].first;
      j.out_col[h] = base + hop_off_[h].second;
    }
    if (!batch_edges_.empty()) j.out_col_cap[0] = std::max<int64_t>(batch_edges_[b], 0);
    if (idx_host_) {
      // the executor thread stages the slice of the caller's idx (kept alive by the Session)
      j.seeds_host = bs ? idx_host_ + start : nullptr;
      j.seeds_dev = s.seeds_dev;
    } else {
      j.seeds_host = nullptr;
      j.seeds_dev = idx_dev_ + start;
    }
    j.batch_size = bs;
    j.rng_seed = (uint64_t)(stop * 17 + 5) & 0xFFFFFFFFull;  // fast_sampler.cpp:994
    j.x_out = has_x_ ? s.x.data_ptr() : nullptr;
    j.y_out = (s.y.defined() && bs) ? s.y.data_ptr() : nullptr;
    if (parts_ >= 0) {
      j.n_id_out = base + nid_off_;
      j.bucket_ids = base + nid_off_ + node_bound_;
      j.perm = base + nid_off_ + 2 * node_bound_;
    }
    const uint64_t t = submit_(executor_, &j);
    if (!t) {
      drop(s);
      fail("spp_executor_submit", 0);
    }
    s.ticket = t;
    s.start = start;
    s.stop = stop;
    free_.pop_front();
    pending_.push_back(&s);
    ++next_;
  }

  py::object finalize(Slot& s) {
    const int64_t* m = s.meta_host;
    if (m[SPP_META_OVERFLOW]) {
      drop(s);
      PyErr_SetString(error_cls_.ptr(), "sampler buffer bound exceeded on the device (SPP_META_OVERFLOW)");
      throw py::error_already_set();
    }
    const int L = n_hops_;
    const int64_t nb = m[SPP_META_NODES(L)];
    // adjacency per hop, outermost hop first (reversed like fast_sampler.cpp:224)
    py::list adjs(L);
    for (int h = 0; h < L; ++h) {
      const int64_t T = m[SPP_META_NODES(h)], E = m[SPP_META_EDGES(h)];
      adjs[L - 1 - h] = py::make_tuple(window(s.arena, hop_off_[h].first, T + 1), window(s.arena, hop_off_[h].second, E),
                                       e_id_, py::make_tuple(T, m[SPP_META_NODES(h + 1)]));
    }
    py::tuple range = py::make_tuple(s.start, s.stop);
    py::object y = py::none(), y_flat = py::none();
    if (s.y.defined()) {
      y = py::cast(s.y);
      // the labels as the training loop takes them: y.squeeze() (fast_trainer/samplers.py:233,254)
      int64_t dims[2];
      int nd = 0;
      for (int64_t d : s.y.sizes())
        if (d != 1) dims[nd++] = d;
      y_flat = py::cast(nd == 2 ? s.y : window(s.y, 0, c10::IntArrayRef(dims, nd)));
    }
    py::object out;
    // every tensor of the batch is a view of the one block: record_stream on it covers them all
    py::tuple owners = py::make_tuple(s.block);
    if (parts_ < 0) {
      out = py::make_tuple(s.x.size(0) > nb ? window(s.x, 0, nb, feat_dim_) : s.x, y, adjs, range, owners, y_flat);
    } else {
      const int64_t* counts = m + SPP_META_WORDS;
      py::list buckets(parts_);
      int64_t pos = nid_off_ + node_bound_;
      for (int p = 0; p < parts_; ++p) {
        buckets[p] = window(s.arena, pos, counts[p]);
        pos += counts[p];
      }
      at::Tensor cached = window(s.arena, pos, counts[parts_]);
      py::object x = py::none();  // no reachable feature tables: the all_to_all prefetcher gathers x
      if (has_x_) x = py::cast(s.x.size(0) > nb ? window(s.x, 0, nb, feat_dim_) : s.x);
      out = py::make_tuple(window(s.arena, nid_off_, nb), buckets, cached, window(s.arena, nid_off_ + 2 * node_bound_, nb),
                           adjs, range, y, x, owners, y_flat);
    }
    drop(s);
    return out;
  }

  const c10::Device device_;
  const int n_hops_, parts_;  // parts_ < 0: not distributed
  const std::vector<std::pair<int64_t, int64_t>> hop_off_;
  const int64_t nid_off_, y_off_, arena_words_, node_bound_;
  const bool has_x_;
  const int64_t feat_dim_;
  const at::ScalarType feat_dtype_;
  const bool has_y_, y_in_arena_;
  const int64_t y_cols_;
  const at::ScalarType y_dtype_;
  const std::vector<std::pair<int64_t, int64_t>> ranges_;
  const std::vector<int64_t> batch_edges_;
  const int64_t* idx_host_;
  int64_t* idx_dev_;
  void* executor_;
  submit_fn submit_;
  poll_fn poll_;
  wait_fn wait_;
  error_fn last_error_;
  at::Tensor e_id_;
  py::object error_cls_;
  std::deque<Slot> slots_;  // stable addresses
  std::deque<Slot*> free_, pending_;
  bool y_separate_ = false;  // labels are not int64: their own window behind x instead of the arena's tail
  int64_t x_byte_off_ = 0, y_byte_off_ = 0, block_bytes_ = 8;
  int64_t total_ = 0, next_ = 0, consumed_ = 0, blocked_us_ = 0, blocked_occasions_ = 0;
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "native per-batch host path of salient_plusplus_b200.fast_sampler.Session";
  m.attr("ABI_VERSION") = SPP_ABI_VERSION;
  m.attr("BATCH_JOB_BYTES") = sizeof(spp_batch_job);
  py::class_<HostSession>(m, "HostSession")
      .def(py::init<const py::dict&>())
      .def("fill", &HostSession::fill)
      .def("get", &HostSession::get, py::arg("blocking"))
      .def("complete_count", &HostSession::complete_count)
      .def("release", &HostSession::release)
      .def_property_readonly("consumed", &HostSession::consumed)
      .def_property_readonly("issued", &HostSession::issued)
      .def_property_readonly("total", &HostSession::total)
      .def_property_readonly("in_flight", &HostSession::in_flight)
      .def_property_readonly("block_bytes", &HostSession::block_bytes)
      .def_property_readonly("blocked_us", &HostSession::blocked_us)
      .def_property_readonly("blocked_occasions", &HostSession::blocked_occasions);
}
