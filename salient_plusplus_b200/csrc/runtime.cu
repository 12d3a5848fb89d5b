// runtime.cu -- error reporting, launch accounting and CUDA-IPC peer mapping.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace spp {

static thread_local char g_err[512] = "";
static std::atomic<int> g_trace_on{0};
static std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)(int64_t)n, std::memory_order_relaxed); }
static std::atomic<uint64_t> g_replays{0};
void count_replay() { g_replays.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<uint64_t> g_captures{0};
void count_capture() { g_captures.fetch_add(1, std::memory_order_relaxed); }
bool tracing_active() { return g_trace_on.load(std::memory_order_relaxed) != 0; }

int num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

// ---- diagnostics: event trace of the launch sequence ------------------------------------------
// A mark is a timing-enabled CUDA event recorded on the launching stream right after a kernel
// (or copy) was issued; the time between consecutive marks of one stream is that operation's
// duration as the stream saw it, i.e. including the time it waited for SM slots while other
// in-flight mini-batches were running (tools/trace_pipeline.py).
struct TraceMark {
  int label;
  int hop;
  cudaStream_t stream;
  cudaEvent_t ev;
};
static std::mutex g_trace_mu;
static std::vector<TraceMark> g_trace;
static size_t g_trace_cap = 0;

void trace_mark(int label, int hop, cudaStream_t st) {
  if (!g_trace_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (g_trace.size() >= g_trace_cap) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_trace.push_back({label, hop, st, ev});
}

// ---- pointers handed out by spp_ipc_import (peer GPUs' memory) ------------------------------------
static std::mutex g_ipc_mu;
static std::vector<const void*> g_ipc_ptrs;

bool ipc_imported(const void* p) {
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (const void* q : g_ipc_ptrs)
    if (q == p) return true;
  return false;
}

// ---- side streams ----------------------------------------------------------------------------
static std::mutex g_aux_mu;
static std::vector<std::pair<cudaStream_t, AuxStreams*>> g_aux;

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

Tunables& tunables() {
  static Tunables t = {env_int("SPP_GATHER_CTAS_PER_SM", 0), env_int("SPP_GATHER_BULK", -1), env_int("SPP_BULK_TILE", 4096),
                       env_int("SPP_BULK_STAGES", 6), env_int("SPP_BULK_CTAS_PER_SM", 2), env_int("SPP_GATHER_SPLIT", 0),
                       env_int("SPP_GATHER_TILE_ROWS", 0), env_int("SPP_GATHER_L2HINT", 0)};
  return t;
}

int pipeline_flags() {
  static int v = -1;
  if (v < 0) v = env_int("SPP_FORK", 0) & 3;
  return v;
}

AuxStreams* aux_streams(cudaStream_t main) {
  std::lock_guard<std::mutex> lk(g_aux_mu);
  for (auto& e : g_aux)
    if (e.first == main) return e.second;
  // stream priorities: numerically lower = scheduled first; 0 is the default (and lowest)
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  auto clampp = [&](int p) { return p > lo ? lo : (p < hi ? hi : p); };
  auto* a = new AuxStreams();
  bool ok = cudaStreamCreateWithPriority(&a->relabel, cudaStreamNonBlocking, clampp(env_int("SPP_RELABEL_PRIO", 0))) == cudaSuccess &&
            cudaStreamCreateWithPriority(&a->gather, cudaStreamNonBlocking, clampp(env_int("SPP_GATHER_PRIO", 0))) == cudaSuccess;
  for (int h = 0; ok && h < SPP_MAX_HOPS; ++h) ok = cudaEventCreateWithFlags(&a->fork[h], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&a->fork_gather, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&a->join_relabel, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&a->join_gather, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    cuda_fail(cudaGetLastError(), "side stream creation");
    delete a;
    return nullptr;
  }
  g_aux.emplace_back(main, a);
  return a;
}

}  // namespace spp

extern "C" {

int spp_trace_begin(int64_t max_marks) {
  std::lock_guard<std::mutex> lk(spp::g_trace_mu);
  for (auto& m : spp::g_trace) cudaEventDestroy(m.ev);
  spp::g_trace.clear();
  spp::g_trace_cap = max_marks > 0 ? (size_t)max_marks : 0;
  spp::g_trace.reserve(spp::g_trace_cap);
  spp::g_trace_on.store(max_marks > 0 ? 1 : 0);
  return 0;
}

int64_t spp_trace_end(int32_t* labels, int32_t* hops, uint64_t* streams, double* ms, int64_t cap) {
  spp::g_trace_on.store(0);
  std::lock_guard<std::mutex> lk(spp::g_trace_mu);
  cudaDeviceSynchronize();
  int64_t n = 0;
  for (auto& m : spp::g_trace) {
    if (n < cap && labels && hops && streams && ms) {
      float f = 0.f;
      cudaEventElapsedTime(&f, spp::g_trace[0].ev, m.ev);
      labels[n] = m.label;
      hops[n] = m.hop;
      streams[n] = (uint64_t)(uintptr_t)m.stream;
      ms[n] = (double)f;
      ++n;
    }
  }
  for (auto& m : spp::g_trace) cudaEventDestroy(m.ev);
  spp::g_trace.clear();
  cudaGetLastError();
  return n;
}

int spp_tune(const char* key, int value) {
  if (!key) return spp::fail(SPP_EINVAL, "spp_tune: null key");
  spp::Tunables& t = spp::tunables();
  if (!strcmp(key, "gather_ctas_per_sm")) t.gather_ctas_per_sm = value;
  else if (!strcmp(key, "gather_bulk")) t.gather_bulk = value;
  else if (!strcmp(key, "bulk_tile")) t.bulk_tile = value;
  else if (!strcmp(key, "bulk_stages")) t.bulk_stages = value;
  else if (!strcmp(key, "bulk_ctas_per_sm")) t.bulk_ctas_per_sm = value;
  else if (!strcmp(key, "gather_split")) t.gather_split = value;
  else if (!strcmp(key, "gather_tile_rows")) t.gather_tile_rows = value;
  else if (!strcmp(key, "gather_l2_hint")) t.gather_l2_hint = value;
  else return spp::fail(SPP_EINVAL, "spp_tune: unknown key '%s'", key);
  return 0;
}

int spp_abi_version(void) { return SPP_ABI_VERSION; }
const char* spp_last_error(void) { return spp::g_err; }
uint64_t spp_launch_count(void) { return spp::g_launches.load(std::memory_order_relaxed); }
uint64_t spp_graph_replays(void) { return spp::g_replays.load(std::memory_order_relaxed); }
uint64_t spp_graph_captures(void) { return spp::g_captures.load(std::memory_order_relaxed); }

// ---- CUDA IPC ---------------------------------------------------------------------------------
// The feature partition of every rank is exported once at set-up; peers map it and the gather
// kernel dereferences the mapped pointer directly (loads travel over NVLink / NVSwitch).
int spp_ipc_export(const void* ptr, uint8_t* handle_host, int64_t* offset_host) {
  if (!ptr || !handle_host || !offset_host) return spp::fail(SPP_EINVAL, "spp_ipc_export: null argument");
  // base address of the allocation `ptr` lives in (the IPC handle always refers to the base)
  typedef int (*getrange_t)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SPP_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return spp::fail(SPP_EUNSUPPORTED, "spp_ipc_export: cuMemGetAddressRange unavailable");
  unsigned long long base = 0;
  size_t size = 0;
  int r = ((getrange_t)fn)(&base, &size, (unsigned long long)(uintptr_t)ptr);
  if (r != 0) return spp::fail(SPP_EINVAL, "spp_ipc_export: cuMemGetAddressRange failed (%d)", r);
  cudaIpcMemHandle_t h;
  SPP_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(handle_host, &h, 64);
  *offset_host = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  return 0;
}

int spp_ipc_import(const uint8_t* handle_host, int64_t offset, void** ptr_host) {
  if (!handle_host || !ptr_host) return spp::fail(SPP_EINVAL, "spp_ipc_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, 64);
  void* base = nullptr;
  SPP_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_host = (void*)((uintptr_t)base + (uintptr_t)offset);
  {
    std::lock_guard<std::mutex> lk(spp::g_ipc_mu);
    spp::g_ipc_ptrs.push_back(*ptr_host);
  }
  return 0;
}

int spp_ipc_close(void* ptr, int64_t offset) {
  if (!ptr) return 0;
  {
    std::lock_guard<std::mutex> lk(spp::g_ipc_mu);
    for (size_t i = 0; i < spp::g_ipc_ptrs.size(); ++i)
      if (spp::g_ipc_ptrs[i] == ptr) {
        spp::g_ipc_ptrs.erase(spp::g_ipc_ptrs.begin() + i);
        break;
      }
  }
  SPP_CUDA(cudaIpcCloseMemHandle((void*)((uintptr_t)ptr - (uintptr_t)offset)));
  return 0;
}

int spp_enable_peer_access(int peer_device) {
  int dev = 0;
  SPP_CUDA(cudaGetDevice(&dev));
  if (dev == peer_device) return 0;
  int can = 0;
  SPP_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) return spp::fail(SPP_EUNSUPPORTED, "device %d cannot access peer %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  SPP_CUDA(e);
  return 0;
}

}  // extern "C"
