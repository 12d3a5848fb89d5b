// runtime.cu -- error reporting, launch accounting and CUDA-IPC peer mapping.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace spp {

static thread_local char g_err[512] = "";
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

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

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

}  // namespace spp

extern "C" {

int spp_abi_version(void) { return SPP_ABI_VERSION; }
const char* spp_last_error(void) { return spp::g_err; }
uint64_t spp_launch_count(void) { return spp::g_launches.load(std::memory_order_relaxed); }

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
  return 0;
}

int spp_ipc_close(void* ptr, int64_t offset) {
  if (!ptr) return 0;
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
