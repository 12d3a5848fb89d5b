// vip.cu -- vertex-inclusion-probability (VIP) propagation, the set-up-time model that decides
// which remote feature rows each rank replicates (driver/drivers/ddp.py:134-239
// get_frequency_tensors_fast, caching/vip.py:123-180).  One hop:
//     wp[u]    = min(1, fanout / deg(u)) * p_in[u]
//     p_out[v] = 1 - exp(-sum_{u in N(v)} wp[u])          (the Taylor form the driver uses, fp64)
//     not_total[v] *= 1 - p_out[v]
// Replaces torch_scatter.segment_csr plus the chunked H2D streaming of the CSR: the graph is
// already resident.  HBM/L2-bound gather-reduce: one random 8-byte read per CSR entry.
#include "common.cuh"

namespace spp {

// exact != 0: the per-neighbour term is -log(1 - w*p) (caching/vip.py:166-172) instead of the
// first-order w*p the driver uses (ddp.py:219-224); both feed p_out = 1 - exp(-sum).
__global__ void k_vip_weight(const int64_t* __restrict__ rowptr, const double* __restrict__ p_in, int64_t n,
                             double fanout, int exact, double* __restrict__ wp) {
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    const double deg = (double)(rowptr[u + 1] - rowptr[u]);
    const double w = fmin(1.0, fanout / deg);  // deg == 0 -> +inf -> 1, like torch.minimum(1, fanout/deg)
    const double t = w * p_in[u];
    wp[u] = exact ? -log(1.0 - t) : t;
  }
}

constexpr int kVipGroup = 8;  // lanes per row

template <bool kCol64>
__global__ void __launch_bounds__(256) k_vip_rowsum(const int64_t* __restrict__ rowptr, const void* __restrict__ col,
                                                    const double* __restrict__ wp, int64_t n, double* __restrict__ p_out,
                                                    double* __restrict__ not_total) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (kVipGroup - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kVipGroup;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / kVipGroup;
  const int64_t rounds = (n + groups - 1) / groups;
  for (int64_t r = 0; r < rounds; ++r) {  // warp-uniform trip count (shuffles below)
    const int64_t v = r * groups + group;
    double s = 0.0;
    if (v < n) {
      const int64_t b = rowptr[v], e = rowptr[v + 1];
      for (int64_t j = b + gl; j < e; j += kVipGroup) {
        const int64_t u = kCol64 ? reinterpret_cast<const int64_t*>(col)[j] : (int64_t)reinterpret_cast<const int32_t*>(col)[j];
        s += __ldg(wp + u);
      }
    }
#pragma unroll
    for (int d = kVipGroup / 2; d > 0; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
    if (v < n && gl == 0) {
      const double p = 1.0 - exp(-s);
      p_out[v] = p;
      if (not_total) not_total[v] *= (1.0 - p);
    }
  }
}

}  // namespace spp

extern "C" int spp_vip_hop(const spp_graph* g, double fanout, int exact, const double* p_in, double* p_out,
                           double* not_total, double* scratch, void* stream) {
  using namespace spp;
  if (!g || !g->rowptr) return fail(SPP_EINVAL, "spp_vip_hop: null graph");
  const int64_t n = g->num_nodes;
  if (n <= 0) return 0;
  if (!p_in || !p_out || !scratch) return fail(SPP_EINVAL, "spp_vip_hop: null pointer");
  if (p_in == p_out) return fail(SPP_EINVAL, "spp_vip_hop: p_in and p_out must differ");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = (int64_t)num_sms() * 8;
  int64_t ctas = ceil_div(n, 256 * 4);
  k_vip_weight<<<(int)(ctas < cap ? ctas : cap), 256, 0, st>>>(g->rowptr, p_in, n, fanout, exact, scratch);
  SPP_KERNEL_CHECK("k_vip_weight");
  ctas = ceil_div(n, 256 / kVipGroup);
  const int grid = (int)(ctas < cap ? ctas : cap);
  if (g->col_is_64) k_vip_rowsum<true><<<grid, 256, 0, st>>>(g->rowptr, g->col, scratch, n, p_out, not_total);
  else k_vip_rowsum<false><<<grid, 256, 0, st>>>(g->rowptr, g->col, scratch, n, p_out, not_total);
  SPP_KERNEL_CHECK("k_vip_rowsum");
  return 0;
}
