// GATConv(heads=1, concat=False) aggregation: edge softmax + weighted neighbour sum, fwd and bwd.
//
// Replaces the stock PyG GATConv used by pygda/nn/gnn_base.py:81-87:
//   e_ij  = leaky_relu(a_src[j] + a_dst[i], slope)           for every edge j -> i (incl. self loops)
//   alpha = softmax over the in-edges of i  (max-subtracted, +1e-16 in the denominator)
//   out_i = sum_j alpha_ij h_j
// One warp per target row on the CSR-by-target of the self-looped, un-normalised graph.  alpha is
// kept (CSR order) for the backward.  API surface only in the reference (no benchmark script
// selects gnn='gat'), so this is a straightforward kernel, not a tuned one.
#include "graph.cuh"

namespace gda {
namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
k_gat_fwd(const int* __restrict__ rowptr, const int* __restrict__ colidx, int N, int C, const float* __restrict__ h,
          const float* __restrict__ a_src, const float* __restrict__ a_dst, float slope, float* __restrict__ out,
          float* __restrict__ alpha) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const int s = rowptr[i], e = rowptr[i + 1];
  const float ad = a_dst[i];
  float m = -INFINITY;
  for (int p = s + lane; p < e; p += 32) {
    float v = a_src[colidx[p]] + ad;
    v = v > 0.f ? v : slope * v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int p = s + lane; p < e; p += 32) {
    float v = a_src[colidx[p]] + ad;
    v = v > 0.f ? v : slope * v;
    const float ex = expf(v - m);
    alpha[p] = ex;
    sum += ex;
  }
  sum = warp_sum(sum) + 1e-16f;
  const float inv = 1.f / sum;
  __syncwarp();
  for (int p = s + lane; p < e; p += 32) alpha[p] *= inv;
  __syncwarp();
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int p = s; p < e; ++p) acc = fmaf(alpha[p], __ldg(h + (int64_t)colidx[p] * C + c), acc);
    out[(int64_t)i * C + c] = acc;
  }
}

__global__ void __launch_bounds__(256)
k_gat_bwd(const int* __restrict__ rowptr, const int* __restrict__ colidx, int N, int C, const float* __restrict__ h,
          const float* __restrict__ a_src, const float* __restrict__ a_dst, float slope,
          const float* __restrict__ alpha, const float* __restrict__ gout, float* __restrict__ dh,
          float* __restrict__ da_src, float* __restrict__ da_dst, float* __restrict__ dalpha) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const int s = rowptr[i], e = rowptr[i + 1];
  const float* gi = gout + (int64_t)i * C;
  float t = 0.f;
  for (int p = s; p < e; ++p) {                        // d alpha_p = <g_i, h_j>; dh_j += alpha_p g_i
    const int j = colidx[p];
    const float a = alpha[p];
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = gi[c];
      dot = fmaf(g, __ldg(h + (int64_t)j * C + c), dot);
      atomicAdd(dh + (int64_t)j * C + c, a * g);
    }
    dot = warp_sum(dot);
    if (lane == 0) dalpha[p] = dot;
    t = fmaf(a, dot, t);
  }
  __syncwarp();
  const float ad = a_dst[i];
  float acc_dst = 0.f;
  for (int p = s + lane; p < e; p += 32) {             // softmax and leaky-relu backward
    const int j = colidx[p];
    const float ds = alpha[p] * (dalpha[p] - t);
    const float raw = a_src[j] + ad;
    const float de = raw > 0.f ? ds : slope * ds;
    atomicAdd(da_src + j, de);
    acc_dst += de;
  }
  acc_dst = warp_sum(acc_dst);
  if (lane == 0) da_dst[i] = acc_dst;
}

}  // namespace
}  // namespace gda

using namespace gda;

extern "C" {

int gda_gat_fwd(const gda_graph_t* g, const float* h, int C, const float* a_src, const float* a_dst,
                float negative_slope, float* out, float* alpha, gda_stream_t stream) {
  GDA_REQUIRE(g && !g->peer_packed, "gda_gat_fwd: needs a whole (unpartitioned) graph");
  GDA_REQUIRE(C > 0, "gda_gat_fwd: C must be positive");
  if (g->N == 0) return GDA_OK;
  GDA_REQUIRE(h && a_src && a_dst && out && alpha, "gda_gat_fwd: NULL pointer");
  k_gat_fwd<<<static_cast<unsigned>(ceil_div(g->N, 8)), 256, 0, as_stream(stream)>>>(
      g->csr.rowptr, g->csr.colidx, static_cast<int>(g->N), C, h, a_src, a_dst, negative_slope, out, alpha);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_gat_bwd(const gda_graph_t* g, const float* h, int C, const float* a_src, const float* a_dst,
                float negative_slope, const float* alpha, const float* gout, float* dh, float* da_src,
                float* da_dst, float* scratch, gda_stream_t stream) {
  GDA_REQUIRE(g && !g->peer_packed, "gda_gat_bwd: needs a whole (unpartitioned) graph");
  if (g->N == 0) return GDA_OK;
  GDA_REQUIRE(h && a_src && a_dst && alpha && gout && dh && da_src && da_dst && scratch, "gda_gat_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(dh, 0, sizeof(float) * g->N * C, st));
  GDA_CUDA(cudaMemsetAsync(da_src, 0, sizeof(float) * g->N, st));
  k_gat_bwd<<<static_cast<unsigned>(ceil_div(g->N, 8)), 256, 0, st>>>(
      g->csr.rowptr, g->csr.colidx, static_cast<int>(g->N), C, h, a_src, a_dst, negative_slope, alpha, gout, dh,
      da_src, da_dst, scratch);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
