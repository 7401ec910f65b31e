// Two-view attention fusion of UDAGCN(ppmi=True): Attention.forward (pygda/nn/attention.py:52-55)
//   stacked = stack([x0, x1], dim=1); weights = softmax(Linear(H, 1)(stacked), dim=1); out = sum(stacked * weights, 1)
// as one pass (HBM bound: read 2 N H, write N H) instead of stack + matmul + softmax + mul + sum, and its backward.
// One warp per row; a0 = softmax weight of view 0 is kept for the backward ([N] floats).
#include "common.cuh"

namespace gda {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void k_attn2_fwd(const float* __restrict__ x0, const float* __restrict__ x1, int64_t N, int H,
                            const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out,
                            float* __restrict__ a0_out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float bias = b ? __ldg(b) : 0.f;
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += nwarps) {
    const float* __restrict__ p0 = x0 + r * H;
    const float* __restrict__ p1 = x1 + r * H;
    float s0 = 0.f, s1 = 0.f;
    for (int c = lane; c < H; c += 32) {
      const float wc = __ldg(w + c);
      s0 = fmaf(wc, p0[c], s0);
      s1 = fmaf(wc, p1[c], s1);
    }
    s0 = warp_sum(s0) + bias;
    s1 = warp_sum(s1) + bias;
    const float m = fmaxf(s0, s1);
    const float e0 = __expf(s0 - m), e1 = __expf(s1 - m);
    const float a0 = e0 / (e0 + e1), a1 = e1 / (e0 + e1);
    for (int c = lane; c < H; c += 32) out[r * H + c] = a0 * p0[c] + a1 * p1[c];
    if (lane == 0) a0_out[r] = a0;
  }
}

// g0 = a0 go + ds w, g1 = a1 go - ds w, ds = a0 a1 (go.x0 - go.x1); dw += ds (x0 - x1); db = 0 (softmax shift)
__global__ void k_attn2_bwd(const float* __restrict__ x0, const float* __restrict__ x1, int64_t N, int H,
                            const float* __restrict__ w, const float* __restrict__ a0v, const float* __restrict__ go,
                            float* __restrict__ g0, float* __restrict__ g1, float* __restrict__ dw) {
  extern __shared__ float sdw[];                          // [H] per block
  for (int c = threadIdx.x; c < H; c += blockDim.x) sdw[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += nwarps) {
    const float* __restrict__ p0 = x0 + r * H;
    const float* __restrict__ p1 = x1 + r * H;
    const float* __restrict__ g = go + r * H;
    float d0 = 0.f, d1 = 0.f;
    for (int c = lane; c < H; c += 32) {
      d0 = fmaf(g[c], p0[c], d0);
      d1 = fmaf(g[c], p1[c], d1);
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    const float a0 = a0v[r], a1 = 1.f - a0;
    const float ds = a0 * a1 * (d0 - d1);
    for (int c = lane; c < H; c += 32) {
      const float wc = __ldg(w + c), gc = g[c];
      if (g0) g0[r * H + c] = fmaf(ds, wc, a0 * gc);
      if (g1) g1[r * H + c] = fmaf(-ds, wc, a1 * gc);
      atomicAdd(&sdw[c], ds * (p0[c] - p1[c]));
    }
  }
  __syncthreads();
  if (dw)
    for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(dw + c, sdw[c]);
}

}  // namespace
}  // namespace gda

extern "C" {

int gda_attention2_fwd(const float* x0, const float* x1, int64_t N, int H, const float* w, const float* b, float* out,
                       float* a0_out, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && H > 0, "gda_attention2_fwd: bad size");
  if (N == 0) return GDA_OK;
  GDA_REQUIRE(x0 && x1 && w && out && a0_out, "gda_attention2_fwd: NULL pointer");
  int64_t blocks = ceil_div(N, 8);
  if (blocks > int64_t(kNumSMs) * 8) blocks = int64_t(kNumSMs) * 8;
  k_attn2_fwd<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(x0, x1, N, H, w, b, out, a0_out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_attention2_bwd(const float* x0, const float* x1, int64_t N, int H, const float* w, const float* a0,
                       const float* gout, float* g0, float* g1, float* dw, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && H > 0 && H <= 8192, "gda_attention2_bwd: bad size (H <= 8192)");
  cudaStream_t st = as_stream(stream);
  if (dw) GDA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * H, st));
  if (N == 0) return GDA_OK;
  GDA_REQUIRE(x0 && x1 && w && a0 && gout, "gda_attention2_bwd: NULL pointer");
  int64_t blocks = ceil_div(N, 8);
  if (blocks > int64_t(kNumSMs) * 4) blocks = int64_t(kNumSMs) * 4;
  k_attn2_bwd<<<static_cast<unsigned>(blocks), 256, sizeof(float) * H, st>>>(x0, x1, N, H, w, a0, gout, g0, g1, dw);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
