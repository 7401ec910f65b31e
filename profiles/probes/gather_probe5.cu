// Gather through the TMA: every neighbour row (512 B) is fetched by one cp.async.bulk into a per-warp
// shared-memory ring (completion on an mbarrier per slot), then read back with one LDS.128 per lane.
// Same workload as gather_probe2 (100k output rows = sums of 12 random rows each, X <-> Y ping-pong), so the
// numbers compare directly with the LDG.128 form (16.5 TB/s at 64 warps/SM x 4 loads in flight).
// Question answered: does taking the landing buffers out of the register file (and the requests off the
// LSU path) lift the ~2 B/clk/warp ceiling the LDG form shows?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe5 gather_probe5.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int S, int B, int PER>   // S ring slots per warp, refilled B at a time; S % B == 0, S <= 32
__global__ void k_tma_gather(const float4* __restrict__ X, const int* __restrict__ idx, int n_rows, float4* __restrict__ Y) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* ring = smem + static_cast<size_t>(warp) * S * 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(nw) * S * 512) + warp * S;
  if (lane < S) mbar_init(smem_u32(bars + lane), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long gw = blockIdx.x * static_cast<long>(nw) + warp, nwarps = static_cast<long>(gridDim.x) * nw;
  if (gw >= n_rows) return;
  const long total = ((n_rows - gw + nwarps - 1) / nwarps) * PER;
  const uint32_t ring_u = smem_u32(ring), bars_u = smem_u32(bars);
  auto issue = [&](long e) {
    const long row = gw + (e / PER) * nwarps;
    const int col = __ldg(idx + row * PER + (e % PER));
    const uint32_t s = static_cast<uint32_t>(e % S);
    mbar_expect_tx(bars_u + 8 * s, 512);
    bulk_g2s(ring_u + 512 * s, X + static_cast<long>(col) * 32, 512, bars_u + 8 * s);
  };
  if (lane < S && lane < total) issue(lane);
  float4 acc = make_float4(0, 0, 0, 0);
  for (long e = 0; e < total; ++e) {
    const uint32_t s = static_cast<uint32_t>(e % S);
    mbar_wait(bars_u + 8 * s, static_cast<uint32_t>((e / S) & 1));
    const float4 v = *reinterpret_cast<const float4*>(ring + 512 * s + 16 * lane);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    if (e % PER == PER - 1) {
      Y[(gw + (e / PER) * nwarps) * 32 + lane] = acc;
      acc = make_float4(0, 0, 0, 0);
    }
    if ((e + 1) % B == 0) {
      __syncwarp();
      const long ne = e + 1 - B + S + lane;
      if (lane < B && ne < total) issue(ne);
    }
  }
}

int main() {
  const int N = 100000, PER = 12, n_idx = N * PER;
  float4 *X, *Y; int* idx;
  cudaMalloc(&X, (size_t)N * 512); cudaMalloc(&Y, (size_t)N * 512); cudaMalloc(&idx, n_idx * 4);
  std::vector<float> hx((size_t)N * 128);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u) & 0xffff) / 65536.0f;
  cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(Y, 0, (size_t)N * 512);
  std::vector<int> h(n_idx);
  unsigned s = 12345;
  for (int i = 0; i < n_idx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % N; }
  cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
  // reference for one output row (row 777) to check the data path
  double ref = 0; for (int k = 0; k < PER; ++k) ref += hx[(size_t)h[777 * PER + k] * 128 + 5];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(S, B, WARPS, CTAS_PER_SM)                                                                       \
  { const int smem = WARPS * S * 512 + WARPS * S * 8;                                                      \
    cudaFuncSetAttribute(k_tma_gather<S, B, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
    float4 *a = X, *b = Y;                                                                                 \
    k_tma_gather<S, B, PER><<<148 * CTAS_PER_SM, WARPS * 32, smem>>>(a, idx, N, b);                        \
    float chk; cudaMemcpy(&chk, reinterpret_cast<float*>(b) + 777 * 128 + 5, 4, cudaMemcpyDeviceToHost);   \
    for (int r = 0; r < 3; ++r) { k_tma_gather<S, B, PER><<<148 * CTAS_PER_SM, WARPS * 32, smem>>>(a, idx, N, b); float4* t = a; a = b; b = t; } \
    cudaEventRecord(e0);                                                                                   \
    for (int r = 0; r < 10; ++r) { k_tma_gather<S, B, PER><<<148 * CTAS_PER_SM, WARPS * 32, smem>>>(a, idx, N, b); float4* t = a; a = b; b = t; } \
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;  \
    printf("slots/warp=%2d refill=%d warps/CTA=%d CTAs/SM=%d (%3d rows in flight/SM, %3d KB smem/SM): %7.1f us/launch  %6.2f TB/s  check %.4f vs %.4f  %s\n", \
           S, B, WARPS, CTAS_PER_SM, S * WARPS * CTAS_PER_SM, smem * CTAS_PER_SM / 1024, ms * 1e3,          \
           (double)n_idx * 512 / ms / 1e9, chk, ref, cudaGetErrorString(cudaGetLastError()));              \
    cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice); }
  RUN(8, 4, 4, 8) RUN(16, 4, 4, 6) RUN(16, 8, 4, 6) RUN(32, 8, 4, 3) RUN(8, 4, 8, 6) RUN(16, 4, 8, 3)
  RUN(12, 4, 4, 8) RUN(24, 12, 4, 4) RUN(4, 4, 8, 8) RUN(8, 8, 4, 12)
  return 0;
}
