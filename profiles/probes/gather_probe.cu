// Ceiling of random 512-byte row gathers from an L2-sized table on this GPU: how fast can ANY kernel
// pull rows X[idx[i], :] (128 floats) when nothing else limits it?  Variants: registers (LDG.128,
// UNROLL independent rows per lane in flight) and cp.async (LDGSTS into a shared-memory ring).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu && ./gather_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int UNROLL>
__global__ void k_gather_reg(const float4* __restrict__ X, const int* __restrict__ idx, int n_idx, float4* out) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long base = warp * UNROLL; base + UNROLL <= n_idx; base += nwarps * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(X + (long)__ldg(idx + base + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 1234.5f) out[0] = acc;
}

template <int DEPTH, int BATCH>   // per warp: DEPTH batches of BATCH rows in a smem ring
__global__ void k_gather_async(const float4* __restrict__ X, const int* __restrict__ idx, int n_idx, float4* out) {
  extern __shared__ float4 ring[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4* my = ring + (size_t)w * DEPTH * BATCH * 32;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long per = n_idx / nwarps / BATCH * BATCH;
  const long begin = warp * per, end = begin + per;
  float4 acc = make_float4(0, 0, 0, 0);
  long pi = begin;
  int issued = 0, consumed = 0;
  for (int d = 0; d < DEPTH - 1; ++d) {
    if (pi < end) {
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const float4* src = X + (long)__ldg(idx + pi + u) * 32 + lane;
        unsigned dst = (unsigned)__cvta_generic_to_shared(my + ((issued % DEPTH) * BATCH + u) * 32 + lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
      }
      pi += BATCH;
    }
    asm volatile("cp.async.commit_group;");
    ++issued;
  }
  while (consumed < (int)(per / BATCH)) {
    if (pi < end) {
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const float4* src = X + (long)__ldg(idx + pi + u) * 32 + lane;
        unsigned dst = (unsigned)__cvta_generic_to_shared(my + ((issued % DEPTH) * BATCH + u) * 32 + lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
      }
      pi += BATCH;
    }
    asm volatile("cp.async.commit_group;");
    ++issued;
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1));
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const float4 v = my[((consumed % DEPTH) * BATCH + u) * 32 + lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    ++consumed;
  }
  if (acc.x == 1234.5f) out[0] = acc;
}

int main() {
  const int N = 100000, n_idx = 1100000 * 8;   // 8 aggregation launches' worth of row gathers
  float4* X; int* idx; float4* out;
  cudaMalloc(&X, (size_t)N * 512); cudaMalloc(&idx, n_idx * 4); cudaMalloc(&out, 64);
  cudaMemset(X, 0, (size_t)N * 512);
  std::vector<int> h(n_idx);
  unsigned s = 12345;
  for (int i = 0; i < n_idx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % N; }
  cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms) {
    printf("%-34s %8.1f us  %7.2f TB/s of gathered rows\n", name, ms * 1e3, (double)n_idx * 512 / ms / 1e9);
  };
#define RUN_REG(U, BLK, GRID)                                                        \
  { for (int r = 0; r < 2; ++r) k_gather_reg<U><<<GRID, BLK>>>(X, idx, n_idx, out);  \
    cudaEventRecord(e0); k_gather_reg<U><<<GRID, BLK>>>(X, idx, n_idx, out); cudaEventRecord(e1); \
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);          \
    char nm[64]; snprintf(nm, 64, "LDG.128 U=%d block=%d grid=%d", U, BLK, GRID); report(nm, ms); }
  RUN_REG(4, 256, 148 * 8) RUN_REG(8, 256, 148 * 8) RUN_REG(8, 256, 148 * 4) RUN_REG(16, 256, 148 * 4)
  RUN_REG(8, 128, 148 * 16) RUN_REG(16, 128, 148 * 8)
#define RUN_ASYNC(D, B, BLK, GRID)                                                   \
  { size_t sm = (size_t)(BLK / 32) * D * B * 512;                                    \
    cudaFuncSetAttribute(k_gather_async<D, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    for (int r = 0; r < 2; ++r) k_gather_async<D, B><<<GRID, BLK, sm>>>(X, idx, n_idx, out); \
    cudaEventRecord(e0); k_gather_async<D, B><<<GRID, BLK, sm>>>(X, idx, n_idx, out); cudaEventRecord(e1); \
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);          \
    char nm[64]; snprintf(nm, 64, "cp.async D=%d B=%d block=%d grid=%d", D, B, BLK, GRID); report(nm, ms); }
  RUN_ASYNC(4, 8, 64, 148 * 6) RUN_ASYNC(8, 8, 64, 148 * 3) RUN_ASYNC(8, 4, 64, 148 * 6)
  RUN_ASYNC(4, 8, 128, 148 * 3) RUN_ASYNC(16, 4, 64, 148 * 3) RUN_ASYNC(8, 8, 128, 148)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
