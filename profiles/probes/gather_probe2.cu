// Gather ceiling under the aggregation's real memory behaviour: each launch gathers 1.1M random rows
// from X (51 MB) and writes 100k output rows to Y (51 MB); launches ping-pong X <-> Y like A_hat^k.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe2 gather_probe2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <int UNROLL, int PER_ROW>   // every output row = sum of PER_ROW gathered rows
__global__ void k_gather_write(const float4* __restrict__ X, const int* __restrict__ idx, int n_rows, float4* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long row = warp; row < n_rows; row += nwarps) {
    float4 acc = make_float4(0, 0, 0, 0);
    const int* my = idx + row * PER_ROW;
#pragma unroll
    for (int b = 0; b < PER_ROW; b += UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(X + (long)__ldg(my + b + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    Y[row * 32 + lane] = acc;
  }
}

__global__ void k_fill_random(float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = (float)(h & 0xffffff) / 16777216.0f - 0.5f;
  }
}

int main(int argc, char** argv) {
  const bool random_data = argc > 1;      // any argument: X holds pseudo-random floats instead of zeros
  const int N = 100000, PER = 12, n_idx = N * PER;
  float4 *X, *Y; int* idx;
  cudaMalloc(&X, (size_t)N * 512); cudaMalloc(&Y, (size_t)N * 512); cudaMalloc(&idx, n_idx * 4);
  cudaMemset(X, 0, (size_t)N * 512); cudaMemset(Y, 0, (size_t)N * 512);
  if (random_data) { k_fill_random<<<1024, 256>>>((float*)X, (size_t)N * 128); k_fill_random<<<1024, 256>>>((float*)Y, (size_t)N * 128); }
  printf("X data: %s\n", random_data ? "pseudo-random floats" : "zeros");
  std::vector<int> h(n_idx);
  unsigned s = 12345;
  for (int i = 0; i < n_idx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % N; }
  cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(U, BLK, GRID, PP)                                                              \
  { float4 *a = X, *b = Y;                                                                   \
    for (int r = 0; r < 4; ++r) { k_gather_write<U, PER><<<GRID, BLK>>>(a, idx, N, b); if (PP) { float4* t = a; a = b; b = t; } } \
    cudaEventRecord(e0);                                                                     \
    for (int r = 0; r < 10; ++r) { k_gather_write<U, PER><<<GRID, BLK>>>(a, idx, N, b); if (PP) { float4* t = a; a = b; b = t; } } \
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10; \
    printf("U=%d block=%d grid=%d %s: %7.1f us/launch  %6.2f TB/s gathered\n", U, BLK, GRID,  \
           PP ? "ping-pong X<->Y" : "same X -> Y   ", ms * 1e3, (double)n_idx * 512 / ms / 1e9); }
  RUN(4, 256, 148 * 8, 0) RUN(4, 256, 148 * 8, 1) RUN(12, 256, 148 * 4, 0) RUN(12, 256, 148 * 4, 1)
  RUN(6, 128, 148 * 16, 1) RUN(4, 64, 148 * 32, 1) RUN(12, 64, 148 * 12, 1)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
