// Why does the aggregation plateau at ~10 TB/s of gathered rows when uniform-random gathers reach
// 16.5 TB/s (gather_probe2)?  Hypothesis: the power-law column distribution -- a few hub rows
// receive a large share of the gathers and their L2 slices serialise.  This probe repeats the
// gather+write experiment with Chung-Lu-like column indices (weight (i+48)^(-2/3), as
// pygda_b200/synthetic.py), then with the top-K hub rows replicated R times at distinct addresses.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe3 gather_probe3.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

template <int UNROLL, int PER_ROW>
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

int main() {
  const int N = 100000, PER = 12, n_idx = N * PER, K = 2048, R = 8;
  const int NX = N + K * R;                         // room for replicas behind the table
  float4 *X, *Y; int* idx;
  cudaMalloc(&X, (size_t)NX * 512); cudaMalloc(&Y, (size_t)NX * 512); cudaMalloc(&idx, n_idx * 4);
  cudaMemset(X, 0, (size_t)NX * 512); cudaMemset(Y, 0, (size_t)NX * 512);
  // cumulative weights of the Chung-Lu distribution, node ids shuffled
  std::vector<double> cum(N);
  double tot = 0;
  for (int i = 0; i < N; ++i) { tot += std::pow(i + 48.0, -2.0 / 3.0); cum[i] = tot; }
  std::vector<int> perm(N);
  for (int i = 0; i < N; ++i) perm[i] = i;
  unsigned s = 777;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0; };
  for (int i = N - 1; i > 0; --i) { int j = (int)(rnd() * (i + 1)); std::swap(perm[i], perm[j]); }
  std::vector<int> rank_of(n_idx), h(n_idx);
  for (int i = 0; i < n_idx; ++i) {
    const double t = rnd() * tot;
    const int r = (int)(std::lower_bound(cum.begin(), cum.end(), t) - cum.begin());
    rank_of[i] = std::min(r, N - 1);
    h[i] = perm[rank_of[i]];
  }
  long hub_hits = 0;
  for (int i = 0; i < n_idx; ++i) hub_hits += rank_of[i] < K;
  printf("top-%d nodes receive %.1f %% of the gathers\n", K, 100.0 * hub_hits / n_idx);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name) {
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    float4 *a = X, *b = Y;
    for (int r = 0; r < 4; ++r) { k_gather_write<4, PER><<<148 * 8, 256>>>(a, idx, N, b); std::swap(a, b); }
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) { k_gather_write<4, PER><<<148 * 8, 256>>>(a, idx, N, b); std::swap(a, b); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-46s %7.1f us/launch  %6.2f TB/s gathered\n", name, ms * 1e3, (double)n_idx * 512 / ms / 1e9);
  };
  run("power-law columns");
  for (int i = 0; i < n_idx; ++i)                    // hub rows read from one of R replicas (by destination row)
    if (rank_of[i] < K) h[i] = N + rank_of[i] * R + (i / PER) % R;
  run("power-law, top-2048 hubs replicated x8");
  for (int i = 0; i < n_idx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % N; }
  run("uniform columns (reference)");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
