// Bisecting the gap between the regular gather probe (gather_probe2: 16.5 TB/s) and the production aggregation
// kernel (k_spmm_tasks: 10.4 TB/s) on the config-2 shape (100k rows, ~1.1M non-zeros, 512-byte rows, ping-pong
// buffers).  Starting from the probe's loop, ONE production feature is switched on at a time:
//   bit 0  LEN   real row lengths (Chung-Lu, exponent 2.5, offset 48, + self loop) instead of 12 per row
//   bit 1  PAIR  (colidx, weight) pairs read coalesced by the lanes and broadcast by shuffle, instead of a
//                warp-uniform __ldg per index
//   bit 2  WGT   multiply by the edge weight (fmaf) instead of a plain add
//   bit 3  SORT  rows visited in descending-length order (the work list) instead of 0..N-1
//   bit 4  TAIL  exact tail batches (1..3 gathers) instead of padding the row length to a multiple of 4
//   bit 5  PLAW  power-law column indices (hubs are gathered far more often) instead of uniform ones
// For every mask in a Gray-code-like walk (0, each single bit, all bits) the launch time and the gather rate are
// printed, at 40 and 64 warps per SM.  Whichever bit drops the rate from ~16 to ~10 TB/s is the feature to redesign.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe6 gather_probe6.cu && ./gather_probe6
// (written in round 1 after the GPU budget was spent: compiled, NOT yet run -- first job of round 2)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

constexpr int U = 4;

// LD: cache policy of the 16-byte row gathers -- 0 = ld.global.nc (__ldg, what the library ships), 1 = ld.global.cg
// (L2 only), 2 = ld.global.nc.L1::no_allocate (the `make VARIANT=ldcg|noalloc` builds of libgda)
template <int LD>
__device__ __forceinline__ float4 gather16(const float4* p) {
  if (LD == 1) return __ldcg(p);
  if (LD == 2) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
  }
  return __ldg(p);
}

template <bool PAIR, bool WGT, bool TAIL, int LD = 0>
__global__ void k_probe(const float4* __restrict__ X, const int* __restrict__ rowptr, const int* __restrict__ col,
                        const float* __restrict__ val, const int* __restrict__ order, int n_rows, float4* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long t = warp; t < n_rows; t += nwarps) {
    const int row = __ldg(order + t);
    const int p0 = __ldg(rowptr + row), p1 = __ldg(rowptr + row + 1);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int base = p0; base < p1; base += 32) {
      const int n = min(32, p1 - base);
      int myc = 0;
      float myv = 0.f;
      if (PAIR && lane < n) { myc = __ldg(col + base + lane); myv = __ldg(val + base + lane); }
      for (int j = 0; j < n; j += U) {
        const int k = TAIL ? min(U, n - j) : U;          // !TAIL: lengths are padded to a multiple of U on the host
        float4 v[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u < k) {
            const int c = PAIR ? __shfl_sync(0xffffffffu, myc, j + u) : __ldg(col + base + j + u);
            w[u] = PAIR ? __shfl_sync(0xffffffffu, myv, j + u) : (WGT ? __ldg(val + base + j + u) : 1.f);
            v[u] = gather16<LD>(X + (long)c * 32 + lane);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u < k) {
            if (WGT) { acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
                       acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w); }
            else { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
          }
        }
      }
    }
    Y[(long)row * 32 + lane] = acc;
  }
}

struct Graph { std::vector<int> rowptr, col, order; std::vector<float> val; long nnz; };

static Graph make_graph(int N, bool real_len, bool plaw, bool sorted, bool pad4) {
  std::mt19937 rng(12345);
  std::vector<double> wgt(N);
  for (int i = 0; i < N; ++i) wgt[i] = std::pow(i + 48.0, -1.0 / 1.5);
  std::discrete_distribution<int> pl(wgt.begin(), wgt.end());
  std::uniform_int_distribution<int> un(0, N - 1);
  std::vector<int> perm(N);
  std::iota(perm.begin(), perm.end(), 0);
  std::shuffle(perm.begin(), perm.end(), rng);
  std::vector<int> len(N, 12);
  if (real_len) {                                        // degrees of a Chung-Lu graph with 1M directed edges, + self loop
    std::fill(len.begin(), len.end(), 1);
    for (long e = 0; e < 1000000; ++e) len[perm[pl(rng)]]++;
    for (int i = 0; i < N; ++i) len[i] = std::min(len[i], 64);       // the production kernel cuts rows at 64
  }
  if (pad4) for (int i = 0; i < N; ++i) len[i] = (len[i] + 3) / 4 * 4;
  Graph g;
  g.rowptr.assign(N + 1, 0);
  for (int i = 0; i < N; ++i) g.rowptr[i + 1] = g.rowptr[i] + len[i];
  g.nnz = g.rowptr[N];
  g.col.resize(g.nnz);
  g.val.resize(g.nnz);
  for (long p = 0; p < g.nnz; ++p) { g.col[p] = plaw ? perm[pl(rng)] : un(rng); g.val[p] = 0.25f + (p % 7) * 0.125f; }
  g.order.resize(N);
  std::iota(g.order.begin(), g.order.end(), 0);
  if (sorted) std::stable_sort(g.order.begin(), g.order.end(), [&](int a, int b) { return len[a] > len[b]; });
  return g;
}

template <bool PAIR, bool WGT, bool TAIL, int LD = 0>
static float run(const Graph& g, int N, float4* X, float4* Y, int* d_rp, int* d_col, float* d_val, int* d_ord, int ctas) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float4 *a = X, *b = Y;
  for (int r = 0; r < 4; ++r) { k_probe<PAIR, WGT, TAIL, LD><<<148 * ctas, 128>>>(a, d_rp, d_col, d_val, d_ord, N, b); std::swap(a, b); }
  cudaEventRecord(e0);
  for (int r = 0; r < 20; ++r) { k_probe<PAIR, WGT, TAIL, LD><<<148 * ctas, 128>>>(a, d_rp, d_col, d_val, d_ord, N, b); std::swap(a, b); }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20;
}

int main() {
  const int N = 100000;
  float4 *X, *Y;
  cudaMalloc(&X, (size_t)N * 512); cudaMalloc(&Y, (size_t)N * 512);
  cudaMemset(X, 0, (size_t)N * 512); cudaMemset(Y, 0, (size_t)N * 512);
  const int masks[] = {0, 1, 2, 4, 8, 16 | 1, 32, 1 | 2 | 4, 1 | 8, 1 | 2 | 4 | 8 | 16, 63};
  for (int mask : masks) {
    const bool LEN = mask & 1, PAIR = mask & 2, WGT = mask & 4, SORT = mask & 8, TAIL = mask & 16, PLAW = mask & 32;
    Graph g = make_graph(N, LEN, PLAW, SORT, !TAIL);
    int *d_rp, *d_col, *d_ord;
    float* d_val;
    cudaMalloc(&d_rp, (N + 1) * 4); cudaMalloc(&d_col, g.nnz * 4); cudaMalloc(&d_val, g.nnz * 4); cudaMalloc(&d_ord, N * 4);
    cudaMemcpy(d_rp, g.rowptr.data(), (N + 1) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_col, g.col.data(), g.nnz * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_val, g.val.data(), g.nnz * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ord, g.order.data(), N * 4, cudaMemcpyHostToDevice);
    for (int ctas : {10, 16}) {
      float ms;
#define GO(P, W, T) ms = run<P, W, T>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas)
      if (PAIR) { if (WGT) { if (TAIL) GO(true, true, true); else GO(true, true, false); }
                  else { if (TAIL) GO(true, false, true); else GO(true, false, false); } }
      else { if (WGT) { if (TAIL) GO(false, true, true); else GO(false, true, false); }
             else { if (TAIL) GO(false, false, true); else GO(false, false, false); } }
#undef GO
      printf("mask %2d [%s%s%s%s%s%s] nnz=%ld  %d warps/SM: %7.1f us/launch  %6.2f TB/s gathered\n", mask, LEN ? "LEN " : "",
             PAIR ? "PAIR " : "", WGT ? "WGT " : "", SORT ? "SORT " : "", TAIL ? "TAIL " : "", PLAW ? "PLAW" : "", g.nnz,
             ctas * 4, ms * 1e3, (double)g.nnz * 512 / ms / 1e9);
    }
    if (mask == 63 || mask == 0) {                       // gather cache policy on the two end points
      for (int ctas : {10, 16}) {
        float t[3];
        if (mask == 63) { t[0] = run<true, true, true, 0>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas);
                          t[1] = run<true, true, true, 1>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas);
                          t[2] = run<true, true, true, 2>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas); }
        else { t[0] = run<false, false, false, 0>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas);
               t[1] = run<false, false, false, 1>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas);
               t[2] = run<false, false, false, 2>(g, N, X, Y, d_rp, d_col, d_val, d_ord, ctas); }
        printf("mask %2d  %d warps/SM  gather policy  nc: %6.1f us  cg: %6.1f us  nc.L1::no_allocate: %6.1f us\n", mask,
               ctas * 4, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3);
      }
    }
    cudaFree(d_rp); cudaFree(d_col); cudaFree(d_val); cudaFree(d_ord);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
