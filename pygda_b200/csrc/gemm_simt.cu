// fp32 SIMT GEMM: C = alpha * op(A) * op(B) + beta * C (row-major, any shape/ld).
//
// The general-shape fallback behind gda_gemm_f32 (skinny heads such as hid->classes
// or hid->2, odd leading dimensions, reductions over the node dimension for weight
// gradients).  Large aligned shapes are routed to the tcgen05 kernel (gemm_tc.cu).
// 128 x BN x 16 block tile, 256 threads, 8 x (BN/16) register tile per thread;
// reductions over a long K with few output tiles are split across blockIdx.z into a
// workspace and summed in a fixed order (deterministic).
#include "common.cuh"
#include "gemm.cuh"

namespace gda {
namespace {

constexpr int BM = 128, BK = 16, THREADS = 256;

template <int BN>
__global__ void __launch_bounds__(THREADS)
k_gemm_simt(int transA, int transB, int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda,
            const float* __restrict__ B, int64_t ldb, float beta, float* __restrict__ C, int64_t ldc,
            int k_chunk, int64_t split_stride) {
  constexpr int TN = BN / 16;                       // 8 or 2
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int s = 0; s < BM * BK / THREADS; ++s) {
      const int i = tid + s * THREADS;
      int m, k;
      if (!transA) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < ke) v = transA ? __ldg(A + (int64_t)gk * lda + gm) : __ldg(A + (int64_t)gm * lda + gk);
      As[k][m] = v;
    }
#pragma unroll
    for (int s = 0; s < BN * BK / THREADS; ++s) {
      const int i = tid + s * THREADS;
      int n, k;
      if (!transB) { n = i % BN; k = i / BN; } else { k = i % BK; n = i / BK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < ke) v = transB ? __ldg(B + (int64_t)gn * ldb + gk) : __ldg(B + (int64_t)gk * ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][BN / 2 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        b[0] = b0.x; b[1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* Cz = C + blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn;
      if constexpr (TN == 8) gn = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
      else gn = n0 + tx * 2 + j;
      if (gn >= N) continue;
      float* c = Cz + (int64_t)gm * ldc + gn;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v += beta * *c;
      *c = v;
    }
  }
}

__global__ void k_splitk_reduce(const float* __restrict__ part, int splits, int64_t MN, int N, float alpha,
                                float beta, float* __restrict__ C, int64_t ldc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[z * MN + i];
  float* c = C + (i / N) * ldc + (i % N);
  float v = alpha * s;
  if (beta != 0.f) v += beta * *c;
  *c = v;
}

}  // namespace

int splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float alpha, float beta, float* C, int64_t ldc,
                  cudaStream_t st) {
  const int64_t MN = M * N;
  if (MN == 0) return GDA_OK;
  k_splitk_reduce<<<static_cast<unsigned>(ceil_div(MN, 256)), 256, 0, st>>>(part, splits, MN, (int)N, alpha, beta, C, ldc);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int simt_splits(int64_t M, int64_t N, int64_t K) {
  const int bn = N > 32 ? 128 : 32;
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, bn);
  // A reduction with few output tiles is split over K until the SMs are covered, down to 128 of K per split.  (Round 1
  // split only for K >= 4096 in chunks of >= 1024: the critic's weight gradients of AdaGCN's mini-batch step --
  // [40 | 128, rows <= 1536] x [rows, 128], ONE 128 x 128 tile -- ran 95 times per step on a single CTA for 268 us each,
  // 70 % of the step's GPU time: profiles/r2_l_config5_host/.)
  if (tiles >= 2 * kNumSMs || K < 256) return 1;
  int64_t want = ceil_div(2 * kNumSMs, tiles);
  int64_t maxs = ceil_div(K, 128);
  int64_t s = want < maxs ? want : maxs;
  return s < 1 ? 1 : static_cast<int>(s);
}

int64_t simt_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const int s = simt_splits(M, N, K);
  return s > 1 ? static_cast<int64_t>(s) * M * N * sizeof(float) : 0;
}

int gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes,
              cudaStream_t st) {
  const int splits = simt_splits(M, N, K);
  const int bn = N > 32 ? 128 : 32;
  dim3 grid(static_cast<unsigned>(ceil_div(N, bn)), static_cast<unsigned>(ceil_div(M, BM)), splits);
  GDA_REQUIRE(grid.y <= 65535, "gda_gemm_f32: M too large for the SIMT path");
  int k_chunk = static_cast<int>(ceil_div(ceil_div(K, splits), BK) * BK);
  if (k_chunk == 0) k_chunk = BK;
  float* out = C;
  int64_t out_ld = ldc, stride = 0;
  float a = alpha, b = beta;
  if (splits > 1) {
    const int64_t need = static_cast<int64_t>(splits) * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return fail(GDA_E_WORKSPACE, "gda_gemm_f32: workspace too small");
    out = static_cast<float*>(ws); out_ld = N; stride = M * N; a = 1.f; b = 0.f;
  }
  if (bn == 128)
    k_gemm_simt<128><<<grid, THREADS, 0, st>>>(transA, transB, (int)M, (int)N, (int)K, a, A, lda, B, ldb, b, out, out_ld, k_chunk, stride);
  else
    k_gemm_simt<32><<<grid, THREADS, 0, st>>>(transA, transB, (int)M, (int)N, (int)K, a, A, lda, B, ldb, b, out, out_ld, k_chunk, stride);
  GDA_LAUNCH_CHECK();
  if (splits > 1) {
    const int64_t MN = M * N;
    k_splitk_reduce<<<static_cast<unsigned>(ceil_div(MN, 256)), 256, 0, st>>>(static_cast<float*>(ws), splits, MN, (int)N, alpha, beta, C, ldc);
    GDA_LAUNCH_CHECK();
  }
  return GDA_OK;
}

}  // namespace gda
