// Skinny GEMMs of the classifier / discriminator heads (hid -> classes, hid -> 2):
//   fwd  Y[M,C]  = X[M,K] W[C,K]^T        (pygda/nn/prop_gcn_conv.py:205 with out_channels = C,
//                                           nn.Linear heads a2gnn_base.py:67,70)
//   dx   dX[M,K] = G[M,C] W[C,K]
//   dw   dW[C,K] = G[M,C]^T X[M,K]         (reduction over the node dimension)
// with C <= 16.  All three are pure HBM streams of the [M,K] matrix (4*M*K bytes); a tiled
// GEMM wastes >90 % of its tile on them (round-1 profile: 455 us for dw on the SIMT kernel).
// 8 lanes own one row (16-byte loads, 128 B contiguous per row and step), 4 rows per warp.
#include "common.cuh"
#include "gemm.cuh"

namespace gda {
namespace {

constexpr int MAXC = 16, LPR = 8, THREADS = 256;

// dynamic smem: W as [C][K]
template <int C_MAX>
__global__ void __launch_bounds__(THREADS)
k_skinny_fwd(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw,
             float* __restrict__ Y, int64_t ldy, int64_t M, int K, int C) {
  extern __shared__ float sW[];
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sW[i] = __ldg(W + (int64_t)(i / K) * ldw + (i % K));
  __syncthreads();
  const int lane = threadIdx.x & 31, l = lane & (LPR - 1), sub = lane / LPR;
  const unsigned gmask = ((1u << LPR) - 1u) << (sub * LPR);
  const int64_t rows_per_block = (THREADS / 32) * (32 / LPR);
  for (int64_t row = blockIdx.x * rows_per_block + (threadIdx.x >> 5) * (32 / LPR) + sub; row < M;
       row += (int64_t)gridDim.x * rows_per_block) {
    float acc[C_MAX];
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) acc[c] = 0.f;
    const float* x = X + row * ldx;
    for (int k = l * 4; k < K; k += LPR * 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + k));
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) {
        if (c < C) {
          const float4 wv = *reinterpret_cast<const float4*>(sW + c * K + k);
          acc[c] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[c]))));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) {
      if (c < C) {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(gmask, acc[c], o, LPR);
      }
    }
    if (l == 0) {
      float* y = Y + row * ldy;
#pragma unroll
      for (int c = 0; c < C_MAX; ++c)
        if (c < C) y[c] = acc[c];
    }
  }
}

template <int C_MAX>
__global__ void __launch_bounds__(THREADS)
k_skinny_dx(const float* __restrict__ G, int64_t ldg, const float* __restrict__ W, int64_t ldw,
            float* __restrict__ DX, int64_t lddx, int64_t M, int K, int C) {
  extern __shared__ float sW[];
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sW[i] = __ldg(W + (int64_t)(i / K) * ldw + (i % K));
  __syncthreads();
  const int lane = threadIdx.x & 31, l = lane & (LPR - 1), sub = lane / LPR;
  const int64_t rows_per_block = (THREADS / 32) * (32 / LPR);
  for (int64_t row = blockIdx.x * rows_per_block + (threadIdx.x >> 5) * (32 / LPR) + sub; row < M;
       row += (int64_t)gridDim.x * rows_per_block) {
    float g[C_MAX];
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) g[c] = (c < C) ? __ldg(G + row * ldg + c) : 0.f;
    float* dx = DX + row * lddx;
    for (int k = l * 4; k < K; k += LPR * 4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) {
        if (c < C) {
          const float4 wv = *reinterpret_cast<const float4*>(sW + c * K + k);
          o.x = fmaf(g[c], wv.x, o.x); o.y = fmaf(g[c], wv.y, o.y);
          o.z = fmaf(g[c], wv.z, o.z); o.w = fmaf(g[c], wv.w, o.w);
        }
      }
      *reinterpret_cast<float4*>(dx + k) = o;
    }
  }
}

// per-block partial dW[C,K] over a contiguous range of rows -> part[block][C][K]
template <int C_MAX>
__global__ void __launch_bounds__(THREADS)
k_skinny_dw(const float* __restrict__ G, int64_t ldg, const float* __restrict__ X, int64_t ldx,
            float* __restrict__ part, int64_t M, int K, int C, int64_t rows_per_block) {
  extern __shared__ float sAcc[];                         // [C][K]
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sAcc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, l = lane & (LPR - 1), sub = lane / LPR;
  const int slot = (threadIdx.x >> 5) * (32 / LPR) + sub;  // 0..31: row slot inside the block
  const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int k0 = 0; k0 < K; k0 += LPR * 4) {                // this lane's 4 columns of this K slab
    const int k = k0 + l * 4;
    float acc[C_MAX][4];
#pragma unroll
    for (int c = 0; c < C_MAX; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
    if (k < K) {
      for (int64_t row = r0 + slot; row < r1; row += 32) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(X + row * ldx + k));
#pragma unroll
        for (int c = 0; c < C_MAX; ++c) {
          if (c < C) {
            const float g = __ldg(G + row * ldg + c);
            acc[c][0] = fmaf(g, xv.x, acc[c][0]); acc[c][1] = fmaf(g, xv.y, acc[c][1]);
            acc[c][2] = fmaf(g, xv.z, acc[c][2]); acc[c][3] = fmaf(g, xv.w, acc[c][3]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C_MAX; ++c) {
        if (c < C) {
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(&sAcc[c * K + k + j], acc[c][j]);   // 32 slots -> smem
        }
      }
    }
  }
  __syncthreads();
  float* dst = part + (int64_t)blockIdx.x * C * K;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) dst[i] = sAcc[i];
}

}  // namespace

int splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float alpha, float beta, float* C, int64_t ldc,
                  cudaStream_t st);   // gemm_simt.cu

// which skinny form (0 = none, 1 fwd, 2 dx, 3 dw) a gda_gemm_f32 call maps to
int skinny_kind(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                const float* A, const float* B, const float* C) {
  auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!transA && transB && N <= MAXC && K % 4 == 0 && lda % 4 == 0 && N * K <= 12288 && al16(A) && M >= 1024)
    return 1;                                    // Y[M,N] = A[M,K] B[N,K]^T
  if (!transA && !transB && K <= MAXC && N % 4 == 0 && ldc % 4 == 0 && ldb >= N && K * N <= 12288 && al16(C) &&
      M >= 1024)
    return 2;                                    // dX[M,N] = A[M,K] B[K,N]
  if (transA && !transB && M <= MAXC && N % 4 == 0 && ldb % 4 == 0 && M * N <= 12288 && al16(B) && K >= 1024)
    return 3;                                    // dW[M,N] = A[K,M]^T B[K,N]
  return 0;
}

int64_t skinny_workspace_bytes(int kind, int64_t M, int64_t N, int64_t K) {
  if (kind != 3) return 0;
  const int64_t blocks = ceil_div(K, ceil_div(K, 4 * kNumSMs));
  return blocks * M * N * static_cast<int64_t>(sizeof(float));
}

int gemm_skinny(int kind, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B,
                int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (alpha != 1.f || beta != 0.f) return GDA_E_UNSUPPORTED;
  const int blocks = 4 * kNumSMs;
  if (kind == 1) {
    const int smem = static_cast<int>(N * K * sizeof(float));
    k_skinny_fwd<MAXC><<<blocks, THREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M, (int)K, (int)N);
    GDA_LAUNCH_CHECK();
    return GDA_OK;
  }
  if (kind == 2) {
    const int smem = static_cast<int>(K * N * sizeof(float));
    k_skinny_dx<MAXC><<<blocks, THREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M, (int)N, (int)K);
    GDA_LAUNCH_CHECK();
    return GDA_OK;
  }
  // kind 3: A = G stored [K_red, M_out], B = X stored [K_red, N]
  const int64_t rpb = ceil_div(K, blocks);
  const int64_t nblocks = ceil_div(K, rpb);
  const int64_t need = nblocks * M * N * static_cast<int64_t>(sizeof(float));
  if (!ws || ws_bytes < need) return fail(GDA_E_WORKSPACE, "gda_gemm_f32: workspace too small (skinny dW)");
  const int smem = static_cast<int>(M * N * sizeof(float));
  k_skinny_dw<MAXC><<<static_cast<unsigned>(nblocks), THREADS, smem, st>>>(A, lda, B, ldb, static_cast<float*>(ws), K,
                                                                          (int)N, (int)M, rpb);
  GDA_LAUNCH_CHECK();
  return splitk_reduce(static_cast<float*>(ws), static_cast<int>(nblocks), M, N, 1.f, 0.f, C, ldc, st);
}

}  // namespace gda
