// Multi-bandwidth Gaussian MMD, forward and backward, without the n x n x d temporary.
//
// Replaces pygda/utils/mmd.py: guassian_kernel (:4-55), get_MMD (:57-107), MMD (:109-158).
// Per sample t (blockIdx.z) with n = 2b sampled rows  total = [src[src_idx[t]] ; tgt[tgt_idx[t]]]:
//   pass 1  L2_ij = sum_k (x_ik - x_jk)^2 (the reference's arithmetic, mmd.py:44-46), tiled
//           64x64 through shared memory; sum(L2) in double.
//   pass 2  bandwidth exactly as mmd.py:50-52 (fp32), K_ij = sum_q exp(-L2_ij / bw_q),
//           loss_t = mean over the b x b blocks of sign_ij K_ij (mmd.py:100-106); stores
//           G_ij = sign_ij / b^2 * dK/dL2 in place of L2_ij and its row sums for the backward.
//   bwd     d loss_t / d x_i = 4 (rowsum_i x_i - sum_j G_ij x_j)   (bandwidth is a constant:
//           mmd.py:50 uses L2.data); scattered to the sampled rows with atomic adds.
#include "common.cuh"

namespace gda {
namespace {

constexpr int TILE = 64, DK = 32, THREADS = 256, JSPLIT = 4;

struct MmdWs {
  float* G;        // [times, n, n]
  float* rowsum;   // [times, n]
  double* sums;    // [times, 2]: sum(L2), loss
};

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline MmdWs carve(void* ws, int times, int n) {
  MmdWs w;
  char* p = static_cast<char*>(ws);
  w.G = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (int64_t)times * n * n, 256);
  w.rowsum = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (int64_t)times * n, 256);
  w.sums = reinterpret_cast<double*>(p);
  return w;
}

__device__ __forceinline__ const float* sample_row(const float* src, int64_t lds, const float* tgt, int64_t ldt,
                                                   const int64_t* src_idx, const int64_t* tgt_idx, int t, int b,
                                                   int i) {
  return (i < b) ? src + src_idx[(int64_t)t * b + i] * lds : tgt + tgt_idx[(int64_t)t * b + (i - b)] * ldt;
}

__global__ void __launch_bounds__(THREADS)
k_mmd_l2(const float* __restrict__ src, int64_t lds, const float* __restrict__ tgt, int64_t ldt, int d,
         const int64_t* __restrict__ src_idx, const int64_t* __restrict__ tgt_idx, int b, MmdWs ws) {
  __shared__ float As[DK][TILE + 1];
  __shared__ float Bs[DK][TILE + 4];
  __shared__ const float* rowA[TILE];
  __shared__ const float* rowB[TILE];
  const int n = 2 * b, t = blockIdx.z;
  if (blockIdx.x < blockIdx.y) return;                  // L2 is symmetric: upper-triangular tiles only
  const int i0 = blockIdx.y * TILE, j0 = blockIdx.x * TILE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid < TILE) {
    const int i = i0 + tid;
    rowA[tid] = i < n ? sample_row(src, lds, tgt, ldt, src_idx, tgt_idx, t, b, i) : nullptr;
  } else if (tid < 2 * TILE) {
    const int j = j0 + tid - TILE;
    rowB[tid - TILE] = j < n ? sample_row(src, lds, tgt, ldt, src_idx, tgt_idx, t, b, j) : nullptr;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += DK) {
#pragma unroll
    for (int s = 0; s < TILE * DK / THREADS; ++s) {
      const int e = tid + s * THREADS;
      const int k = e % DK, r = e / DK;
      const float* pa = rowA[r];
      const float* pb = rowB[r];
      As[k][r] = (pa && k0 + k < d) ? __ldg(pa + k0 + k) : 0.f;
      Bs[k][r] = (pb && k0 + k < d) ? __ldg(pb + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float df = a[i] - bb[j];
          acc[i][j] = fmaf(df, df, acc[i][j]);
        }
    }
    __syncthreads();
  }

  float* G = ws.G + (int64_t)t * n * n;
  double local = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      if (gj >= n) continue;
      G[(int64_t)gi * n + gj] = acc[i][j];
      local += acc[i][j];
    }
  }
  if (blockIdx.x != blockIdx.y) {                       // mirror tile: G[j, i] = G[i, j], written coalesced
    local *= 2.0;
    __syncthreads();                                    // As is free again: reuse it as a 64 x 64 staging tile
    float (*T)[TILE + 1] = reinterpret_cast<float (*)[TILE + 1]>(&As[0][0]);   // DK * (TILE+1) >= 32 * 65 floats
    for (int half = 0; half < 2; ++half) {              // 32 tile rows (i) at a time
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int li = ty * 4 + i;
        if (li / 32 == half) {
#pragma unroll
          for (int j = 0; j < 4; ++j) T[li % 32][tx * 4 + j] = acc[i][j];
        }
      }
      __syncthreads();
      for (int e = tid; e < 32 * TILE; e += THREADS) {
        const int li = e % 32, lj = e / 32;             // consecutive threads -> consecutive i (contiguous in row j)
        const int gi = i0 + half * 32 + li, gj = j0 + lj;
        if (gi < n && gj < n) G[(int64_t)gj * n + gi] = T[li][lj];
      }
      __syncthreads();
    }
  }
  // block reduction in double
  __shared__ double red[THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((tid & 31) == 0) red[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) s += red[w];
    atomicAdd(ws.sums + 2 * t, s);
  }
}

constexpr int kMaxKernels = 8;

// one block per (row i, sample t)
__global__ void __launch_bounds__(THREADS)
k_mmd_kernel(int b, float kernel_mul, int kernel_num, MmdWs ws) {
  const int n = 2 * b, t = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
  // bandwidth, mmd.py:50-52, in fp32 like the reference
  const float sum_l2 = static_cast<float>(ws.sums[2 * t]);
  float bw = (sum_l2 + 1e-6f) / static_cast<float>((int64_t)n * n - n);
  bw = bw / powf(kernel_mul, static_cast<float>(kernel_num / 2));
  float inv_bwq[kMaxKernels];
  for (int q = 0; q < kernel_num; ++q) inv_bwq[q] = 1.0f / (bw * powf(kernel_mul, static_cast<float>(q)));
  const bool pow2_chain = (kernel_mul == 2.0f);
  float* G = ws.G + ((int64_t)t * n + i) * n;
  const float inv_bb = 1.f / (static_cast<float>(b) * static_cast<float>(b));
  float loss = 0.f, rs = 0.f;
  for (int j = tid; j < n; j += THREADS) {
    const float l2 = G[j];
    float k = 0.f, dk = 0.f;
    if (pow2_chain) {
      // bandwidths are a factor 2 apart: exp(-l2/bw_q) = exp(-l2/bw_{q+1})^2 -- one expf, then squarings
      float e = expf(-l2 * inv_bwq[kernel_num - 1]);
      for (int q = kernel_num - 1; q >= 0; --q) {
        k += e;
        dk -= e * inv_bwq[q];
        e *= e;
      }
    } else {
      for (int q = 0; q < kernel_num; ++q) {
        const float e = expf(-l2 * inv_bwq[q]);
        k += e;
        dk -= e * inv_bwq[q];
      }
    }
    const float sign = ((i < b) == (j < b)) ? 1.f : -1.f;
    loss += sign * k;
    const float g = sign * inv_bb * dk;
    G[j] = g;
    rs += g;
  }
  __shared__ float shl[THREADS / 32], shr[THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    rs += __shfl_xor_sync(0xffffffffu, rs, o);
  }
  if ((tid & 31) == 0) { shl[tid >> 5] = loss; shr[tid >> 5] = rs; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f, r = 0.f;
    for (int w = 0; w < THREADS / 32; ++w) { l += shl[w]; r += shr[w]; }
    ws.rowsum[(int64_t)t * n + i] = r;
    atomicAdd(ws.sums + 2 * t + 1, static_cast<double>(l) * inv_bb);
  }
}

__global__ void k_mmd_finalize(MmdWs ws, int times, float* loss_out) {
  double s = 0.0;
  for (int t = 0; t < times; ++t) s += static_cast<float>(ws.sums[2 * t + 1]);   // each get_MMD is fp32
  *loss_out = static_cast<float>(s / times);
}

// grad tile: rows i0..i0+127 of sample t, feature columns c0..c0+127, one slice of the j (reduction) range.
// 256 threads, 8 x 8 outputs each (two 4-row and two 4-column groups, 64 apart): per j step a thread reads four float4
// from shared memory for 64 FMAs.  (Round 1's 64 x 64 tile with 4 x 4 outputs per thread read five words per 16 FMAs and
// ran at 15 TFLOP/s: 330 us for the 5.1 GFLOP of the config-2 step, 5 % of it, on the critical path after both branches.)
constexpr int BT = 128, BJ = 16;

__global__ void __launch_bounds__(THREADS, 2)
k_mmd_bwd(const float* __restrict__ src, int64_t lds, const float* __restrict__ tgt, int64_t ldt, int d,
          const int64_t* __restrict__ src_idx, const int64_t* __restrict__ tgt_idx, int b, int times,
          const float* __restrict__ grad_scale, float* __restrict__ gsrc, int64_t ldgs,
          float* __restrict__ gtgt, int64_t ldgt, MmdWs ws, int jsplit) {
  // two shared-memory stages: the global loads of block k + 1 are in flight (in registers) while block k is multiplied
  __shared__ __align__(16) float Gs[2][BJ][BT + 4];       // G[i, j] stored [j][i]
  __shared__ __align__(16) float Xs[2][BJ][BT + 4];       // x_j[c]  stored [j][c]
  const int n = 2 * b, t = blockIdx.z / jsplit, js = blockIdx.z % jsplit;
  const int jchunk = ((n + jsplit - 1) / jsplit + BJ - 1) / BJ * BJ;
  const int jbeg = js * jchunk, jend = min(n, jbeg + jchunk);
  const int i0 = blockIdx.y * BT, c0 = blockIdx.x * BT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* G = ws.G + (int64_t)t * n * n;
  constexpr int PER = BT * BJ / THREADS;                   // 8 elements of each tile per thread
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float rg[PER], rx[PER];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int s = 0; s < PER; ++s) {
      const int e = tid + s * THREADS;
      {   // G tile: j fastest in memory
        const int gi = i0 + e / BJ, gj = j0 + e % BJ;
        rg[s] = (gi < n && gj < jend) ? __ldg(G + (int64_t)gi * n + gj) : 0.f;
      }
      {   // X tile: c fastest in memory
        const int gj = j0 + e / BT, gc = c0 + e % BT;
        rx[s] = (gj < jend && gc < d) ? __ldg(sample_row(src, lds, tgt, ldt, src_idx, tgt_idx, t, b, gj) + gc) : 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int s = 0; s < PER; ++s) {
      const int e = tid + s * THREADS;
      Gs[buf][e % BJ][e / BJ] = rg[s];
      Xs[buf][e / BT][e % BT] = rx[s];
    }
  };

  int buf = 0;
  if (jbeg < jend) { fetch(jbeg); stash(0); }
  __syncthreads();
  for (int j0 = jbeg; j0 < jend; j0 += BJ) {
    const bool more = j0 + BJ < jend;
    if (more) fetch(j0 + BJ);
#pragma unroll
    for (int jj = 0; jj < BJ; ++jj) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Gs[buf][jj][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Gs[buf][jj][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Xs[buf][jj][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Xs[buf][jj][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float x[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], x[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);          // the other stage: nobody reads it during this block
    __syncthreads();
    buf ^= 1;
  }

  const float scale = 4.f * (*grad_scale) / static_cast<float>(times);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gi = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gi >= n) continue;
    const float rs = (js == 0) ? ws.rowsum[(int64_t)t * n + gi] : 0.f;   // row-sum term added once
    const float* xi = sample_row(src, lds, tgt, ldt, src_idx, tgt_idx, t, b, gi);
    float* dst = (gi < b) ? gsrc + src_idx[(int64_t)t * b + gi] * ldgs
                          : gtgt + tgt_idx[(int64_t)t * b + (gi - b)] * ldgt;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gc = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gc >= d) continue;
      const float own = (js == 0) ? rs * __ldg(xi + gc) : 0.f;
      atomicAdd(dst + gc, scale * (own - acc[i][j]));
    }
  }
}

}  // namespace
}  // namespace gda

using namespace gda;

extern "C" {

int64_t gda_mmd_workspace_bytes(int times, int b, int d) {
  (void)d;
  if (times <= 0 || b <= 0) return 0;
  const int64_t n = 2 * (int64_t)b;
  return align_up(sizeof(float) * times * n * n, 256) + align_up(sizeof(float) * times * n, 256) +
         align_up(sizeof(double) * times * 2, 256);
}

int gda_mmd_fwd(const float* src, int64_t lds, const float* tgt, int64_t ldt, int d, const int64_t* src_idx,
                const int64_t* tgt_idx, int times, int b, float kernel_mul, int kernel_num, float* loss_out,
                void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(times > 0 && b > 0 && d > 0, "gda_mmd_fwd: bad size");
  GDA_REQUIRE(kernel_num > 0 && kernel_num <= kMaxKernels, "gda_mmd_fwd: kernel_num must be in 1..8");
  GDA_REQUIRE(src && tgt && src_idx && tgt_idx && loss_out, "gda_mmd_fwd: NULL pointer");
  GDA_REQUIRE(lds >= d && ldt >= d, "gda_mmd_fwd: leading dimension < d");
  if (!workspace || workspace_bytes < gda_mmd_workspace_bytes(times, b, d))
    return fail(GDA_E_WORKSPACE, "gda_mmd_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int n = 2 * b;
  MmdWs ws = carve(workspace, times, n);
  GDA_CUDA(cudaMemsetAsync(ws.sums, 0, sizeof(double) * times * 2, st));
  const unsigned tiles = static_cast<unsigned>(ceil_div(n, TILE));
  k_mmd_l2<<<dim3(tiles, tiles, times), THREADS, 0, st>>>(src, lds, tgt, ldt, d, src_idx, tgt_idx, b, ws);
  GDA_LAUNCH_CHECK();
  k_mmd_kernel<<<dim3(n, times), THREADS, 0, st>>>(b, kernel_mul, kernel_num, ws);
  GDA_LAUNCH_CHECK();
  k_mmd_finalize<<<1, 1, 0, st>>>(ws, times, loss_out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_mmd_bwd(const float* src, int64_t lds, const float* tgt, int64_t ldt, int d, const int64_t* src_idx,
                const int64_t* tgt_idx, int times, int b, const float* grad_scale, float* gsrc, int64_t ldgs,
                float* gtgt, int64_t ldgt, void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(times > 0 && b > 0 && d > 0, "gda_mmd_bwd: bad size");
  GDA_REQUIRE(src && tgt && src_idx && tgt_idx && grad_scale && gsrc && gtgt, "gda_mmd_bwd: NULL pointer");
  if (!workspace || workspace_bytes < gda_mmd_workspace_bytes(times, b, d))
    return fail(GDA_E_WORKSPACE, "gda_mmd_bwd: workspace too small");
  const int n = 2 * b;
  MmdWs ws = carve(workspace, times, n);
  // slices of the reduction range: fill the 2 x 148 resident CTAs without going far past them
  const int64_t tiles = ceil_div(d, BT) * ceil_div(n, BT) * times;
  int jsplit = static_cast<int>((2 * kNumSMs) / (tiles > 0 ? tiles : 1));
  if (jsplit < 1) jsplit = 1;
  if (jsplit > 8) jsplit = 8;
  while (jsplit > 1 && ceil_div(n, jsplit) < 4 * BJ) --jsplit;
  dim3 grid(static_cast<unsigned>(ceil_div(d, BT)), static_cast<unsigned>(ceil_div(n, BT)), times * jsplit);
  k_mmd_bwd<<<grid, THREADS, 0, as_stream(stream)>>>(src, lds, tgt, ldt, d, src_idx, tgt_idx, b, times, grad_scale,
                                                    gsrc, ldgs, gtgt, ldgt, ws, jsplit);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
