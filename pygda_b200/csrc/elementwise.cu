// Elementwise / reduction kernels on the hot path (all HBM-bound, one pass each):
//   bias + activation + dropout        pygda/nn/a2gnn_base.py:136-138, udagcn_base.py:86-88
//   softmax cross-entropy fwd+bwd      pygda/models/a2gnn.py:182,200-204; udagcn.py:172-187
//   softmax entropy fwd+bwd            pygda/models/udagcn.py:193-199
//   segment mean (global_mean_pool)    pygda/nn/a2gnn_base.py:141, adagcn_base.py:94
//   Adam                               pygda/models/a2gnn.py:292-296,317-319 (torch.optim.Adam)
#include <cuda_bf16.h>

#include "common.cuh"

namespace gda {
namespace {

constexpr int kThreads = 256;
inline unsigned grid_for(int64_t n, int per_thread = 1) {
  int64_t b = ceil_div(ceil_div(n > 0 ? n : 1, per_thread), kThreads);
  const int64_t cap = int64_t(kNumSMs) * 16;
  return static_cast<unsigned>(b < cap ? b : cap);
}

// y_r = dropout_r(act(x + bias)) for r < rep: x [n_in] is read once, y is the stacked [rep * n_in] output and the
// keep mask of element e of copy r is hash(seed, r * n_in + e) -- `rep` independent masks over one input
// (the two bottleneck evaluations per domain share layer 1: pygda/models/a2gnn.py:181/192, 193/211).
// V = 4: 16-byte accesses (cols % 4 == 0, aligned pointers); V = 1: any shape.
template <int V>
__global__ void k_bias_act_dropout_fwd(const float* x /* may alias y when rep == 1 */, const float* __restrict__ bias,
                                       float* y, int64_t n_in, int cols, int rep, int act,
                                       uint32_t thresh, float scale, uint64_t seed,
                                       const uint64_t* __restrict__ seed_offset, int do_drop) {
  if (do_drop && seed_offset) seed += __ldg(seed_offset);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
  int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V;
  unsigned col = static_cast<unsigned>(i % cols);
  const unsigned step = static_cast<unsigned>(stride % cols);
  for (; i < n_in; i += stride) {
    float v[V];
    if (V == 4) {
      const float4 t = *reinterpret_cast<const float4*>(x + i);
      v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
    } else {
      v[0] = x[i];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (bias) v[j] += __ldg(bias + col + j);
      if (act == 1) v[j] = fmaxf(v[j], 0.f);
    }
    for (int r = 0; r < rep; ++r) {
      const int64_t o = r * n_in + i;
      float w[V];
#pragma unroll
      for (int j = 0; j < V; ++j)
        w[j] = do_drop ? (dropout_keep(seed, static_cast<uint64_t>(o + j), thresh) ? v[j] * scale : 0.f) : v[j];
      if (V == 4) *reinterpret_cast<float4*>(y + o) = make_float4(w[0], w[1 % V], w[2 % V], w[3 % V]);
      else y[o] = w[0];
    }
    col += step;
    if (col >= static_cast<unsigned>(cols)) col -= cols;
  }
}

// backward of the above.  gy_r = gy0 / gy1 (NULL = no gradient reached that copy); y is the stacked output.
// sum = 1: gx [n_in] = sum_r mask_r(gy_r) (one input, rep outputs); sum = 0: gx stacked [rep * n_in].
template <int V>
__global__ void k_bias_act_dropout_bwd(const float* __restrict__ gy0, const float* __restrict__ gy1,
                                       const float* __restrict__ y, float* __restrict__ gx, int64_t n_in, int rep,
                                       int sum, int act, uint32_t thresh, float scale, uint64_t seed,
                                       const uint64_t* __restrict__ seed_offset, int do_drop) {
  if (do_drop && seed_offset) seed += __ldg(seed_offset);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * V; i < n_in; i += stride) {
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int r = 0; r < rep; ++r) {
      const float* __restrict__ gy = r == 0 ? gy0 : gy1;
      const int64_t o = r * n_in + i;
      float g[V], yo[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { g[j] = 0.f; yo[j] = 1.f; }
      if (gy) {
        if (V == 4) {
          const float4 t = *reinterpret_cast<const float4*>(gy + i);
          g[0] = t.x; g[1 % V] = t.y; g[2 % V] = t.z; g[3 % V] = t.w;
          if (act == 1) {
            const float4 u = *reinterpret_cast<const float4*>(y + o);
            yo[0] = u.x; yo[1 % V] = u.y; yo[2 % V] = u.z; yo[3 % V] = u.w;
          }
        } else {
          g[0] = gy[i];
          if (act == 1) yo[0] = y[o];
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (do_drop) g[j] = dropout_keep(seed, static_cast<uint64_t>(o + j), thresh) ? g[j] * scale : 0.f;
          if (act == 1 && !(yo[j] > 0.f)) g[j] = 0.f;      // relu'(0) = 0, like torch
        }
      }
      if (sum) {
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += g[j];
      } else if (V == 4) {
        *reinterpret_cast<float4*>(gx + o) = make_float4(g[0], g[1 % V], g[2 % V], g[3 % V]);
      } else {
        gx[o] = g[0];
      }
    }
    if (sum) {
      if (V == 4) *reinterpret_cast<float4*>(gx + i) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
      else gx[i] = acc[0];
    }
  }
}

// ---- bf16 feature path (BASELINE config 3): casts, act/dropout and column sums on bf16 activations ----
__global__ void k_cast_f32_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

__global__ void k_cast_bf16_f32(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

// y = dropout(act(x)) on bf16; same (seed, element index) mask as the fp32 kernels and the aggregation epilogue
__global__ void k_act_dropout_bf16_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, int64_t n, int act, uint32_t thresh,
                                       float scale, uint64_t seed, const uint64_t* __restrict__ seed_offset, int do_drop) {
  if (do_drop && seed_offset) seed += __ldg(seed_offset);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __bfloat162float(x[i]);
    if (act == 1) v = fmaxf(v, 0.f);
    if (do_drop) v = dropout_keep(seed, static_cast<uint64_t>(i), thresh) ? v * scale : 0.f;
    y[i] = __float2bfloat16_rn(v);
  }
}

__global__ void k_act_dropout_bf16_bwd(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                                       __nv_bfloat16* __restrict__ gx, int64_t n, int act, uint32_t thresh, float scale,
                                       uint64_t seed, const uint64_t* __restrict__ seed_offset, int do_drop) {
  if (do_drop && seed_offset) seed += __ldg(seed_offset);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = __bfloat162float(gy[i]);
    if (do_drop) g = dropout_keep(seed, static_cast<uint64_t>(i), thresh) ? g * scale : 0.f;
    if (act == 1 && !(__bfloat162float(y[i]) > 0.f)) g = 0.f;
    gx[i] = __float2bfloat16_rn(g);
  }
}

__global__ void k_colsum_bf16(const __nv_bfloat16* __restrict__ x, int64_t rows, int cols, int64_t ldx,
                              float* __restrict__ out, int64_t rows_per_block) {
  const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += __bfloat162float(x[r * ldx + c]);
    atomicAdd(out + c, s);
  }
}

// ---- Bernstein propagation glue (pygda/nn/dgsda_base.py:135-153): y = beta * y + coef * relu(*t) * x with the
// temperature read on the device (no host sync), and d loss / d temp_k = coef * [temp_k > 0] * <g, x> ----
__global__ void k_bern_axpy(float* __restrict__ y, const float* __restrict__ x, int64_t n, float beta, float coef,
                            const float* __restrict__ t) {
  const float a = coef * fmaxf(__ldg(t), 0.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = (beta == 0.f ? 0.f : beta * y[i]) + a * x[i];
}

__global__ void k_bern_dot(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ acc) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += static_cast<double>(a[i]) * static_cast<double>(b[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

__global__ void k_bern_dtemp(const double* __restrict__ acc, float coef, const float* __restrict__ t, float* __restrict__ out) {
  *out = (__ldg(t) > 0.f) ? static_cast<float>(static_cast<double>(coef) * *acc) : 0.f;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// out[c] += sum over this block's row range; out pre-zeroed
__global__ void k_colsum(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx, float* __restrict__ out,
                         int64_t rows_per_block) {
  const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += x[r * ldx + c];
    atomicAdd(out + c, s);
  }
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

constexpr int kMaxC = 64;

__global__ void k_softmax_ce(const float* __restrict__ logits, int64_t rows, int C, int64_t ld,
                             const int64_t* __restrict__ labels, int64_t split, float* __restrict__ loss_out,
                             float* __restrict__ dlogits, float inv_rows) {
  float local = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* z = logits + r * ld;
    float v[kMaxC];
    float m = -INFINITY;
#pragma unroll 8
    for (int c = 0; c < C; ++c) { v[c] = z[c]; m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < C; ++c) s += expf(v[c] - m);
    const float lse = m + logf(s);
    int y = labels ? static_cast<int>(labels[r]) : (r >= split ? 1 : 0);
    y = min(max(y, 0), C - 1);                           // never index outside v[]; range is validated by the caller
    local += lse - v[y];
    if (dlogits) {
      float* g = dlogits + r * C;
      for (int c = 0; c < C; ++c) g[c] = (expf(v[c] - lse) - (c == y ? 1.f : 0.f)) * inv_rows;
    }
  }
  const float tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(loss_out, tot * inv_rows);
}

// Any number of classes: one warp per row, the lanes stride over the classes (max, sum of exponentials, gradient: three
// passes over the row, which stays in L1).  Used above kMaxC classes; same values as k_softmax_ce.
__global__ void k_softmax_ce_wide(const float* __restrict__ logits, int64_t rows, int C, int64_t ld,
                                  const int64_t* __restrict__ labels, int64_t split, float* __restrict__ loss_out,
                                  float* __restrict__ dlogits, float inv_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float local = 0.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* z = logits + r * ld;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = m + logf(s);
    int y = labels ? static_cast<int>(labels[r]) : (r >= split ? 1 : 0);
    y = min(max(y, 0), C - 1);                           // range is validated on the host side (ops.SoftmaxCEFn)
    if (lane == 0) local += lse - z[y];
    if (dlogits) {
      float* g = dlogits + r * C;
      for (int c = lane; c < C; c += 32) g[c] = (expf(z[c] - lse) - (c == y ? 1.f : 0.f)) * inv_rows;
    }
  }
  const float tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(loss_out, tot * inv_rows);
}

__global__ void k_softmax_entropy(const float* __restrict__ logits, int64_t rows, int C, int64_t ld,
                                  float* __restrict__ loss_out, float* __restrict__ dlogits, float inv_rows) {
  float local = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* z = logits + r * ld;
    float p[kMaxC];
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) { p[c] = z[c]; m = fmaxf(m, p[c]); }
    float s = 0.f;
    for (int c = 0; c < C; ++c) { p[c] = expf(p[c] - m); s += p[c]; }
    const float inv = 1.f / s;
    // H = sum_c -q log q with q = clamp(p, 1e-9, 1); dH/dq = -(log q + 1) where not clamped
    float ent = 0.f, dot = 0.f;
    float dq[kMaxC];
    for (int c = 0; c < C; ++c) {
      const float pc = p[c] * inv;
      p[c] = pc;
      const bool clamped = pc < 1e-9f || pc > 1.0f;
      const float q = fminf(fmaxf(pc, 1e-9f), 1.0f);
      const float lq = logf(q);
      ent -= q * lq;
      dq[c] = clamped ? 0.f : -(lq + 1.f);
      dot += dq[c] * pc;
    }
    local += ent;
    if (dlogits) {
      float* g = dlogits + r * C;
      for (int c = 0; c < C; ++c) g[c] = p[c] * (dq[c] - dot) * inv_rows;   // softmax Jacobian
    }
  }
  const float tot = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(loss_out, tot * inv_rows);
}

// one warp per graph
__global__ void k_segment_mean_fwd(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ ptr,
                                   int64_t G, int H, float* __restrict__ out) {
  const int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= G) return;
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ptr[g], r1 = ptr[g + 1];
  const float inv = 1.f / static_cast<float>(max((int64_t)1, r1 - r0));
  for (int c = lane; c < H; c += 32) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += x[r * ldx + c];
    out[g * H + c] = s * inv;
  }
}

__global__ void k_segment_mean_bwd(const float* __restrict__ gout, const int64_t* __restrict__ ptr, int64_t G,
                                   int H, float* __restrict__ gx, int64_t ldgx) {
  const int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= G) return;
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ptr[g], r1 = ptr[g + 1];
  const float inv = 1.f / static_cast<float>(max((int64_t)1, r1 - r0));
  for (int c = lane; c < H; c += 32) {
    const float v = gout[g * H + c] * inv;
    for (int64_t r = r0; r < r1; ++r) gx[r * ldgx + c] = v;
  }
}

struct AdamArgs {
  float* p[GDA_ADAM_MAX_TENSORS];
  const float* g[GDA_ADAM_MAX_TENSORS];
  float* m[GDA_ADAM_MAX_TENSORS];
  float* v[GDA_ADAM_MAX_TENSORS];
  int64_t n[GDA_ADAM_MAX_TENSORS];
};

__global__ void k_adam_tick(float* state, float beta1, float beta2) {
  const double t = static_cast<double>(state[0]) + 1.0;
  state[0] = static_cast<float>(t);
  state[1] = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
  state[2] = static_cast<float>(1.0 - pow(static_cast<double>(beta2), t));
}

// torch.optim.Adam (single-tensor formulation): g += wd*p; m = b1 m + (1-b1) g;
// v = b2 v + (1-b2) g^2; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void k_adam(AdamArgs a, float lr, float beta1, float beta2, float eps, float wd,
                       const float* __restrict__ state) {
  const int t = blockIdx.y;
  const int64_t n = a.n[t];
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const float bc1 = state[1], bc2_sqrt = sqrtf(state[2]);
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;          // lerp form, like torch
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

__global__ void k_fill(float* __restrict__ x, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void k_axpy(float* __restrict__ y, const float* __restrict__ x, int64_t n, float a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

__global__ void k_scale(float* __restrict__ y, const float* __restrict__ x, int64_t n, float a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a * x[i];
}

__global__ void k_scale_dev(float* __restrict__ y, const float* __restrict__ x, int64_t n, float a,
                            const float* __restrict__ a_dev) {
  const float s = a * __ldg(a_dev);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = s * x[i];
}

__global__ void k_counter_inc(uint64_t* c) { *c += 1; }

struct CombineArgs { const float* t[8]; float w[8]; };
__global__ void k_combine(CombineArgs a, int n, float* out) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += a.w[i] * *a.t[i];
  *out = s;
}

}  // namespace
}  // namespace gda

using namespace gda;

extern "C" {

int gda_bias_act_dropout_rep_fwd(const float* x, const float* bias, float* y, int64_t rows, int64_t cols, int rep,
                                 int act, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                                 gda_stream_t stream) {
  GDA_REQUIRE(rows >= 0 && cols >= 0, "gda_bias_act_dropout_fwd: negative size");
  GDA_REQUIRE(rep >= 1, "gda_bias_act_dropout_fwd: rep must be >= 1");
  const int64_t n = rows * cols;
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_bias_act_dropout_fwd: NULL pointer");
  GDA_REQUIRE(cols < (int64_t(1) << 31), "gda_bias_act_dropout_fwd: too many columns");
  GDA_REQUIRE(rep == 1 || x != y, "gda_bias_act_dropout_fwd: in-place needs rep == 1");
  GDA_REQUIRE(act == 0 || act == 1, "gda_bias_act_dropout_fwd: act must be 0 (none) or 1 (relu)");
  GDA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gda_bias_act_dropout_fwd: dropout_p outside [0,1)");
  const uint32_t th = dropout_threshold(dropout_p);
  const float sc = 1.f / (1.f - dropout_p);
  const int drop = dropout_p > 0.f ? 1 : 0;
  if (cols % 4 == 0 && aligned16(x) && aligned16(y))
    k_bias_act_dropout_fwd<4><<<grid_for(n, 8), kThreads, 0, as_stream(stream)>>>(
        x, bias, y, n, static_cast<int>(cols), rep, act, th, sc, seed, seed_offset, drop);
  else
    k_bias_act_dropout_fwd<1><<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(
        x, bias, y, n, static_cast<int>(cols), rep, act, th, sc, seed, seed_offset, drop);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_bias_act_dropout_fwd(const float* x, const float* bias, float* y, int64_t rows, int64_t cols, int act,
                             float dropout_p, uint64_t seed, const uint64_t* seed_offset, gda_stream_t stream) {
  return gda_bias_act_dropout_rep_fwd(x, bias, y, rows, cols, 1, act, dropout_p, seed, seed_offset, stream);
}

int gda_colsum_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* out, gda_stream_t stream) {
  GDA_REQUIRE(rows >= 0 && cols >= 0 && ldx >= cols, "gda_colsum_f32: bad size");
  if (cols == 0) return GDA_OK;
  GDA_REQUIRE(out != nullptr, "gda_colsum_f32: NULL output");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows == 0) return GDA_OK;
  GDA_REQUIRE(x != nullptr, "gda_colsum_f32: NULL input");
  const int64_t want = ceil_div(rows, 64);                      // ~64 rows per block
  const int64_t blocks = want < 16 * kNumSMs ? (want > 0 ? want : 1) : 16 * kNumSMs;
  const int64_t rpb = ceil_div(rows, blocks);
  const int threads = cols >= 256 ? 256 : (cols >= 128 ? 128 : 64);
  k_colsum<<<static_cast<unsigned>(ceil_div(rows, rpb)), threads, 0, st>>>(x, rows, static_cast<int>(cols), ldx, out, rpb);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_bias_act_dropout_rep_bwd(const float* gy0, const float* gy1, const float* y, float* gx, float* gbias,
                                 int64_t rows, int64_t cols, int rep, int sum, int act, float dropout_p, uint64_t seed,
                                 const uint64_t* seed_offset, gda_stream_t stream) {
  GDA_REQUIRE(rows >= 0 && cols >= 0, "gda_bias_act_dropout_bwd: negative size");
  GDA_REQUIRE(rep == 1 || rep == 2, "gda_bias_act_dropout_bwd: rep must be 1 or 2");
  const int64_t n = rows * cols;
  if (n > 0) {
    GDA_REQUIRE((gy0 || (rep == 2 && gy1)) && gx, "gda_bias_act_dropout_bwd: NULL pointer");
    GDA_REQUIRE(act == 0 || (act == 1 && y), "gda_bias_act_dropout_bwd: act must be 0 or 1 (1 needs y)");
    const uint32_t th = dropout_threshold(dropout_p);
    const float sc = 1.f / (1.f - dropout_p);
    const int drop = dropout_p > 0.f ? 1 : 0;
    if (cols % 4 == 0 && aligned16(gy0) && aligned16(gy1) && aligned16(y) && aligned16(gx))
      k_bias_act_dropout_bwd<4><<<grid_for(n, 8), kThreads, 0, as_stream(stream)>>>(gy0, gy1, y, gx, n, rep, sum, act, th,
                                                                                    sc, seed, seed_offset, drop);
    else
      k_bias_act_dropout_bwd<1><<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(gy0, gy1, y, gx, n, rep, sum, act, th,
                                                                                    sc, seed, seed_offset, drop);
    GDA_LAUNCH_CHECK();
  }
  if (gbias) return gda_colsum_f32(gx, (sum ? 1 : rep) * rows, cols, cols, gbias, stream);
  return GDA_OK;
}

int gda_bias_act_dropout_bwd(const float* gy, const float* y, float* gx, float* gbias, int64_t rows, int64_t cols,
                             int act, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                             gda_stream_t stream) {
  return gda_bias_act_dropout_rep_bwd(gy, nullptr, y, gx, gbias, rows, cols, 1, 0, act, dropout_p, seed, seed_offset,
                                      stream);
}

int gda_softmax_ce_fwd_bwd(const float* logits, int64_t rows, int C, int64_t ld, const int64_t* labels,
                           int64_t split, float* loss_out, float* dlogits, gda_stream_t stream) {
  GDA_REQUIRE(rows > 0 && C > 0 && ld >= C, "gda_softmax_ce_fwd_bwd: bad shape");
  GDA_REQUIRE(logits && loss_out, "gda_softmax_ce_fwd_bwd: NULL pointer");
  GDA_REQUIRE(labels || C >= 2, "gda_softmax_ce_fwd_bwd: implicit domain labels need C >= 2");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (C <= kMaxC)
    k_softmax_ce<<<grid_for(rows), kThreads, 0, st>>>(logits, rows, C, ld, labels, split, loss_out, dlogits,
                                                       1.0f / static_cast<float>(rows));
  else
    k_softmax_ce_wide<<<grid_for(rows * 32), kThreads, 0, st>>>(logits, rows, C, ld, labels, split, loss_out, dlogits,
                                                                 1.0f / static_cast<float>(rows));
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_softmax_entropy_fwd_bwd(const float* logits, int64_t rows, int C, int64_t ld, float* loss_out,
                                float* dlogits, gda_stream_t stream) {
  GDA_REQUIRE(rows > 0 && C > 0 && C <= kMaxC && ld >= C, "gda_softmax_entropy_fwd_bwd: bad shape");
  GDA_REQUIRE(logits && loss_out, "gda_softmax_entropy_fwd_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  k_softmax_entropy<<<grid_for(rows), kThreads, 0, st>>>(logits, rows, C, ld, loss_out, dlogits,
                                                          1.0f / static_cast<float>(rows));
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_segment_mean_fwd(const float* x, int64_t ldx, const int64_t* ptr, int64_t G, int H, float* out,
                         gda_stream_t stream) {
  GDA_REQUIRE(G >= 0 && H > 0, "gda_segment_mean_fwd: bad size");
  if (G == 0) return GDA_OK;
  GDA_REQUIRE(x && ptr && out, "gda_segment_mean_fwd: NULL pointer");
  k_segment_mean_fwd<<<static_cast<unsigned>(ceil_div(G, 8)), 256, 0, as_stream(stream)>>>(x, ldx, ptr, G, H, out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_segment_mean_bwd(const float* gout, const int64_t* ptr, int64_t G, int H, float* gx, int64_t ldgx,
                         gda_stream_t stream) {
  GDA_REQUIRE(G >= 0 && H > 0, "gda_segment_mean_bwd: bad size");
  if (G == 0) return GDA_OK;
  GDA_REQUIRE(gout && ptr && gx, "gda_segment_mean_bwd: NULL pointer");
  k_segment_mean_bwd<<<static_cast<unsigned>(ceil_div(G, 8)), 256, 0, as_stream(stream)>>>(gout, ptr, G, H, gx, ldgx);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_adam_step(int num_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float* state, gda_stream_t stream) {
  GDA_REQUIRE(num_tensors >= 0 && num_tensors <= GDA_ADAM_MAX_TENSORS, "gda_adam_step: too many tensors");
  GDA_REQUIRE(state != nullptr, "gda_adam_step: state is NULL");
  cudaStream_t st = as_stream(stream);
  k_adam_tick<<<1, 1, 0, st>>>(state, beta1, beta2);
  GDA_LAUNCH_CHECK();
  if (num_tensors == 0) return GDA_OK;
  GDA_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel, "gda_adam_step: NULL pointer list");
  AdamArgs a;
  int64_t maxn = 0;
  for (int i = 0; i < num_tensors; ++i) {
    GDA_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] >= 0, "gda_adam_step: NULL tensor");
    a.p[i] = params[i]; a.g[i] = grads[i]; a.m[i] = exp_avg[i]; a.v[i] = exp_avg_sq[i]; a.n[i] = numel[i];
    if (numel[i] > maxn) maxn = numel[i];
  }
  if (maxn == 0) return GDA_OK;
  int64_t bx = ceil_div(maxn, kThreads * 4);
  if (bx > 4 * kNumSMs) bx = 4 * kNumSMs;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(num_tensors));
  k_adam<<<grid, kThreads, 0, st>>>(a, lr, beta1, beta2, eps, weight_decay, state);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_fill_f32(float* x, int64_t n, float value, gda_stream_t stream) {
  if (n <= 0) return GDA_OK;
  GDA_REQUIRE(x != nullptr, "gda_fill_f32: NULL pointer");
  k_fill<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(x, n, value);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_axpy_f32(float* y, const float* x, int64_t n, float alpha, gda_stream_t stream) {
  if (n <= 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_axpy_f32: NULL pointer");
  k_axpy<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(y, x, n, alpha);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_scale_f32(float* y, const float* x, int64_t n, float alpha, gda_stream_t stream) {
  if (n <= 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_scale_f32: NULL pointer");
  k_scale<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(y, x, n, alpha);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_scale_dev_f32(float* y, const float* x, int64_t n, float alpha, const float* alpha_dev,
                      gda_stream_t stream) {
  if (n <= 0) return GDA_OK;
  GDA_REQUIRE(x && y && alpha_dev, "gda_scale_dev_f32: NULL pointer");
  k_scale_dev<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(y, x, n, alpha, alpha_dev);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_counter_inc(uint64_t* counter, gda_stream_t stream) {
  GDA_REQUIRE(counter != nullptr, "gda_counter_inc: NULL pointer");
  k_counter_inc<<<1, 1, 0, as_stream(stream)>>>(counter);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_combine_scalars(int n, const float* const* terms, const float* weights, float* out, gda_stream_t stream) {
  GDA_REQUIRE(n >= 1 && n <= 8 && terms && weights && out, "gda_combine_scalars: bad arguments");
  CombineArgs a;
  for (int i = 0; i < n; ++i) {
    GDA_REQUIRE(terms[i] != nullptr, "gda_combine_scalars: NULL term");
    a.t[i] = terms[i]; a.w[i] = weights[i];
  }
  k_combine<<<1, 1, 0, as_stream(stream)>>>(a, n, out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_cast_f32_bf16(const float* x, void* y, int64_t n, gda_stream_t stream) {
  GDA_REQUIRE(n >= 0, "gda_cast_f32_bf16: negative size");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_cast_f32_bf16: NULL pointer");
  k_cast_f32_bf16<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(x, static_cast<__nv_bfloat16*>(y), n);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_cast_bf16_f32(const void* x, float* y, int64_t n, gda_stream_t stream) {
  GDA_REQUIRE(n >= 0, "gda_cast_bf16_f32: negative size");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_cast_bf16_f32: NULL pointer");
  k_cast_bf16_f32<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), y, n);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_act_dropout_bf16_fwd(const void* x, void* y, int64_t n, int act, float dropout_p, uint64_t seed,
                             const uint64_t* seed_offset, gda_stream_t stream) {
  GDA_REQUIRE(n >= 0 && (act == 0 || act == 1) && dropout_p >= 0.f && dropout_p < 1.f, "gda_act_dropout_bf16_fwd: bad argument");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(x && y, "gda_act_dropout_bf16_fwd: NULL pointer");
  k_act_dropout_bf16_fwd<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, act, dropout_threshold(dropout_p),
      1.f / (1.f - dropout_p), seed, seed_offset, dropout_p > 0.f ? 1 : 0);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_act_dropout_bf16_bwd(const void* gy, const void* y, void* gx, int64_t n, int act, float dropout_p, uint64_t seed,
                             const uint64_t* seed_offset, gda_stream_t stream) {
  GDA_REQUIRE(n >= 0 && (act == 0 || act == 1) && dropout_p >= 0.f && dropout_p < 1.f, "gda_act_dropout_bf16_bwd: bad argument");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(gy && gx && (act == 0 || y), "gda_act_dropout_bf16_bwd: NULL pointer");
  k_act_dropout_bf16_bwd<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(gy), static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(gx), n, act,
      dropout_threshold(dropout_p), 1.f / (1.f - dropout_p), seed, seed_offset, dropout_p > 0.f ? 1 : 0);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ldx, float* out, gda_stream_t stream) {
  GDA_REQUIRE(rows >= 0 && cols >= 0 && ldx >= cols, "gda_colsum_bf16: bad size");
  if (cols == 0) return GDA_OK;
  GDA_REQUIRE(out != nullptr, "gda_colsum_bf16: NULL output");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows == 0) return GDA_OK;
  GDA_REQUIRE(x != nullptr, "gda_colsum_bf16: NULL input");
  const int64_t want = ceil_div(rows, 64);
  const int64_t blocks = want < 16 * kNumSMs ? (want > 0 ? want : 1) : 16 * kNumSMs;
  const int64_t rpb = ceil_div(rows, blocks);
  const int threads = cols >= 256 ? 256 : (cols >= 128 ? 128 : 64);
  k_colsum_bf16<<<static_cast<unsigned>(ceil_div(rows, rpb)), threads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), rows,
                                                                                static_cast<int>(cols), ldx, out, rpb);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_bern_axpy_f32(float* y, const float* x, int64_t n, float beta, float coef, const float* temp_k,
                      gda_stream_t stream) {
  GDA_REQUIRE(n >= 0, "gda_bern_axpy_f32: negative size");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(y && x && temp_k, "gda_bern_axpy_f32: NULL pointer");
  k_bern_axpy<<<grid_for(n, 4), kThreads, 0, as_stream(stream)>>>(y, x, n, beta, coef, temp_k);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_bern_dtemp_f32(const float* g, const float* x, int64_t n, float coef, const float* temp_k, float* out,
                       double* scratch, gda_stream_t stream) {
  GDA_REQUIRE(n >= 0 && temp_k && out && scratch, "gda_bern_dtemp_f32: bad argument");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  if (n > 0) {
    GDA_REQUIRE(g && x, "gda_bern_dtemp_f32: NULL pointer");
    k_bern_dot<<<grid_for(n, 8), kThreads, 0, st>>>(g, x, n, scratch);
    GDA_LAUNCH_CHECK();
  }
  k_bern_dtemp<<<1, 1, 0, st>>>(scratch, coef, temp_k, out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
