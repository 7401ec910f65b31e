// The steps either side of the hot path, on the device:
//   gda_collate_graphs   the graph-mode mini-batch collation the reference gets from PyG's
//                        DataLoader / Batch.from_data_list (pygda/models/a2gnn.py:276-286, same block in
//                        udagcn / grade / adagcn): node-wise cat of x, edge_index offset by the running node
//                        count, `batch` = graph id per node, from a dataset kept resident in HBM
//   gda_unpack_rows_f32  the per-step `batch.to(device)` of the fit loop (pygda/models/a2gnn.py:311-312) for sparse
//                        (bag-of-words) feature matrices: the pinned host copy is kept row-compressed, only the
//                        non-zeros cross PCIe and the dense [N, F] matrix is rebuilt on the device, bit for bit
//   gda_argmax_confusion the per-epoch training score (pygda/models/a2gnn.py:328-329:
//                        eval_micro_f1(labels, logits.argmax(dim=1)) -> .cpu().numpy() -> sklearn): argmax fused
//                        with a C x C confusion count, so that C*C integers cross PCIe instead of 2N labels
#include "common.cuh"

namespace gda {
namespace {

// first i in [0, n) with ptr[i + 1] > v, for a non-decreasing ptr[0..n] with ptr[0] <= v < ptr[n]
__device__ __forceinline__ int64_t owner_of(const int64_t* __restrict__ ptr, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid + 1) > v) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// one warp per output node: copies the feature row, writes batch[i]
__global__ void k_collate_nodes(const float* __restrict__ x_all, int F, const int64_t* __restrict__ node_ptr,
                                const int64_t* __restrict__ ids, int64_t B, const int64_t* __restrict__ out_node_ptr,
                                float* __restrict__ x_out, int64_t* __restrict__ batch_out) {
  const int64_t total = __ldg(out_node_ptr + B);
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += nwarps) {
    const int64_t g = owner_of(out_node_ptr, B, i);
    const int64_t src = __ldg(node_ptr + __ldg(ids + g)) + (i - __ldg(out_node_ptr + g));
    const float* __restrict__ s = x_all + src * F;
    float* __restrict__ d = x_out + i * F;
    for (int c = lane; c < F; c += 32) d[c] = __ldg(s + c);
    if (lane == 0) batch_out[i] = g;
  }
}

// one thread per output edge: both endpoints re-based from dataset node ids to batch node ids
__global__ void k_collate_edges(const int64_t* __restrict__ ei_all, int64_t E_all, const int64_t* __restrict__ node_ptr,
                                const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ ids, int64_t B,
                                const int64_t* __restrict__ out_node_ptr, const int64_t* __restrict__ out_edge_ptr,
                                int64_t* __restrict__ ei_out) {
  const int64_t total = __ldg(out_edge_ptr + B);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = owner_of(out_edge_ptr, B, e);
    const int64_t id = __ldg(ids + g);
    const int64_t src = __ldg(edge_ptr + id) + (e - __ldg(out_edge_ptr + g));
    const int64_t shift = __ldg(out_node_ptr + g) - __ldg(node_ptr + id);
    ei_out[e] = __ldg(ei_all + src) + shift;
    ei_out[total + e] = __ldg(ei_all + E_all + src) + shift;
  }
}

// one warp per row of a row-compressed feature matrix: out[r, cols[p]] = vals[p]; the rest of the row is zero-filled
// by the memset that precedes the launch
template <typename CI>
__global__ void k_unpack_rows(const float* __restrict__ vals, const CI* __restrict__ cols, const int64_t* __restrict__ rowptr,
                              int64_t N, int64_t ldo, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += nwarps) {
    const int64_t p0 = __ldg(rowptr + r), p1 = __ldg(rowptr + r + 1);
    float* __restrict__ row = out + r * ldo;
    for (int64_t p = p0 + lane; p < p1; p += 32) row[static_cast<int64_t>(__ldg(cols + p))] = __ldg(vals + p);
  }
}

// The same from DELTA-coded column ids, one byte per entry: within a row the running column starts at 0 and every byte
// adds its value; a byte of 255 only advances (escape), any other byte b advances by b and emits the row's next value.
// One warp per row, 32 bytes per round: inclusive scan of the increments, ballot / popcount for the rank of the
// emitting bytes.  Half the index bytes of the uint16 form (bag-of-words rows have gaps of ~1/density << 255).
__global__ void k_unpack_rows_delta(const float* __restrict__ vals, const uint8_t* __restrict__ deltas,
                                    const int32_t* __restrict__ val_ptr, const int32_t* __restrict__ byte_ptr,
                                    int64_t N, int64_t F, int64_t ldo, float* __restrict__ out, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += nwarps) {
    const int b0 = __ldg(byte_ptr + r), b1 = __ldg(byte_ptr + r + 1);
    const float* __restrict__ v = vals + __ldg(val_ptr + r);
    float* __restrict__ row = out + r * ldo;
    int col = 0, cnt = 0;
    for (int base = b0; base < b1; base += 32) {
      const bool valid = base + lane < b1;
      const int byte = valid ? static_cast<int>(__ldg(deltas + base + lane)) : 0;
      int incl = byte;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const bool emit = valid && byte != 255;
      const unsigned m = __ballot_sync(0xffffffffu, emit);
      if (emit) {
        const int c = col + incl;
        if (c < F) row[c] = __ldg(v + cnt + __popc(m & ((1u << lane) - 1u)));
        else if (err) *err = 1;
      }
      col += __shfl_sync(0xffffffffu, incl, 31);
      cnt += __popc(m);
    }
  }
}

// Dense rebuild from the TILE-PACKED form (gemm_xt.cu / pygda_b200.data.PackedTiles): one warp per sub-tile of
// 32 rows x 64 columns; entry i of a sub-tile sits at position ((#segment boundaries <= i) << 8) | code.
__global__ void k_unpack_tiles(const float* __restrict__ vals, const uint8_t* __restrict__ codes,
                               const int32_t* __restrict__ ptr, const uint16_t* __restrict__ seg,
                               int64_t N, int64_t F, int nkb, int64_t ntiles, int64_t ldo, float* __restrict__ out,
                               int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); t < ntiles; t += nwarps) {
    const int e0 = __ldg(ptr + t), n = __ldg(ptr + t + 1) - e0;
    int sg[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) sg[k] = __ldg(seg + t * 8 + k);
    const int64_t row0 = (t / nkb) * 32, col0 = (t % nkb) * 64;
    for (int i = lane; i < n; i += 32) {
      int h = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) h += i >= sg[k];
      const int p = (h << 8) | static_cast<int>(__ldg(codes + e0 + i));
      const int64_t r = row0 + (p >> 6), c = col0 + (p & 63);
      if (r < N && c < F) out[r * ldo + c] = __ldg(vals + e0 + i);
      else if (err) *err = 1;
    }
  }
}

// fp32 values staged with their exponents entropy-packed (pygda_b200.data.PackedTiles, the pinned staging form of a
// sparse x): per value 3 bytes = sign << 23 | mantissa, and a 4-bit exponent code = exponent - meta[0] (15 = escape:
// the value is in the escape list).  Lossless for ANY fp32 input; 3.5 instead of 4 bytes per value when the exponents
// cluster (row-normalised bag-of-words features span ~12 binades).  One thread rebuilds four values.
__global__ void k_unpack_values(const uint32_t* __restrict__ m24, const uint16_t* __restrict__ ecode,
                                const int32_t* __restrict__ meta, int64_t n, float* __restrict__ out) {
  const uint32_t base = static_cast<uint32_t>(__ldg(meta));
  const int64_t groups = (n + 3) / 4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < groups; t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w0 = __ldg(m24 + 3 * t), w1 = __ldg(m24 + 3 * t + 1), w2 = __ldg(m24 + 3 * t + 2);
    const uint32_t e = __ldg(ecode + t);
    const uint32_t v[4] = {w0 & 0xFFFFFFu, (w0 >> 24) | ((w1 & 0xFFFFu) << 8), (w1 >> 16) | ((w2 & 0xFFu) << 16), w2 >> 8};
    uint32_t bits[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t code = (e >> (4 * k)) & 15u;
      bits[k] = ((v[k] >> 23) << 31) | ((base + code) << 23) | (v[k] & 0x7FFFFFu);     // escapes are patched afterwards
    }
    if (4 * t + 4 <= n) {
      reinterpret_cast<uint4*>(out)[t] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
    } else {
      for (int k = 0; 4 * t + k < n; ++k) out[4 * t + k] = __uint_as_float(bits[k]);
    }
  }
}

__global__ void k_patch_values(const int32_t* __restrict__ idx, const float* __restrict__ val, const int32_t* __restrict__ meta,
                               int64_t cap, int64_t n, float* __restrict__ out) {
  const int64_t cnt = min(static_cast<int64_t>(__ldg(meta + 1)), cap);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < cnt; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = __ldg(idx + j);
    if (i >= 0 && i < n) out[i] = __ldg(val + j);
  }
}

constexpr int kMaxClasses = 64;

// argmax (first maximal index, like torch.argmax on the CPU) + confusion counts[label * C + pred]
__global__ void k_argmax_confusion(const float* __restrict__ logits, int64_t rows, int C, int64_t ld,
                                   const int64_t* __restrict__ labels, int64_t* __restrict__ pred_out,
                                   unsigned long long* __restrict__ counts, int* __restrict__ bad_label) {
  extern __shared__ unsigned int hist[];                  // C * C
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* __restrict__ z = logits + r * ld;
    int best = 0;
    float bv = z[0];
    for (int c = 1; c < C; ++c) {
      const float v = z[c];
      if (v > bv || (v != v && !(bv != bv))) { bv = v; best = c; }      // NaN counts as maximal, like torch
    }
    if (pred_out) pred_out[r] = best;
    const int64_t y = labels[r];
    if (y < 0 || y >= C) { atomicExch(bad_label, 1); continue; }
    atomicAdd(&hist[static_cast<int>(y) * C + best], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (hist[i]) atomicAdd(counts + i, static_cast<unsigned long long>(hist[i]));
}

}  // namespace
}  // namespace gda

extern "C" {

int gda_collate_graphs(const float* x_all, int F, const int64_t* edge_index_all, int64_t E_all,
                       const int64_t* node_ptr, const int64_t* edge_ptr, const int64_t* graph_ids, int64_t B,
                       const int64_t* out_node_ptr, const int64_t* out_edge_ptr, int64_t num_nodes_out,
                       int64_t num_edges_out, float* x_out, int64_t* edge_index_out, int64_t* batch_out,
                       gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(B >= 0 && F > 0 && E_all >= 0 && num_nodes_out >= 0 && num_edges_out >= 0, "gda_collate_graphs: bad size");
  if (B == 0) return GDA_OK;
  GDA_REQUIRE(node_ptr && edge_ptr && graph_ids && out_node_ptr && out_edge_ptr, "gda_collate_graphs: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (num_nodes_out > 0) {
    GDA_REQUIRE(x_all && x_out && batch_out, "gda_collate_graphs: NULL node buffer");
    int64_t blocks = ceil_div(num_nodes_out, 8);
    if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
    k_collate_nodes<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x_all, F, node_ptr, graph_ids, B, out_node_ptr, x_out,
                                                                  batch_out);
    GDA_LAUNCH_CHECK();
  }
  if (num_edges_out > 0) {
    GDA_REQUIRE(edge_index_all && edge_index_out, "gda_collate_graphs: NULL edge buffer");
    int64_t blocks = ceil_div(num_edges_out, 256);
    if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
    k_collate_edges<<<static_cast<unsigned>(blocks), 256, 0, st>>>(edge_index_all, E_all, node_ptr, edge_ptr, graph_ids, B,
                                                                  out_node_ptr, out_edge_ptr, edge_index_out);
    GDA_LAUNCH_CHECK();
  }
  return GDA_OK;
}

int gda_argmax_confusion(const float* logits, int64_t rows, int C, int64_t ld, const int64_t* labels,
                         int64_t* pred_out, int64_t* counts, int* bad_label, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(rows >= 0 && C > 0 && C <= kMaxClasses && ld >= C, "gda_argmax_confusion: bad shape (need 0 < C <= 64)");
  GDA_REQUIRE(counts && bad_label, "gda_argmax_confusion: NULL output");
  cudaStream_t st = as_stream(stream);
  GDA_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * C * C, st));
  GDA_CUDA(cudaMemsetAsync(bad_label, 0, sizeof(int), st));
  if (rows == 0) return GDA_OK;
  GDA_REQUIRE(logits && labels, "gda_argmax_confusion: NULL input");
  int64_t blocks = ceil_div(rows, 256 * 4);
  if (blocks > int64_t(kNumSMs) * 4) blocks = int64_t(kNumSMs) * 4;
  if (blocks < 1) blocks = 1;
  k_argmax_confusion<<<static_cast<unsigned>(blocks), 256, sizeof(unsigned int) * C * C, st>>>(
      logits, rows, C, ld, labels, pred_out, reinterpret_cast<unsigned long long*>(counts), bad_label);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_unpack_rows_f32(const float* vals, const void* cols, int col_bytes, const int64_t* rowptr, int64_t N, int64_t F,
                        float* out, int64_t ldo, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && F >= 0 && ldo >= F, "gda_unpack_rows_f32: bad size");
  GDA_REQUIRE(col_bytes == 2 || col_bytes == 4, "gda_unpack_rows_f32: column ids are uint16 or int32");
  if (N == 0 || F == 0) return GDA_OK;
  GDA_REQUIRE(rowptr && out, "gda_unpack_rows_f32: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (ldo == F) {
    GDA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(N) * F, st));
  } else {
    GDA_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * F, static_cast<size_t>(N), st));
  }
  int64_t blocks = ceil_div(N, 8);
  if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
  if (col_bytes == 2)
    k_unpack_rows<uint16_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(vals, static_cast<const uint16_t*>(cols), rowptr, N,
                                                                          ldo, out);
  else
    k_unpack_rows<int32_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(vals, static_cast<const int32_t*>(cols), rowptr, N,
                                                                         ldo, out);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_unpack_rows_delta_f32(const float* vals, const uint8_t* deltas, const int32_t* val_ptr, const int32_t* byte_ptr,
                              int64_t N, int64_t F, float* out, int64_t ldo, int* error_flag, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && F >= 0 && ldo >= F, "gda_unpack_rows_delta_f32: bad size");
  if (N == 0 || F == 0) return GDA_OK;
  GDA_REQUIRE(val_ptr && byte_ptr && out, "gda_unpack_rows_delta_f32: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (ldo == F) {
    GDA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(N) * F, st));
  } else {
    GDA_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * F, static_cast<size_t>(N), st));
  }
  int64_t blocks = ceil_div(N, 8);
  if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
  k_unpack_rows_delta<<<static_cast<unsigned>(blocks), 256, 0, st>>>(vals, deltas, val_ptr, byte_ptr, N, F, ldo, out,
                                                                     error_flag);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_unpack_tiles_f32(const float* vals, const uint8_t* codes, const int32_t* ptr, const void* seg,
                         int64_t N, int64_t F, float* out, int64_t ldo, int* error_flag, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && F >= 0 && ldo >= F, "gda_unpack_tiles_f32: bad size");
  if (N == 0 || F == 0) return GDA_OK;
  GDA_REQUIRE(ptr && seg && out, "gda_unpack_tiles_f32: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (ldo == F) {
    GDA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(N) * F, st));
  } else {
    GDA_CUDA(cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * F, static_cast<size_t>(N), st));
  }
  const int nkb = static_cast<int>(ceil_div(F, 64));
  const int64_t ntiles = ceil_div(N, 32) * nkb;          // the padding strips hold no entries
  int64_t blocks = ceil_div(ntiles, 8);
  if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
  k_unpack_tiles<<<static_cast<unsigned>(blocks), 256, 0, st>>>(vals, codes, ptr, static_cast<const uint16_t*>(seg), N, F,
                                                                nkb, ntiles, ldo, out, error_flag);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_unpack_values_f32(const void* m24, const void* ecode, const int32_t* meta, const int32_t* esc_idx,
                          const float* esc_val, int64_t esc_capacity, int64_t n, float* out, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(n >= 0 && esc_capacity >= 0, "gda_unpack_values_f32: negative size");
  if (n == 0) return GDA_OK;
  GDA_REQUIRE(m24 && ecode && meta && out, "gda_unpack_values_f32: NULL pointer");
  GDA_REQUIRE(esc_capacity == 0 || (esc_idx && esc_val), "gda_unpack_values_f32: NULL escape list");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(m24) % 4 == 0 && reinterpret_cast<uintptr_t>(ecode) % 2 == 0 &&
                  reinterpret_cast<uintptr_t>(out) % 16 == 0,
              "gda_unpack_values_f32: m24 must be 4-byte, ecode 2-byte, out 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  int64_t blocks = ceil_div(ceil_div(n, 4), 256);
  if (blocks > int64_t(kNumSMs) * 16) blocks = int64_t(kNumSMs) * 16;
  k_unpack_values<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const uint32_t*>(m24),
                                                                 static_cast<const uint16_t*>(ecode), meta, n, out);
  GDA_LAUNCH_CHECK();
  if (esc_capacity > 0) {
    int64_t pb = ceil_div(esc_capacity, 256);
    if (pb > 1024) pb = 1024;
    k_patch_values<<<static_cast<unsigned>(pb), 256, 0, st>>>(esc_idx, esc_val, meta, esc_capacity, n, out);
    GDA_LAUNCH_CHECK();
  }
  return GDA_OK;
}

}  // extern "C"
