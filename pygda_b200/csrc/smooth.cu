// TDSS smoothness graph and Laplacian regulariser.
//
//  * gda_khop_create  replaces TDSS.smoothness(smooth_mode='K-hop') = (k-1) x TwoHopNeighbor
//    (torch_sparse.spspmm(A, A) pattern -> remove_self_loops -> cat with the input -> coalesce;
//    pygda/models/tdss.py:66-90) followed by add_remaining_self_loops (:376-385).  Integer work,
//    bit-exact: expand every 2-path i -> j -> k into the key i*N + k, radix sort, unique.
//  * gda_rw_create    replaces smooth_mode='RW' (tdss.py:367-373): one uniform random walk of
//    `walk_length` steps from every node, edge (v, i) for every node v visited from i, in
//    dense_to_sparse (row-major) order -- without the reference's dense N x N matrix.  The upstream
//    torch_cluster RNG stream cannot be reproduced; the walk uses the library's counter hash.
//  * gda_graph_degrees / gda_row_scale_rsqrt_f32 / gda_laplacian_finish_f32: the pieces of
//    compute_laplacian_loss (tdss.py:385-449) around two aggregation launches:
//        g = D^-1/2 f,  u = (A + A^T) g,  r = (d_out + d_in) g - u,
//        loss = 1/2 sum_e |g_row - g_col|^2 = 1/2 sum_i g_i . r_i,   d loss / d f = D^-1/2 r
//    so the [E, H] edge tensors of the reference (features[row], features[col], their difference)
//    are never formed.
#include <cub/cub.cuh>

#include <memory>
#include <vector>

#include "graph.cuh"

struct gda_edges {
  int64_t E = 0, N = 0;
  int64_t* ei = nullptr;      // [2, E] row-major, device
  ~gda_edges() { cudaFree(ei); }
};

namespace gda {
namespace {

constexpr int kT = 256;
inline unsigned nblk(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kT)); }
constexpr unsigned long long kDropKey = ~0ull;

struct Tmp {
  std::vector<void*> ptrs;
  ~Tmp() { for (void* p : ptrs) cudaFree(p); }
  template <typename T> int get(T** p, int64_t n) {
    *p = nullptr;
    GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * static_cast<size_t>(n > 0 ? n : 1)));
    ptrs.push_back(*p);
    return GDA_OK;
  }
};

__global__ void k_check_split(const int64_t* __restrict__ ei, int64_t E, int64_t N, int* __restrict__ src,
                              int* __restrict__ dst, int* __restrict__ err) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) { *err = 1; src[e] = 0; dst[e] = 0; return; }
  src[e] = static_cast<int>(s);
  dst[e] = static_cast<int>(d);
}

__global__ void k_lower_bound(const int* __restrict__ keys, int64_t n, int64_t N, int* __restrict__ ptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > N) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] < i) lo = mid + 1; else hi = mid; }
  ptr[i] = static_cast<int>(lo);
}

// number of 2-paths that start with edge e = out-degree of its head
__global__ void k_path_counts(const int* __restrict__ dst, int64_t E, const int* __restrict__ outptr,
                              long long* __restrict__ cnt) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > E) return;
  cnt[e] = e < E ? outptr[dst[e] + 1] - outptr[dst[e]] : 0;
}

__global__ void k_emit_keys(const int* __restrict__ src, const int* __restrict__ dst, int64_t E, int64_t N,
                            const int* __restrict__ outptr, const int* __restrict__ outcol, const long long* __restrict__ pos,
                            unsigned long long* __restrict__ keys) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const unsigned long long i = static_cast<unsigned long long>(src[e]);
  keys[e] = i * N + dst[e];                                       // the original edge (tdss.py:77)
  const int j = dst[e];
  unsigned long long* out = keys + E + pos[e];
  for (int p = outptr[j], q = 0; p < outptr[j + 1]; ++p, ++q) {
    const int k = outcol[p];
    out[q] = (static_cast<unsigned long long>(k) == i) ? kDropKey : i * N + k;   // remove_self_loops (:75)
  }
}

__global__ void k_keys_to_pairs(const unsigned long long* __restrict__ keys, int64_t n, int64_t N, int* __restrict__ src,
                                int* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  src[i] = static_cast<int>(keys[i] / N);
  dst[i] = static_cast<int>(keys[i] % N);
}

__global__ void k_nonloop_flag(const int* __restrict__ src, const int* __restrict__ dst, int64_t E, int* __restrict__ flag) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > E) return;
  flag[e] = (e < E && src[e] != dst[e]) ? 1 : 0;
}

// add_remaining_self_loops without attributes: kept edges in order, then loops 0..N-1
__global__ void k_write_with_loops(const int* __restrict__ src, const int* __restrict__ dst, int64_t E, const int* __restrict__ flag,
                                   const int* __restrict__ pos, int64_t kept, int64_t N, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = kept + N;
  if (i < E && flag[i]) { out[pos[i]] = src[i]; out[total + pos[i]] = dst[i]; }
  if (i < N) { out[kept + i] = i; out[total + kept + i] = i; }
}

__global__ void k_write_pairs(const int* __restrict__ src, const int* __restrict__ dst, int64_t E, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < E) { out[i] = src[i]; out[E + i] = dst[i]; }
}

__global__ void k_walk_keys(const int* __restrict__ outptr, const int* __restrict__ outcol, int64_t N, int L, uint64_t seed,
                            unsigned long long* __restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int cur = static_cast<int>(i);
  keys[i * (L + 1)] = static_cast<unsigned long long>(cur) * N + i;           // walk[:, 0] = start
  for (int t = 1; t <= L; ++t) {
    const int d = outptr[cur + 1] - outptr[cur];
    if (d > 0) {                                                             // no out-edge: stay put
      const uint32_t r = mix_hash(seed, static_cast<uint64_t>(i) * L + (t - 1));
      cur = outcol[outptr[cur] + static_cast<int>((static_cast<uint64_t>(r) * d) >> 32)];
    }
    keys[i * (L + 1) + t] = static_cast<unsigned long long>(cur) * N + i;     // adj[walk, start] = 1 (tdss.py:372)
  }
}

// CSR by source of an (src, dst) list: outptr [N+1], outcol [E] (stable)
int csr_by_source(const int* src, const int* dst, int64_t E, int64_t N, int* outptr, int* outcol, Tmp& t, cudaStream_t st) {
  int* keys_sorted;
  int rc;
  if ((rc = t.get(&keys_sorted, E))) return rc;
  if (E > 0) {
    int end_bit = 1;
    while ((int64_t(1) << end_bit) < N && end_bit < 31) ++end_bit;
    size_t bytes = 0;
    GDA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, src, keys_sorted, dst, outcol, static_cast<int>(E), 0, end_bit, st));
    char* tmp;
    if ((rc = t.get(&tmp, static_cast<int64_t>(bytes)))) return rc;
    GDA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, src, keys_sorted, dst, outcol, static_cast<int>(E), 0, end_bit, st));
  }
  k_lower_bound<<<nblk(N + 1), kT, 0, st>>>(keys_sorted, E, N, outptr);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

// sort + unique; drops kDropKey.  Returns the number of distinct keys in *n_out, keys in *out (owned by t).
int sort_unique(unsigned long long* keys, int64_t n, int64_t N, unsigned long long** out, int64_t* n_out, Tmp& t,
                cudaStream_t st) {
  unsigned long long *sorted, *uniq;
  int* d_count;
  int rc;
  if ((rc = t.get(&sorted, n)) || (rc = t.get(&uniq, n)) || (rc = t.get(&d_count, 1))) return rc;
  *out = uniq;
  *n_out = 0;
  if (n == 0) return GDA_OK;
  size_t b1 = 0, b2 = 0;
  GDA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, keys, sorted, static_cast<int>(n), 0, 64, st));
  GDA_CUDA(cub::DeviceSelect::Unique(nullptr, b2, sorted, uniq, d_count, static_cast<int>(n), st));
  char* tmp;
  if ((rc = t.get(&tmp, static_cast<int64_t>(b1 > b2 ? b1 : b2)))) return rc;
  GDA_CUDA(cub::DeviceRadixSort::SortKeys(tmp, b1, keys, sorted, static_cast<int>(n), 0, 64, st));
  GDA_CUDA(cub::DeviceSelect::Unique(tmp, b2, sorted, uniq, d_count, static_cast<int>(n), st));
  int h_count = 0;
  unsigned long long last = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  if (h_count > 0) {
    GDA_CUDA(cudaMemcpyAsync(&last, uniq + h_count - 1, sizeof(last), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    if (last == kDropKey) --h_count;
  }
  (void)N;
  *n_out = h_count;
  return GDA_OK;
}

}  // namespace
}  // namespace gda

extern "C" {

int gda_khop_create(const int64_t* edge_index, int64_t E, int64_t N, int k, gda_stream_t stream, gda_edges_t** out) {
  using namespace gda;
  GDA_REQUIRE(out != nullptr, "gda_khop_create: out is NULL");
  *out = nullptr;
  GDA_REQUIRE(N >= 0 && E >= 0 && k >= 1, "gda_khop_create: bad size / k");
  GDA_REQUIRE(E == 0 || edge_index != nullptr, "gda_khop_create: edge_index is NULL");
  GDA_REQUIRE(N < (int64_t(1) << 31) && E < (int64_t(1) << 31) - 1, "gda_khop_create: N and E must fit in int32");
  cudaStream_t st = as_stream(stream);
  Tmp t;
  int rc;
  int *src, *dst, *err;
  if ((rc = t.get(&src, E)) || (rc = t.get(&dst, E)) || (rc = t.get(&err, 1))) return rc;
  GDA_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (E > 0) { k_check_split<<<nblk(E), kT, 0, st>>>(edge_index, E, N, src, dst, err); GDA_LAUNCH_CHECK(); }
  int h_err = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  if (h_err) return fail(GDA_E_INDEX, "gda_khop_create: edge_index entry outside [0, N)");

  int64_t cur_E = E;
  for (int hop = 1; hop < k; ++hop) {                                 // one TwoHopNeighbor application
    int *outptr, *outcol;
    long long *cnt, *pos;                                             // 64-bit: hub-rich graphs overflow int32 path counts
    if ((rc = t.get(&outptr, N + 1)) || (rc = t.get(&outcol, cur_E)) || (rc = t.get(&cnt, cur_E + 1)) ||
        (rc = t.get(&pos, cur_E + 1))) return rc;
    if ((rc = csr_by_source(src, dst, cur_E, N, outptr, outcol, t, st))) return rc;
    k_path_counts<<<nblk(cur_E + 1), kT, 0, st>>>(dst, cur_E, outptr, cnt);
    GDA_LAUNCH_CHECK();
    size_t sb = 0;
    GDA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, cnt, pos, static_cast<int>(cur_E + 1), st));
    char* stmp;
    if ((rc = t.get(&stmp, static_cast<int64_t>(sb)))) return rc;
    GDA_CUDA(cub::DeviceScan::ExclusiveSum(stmp, sb, cnt, pos, static_cast<int>(cur_E + 1), st));
    long long h_paths = 0;
    GDA_CUDA(cudaMemcpyAsync(&h_paths, pos + cur_E, sizeof(long long), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    GDA_REQUIRE(h_paths >= 0 && cur_E + static_cast<int64_t>(h_paths) < (int64_t(1) << 31) - 1,
                "gda_khop_create: more than 2^31 two-hop paths (graph too dense for the k-hop smoothing)");
    const int64_t nk = cur_E + h_paths;
    unsigned long long *keys, *uniq;
    if ((rc = t.get(&keys, nk))) return rc;
    if (cur_E > 0) {
      k_emit_keys<<<nblk(cur_E), kT, 0, st>>>(src, dst, cur_E, N, outptr, outcol, pos, keys);
      GDA_LAUNCH_CHECK();
    }
    int64_t nu = 0;
    if ((rc = sort_unique(keys, nk, N, &uniq, &nu, t, st))) return rc;
    if ((rc = t.get(&src, nu)) || (rc = t.get(&dst, nu))) return rc;
    if (nu > 0) { k_keys_to_pairs<<<nblk(nu), kT, 0, st>>>(uniq, nu, N, src, dst); GDA_LAUNCH_CHECK(); }
    cur_E = nu;
  }

  // add_remaining_self_loops(edge_index, None, num_nodes=N)  (tdss.py:376 / :385)
  int *flag, *pos;
  if ((rc = t.get(&flag, cur_E + 1)) || (rc = t.get(&pos, cur_E + 1))) return rc;
  k_nonloop_flag<<<nblk(cur_E + 1), kT, 0, st>>>(src, dst, cur_E, flag);
  GDA_LAUNCH_CHECK();
  size_t sb = 0;
  GDA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, flag, pos, static_cast<int>(cur_E + 1), st));
  char* stmp;
  if ((rc = t.get(&stmp, static_cast<int64_t>(sb)))) return rc;
  GDA_CUDA(cub::DeviceScan::ExclusiveSum(stmp, sb, flag, pos, static_cast<int>(cur_E + 1), st));
  int kept = 0;
  GDA_CUDA(cudaMemcpyAsync(&kept, pos + cur_E, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  std::unique_ptr<gda_edges> res(new gda_edges());
  res->N = N;
  res->E = static_cast<int64_t>(kept) + N;
  GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(&res->ei), sizeof(int64_t) * 2 * static_cast<size_t>(res->E > 0 ? res->E : 1)));
  const int64_t span = cur_E > N ? cur_E : N;
  if (span > 0) {
    k_write_with_loops<<<nblk(span), kT, 0, st>>>(src, dst, cur_E, flag, pos, kept, N, res->ei);
    GDA_LAUNCH_CHECK();
  }
  GDA_CUDA(cudaStreamSynchronize(st));
  *out = res.release();
  return GDA_OK;
}

int gda_rw_create(const int64_t* edge_index, int64_t E, int64_t N, int walk_length, uint64_t seed, gda_stream_t stream,
                  gda_edges_t** out) {
  using namespace gda;
  GDA_REQUIRE(out != nullptr, "gda_rw_create: out is NULL");
  *out = nullptr;
  GDA_REQUIRE(N >= 0 && E >= 0 && walk_length >= 0, "gda_rw_create: bad size / walk length");
  GDA_REQUIRE(E == 0 || edge_index != nullptr, "gda_rw_create: edge_index is NULL");
  GDA_REQUIRE(N < (int64_t(1) << 31) && E < (int64_t(1) << 31) - 1 && N * (walk_length + 1) < (int64_t(1) << 31) - 1,
              "gda_rw_create: sizes must fit in int32");
  cudaStream_t st = as_stream(stream);
  Tmp t;
  int rc;
  int *src, *dst, *err, *outptr, *outcol;
  if ((rc = t.get(&src, E)) || (rc = t.get(&dst, E)) || (rc = t.get(&err, 1)) || (rc = t.get(&outptr, N + 1)) ||
      (rc = t.get(&outcol, E))) return rc;
  GDA_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (E > 0) { k_check_split<<<nblk(E), kT, 0, st>>>(edge_index, E, N, src, dst, err); GDA_LAUNCH_CHECK(); }
  int h_err = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  if (h_err) return fail(GDA_E_INDEX, "gda_rw_create: edge_index entry outside [0, N)");
  if ((rc = csr_by_source(src, dst, E, N, outptr, outcol, t, st))) return rc;
  const int64_t nk = N * (walk_length + 1);
  unsigned long long *keys, *uniq;
  if ((rc = t.get(&keys, nk))) return rc;
  if (N > 0) { k_walk_keys<<<nblk(N), kT, 0, st>>>(outptr, outcol, N, walk_length, seed, keys); GDA_LAUNCH_CHECK(); }
  int64_t nu = 0;
  if ((rc = sort_unique(keys, nk, N, &uniq, &nu, t, st))) return rc;
  int *rs, *rd;
  if ((rc = t.get(&rs, nu)) || (rc = t.get(&rd, nu))) return rc;
  std::unique_ptr<gda_edges> res(new gda_edges());
  res->N = N;
  res->E = nu;
  GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(&res->ei), sizeof(int64_t) * 2 * static_cast<size_t>(nu > 0 ? nu : 1)));
  if (nu > 0) {
    k_keys_to_pairs<<<nblk(nu), kT, 0, st>>>(uniq, nu, N, rs, rd);
    GDA_LAUNCH_CHECK();
    k_write_pairs<<<nblk(nu), kT, 0, st>>>(rs, rd, nu, res->ei);
    GDA_LAUNCH_CHECK();
  }
  GDA_CUDA(cudaStreamSynchronize(st));
  *out = res.release();
  return GDA_OK;
}

int64_t gda_edges_size(const gda_edges_t* e) { return e ? e->E : -1; }

int gda_edges_export(const gda_edges_t* e, int64_t* out, gda_stream_t stream) {
  GDA_REQUIRE(e != nullptr && (e->E == 0 || out != nullptr), "gda_edges_export: NULL argument");
  if (e->E == 0) return GDA_OK;
  GDA_CUDA(cudaMemcpyAsync(out, e->ei, sizeof(int64_t) * 2 * static_cast<size_t>(e->E), cudaMemcpyDeviceToDevice,
                           gda::as_stream(stream)));
  return GDA_OK;
}

int gda_edges_destroy(gda_edges_t* e) {
  delete e;
  return GDA_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ Laplacian loss
namespace gda {
namespace {

__global__ void k_degrees(const int* __restrict__ rowptr_in, const int* __restrict__ rowptr_out, int64_t N,
                          float* __restrict__ out_deg, float* __restrict__ in_deg) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (in_deg) in_deg[i] = static_cast<float>(rowptr_in[i + 1] - rowptr_in[i]);
  if (out_deg) out_deg[i] = static_cast<float>(rowptr_out[i + 1] - rowptr_out[i]);
}

// y[i, :] = x[i, :] * deg[i]^-1/2   (inf -> 0: tdss.py:432-433)
__global__ void k_row_scale_rsqrt(const float* __restrict__ x, int64_t ldx, const float* __restrict__ deg,
                                  float* __restrict__ y, int64_t N, int H) {
  const int64_t total = N * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H;
    const int c = static_cast<int>(i - r * H);
    const float d = __ldg(deg + r);
    const float s = d > 0.f ? 1.0f / sqrtf(d) : 0.f;
    y[i] = __ldg(x + r * ldx + c) * s;
  }
}

// r = (d_out + d_in) g - u_in - u_out ; df = D_out^-1/2 r ; block partial of sum_i g_i . r_i
__global__ void __launch_bounds__(256)
k_laplacian_finish(const float* __restrict__ g, const float* __restrict__ u_in, const float* __restrict__ u_out,
                   const float* __restrict__ out_deg, const float* __restrict__ in_deg, int64_t N, int H,
                   double* __restrict__ partial, float* __restrict__ df) {
  const int64_t total = N * H;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H;
    const float dout = __ldg(out_deg + r), c = dout + __ldg(in_deg + r);
    const float gi = g[i];
    const float ri = c * gi - u_in[i] - u_out[i];
    acc += static_cast<double>(gi) * static_cast<double>(ri);
    df[i] = (dout > 0.f ? 1.0f / sqrtf(dout) : 0.f) * ri;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void k_laplacian_sum(const double* __restrict__ partial, int n, float* __restrict__ loss) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];      // fixed order: deterministic
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = static_cast<float>(0.5 * sh[0]);
}

constexpr int kLapBlocks = 2 * kNumSMs;

}  // namespace
}  // namespace gda

extern "C" {

int gda_graph_degrees(const gda_graph_t* g, float* out_deg, float* in_deg, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(g != nullptr && (out_deg || in_deg), "gda_graph_degrees: NULL argument");
  GDA_REQUIRE(!g->peer_packed, "gda_graph_degrees: not available for a partition");
  if (g->N == 0) return GDA_OK;
  k_degrees<<<nblk(g->N), kT, 0, as_stream(stream)>>>(g->csr.rowptr, g->csr_t.rowptr, g->N, out_deg, in_deg);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_row_scale_rsqrt_f32(const float* x, int64_t ldx, const float* deg, float* y, int64_t N, int H, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && H > 0 && ldx >= H, "gda_row_scale_rsqrt_f32: bad shape");
  if (N == 0) return GDA_OK;
  GDA_REQUIRE(x && deg && y, "gda_row_scale_rsqrt_f32: NULL pointer");
  int64_t blocks = ceil_div(N * H, kT * 4);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  k_row_scale_rsqrt<<<static_cast<unsigned>(blocks), kT, 0, as_stream(stream)>>>(x, ldx, deg, y, N, H);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int64_t gda_laplacian_workspace_bytes(void) { return static_cast<int64_t>(gda::kLapBlocks) * sizeof(double); }

int gda_laplacian_finish_f32(const float* g, const float* u_in, const float* u_out, const float* out_deg,
                             const float* in_deg, int64_t N, int H, float* loss, float* df, void* workspace,
                             int64_t workspace_bytes, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && H > 0, "gda_laplacian_finish_f32: bad shape");
  GDA_REQUIRE(loss != nullptr, "gda_laplacian_finish_f32: loss is NULL");
  GDA_REQUIRE(workspace && workspace_bytes >= gda_laplacian_workspace_bytes(), "gda_laplacian_finish_f32: workspace too small");
  cudaStream_t st = as_stream(stream);
  double* partial = static_cast<double*>(workspace);
  if (N == 0) {
    GDA_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return GDA_OK;
  }
  GDA_REQUIRE(g && u_in && u_out && out_deg && in_deg && df, "gda_laplacian_finish_f32: NULL pointer");
  k_laplacian_finish<<<kLapBlocks, 256, 0, st>>>(g, u_in, u_out, out_deg, in_deg, N, H, partial, df);
  GDA_LAUNCH_CHECK();
  k_laplacian_sum<<<1, 256, 0, st>>>(partial, kLapBlocks, loss);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
