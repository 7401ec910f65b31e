// gda_graph_t: one-time device build of the normalised adjacency.
//
// Replaces gcn_norm (pygda/nn/prop_gcn_conv.py:24-81), CachedGCNConv.norm
// (pygda/nn/cached_gcn_conv.py:63-103) and the add_remaining_self_loops /
// scatter_add calls inside them, and builds CSR (rows = targets) and CSC-as-CSR
// (rows = sources) for the aggregation kernels.  Radix sort / scans come from CUB
// (header library in the CUDA toolkit): this is one-time preparation, not the
// per-step path.
#include <cstdlib>
#include <cub/cub.cuh>

#include "graph.cuh"

namespace gda {

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kThreads)); }

__global__ void k_flag_edges(const int64_t* __restrict__ ei, int64_t E, int64_t N, int drop_loops,
                             int* __restrict__ flag, int* __restrict__ err) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    *err = 1;
    flag[e] = 0;
    return;
  }
  flag[e] = (drop_loops && s == d) ? 0 : 1;
}

// existing self loop keeps ITS weight; with duplicates the last one in edge order wins
// (what a sequential index assignment does)
__global__ void k_loop_eid(const int64_t* __restrict__ ei, int64_t E, int* __restrict__ loop_eid) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = ei[e], d = ei[E + e];
  if (s == d) atomicMax(&loop_eid[s], static_cast<int>(e));
}

__global__ void k_fill_coo_edges(const int64_t* __restrict__ ei, const float* __restrict__ w, int64_t E,
                                 const int* __restrict__ flag, const int* __restrict__ pos,
                                 int* __restrict__ coo_src, int* __restrict__ coo_dst,
                                 float* __restrict__ coo_w) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E || !flag[e]) return;
  int p = pos[e];
  coo_src[p] = static_cast<int>(ei[e]);
  coo_dst[p] = static_cast<int>(ei[E + e]);
  coo_w[p] = w ? w[e] : 1.0f;
}

__global__ void k_fill_coo_loops(const float* __restrict__ w, const int* __restrict__ loop_eid, int64_t N,
                                 int64_t base, float fill, int* __restrict__ coo_src,
                                 int* __restrict__ coo_dst, float* __restrict__ coo_w) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  coo_src[base + i] = static_cast<int>(i);
  coo_dst[base + i] = static_cast<int>(i);
  int e = loop_eid[i];
  coo_w[base + i] = (e >= 0) ? (w ? w[e] : 1.0f) : fill;
}

__global__ void k_iota(int* __restrict__ x, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i] = static_cast<int>(i);
}

__global__ void k_gather_sorted(const int* __restrict__ perm, const int* __restrict__ other,
                                const float* __restrict__ w, int64_t nnz, int* __restrict__ colidx,
                                float* __restrict__ raw) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  int e = perm[p];
  colidx[p] = other[e];
  raw[p] = w[e];
}

// rowptr[i] = first position whose sorted key >= i
__global__ void k_rowptr(const int* __restrict__ keys, int64_t nnz, int64_t N, int* __restrict__ rowptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > N) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < i) lo = mid + 1; else hi = mid;
  }
  rowptr[i] = static_cast<int>(lo);
}

// sequential per-row sum in CSR (== COO) order: the order a sequential scatter_add uses
__global__ void k_degree(const int* __restrict__ rowptr, const float* __restrict__ raw, int64_t N,
                         float* __restrict__ dinv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float deg = 0.f;
  for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) deg += raw[p];
  // deg.pow(-0.5) with +inf -> 0 (prop_gcn_conv.py:79-80); NaN for negative degrees is kept
  float r = static_cast<float>(1.0 / sqrt(static_cast<double>(deg)));
  if (isinf(r)) r = 0.f;
  dinv[i] = r;
}

// w = dinv[src] * w * dinv[dst], evaluated left to right like the reference expression
__global__ void k_norm_csr(const int* __restrict__ rowkeys, const int* __restrict__ colidx,
                           const float* __restrict__ raw, const float* __restrict__ dinv, int64_t nnz,
                           int rows_are_src, float* __restrict__ vals) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  int r = rowkeys[p], c = colidx[p];
  int src = rows_are_src ? r : c, dst = rows_are_src ? c : r;
  vals[p] = (dinv[src] * raw[p]) * dinv[dst];
}

__global__ void k_norm_coo(const int* __restrict__ src, const int* __restrict__ dst, const float* __restrict__ dinv,
                           int64_t nnz, float* __restrict__ w) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  w[p] = (dinv[src[p]] * w[p]) * dinv[dst[p]];
}

// smallest row length of a CSR: rows are all non-empty iff it is >= 1
__global__ void k_min_degree(const int* __restrict__ rowptr, int64_t N, int* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int d = rowptr[i + 1] - rowptr[i];
  if (d == 0) atomicMin(out, 0);
}

__global__ void k_long_flags(const int* __restrict__ rowptr, int64_t N, int seg, int* __restrict__ is_long,
                             int* __restrict__ nsegs) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int deg = rowptr[i + 1] - rowptr[i];
  int l = deg > seg;
  is_long[i] = l;
  nsegs[i] = l ? (deg + seg - 1) / seg : 0;
}

__global__ void k_long_fill(const int* __restrict__ is_long, const int* __restrict__ long_pos,
                            const int* __restrict__ nsegs, const int* __restrict__ seg_pos, int64_t N,
                            int* __restrict__ long_rows, int* __restrict__ long_seg_ptr,
                            int* __restrict__ seg_long) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N || !is_long[i]) return;
  int L = long_pos[i];
  long_rows[L] = static_cast<int>(i);
  long_seg_ptr[L] = seg_pos[i];
  for (int j = 0; j < nsegs[i]; ++j) seg_long[seg_pos[i] + j] = L;
}

__global__ void k_count_key(const unsigned* __restrict__ keys, int64_t n, unsigned key, int* __restrict__ count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool hit = i < n && keys[i] == key;
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// Work items of k_spmm_tasks: item < num_segs is a segment of a long row, item - num_segs a row.
// Rows covered by segments (and empty rows) get length 0 and sort to the end of the list.
__global__ void k_task_make(const int* __restrict__ rowptr, int64_t N, int num_segs, int seg,
                            const int* __restrict__ long_rows, const int* __restrict__ long_seg_ptr,
                            const int* __restrict__ seg_long, unsigned* __restrict__ keys,
                            unsigned long long* __restrict__ descs) {
  int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (item >= num_segs + N) return;
  int p, len;
  unsigned y;
  if (item < num_segs) {
    const int L = seg_long[item];
    const int row = long_rows[L];
    p = rowptr[row] + (static_cast<int>(item) - long_seg_ptr[L]) * seg;
    len = min(seg, rowptr[row + 1] - p);
    y = 0x80000000u | static_cast<unsigned>(item);
  } else {
    const int row = static_cast<int>(item - num_segs);
    p = rowptr[row];
    len = rowptr[row + 1] - p;
    if (len > seg) len = 0;
    y = static_cast<unsigned>(row);
  }
  keys[item] = static_cast<unsigned>(seg - len);                      // ascending key = descending length
  y |= static_cast<unsigned>(len > 0 ? len - 1 : 0) << 25;
  descs[item] = (static_cast<unsigned long long>(y) << 32) | static_cast<unsigned>(p);
}

__global__ void k_export_coo(const int* __restrict__ src, const int* __restrict__ dst, int64_t nnz,
                             int64_t* __restrict__ out) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  out[p] = src[p];
  out[nnz + p] = dst[p];
}

// Graph storage comes from the device's stream-ordered memory pool (cudaMallocAsync) with the release threshold
// raised, so that the ~25 allocations of a graph build cost microseconds instead of ~0.2 ms each: graph-level
// estimators build two graphs per MINI-BATCH (a new edge_index every step) -- 4.9 ms per build, a fifth of the
// host-bound AdaGCN step, with plain cudaMalloc / cudaFree (profiles/r2_l_config5_host/).
thread_local cudaStream_t t_alloc_stream = nullptr;

struct AllocScope {          // the stream the build runs on: allocations and scratch frees are ordered on it
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t s) : prev(t_alloc_stream) {
    t_alloc_stream = s;
    static thread_local int pool_ready_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != pool_ready_dev) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~uint64_t(0);
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready_dev = dev;
    }
  }
  ~AllocScope() { t_alloc_stream = prev; }
};

template <typename T>
int dev_alloc(T** p, int64_t n) {
  *p = nullptr;
  // + 64 bytes of slack: the aggregation kernels prefetch the next batch's column indices with unconditional loads
  // that may run up to 3 entries past the last non-zero (spmm.cu: k_spmm_unw)
  GDA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(p), sizeof(T) * static_cast<size_t>(n > 0 ? n : 1) + 64, t_alloc_stream));
  return GDA_OK;
}

inline void dev_free(void* p) {        // graph members: the owner has synchronised the device (see ~gda_graph)
  if (p) cudaFreeAsync(p, nullptr);
}

struct Scratch {   // frees temporaries on every exit path (ordered after the build's kernels on its stream)
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) if (p) cudaFreeAsync(p, t_alloc_stream); }
  template <typename T> int get(T** p, int64_t n) {
    int rc = dev_alloc(p, n);
    if (rc == GDA_OK) ptrs.push_back(*p);
    return rc;
  }
};

int exclusive_scan(const int* in, int* out, int64_t n, Scratch& sc, cudaStream_t st) {
  size_t bytes = 0;
  GDA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, static_cast<int>(n), st));
  void* tmp; int rc = sc.get(reinterpret_cast<char**>(&tmp), static_cast<int64_t>(bytes));
  if (rc) return rc;
  GDA_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, static_cast<int>(n), st));
  return GDA_OK;
}

// stable sort of (key = node id, value = COO position)
int sort_by_key(const int* keys_in, int* keys_out, int* perm_out, int64_t nnz, int64_t N, Scratch& sc,
                cudaStream_t st) {
  int* iota; int rc = sc.get(&iota, nnz);
  if (rc) return rc;
  k_iota<<<blocks_for(nnz), kThreads, 0, st>>>(iota, nnz);
  GDA_LAUNCH_CHECK();
  int end_bit = 1;
  while ((int64_t(1) << end_bit) < N && end_bit < 31) ++end_bit;
  size_t bytes = 0;
  GDA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, iota, perm_out,
                                           static_cast<int>(nnz), 0, end_bit, st));
  void* tmp; rc = sc.get(reinterpret_cast<char**>(&tmp), static_cast<int64_t>(bytes));
  if (rc) return rc;
  GDA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, iota, perm_out,
                                           static_cast<int>(nnz), 0, end_bit, st));
  return GDA_OK;
}

// Length-sorted work list (graph.cuh: Csr::tasks).  Only for graphs whose rows are all non-empty
// (a self loop in every row): then exactly the num_long covered rows have length 0 and are cut off.
int build_tasks(Csr& c, int64_t N, int seg, Scratch& sc, cudaStream_t st) {
  c.tasks = nullptr;
  c.num_tasks = 0;
  const int64_t total = static_cast<int64_t>(c.num_segs) + N;
  if ((c.may_have_empty_rows && !c.skip_empty_rows) || total == 0 || N >= (int64_t(1) << 25) ||
      c.num_segs >= (1 << 25) || seg > 64)
    return GDA_OK;
  unsigned *keys, *keys_sorted;
  unsigned long long *descs, *sorted;
  int rc;
  if ((rc = sc.get(&keys, total)) || (rc = sc.get(&keys_sorted, total)) || (rc = sc.get(&descs, total))) return rc;
  if ((rc = dev_alloc(&sorted, total))) return rc;
  c.tasks = reinterpret_cast<uint2*>(sorted);
  k_task_make<<<blocks_for(total), kThreads, 0, st>>>(c.rowptr, N, c.num_segs, seg, c.long_rows, c.long_seg_ptr,
                                                      c.seg_long, keys, descs);
  GDA_LAUNCH_CHECK();
  size_t bytes = 0;
  GDA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, keys_sorted, descs, sorted,
                                           static_cast<int>(total), 0, 8, st));
  void* tmp;
  if ((rc = sc.get(reinterpret_cast<char**>(&tmp), static_cast<int64_t>(bytes)))) return rc;
  GDA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, keys_sorted, descs, sorted,
                                           static_cast<int>(total), 0, 8, st));
  int* zero_len = nullptr;                               // items of length 0: split long rows, and empty rows
  if ((rc = sc.get(&zero_len, 1))) return rc;
  GDA_CUDA(cudaMemsetAsync(zero_len, 0, sizeof(int), st));
  k_count_key<<<blocks_for(total), kThreads, 0, st>>>(keys, total, static_cast<unsigned>(seg), zero_len);
  GDA_LAUNCH_CHECK();
  int h_zero = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_zero, zero_len, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  c.num_tasks = static_cast<int32_t>(total - h_zero);    // == total - num_long when no row is empty
  if (c.skip_empty_rows) c.may_have_empty_rows = false;  // empty rows are simply not in the list: lean kernels apply
  return GDA_OK;
}

int build_long_rows(Csr& c, int64_t N, int seg, Scratch& sc, cudaStream_t st) {
  int *is_long, *nsegs, *long_pos, *seg_pos;
  int rc;
  if ((rc = sc.get(&is_long, N + 1)) || (rc = sc.get(&nsegs, N + 1)) ||
      (rc = sc.get(&long_pos, N + 1)) || (rc = sc.get(&seg_pos, N + 1))) return rc;
  GDA_CUDA(cudaMemsetAsync(is_long, 0, sizeof(int) * (N + 1), st));
  GDA_CUDA(cudaMemsetAsync(nsegs, 0, sizeof(int) * (N + 1), st));
  k_long_flags<<<blocks_for(N), kThreads, 0, st>>>(c.rowptr, N, seg, is_long, nsegs);
  GDA_LAUNCH_CHECK();
  if ((rc = exclusive_scan(is_long, long_pos, N + 1, sc, st))) return rc;
  if ((rc = exclusive_scan(nsegs, seg_pos, N + 1, sc, st))) return rc;
  int totals[2];
  GDA_CUDA(cudaMemcpyAsync(&totals[0], long_pos + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaMemcpyAsync(&totals[1], seg_pos + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  c.num_long = totals[0];
  c.num_segs = totals[1];
  if ((rc = dev_alloc(&c.long_rows, c.num_long)) || (rc = dev_alloc(&c.long_seg_ptr, c.num_long + 1)) ||
      (rc = dev_alloc(&c.seg_long, c.num_segs)) || (rc = dev_alloc(&c.counters, c.num_long))) return rc;
  GDA_CUDA(cudaMemsetAsync(c.counters, 0, sizeof(int) * (c.num_long > 0 ? c.num_long : 1), st));
  if (c.num_long > 0) {
    k_long_fill<<<blocks_for(N), kThreads, 0, st>>>(is_long, long_pos, nsegs, seg_pos, N, c.long_rows,
                                                    c.long_seg_ptr, c.seg_long);
    GDA_LAUNCH_CHECK();
  }
  GDA_CUDA(cudaMemcpyAsync(c.long_seg_ptr + c.num_long, &c.num_segs, sizeof(int),
                           cudaMemcpyHostToDevice, st));
  GDA_CUDA(cudaStreamSynchronize(st));   // &c.num_segs must outlive the copy
  return build_tasks(c, N, seg, sc, st);
}

void free_csr(Csr& c) {
  dev_free(c.rowptr); dev_free(c.colidx); dev_free(c.vals);
  dev_free(c.long_rows); dev_free(c.long_seg_ptr); dev_free(c.seg_long); dev_free(c.counters);
  dev_free(c.tasks);
  c = Csr{};
}

void free_graph_members(gda_graph* g) {
  dev_free(g->coo_src); dev_free(g->coo_dst); dev_free(g->coo_w); dev_free(g->dinv);
  free_csr(g->csr); free_csr(g->csr_t);
}

}  // namespace

int graph_create(const int64_t* ei, int64_t E, int64_t N, const float* w, int flags, cudaStream_t st,
                 gda_graph** out) {
  GDA_REQUIRE(out != nullptr, "gda_graph_create: out is NULL");
  *out = nullptr;
  AllocScope alloc_scope(st);
  GDA_REQUIRE(N >= 0 && E >= 0, "gda_graph_create: negative size");
  GDA_REQUIRE(E == 0 || ei != nullptr, "gda_graph_create: edge_index is NULL");
  GDA_REQUIRE(E + N < (int64_t(1) << 31) - 1, "gda_graph_create: E + N must fit in int32");
  GDA_REQUIRE(!((flags & GDA_NORM_SYM_COL) && (flags & GDA_NORM_SYM_ROW)),
              "gda_graph_create: choose one normalisation");

  std::unique_ptr<gda_graph> g(new gda_graph());
  g->N = N; g->E = E; g->flags = flags; g->seg = kLongRowSegment;
  if (const char* e = std::getenv("GDA_SEG")) {           // experiments: segment length of split long rows (8..64)
    const int v = std::atoi(e);
    if (v >= 8 && v <= 64 && v % 4 == 0) g->seg = v;
  }
  GDA_CUDA(cudaGetDevice(&g->device));
  const bool loops = flags & GDA_SELF_LOOPS;
  Scratch sc;
  int rc;

  // 1. validate + flag surviving edges, stable positions
  int *flag, *pos, *err, *loop_eid;
  if ((rc = sc.get(&flag, E + 1)) || (rc = sc.get(&pos, E + 1)) || (rc = sc.get(&err, 1)) ||
      (rc = sc.get(&loop_eid, N))) return rc;
  GDA_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  GDA_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (E + 1), st));
  if (E > 0) {
    k_flag_edges<<<blocks_for(E), kThreads, 0, st>>>(ei, E, N, loops ? 1 : 0, flag, err);
    GDA_LAUNCH_CHECK();
  }
  if ((rc = exclusive_scan(flag, pos, E + 1, sc, st))) return rc;
  int h_err = 0, kept = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaMemcpyAsync(&kept, pos + E, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  if (h_err) return fail(GDA_E_INDEX, "gda_graph_create: edge_index entry outside [0, N)");
  const int64_t nnz = kept + (loops ? N : 0);
  g->nnz = nnz;

  // 2. COO in the reference's order: kept edges, then loops 0..N-1
  if ((rc = dev_alloc(&g->coo_src, nnz)) || (rc = dev_alloc(&g->coo_dst, nnz)) ||
      (rc = dev_alloc(&g->coo_w, nnz)) || (rc = dev_alloc(&g->dinv, N))) return rc;
  if (E > 0) {
    k_fill_coo_edges<<<blocks_for(E), kThreads, 0, st>>>(ei, w, E, flag, pos, g->coo_src, g->coo_dst, g->coo_w);
    GDA_LAUNCH_CHECK();
  }
  if (loops && N > 0) {
    GDA_CUDA(cudaMemsetAsync(loop_eid, 0xff, sizeof(int) * N, st));   // -1
    if (E > 0) {
      k_loop_eid<<<blocks_for(E), kThreads, 0, st>>>(ei, E, loop_eid);
      GDA_LAUNCH_CHECK();
    }
    const float fill = (flags & GDA_IMPROVED) ? 2.0f : 1.0f;
    k_fill_coo_loops<<<blocks_for(N), kThreads, 0, st>>>(w, loop_eid, N, kept, fill, g->coo_src, g->coo_dst, g->coo_w);
    GDA_LAUNCH_CHECK();
  }

  // 3. CSR by target (forward) and by source (transpose), stable
  int *keys_d, *perm_d, *keys_s, *perm_s;
  float *raw_d, *raw_s;
  if ((rc = sc.get(&keys_d, nnz)) || (rc = sc.get(&perm_d, nnz)) || (rc = sc.get(&keys_s, nnz)) ||
      (rc = sc.get(&perm_s, nnz)) || (rc = sc.get(&raw_d, nnz)) || (rc = sc.get(&raw_s, nnz))) return rc;
  for (int t = 0; t < 2; ++t) {
    Csr& c = t ? g->csr_t : g->csr;
    const int* key_in = t ? g->coo_src : g->coo_dst;
    const int* other = t ? g->coo_dst : g->coo_src;
    int* keys = t ? keys_s : keys_d;
    int* perm = t ? perm_s : perm_d;
    float* raw = t ? raw_s : raw_d;
    if ((rc = dev_alloc(&c.rowptr, N + 1)) || (rc = dev_alloc(&c.colidx, nnz)) ||
        (rc = dev_alloc(&c.vals, nnz))) return rc;
    if (nnz > 0) {
      if ((rc = sort_by_key(key_in, keys, perm, nnz, N, sc, st))) return rc;
      k_gather_sorted<<<blocks_for(nnz), kThreads, 0, st>>>(perm, other, g->coo_w, nnz, c.colidx, raw);
      GDA_LAUNCH_CHECK();
    }
    k_rowptr<<<blocks_for(N + 1), kThreads, 0, st>>>(keys, nnz, N, c.rowptr);
    GDA_LAUNCH_CHECK();
  }

  // 4. symmetric normalisation
  const bool norm = flags & (GDA_NORM_SYM_COL | GDA_NORM_SYM_ROW);
  if (norm && N > 0) {
    const bool by_row = flags & GDA_NORM_SYM_ROW;
    k_degree<<<blocks_for(N), kThreads, 0, st>>>(by_row ? g->csr_t.rowptr : g->csr.rowptr,
                                                 by_row ? raw_s : raw_d, N, g->dinv);
    GDA_LAUNCH_CHECK();
    if (nnz > 0) {
      k_norm_csr<<<blocks_for(nnz), kThreads, 0, st>>>(keys_d, g->csr.colidx, raw_d, g->dinv, nnz, 0, g->csr.vals);
      GDA_LAUNCH_CHECK();
      k_norm_csr<<<blocks_for(nnz), kThreads, 0, st>>>(keys_s, g->csr_t.colidx, raw_s, g->dinv, nnz, 1, g->csr_t.vals);
      GDA_LAUNCH_CHECK();
      k_norm_coo<<<blocks_for(nnz), kThreads, 0, st>>>(g->coo_src, g->coo_dst, g->dinv, nnz, g->coo_w);
      GDA_LAUNCH_CHECK();
    }
  } else if (nnz > 0) {
    GDA_CUDA(cudaMemcpyAsync(g->csr.vals, raw_d, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st));
    GDA_CUDA(cudaMemcpyAsync(g->csr_t.vals, raw_s, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st));
  }

  g->unit_weights = norm && w == nullptr && !(flags & GDA_IMPROVED);
  g->csr.may_have_empty_rows = g->csr_t.may_have_empty_rows = !loops;   // a self loop in every row
  g->csr.skip_empty_rows = g->csr_t.skip_empty_rows = (flags & GDA_SKIP_EMPTY_ROWS) != 0;
  if (!loops && N > 0) {
    // graphs given with their loops already in place (TDSS's smoothing graph, tdss.py:376-385) have no empty
    // row either: look, so that they take the lean aggregation kernels as well
    int* mins; if ((rc = sc.get(&mins, 2))) return rc;
    int ones[2] = {1, 1};
    GDA_CUDA(cudaMemcpyAsync(mins, ones, sizeof(ones), cudaMemcpyHostToDevice, st));
    k_min_degree<<<blocks_for(N), kThreads, 0, st>>>(g->csr.rowptr, N, mins);
    GDA_LAUNCH_CHECK();
    k_min_degree<<<blocks_for(N), kThreads, 0, st>>>(g->csr_t.rowptr, N, mins + 1);
    GDA_LAUNCH_CHECK();
    int h_min[2] = {0, 0};
    GDA_CUDA(cudaMemcpyAsync(h_min, mins, sizeof(h_min), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    g->csr.may_have_empty_rows = h_min[0] == 0;
    g->csr_t.may_have_empty_rows = h_min[1] == 0;
  }
  // 5. long-row segments for both orientations
  if ((rc = build_long_rows(g->csr, N, g->seg, sc, st))) return rc;
  if ((rc = build_long_rows(g->csr_t, N, g->seg, sc, st))) return rc;
  GDA_CUDA(cudaStreamSynchronize(st));
  *out = g.release();
  return GDA_OK;
}

namespace {
__global__ void k_part_rowptr(const int* __restrict__ rowptr, int64_t lo, int64_t n, int base, int* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= n) out[i] = rowptr[lo + i] - base;
}
__global__ void k_part_cols(const int* __restrict__ colidx, const float* __restrict__ vals, int64_t base, int64_t nnz,
                            int rows_per_rank, int* __restrict__ col_out, float* __restrict__ val_out) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int c = colidx[base + e];
  const int owner = c / rows_per_rank;
  col_out[e] = (owner << 28) | (c - owner * rows_per_rank);
  val_out[e] = vals[base + e];
}
}  // namespace

// Row block [row_lo, row_hi) of a normalised graph, columns re-encoded as (owner rank, row inside
// the owner's block) for the peer-memory aggregation kernel.
int graph_partition(const gda_graph* g, int64_t row_lo, int64_t row_hi, int64_t rows_per_rank, cudaStream_t st,
                    gda_graph** out) {
  GDA_REQUIRE(g && out, "gda_graph_partition: NULL argument");
  *out = nullptr;
  AllocScope alloc_scope(st);
  GDA_REQUIRE(!g->peer_packed, "gda_graph_partition: already a partition");
  GDA_REQUIRE(!g->csr.may_have_empty_rows, "gda_graph_partition: needs a graph built with self loops");
  GDA_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= g->N, "gda_graph_partition: bad row range");
  GDA_REQUIRE(rows_per_rank > 0 && rows_per_rank < (int64_t(1) << 28) &&
              ceil_div(g->N, rows_per_rank) <= GDA_MAX_PEERS, "gda_graph_partition: bad rows_per_rank");
  std::unique_ptr<gda_graph> p(new gda_graph());
  const int64_t n = row_hi - row_lo;
  p->N = n; p->E = 0; p->flags = g->flags; p->seg = g->seg; p->device = g->device;
  p->peer_packed = true; p->rows_per_rank = rows_per_rank; p->global_N = g->N; p->row_lo = row_lo;
  Scratch sc;
  int rc;
  for (int t = 0; t < 2; ++t) {
    const Csr& s = t ? g->csr_t : g->csr;
    Csr& d = t ? p->csr_t : p->csr;
    int bounds[2] = {0, 0};
    GDA_CUDA(cudaMemcpyAsync(&bounds[0], s.rowptr + row_lo, sizeof(int), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaMemcpyAsync(&bounds[1], s.rowptr + row_hi, sizeof(int), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz = bounds[1] - bounds[0];
    if (t == 0) p->nnz = nnz;
    if ((rc = dev_alloc(&d.rowptr, n + 1)) || (rc = dev_alloc(&d.colidx, nnz)) || (rc = dev_alloc(&d.vals, nnz)))
      return rc;
    k_part_rowptr<<<blocks_for(n + 1), kThreads, 0, st>>>(s.rowptr, row_lo, n, bounds[0], d.rowptr);
    GDA_LAUNCH_CHECK();
    if (nnz > 0) {
      k_part_cols<<<blocks_for(nnz), kThreads, 0, st>>>(s.colidx, s.vals, bounds[0], nnz, static_cast<int>(rows_per_rank),
                                                        d.colidx, d.vals);
      GDA_LAUNCH_CHECK();
    }
    d.may_have_empty_rows = false;
    if ((rc = build_long_rows(d, n, p->seg, sc, st))) return rc;
  }
  GDA_CUDA(cudaStreamSynchronize(st));
  *out = p.release();
  return GDA_OK;
}

int graph_export_coo(const gda_graph* g, int64_t* ei_out, float* w_out, cudaStream_t st) {
  GDA_REQUIRE(g && (g->nnz == 0 || (ei_out && w_out)), "gda_graph_export_coo: NULL argument");
  GDA_REQUIRE(!g->peer_packed, "gda_graph_export_coo: not available for a partition");
  if (g->nnz == 0) return GDA_OK;
  k_export_coo<<<blocks_for(g->nnz), kThreads, 0, st>>>(g->coo_src, g->coo_dst, g->nnz, ei_out);
  GDA_LAUNCH_CHECK();
  GDA_CUDA(cudaMemcpyAsync(w_out, g->coo_w, sizeof(float) * g->nnz, cudaMemcpyDeviceToDevice, st));
  return GDA_OK;
}

int graph_export_csr(const gda_graph* g, int transpose, int32_t* rowptr, int32_t* colidx, float* vals,
                     cudaStream_t st) {
  GDA_REQUIRE(g && rowptr, "gda_graph_export_csr: NULL argument");
  const Csr& c = transpose ? g->csr_t : g->csr;
  GDA_CUDA(cudaMemcpyAsync(rowptr, c.rowptr, sizeof(int) * (g->N + 1), cudaMemcpyDeviceToDevice, st));
  if (g->nnz > 0) {
    GDA_REQUIRE(colidx && vals, "gda_graph_export_csr: NULL argument");
    GDA_CUDA(cudaMemcpyAsync(colidx, c.colidx, sizeof(int) * g->nnz, cudaMemcpyDeviceToDevice, st));
    GDA_CUDA(cudaMemcpyAsync(vals, c.vals, sizeof(float) * g->nnz, cudaMemcpyDeviceToDevice, st));
  }
  return GDA_OK;
}

}  // namespace gda

gda_graph::~gda_graph() {
  // kernels on ANY stream may still read the graph (cudaFree used to wait for them implicitly, once per pointer)
  cudaDeviceSynchronize();
  gda::free_graph_members(this);
}
