// PPMI graph construction on the GPU.
//
// Replaces PPMIConv.norm (pygda/nn/ppmi_conv.py:56-184), which the reference runs in pure Python once per
// (layer, cache_name): a dict-of-sets adjacency (:98-117), 40 rounds of random walks of random length in
// [1, path_len] from every node that has an edge (:134-148), per-start visit counters normalised to
// probabilities (:150), column sums (:152-156), ppmi = max(log(p / colsum * #visited / path_len), 0) (:158-163)
// and an edge (start, visited) for EVERY visited pair, zero scores included (:165-172).  The symmetric
// normalisation with remaining self loops (:174-184) is gda_graph_create's job (GDA_GRAPH_NORM_SYM_ROW).
//
// Here: undirected de-duplicated CSR by radix sort + unique; one thread per (round, start) walker emitting
// the key start * N + visited for every step; radix sort + run-length encoding = the visit counters; integer
// atomics for the row sums; fp64 for the probabilities, column sums and scores (the reference computes them in
// Python floats).  np.random's stream cannot be reproduced: walks use the library's counter hash (common.cuh),
// so results match the reference in distribution; given the same visit counts they match to fp64 round-off
// (tests/test_gpu_ppmi.py checks both).
#include <cub/cub.cuh>

#include <memory>
#include <vector>

#include "common.cuh"

struct gda_wedges {
  int64_t M = 0, N = 0;
  int64_t* ei = nullptr;      // [2, M] row-major, device
  float* w = nullptr;         // [M]
  int* cnt = nullptr;         // [M] visit counts
  ~gda_wedges() { cudaFree(ei); cudaFree(w); cudaFree(cnt); }
};

namespace gda {
namespace {

constexpr int kT = 256;
inline unsigned nblk(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kT)); }
constexpr unsigned long long kDropKey = ~0ull;
constexpr int64_t kMaxBatchKeys = int64_t(1) << 28;

struct Tmp {
  std::vector<void*> ptrs;
  ~Tmp() { for (void* p : ptrs) cudaFree(p); }
  template <typename T> int get(T** p, int64_t n) {
    *p = nullptr;
    GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * static_cast<size_t>(n > 0 ? n : 1)));
    ptrs.push_back(*p);
    return GDA_OK;
  }
  void release(void* p) {
    for (auto& q : ptrs) if (q == p) { cudaFree(q); q = nullptr; }
  }
};

// both directions of every edge: add_edge(a, b); add_edge(b, a)  (ppmi_conv.py:111-115)
__global__ void k_und_keys(const int64_t* __restrict__ ei, int64_t E, int64_t N, unsigned long long* __restrict__ keys,
                           int* __restrict__ err) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) { *err = 1; keys[e] = kDropKey; keys[E + e] = kDropKey; return; }
  keys[e] = static_cast<unsigned long long>(s) * N + d;
  keys[E + e] = static_cast<unsigned long long>(d) * N + s;
}

__global__ void k_split_keys(const unsigned long long* __restrict__ keys, int64_t n, int64_t N, int* __restrict__ a,
                             int* __restrict__ b) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[i] = static_cast<int>(keys[i] / N);
  b[i] = static_cast<int>(keys[i] % N);
}

__global__ void k_lower_bound(const int* __restrict__ keys, int64_t n, int64_t N, int* __restrict__ ptr) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > N) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] < i) lo = mid + 1; else hi = mid; }
  ptr[i] = static_cast<int>(lo);
}

// Walker w = round * N + start.  Length 1 + U{0..L-1} (np.random.randint(1, path_len + 1), :137); every step moves
// to a uniform neighbour of the current node (:119-122,139,148).  visit(t, node) is called for each step.
template <typename F>
__device__ __forceinline__ void walk(const int* __restrict__ adjptr, const int* __restrict__ adjcol, int start, int64_t w,
                                     int L, uint64_t seed, F visit) {
  const uint64_t base = static_cast<uint64_t>(w) * (L + 1);
  const int len = 1 + static_cast<int>((static_cast<uint64_t>(mix_hash(seed, base)) * L) >> 32);
  int cur = start;
  for (int t = 0; t < L; ++t) {
    if (t < len) {
      const int p = adjptr[cur], d = adjptr[cur + 1] - p;       // d >= 1: every node reached has the edge it came by
      cur = adjcol[p + static_cast<int>((static_cast<uint64_t>(mix_hash(seed, base + 1 + t)) * d) >> 32)];
      visit(t, cur);
    } else {
      visit(t, -1);
    }
  }
}

__global__ void k_walk_keys(const int* __restrict__ adjptr, const int* __restrict__ adjcol, int64_t N, int L, int round0,
                            int rounds, uint64_t seed, unsigned long long* __restrict__ keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(rounds) * N) return;
  const int start = static_cast<int>(i % N);
  unsigned long long* out = keys + i * L;
  if (adjptr[start + 1] == adjptr[start]) {                     // `for a in adj_dict`: nodes without an edge do not walk
    for (int t = 0; t < L; ++t) out[t] = kDropKey;
    return;
  }
  const int64_t w = (static_cast<int64_t>(round0) + i / N) * N + start;
  walk(adjptr, adjcol, start, w, L, seed,
       [&](int t, int node) { out[t] = node < 0 ? kDropKey : static_cast<unsigned long long>(start) * N + node; });
}

__global__ void k_walk_nodes(const int* __restrict__ adjptr, const int* __restrict__ adjcol, int64_t N, int L, int rounds,
                             uint64_t seed, int* __restrict__ walks) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(rounds) * N) return;
  const int start = static_cast<int>(i % N);
  int* out = walks + i * L;
  if (adjptr[start + 1] == adjptr[start]) {
    for (int t = 0; t < L; ++t) out[t] = -1;
    return;
  }
  walk(adjptr, adjcol, start, i, L, seed, [&](int t, int node) { out[t] = node; });
}

__global__ void k_row_sums(const unsigned long long* __restrict__ keys, const int* __restrict__ cnt, int64_t M, int64_t N,
                           int* __restrict__ rowsum) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < M) atomicAdd(rowsum + keys[i] / N, cnt[i]);
}

// p = count / rowsum (norm(), :126-131); prob_sums[b] += p (:152-156)
__global__ void k_col_sums(const unsigned long long* __restrict__ keys, const int* __restrict__ cnt, int64_t M, int64_t N,
                           const int* __restrict__ rowsum, double* __restrict__ colsum) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long k = keys[i];
  atomicAdd(colsum + k % N, static_cast<double>(cnt[i]) / static_cast<double>(rowsum[k / N]));
}

__global__ void k_count_visited(const double* __restrict__ colsum, int64_t N, int* __restrict__ nvisited) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int v = (i < N && colsum[i] > 0.0) ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nvisited, __popc(m));
}

// ppmi = max(log(p / prob_sums[b] * len(prob_sums) / path_len), 0)  (:160-163); edge (a, b) (:165-170)
__global__ void k_scores(const unsigned long long* __restrict__ keys, const int* __restrict__ cnt, int64_t M, int64_t N,
                         const int* __restrict__ rowsum, const double* __restrict__ colsum, const int* __restrict__ nvisited,
                         int L, int64_t* __restrict__ ei, float* __restrict__ w) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long k = keys[i];
  const int64_t a = static_cast<int64_t>(k / N), b = static_cast<int64_t>(k % N);
  const double p = static_cast<double>(cnt[i]) / static_cast<double>(rowsum[a]);
  const double s = log(p / colsum[b] * static_cast<double>(*nvisited) / static_cast<double>(L));
  ei[i] = a;
  ei[M + i] = b;
  w[i] = static_cast<float>(s > 0.0 ? s : 0.0);
}

int sort_keys(unsigned long long* in, unsigned long long* out, int64_t n, Tmp& t, cudaStream_t st) {
  size_t bytes = 0;
  GDA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, static_cast<int>(n), 0, 64, st));
  char* tmp;
  int rc;
  if ((rc = t.get(&tmp, static_cast<int64_t>(bytes)))) return rc;
  GDA_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, in, out, static_cast<int>(n), 0, 64, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  t.release(tmp);
  return GDA_OK;
}

// sorted keys -> (unique keys, run lengths); a trailing kDropKey run is dropped
int run_lengths(const unsigned long long* sorted, int64_t n, unsigned long long* uniq, int* counts, int64_t* n_out, Tmp& t,
                cudaStream_t st) {
  int* d_runs;
  int rc;
  if ((rc = t.get(&d_runs, 1))) return rc;
  size_t bytes = 0;
  GDA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, sorted, uniq, counts, d_runs, static_cast<int>(n), st));
  char* tmp;
  if ((rc = t.get(&tmp, static_cast<int64_t>(bytes)))) return rc;
  GDA_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, bytes, sorted, uniq, counts, d_runs, static_cast<int>(n), st));
  int h_runs = 0;
  GDA_CUDA(cudaMemcpyAsync(&h_runs, d_runs, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDA_CUDA(cudaStreamSynchronize(st));
  if (h_runs > 0) {
    unsigned long long last = 0;
    GDA_CUDA(cudaMemcpyAsync(&last, uniq + h_runs - 1, sizeof(last), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    if (last == kDropKey) --h_runs;
  }
  t.release(tmp);
  *n_out = h_runs;
  return GDA_OK;
}

// undirected, de-duplicated adjacency in CSR form (neighbours ascending)
int build_adjacency(const int64_t* edge_index, int64_t E, int64_t N, int** adjptr, int** adjcol, int64_t* nnz, Tmp& t,
                    cudaStream_t st, const char* who) {
  int rc;
  unsigned long long *keys, *sorted, *uniq;
  int *err, *cnts;
  if ((rc = t.get(&keys, 2 * E)) || (rc = t.get(&sorted, 2 * E)) || (rc = t.get(&uniq, 2 * E)) ||
      (rc = t.get(&cnts, 2 * E)) || (rc = t.get(&err, 1)) || (rc = t.get(adjptr, N + 1)))
    return rc;
  GDA_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  int64_t nu = 0;
  if (E > 0) {
    k_und_keys<<<nblk(E), kT, 0, st>>>(edge_index, E, N, keys, err);
    GDA_LAUNCH_CHECK();
    int h_err = 0;
    GDA_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    if (h_err) return fail(GDA_E_INDEX, std::string(who) + ": edge_index entry outside [0, N)");
    if ((rc = sort_keys(keys, sorted, 2 * E, t, st))) return rc;
    if ((rc = run_lengths(sorted, 2 * E, uniq, cnts, &nu, t, st))) return rc;
  }
  int* a;
  if ((rc = t.get(&a, nu)) || (rc = t.get(adjcol, nu))) return rc;
  if (nu > 0) { k_split_keys<<<nblk(nu), kT, 0, st>>>(uniq, nu, N, a, *adjcol); GDA_LAUNCH_CHECK(); }
  k_lower_bound<<<nblk(N + 1), kT, 0, st>>>(a, nu, N, *adjptr);
  GDA_LAUNCH_CHECK();
  GDA_CUDA(cudaStreamSynchronize(st));
  t.release(keys); t.release(sorted); t.release(uniq); t.release(cnts);
  *nnz = nu;
  return GDA_OK;
}

}  // namespace
}  // namespace gda

extern "C" {

int gda_ppmi_create(const int64_t* edge_index, int64_t E, int64_t N, int path_len, int rounds, uint64_t seed,
                    gda_stream_t stream, gda_wedges_t** out) {
  using namespace gda;
  GDA_REQUIRE(out != nullptr, "gda_ppmi_create: out is NULL");
  *out = nullptr;
  GDA_REQUIRE(N >= 0 && E >= 0 && path_len >= 1 && rounds >= 1, "gda_ppmi_create: bad size / path_len / rounds");
  GDA_REQUIRE(E == 0 || edge_index != nullptr, "gda_ppmi_create: edge_index is NULL");
  GDA_REQUIRE(N < (int64_t(1) << 31) && 2 * E < (int64_t(1) << 31) - 1, "gda_ppmi_create: N and 2E must fit in int32");
  GDA_REQUIRE(N * path_len <= kMaxBatchKeys, "gda_ppmi_create: N * path_len above 2^28 (one round per batch at most)");
  cudaStream_t st = as_stream(stream);
  Tmp t;
  int rc;
  int *adjptr, *adjcol;
  int64_t adj_nnz = 0;
  if ((rc = build_adjacency(edge_index, E, N, &adjptr, &adjcol, &adj_nnz, t, st, "gda_ppmi_create"))) return rc;

  // visit counters: batches of rounds through sort + run-length encoding, merged by key
  unsigned long long* acc_keys = nullptr;
  int* acc_cnt = nullptr;
  int64_t acc_n = 0;
  const int64_t per_round = N * path_len;
  int rounds_per_batch = per_round > 0 ? static_cast<int>(kMaxBatchKeys / per_round) : rounds;
  if (rounds_per_batch < 1) rounds_per_batch = 1;
  if (rounds_per_batch > rounds) rounds_per_batch = rounds;
  for (int r0 = 0; r0 < rounds && per_round > 0 && adj_nnz > 0; r0 += rounds_per_batch) {
    const int rb = rounds - r0 < rounds_per_batch ? rounds - r0 : rounds_per_batch;
    const int64_t nk = static_cast<int64_t>(rb) * per_round;
    unsigned long long *keys, *sorted, *uniq;
    int* cnts;
    if ((rc = t.get(&keys, nk)) || (rc = t.get(&sorted, nk))) return rc;
    k_walk_keys<<<nblk(static_cast<int64_t>(rb) * N), kT, 0, st>>>(adjptr, adjcol, N, path_len, r0, rb, seed, keys);
    GDA_LAUNCH_CHECK();
    if ((rc = sort_keys(keys, sorted, nk, t, st))) return rc;
    t.release(keys);
    if ((rc = t.get(&uniq, nk)) || (rc = t.get(&cnts, nk))) return rc;
    int64_t nu = 0;
    if ((rc = run_lengths(sorted, nk, uniq, cnts, &nu, t, st))) return rc;
    t.release(sorted);
    if (acc_keys == nullptr) {
      acc_keys = uniq; acc_cnt = cnts; acc_n = nu;
      continue;
    }
    // merge: concat -> sort pairs -> reduce by key
    const int64_t nm = acc_n + nu;
    GDA_REQUIRE(nm < (int64_t(1) << 31) - 1, "gda_ppmi_create: more than 2^31 distinct (start, visited) pairs");
    unsigned long long *ck, *sk, *rk;
    int *cc, *sc, *rcnt, *d_runs;
    if ((rc = t.get(&ck, nm)) || (rc = t.get(&cc, nm)) || (rc = t.get(&sk, nm)) || (rc = t.get(&sc, nm)) ||
        (rc = t.get(&rk, nm)) || (rc = t.get(&rcnt, nm)) || (rc = t.get(&d_runs, 1)))
      return rc;
    GDA_CUDA(cudaMemcpyAsync(ck, acc_keys, sizeof(*ck) * acc_n, cudaMemcpyDeviceToDevice, st));
    GDA_CUDA(cudaMemcpyAsync(ck + acc_n, uniq, sizeof(*ck) * nu, cudaMemcpyDeviceToDevice, st));
    GDA_CUDA(cudaMemcpyAsync(cc, acc_cnt, sizeof(int) * acc_n, cudaMemcpyDeviceToDevice, st));
    GDA_CUDA(cudaMemcpyAsync(cc + acc_n, cnts, sizeof(int) * nu, cudaMemcpyDeviceToDevice, st));
    size_t b1 = 0, b2 = 0;
    GDA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, ck, sk, cc, sc, static_cast<int>(nm), 0, 64, st));
    GDA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, b2, sk, rk, sc, rcnt, d_runs, cub::Sum(), static_cast<int>(nm), st));
    char* tmp;
    if ((rc = t.get(&tmp, static_cast<int64_t>(b1 > b2 ? b1 : b2)))) return rc;
    GDA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, b1, ck, sk, cc, sc, static_cast<int>(nm), 0, 64, st));
    GDA_CUDA(cub::DeviceReduce::ReduceByKey(tmp, b2, sk, rk, sc, rcnt, d_runs, cub::Sum(), static_cast<int>(nm), st));
    int h_runs = 0;
    GDA_CUDA(cudaMemcpyAsync(&h_runs, d_runs, sizeof(int), cudaMemcpyDeviceToHost, st));
    GDA_CUDA(cudaStreamSynchronize(st));
    t.release(tmp); t.release(ck); t.release(cc); t.release(sk); t.release(sc);
    t.release(acc_keys); t.release(acc_cnt); t.release(uniq); t.release(cnts);
    acc_keys = rk; acc_cnt = rcnt; acc_n = h_runs;
  }

  std::unique_ptr<gda_wedges> res(new gda_wedges());
  res->N = N;
  res->M = acc_n;
  const size_t m1 = static_cast<size_t>(acc_n > 0 ? acc_n : 1);
  GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(&res->ei), sizeof(int64_t) * 2 * m1));
  GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(&res->w), sizeof(float) * m1));
  GDA_CUDA(cudaMalloc(reinterpret_cast<void**>(&res->cnt), sizeof(int) * m1));
  if (acc_n > 0) {
    int *rowsum, *nvisited;
    double* colsum;
    if ((rc = t.get(&rowsum, N)) || (rc = t.get(&colsum, N)) || (rc = t.get(&nvisited, 1))) return rc;
    GDA_CUDA(cudaMemsetAsync(rowsum, 0, sizeof(int) * N, st));
    GDA_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * N, st));
    GDA_CUDA(cudaMemsetAsync(nvisited, 0, sizeof(int), st));
    k_row_sums<<<nblk(acc_n), kT, 0, st>>>(acc_keys, acc_cnt, acc_n, N, rowsum);
    GDA_LAUNCH_CHECK();
    k_col_sums<<<nblk(acc_n), kT, 0, st>>>(acc_keys, acc_cnt, acc_n, N, rowsum, colsum);
    GDA_LAUNCH_CHECK();
    k_count_visited<<<nblk(N), kT, 0, st>>>(colsum, N, nvisited);
    GDA_LAUNCH_CHECK();
    k_scores<<<nblk(acc_n), kT, 0, st>>>(acc_keys, acc_cnt, acc_n, N, rowsum, colsum, nvisited, path_len, res->ei, res->w);
    GDA_LAUNCH_CHECK();
    GDA_CUDA(cudaMemcpyAsync(res->cnt, acc_cnt, sizeof(int) * acc_n, cudaMemcpyDeviceToDevice, st));
  }
  GDA_CUDA(cudaStreamSynchronize(st));
  *out = res.release();
  return GDA_OK;
}

int gda_ppmi_walks(const int64_t* edge_index, int64_t E, int64_t N, int path_len, int rounds, uint64_t seed,
                   int32_t* walks_out, gda_stream_t stream) {
  using namespace gda;
  GDA_REQUIRE(N >= 0 && E >= 0 && path_len >= 1 && rounds >= 1, "gda_ppmi_walks: bad size / path_len / rounds");
  GDA_REQUIRE(E == 0 || edge_index != nullptr, "gda_ppmi_walks: edge_index is NULL");
  GDA_REQUIRE(N < (int64_t(1) << 31) && 2 * E < (int64_t(1) << 31) - 1, "gda_ppmi_walks: N and 2E must fit in int32");
  if (N == 0) return GDA_OK;
  GDA_REQUIRE(walks_out != nullptr, "gda_ppmi_walks: walks_out is NULL");
  cudaStream_t st = as_stream(stream);
  Tmp t;
  int rc;
  int *adjptr, *adjcol;
  int64_t adj_nnz = 0;
  if ((rc = build_adjacency(edge_index, E, N, &adjptr, &adjcol, &adj_nnz, t, st, "gda_ppmi_walks"))) return rc;
  k_walk_nodes<<<nblk(static_cast<int64_t>(rounds) * N), kT, 0, st>>>(adjptr, adjcol, N, path_len, rounds, seed, walks_out);
  GDA_LAUNCH_CHECK();
  GDA_CUDA(cudaStreamSynchronize(st));
  return GDA_OK;
}

int64_t gda_wedges_size(const gda_wedges_t* e) { return e ? e->M : -1; }

int gda_wedges_export(const gda_wedges_t* e, int64_t* edge_index_out, float* weight_out, int32_t* counts_out,
                      gda_stream_t stream) {
  GDA_REQUIRE(e != nullptr, "gda_wedges_export: NULL handle");
  if (e->M == 0) return GDA_OK;
  GDA_REQUIRE(edge_index_out && weight_out, "gda_wedges_export: NULL output");
  cudaStream_t st = gda::as_stream(stream);
  GDA_CUDA(cudaMemcpyAsync(edge_index_out, e->ei, sizeof(int64_t) * 2 * static_cast<size_t>(e->M), cudaMemcpyDeviceToDevice, st));
  GDA_CUDA(cudaMemcpyAsync(weight_out, e->w, sizeof(float) * static_cast<size_t>(e->M), cudaMemcpyDeviceToDevice, st));
  if (counts_out)
    GDA_CUDA(cudaMemcpyAsync(counts_out, e->cnt, sizeof(int) * static_cast<size_t>(e->M), cudaMemcpyDeviceToDevice, st));
  return GDA_OK;
}

int gda_wedges_destroy(gda_wedges_t* e) {
  delete e;
  return GDA_OK;
}

}  // extern "C"
