// Internal layout of gda_graph_t (HBM-resident, int32 indices, fp32 weights).
#pragma once
#include <memory>
#include <vector>

#include "common.cuh"

namespace gda {

// Rows with more than this many non-zeros are split into segments of this size so
// that no warp walks a hub row alone (see spmm.cu).
constexpr int kLongRowSegment = 64;

struct Csr {
  int32_t* rowptr = nullptr;   // [N+1]
  int32_t* colidx = nullptr;   // [nnz]
  float*   vals = nullptr;     // [nnz]
  // long-row split
  int32_t  num_long = 0;       // rows with degree > seg
  int32_t  num_segs = 0;       // total segments over those rows
  int32_t* long_rows = nullptr;     // [num_long] row ids
  int32_t* long_seg_ptr = nullptr;  // [num_long+1] first segment of each long row
  int32_t* seg_long = nullptr;      // [num_segs] long-row index of each segment
  int32_t* counters = nullptr;      // [num_long] arrival counters, zero between launches
  // work list for k_spmm_tasks (spmm.cu): one entry per short row and per long-row segment, sorted by
  // descending length so that a grid-stride walk hands every warp the same mix of lengths.
  //   x = first non-zero;  y = segment flag << 31 | (len - 1) << 25 | row id (or global segment index)
  uint2*   tasks = nullptr;         // [num_tasks]; null when not built (empty rows possible, ids >= 2^25)
  int32_t  num_tasks = 0;
  bool     may_have_empty_rows = true;   // false when every row holds a self loop
  bool     skip_empty_rows = false;      // GDA_SKIP_EMPTY_ROWS: empty rows are left out of the work list and never written
};

}  // namespace gda

struct gda_graph {
  int64_t N = 0, E = 0, nnz = 0;
  int flags = 0, device = 0, seg = gda::kLongRowSegment;
  // COO in the reference's order (kept edges, then loops), normalised weights
  int32_t* coo_src = nullptr;
  int32_t* coo_dst = nullptr;
  float*   coo_w = nullptr;
  float*   dinv = nullptr;     // [N] deg^-1/2
  // every raw edge weight is 1 and the normalisation is symmetric: vals[e] == dinv[src] * dinv[dst], so the
  // aggregation factors as D (A + I) D and chains can run on the weight-free kernel (spmm.cu: k_spmm_unw)
  bool     unit_weights = false;
  gda::Csr csr;                // rows = targets: Y = A_hat X
  gda::Csr csr_t;              // rows = sources: Y = A_hat^T X
  // row-block partition of a larger graph (peer path): colidx = owner << 28 | local row
  bool    peer_packed = false;
  int64_t rows_per_rank = 0, global_N = 0, row_lo = 0;
  ~gda_graph();
};

namespace gda {
int graph_create(const int64_t* ei, int64_t E, int64_t N, const float* w, int flags, cudaStream_t st,
                 gda_graph** out);
int graph_export_coo(const gda_graph* g, int64_t* ei_out, float* w_out, cudaStream_t st);
int graph_export_csr(const gda_graph* g, int transpose, int32_t* rowptr, int32_t* colidx, float* vals,
                     cudaStream_t st);
int graph_partition(const gda_graph* g, int64_t row_lo, int64_t row_hi, int64_t rows_per_rank, cudaStream_t st,
                    gda_graph** out);
}  // namespace gda
