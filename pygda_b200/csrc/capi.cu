// C-ABI glue: error string, version, graph handle wrappers (see include/gda.h).
#include <atomic>

#include "graph.cuh"

namespace gda {
namespace {
thread_local std::string g_last_error;
}
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace gda

extern "C" {

int gda_version(void) { return GDA_VERSION; }
int gda_sm_arch(void) { return 100; }
const char* gda_last_error(void) { return gda::g_last_error.c_str(); }
uint64_t gda_launch_count(void) { return gda::g_launches.load(std::memory_order_relaxed); }

int gda_graph_create(const int64_t* edge_index, int64_t E, int64_t N, const float* edge_weight, int flags,
                     gda_stream_t stream, gda_graph_t** out) {
  try {
    return gda::graph_create(edge_index, E, N, edge_weight, flags, gda::as_stream(stream), out);
  } catch (const std::exception& e) {
    return gda::fail(GDA_E_CUDA, std::string("gda_graph_create: ") + e.what());
  }
}

int gda_graph_destroy(gda_graph_t* g) {
  delete g;
  return GDA_OK;
}

int gda_graph_info(const gda_graph_t* g, int64_t* N, int64_t* nnz, int64_t* num_long_rows,
                   int64_t* num_long_rows_t) {
  GDA_REQUIRE(g != nullptr, "gda_graph_info: graph is NULL");
  if (N) *N = g->N;
  if (nnz) *nnz = g->nnz;
  if (num_long_rows) *num_long_rows = g->csr.num_long;
  if (num_long_rows_t) *num_long_rows_t = g->csr_t.num_long;
  return GDA_OK;
}

int gda_graph_export_coo(const gda_graph_t* g, int64_t* edge_index_out, float* weight_out, gda_stream_t stream) {
  return gda::graph_export_coo(g, edge_index_out, weight_out, gda::as_stream(stream));
}

int gda_graph_export_csr(const gda_graph_t* g, int transpose, int32_t* rowptr, int32_t* colidx, float* vals,
                         gda_stream_t stream) {
  return gda::graph_export_csr(g, transpose, rowptr, colidx, vals, gda::as_stream(stream));
}

int gda_graph_partition(const gda_graph_t* g, int64_t row_lo, int64_t row_hi, int64_t rows_per_rank,
                        gda_stream_t stream, gda_graph_t** out) {
  try {
    return gda::graph_partition(g, row_lo, row_hi, rows_per_rank, gda::as_stream(stream), out);
  } catch (const std::exception& e) {
    return gda::fail(GDA_E_CUDA, std::string("gda_graph_partition: ") + e.what());
  }
}

}  // extern "C"
