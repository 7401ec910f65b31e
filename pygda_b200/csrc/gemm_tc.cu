// tcgen05 split-bf16 GEMM -- placeholder until the kernel lands (next milestone):
// reports "not supported" so gda_gemm_f32 routes everything to the SIMT kernel.
#include "gemm.cuh"

namespace gda {
bool tc_supported(int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const float*, const float*,
                  const float*) { return false; }
int64_t tc_workspace_bytes(int, int, int64_t, int64_t, int64_t) { return 0; }
int gemm_tc(int, int, int64_t, int64_t, int64_t, float, const float*, int64_t, const float*, int64_t, float,
            float*, int64_t, void*, int64_t, cudaStream_t) {
  return fail(GDA_E_UNSUPPORTED, "tcgen05 GEMM not built");
}
}  // namespace gda
