// Internal GEMM entry points (see gemm.cu for the dispatcher).
#pragma once
#include "common.cuh"

namespace gda {
int simt_splits(int64_t M, int64_t N, int64_t K);
int64_t simt_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes,
              cudaStream_t st);

// tcgen05 split-bf16 path (gemm_tc.cu): returns true when the shape is handled there
bool tc_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                  const float* A, const float* B, const float* C);
int64_t tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K);
int gemm_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
            const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes,
            cudaStream_t st);
int split_bf16(const float* x, int64_t rows, int64_t cols, int64_t ldx, void* hi, void* lo, int64_t ldo,
               cudaStream_t st);
bool bf16x3_shape_ok(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb);
int64_t bf16x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gemm_bf16x3(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_lo,
                int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb, float* C, int64_t ldc, void* ws,
                int64_t ws_bytes, cudaStream_t st);
// plain bf16 operands (one UMMA per K step), fp32 accumulation, C fp32 or bf16
int64_t bf16_workspace_bytes(int64_t M, int64_t N, int64_t K, int out_bf16);
int gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
              int64_t ldb, void* C, int64_t ldc, int out_bf16, void* ws, int64_t ws_bytes, cudaStream_t st);
}  // namespace gda

namespace gda {
// skinny heads (gemm_skinny.cu): kind 0 = not applicable
int skinny_kind(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                const float* A, const float* B, const float* C);
int64_t skinny_workspace_bytes(int kind, int64_t M, int64_t N, int64_t K);
int gemm_skinny(int kind, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B,
                int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes, cudaStream_t st);
}  // namespace gda
