// Shared helpers for libgda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/gda.h"

namespace gda {

void set_error(const std::string& msg);          // capi.cu (thread-local)
void count_launch();                             // capi.cu (statistics only)

inline int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

#define GDA_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      return ::gda::fail(GDA_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    }                                                                               \
  } while (0)

#define GDA_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ::gda::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      return ::gda::fail(GDA_E_CUDA, std::string(__FILE__) + ":" + std::to_string(__LINE__) + \
                                         ": launch: " + cudaGetErrorString(_e));   \
    }                                                                               \
  } while (0)

#define GDA_REQUIRE(cond, msg)                                     \
  do {                                                             \
    if (!(cond)) return ::gda::fail(GDA_E_INVALID, std::string(msg)); \
  } while (0)

inline cudaStream_t as_stream(gda_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;   // B200

// ---- counter-based keep mask for dropout: the same (seed, index) always gives the
// same decision, so the backward regenerates the mask instead of storing it.
__host__ __device__ __forceinline__ uint32_t mix_hash(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}

// keep iff uniform(0,1) >= p
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thresh) {
  return mix_hash(seed, idx) >= thresh;
}

inline uint32_t dropout_threshold(float p) {
  double t = static_cast<double>(p) * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return static_cast<uint32_t>(t);
}

}  // namespace gda
