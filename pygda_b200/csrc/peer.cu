// Peer (NVLink) memory for the node-partitioned multi-GPU path.
//
// Each rank owns a row block of every [N, H] activation.  Instead of all-gathering the whole
// matrix before every propagation step, the aggregation kernel reads the neighbour rows it
// needs straight out of the owners' HBM over NVLink (P2P loads inside k_spmm_fast<PEER>), so the
// transfer overlaps the gather/FMA work row by row and only rows that are actually referenced
// cross the fabric.  This file provides the plumbing the reference never had (it is
// single-device): CUDA-IPC symmetric buffers and a device-side
// all-ranks barrier over flags in peer memory (one tiny kernel, no host round trip).
#include <cstring>

#include "common.cuh"

namespace gda {
namespace {

struct FlagPtrs { uint64_t* p[GDA_MAX_PEERS]; };

// thread q: tell rank q that `rank` reached `epoch`, then wait until rank q has told us.
__global__ void k_peer_barrier(FlagPtrs flags, int rank, int num_peers, uint64_t epoch, int* error) {
  const int q = threadIdx.x;
  if (q >= num_peers) return;
  uint64_t* theirs = flags.p[q] + rank;            // slot `rank` of rank q's flag array
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const uint64_t* mine = flags.p[rank] + q;
  const long long t0 = clock64();
  while (true) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > 20000000000LL) {           // ~10 s: a peer died; fail instead of hanging the GPU
      if (error) *error = 1;
      break;
    }
  }
}

// The same barrier with the epoch kept in DEVICE memory: thread 0 advances the rank's own counter and the block
// uses the new value.  Nothing about the launch depends on host state, so a captured CUDA graph can replay it
// (every rank replays the same sequence of barriers on a channel, so the counters stay in step).
__global__ void k_peer_barrier_dev(FlagPtrs flags, int rank, int num_peers, uint64_t* epoch_dev, int* error) {
  __shared__ uint64_t e_sh;
  if (threadIdx.x == 0) {
    const uint64_t e = *epoch_dev + 1;
    *epoch_dev = e;
    e_sh = e;
  }
  __syncthreads();
  const uint64_t epoch = e_sh;
  const int q = threadIdx.x;
  if (q >= num_peers) return;
  uint64_t* theirs = flags.p[q] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const uint64_t* mine = flags.p[rank] + q;
  const long long t0 = clock64();
  while (true) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > 20000000000LL) {
      if (error) *error = 1;
      break;
    }
  }
}

// Halo exchange of a partitioned matrix (dist.py, halo mode): entry i copies row rows[i] of the local matrix into row
// slots[i] of rank peer[i]'s buffer (P2P stores over NVLink).  One warp per entry, 16 bytes per lane and round.
struct PushPtrs { char* p[GDA_MAX_PEERS]; };

__global__ void k_push_rows(const float* __restrict__ src, int64_t ld, const int32_t* __restrict__ rows,
                            const int32_t* __restrict__ slots, const int32_t* __restrict__ peer, int64_t count,
                            PushPtrs dst, int64_t dst_ld, int H4) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < count; i += nwarps) {
    const float4* s = reinterpret_cast<const float4*>(src + static_cast<int64_t>(__ldg(rows + i)) * ld);
    float4* d = reinterpret_cast<float4*>(dst.p[__ldg(peer + i)] + static_cast<int64_t>(__ldg(slots + i)) * dst_ld * 4);
    for (int c = lane; c < H4; c += 32) d[c] = s[c];
  }
}

}  // namespace
}  // namespace gda

using namespace gda;

extern "C" {

int gda_sym_alloc(int64_t bytes, void** ptr, unsigned char* handle_out) {
  GDA_REQUIRE(bytes > 0 && ptr && handle_out, "gda_sym_alloc: bad arguments");
  *ptr = nullptr;
  GDA_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));
  GDA_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  GDA_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  return GDA_OK;
}

int gda_sym_open(const unsigned char* handle, void** ptr) {
  GDA_REQUIRE(handle && ptr, "gda_sym_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  GDA_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GDA_OK;
}

int gda_sym_close(void* ptr) {
  if (ptr) GDA_CUDA(cudaIpcCloseMemHandle(ptr));
  return GDA_OK;
}

int gda_sym_free(void* ptr) {
  if (ptr) GDA_CUDA(cudaFree(ptr));
  return GDA_OK;
}

int gda_peer_barrier(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t epoch, int* error_flag,
                     gda_stream_t stream) {
  GDA_REQUIRE(peer_flags && num_peers >= 1 && num_peers <= GDA_MAX_PEERS && rank >= 0 && rank < num_peers,
              "gda_peer_barrier: bad arguments");
  FlagPtrs f;
  for (int i = 0; i < num_peers; ++i) {
    GDA_REQUIRE(peer_flags[i] != nullptr, "gda_peer_barrier: NULL flag array");
    f.p[i] = peer_flags[i];
  }
  k_peer_barrier<<<1, 32, 0, as_stream(stream)>>>(f, rank, num_peers, epoch, error_flag);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_push_rows_f32(const float* src, int64_t ld, const int32_t* rows, const int32_t* slots, const int32_t* peer,
                      int64_t count, void* const* peer_base, int num_peers, int64_t peer_ld, int H, gda_stream_t stream) {
  GDA_REQUIRE(count >= 0 && H > 0 && H % 4 == 0 && ld % 4 == 0 && peer_ld % 4 == 0,
              "gda_push_rows_f32: H and leading dimensions must be multiples of 4");
  if (count == 0) return GDA_OK;
  GDA_REQUIRE(src && rows && slots && peer && peer_base && num_peers >= 1 && num_peers <= GDA_MAX_PEERS,
              "gda_push_rows_f32: bad arguments");
  PushPtrs d;
  for (int i = 0; i < GDA_MAX_PEERS; ++i) d.p[i] = static_cast<char*>(peer_base[i < num_peers ? i : 0]);
  int64_t blocks = ceil_div(count, 8);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  k_push_rows<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(src, ld, rows, slots, peer, count, d, peer_ld,
                                                                            H / 4);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_peer_barrier_dev(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t* epoch_dev, int* error_flag,
                         gda_stream_t stream) {
  GDA_REQUIRE(peer_flags && epoch_dev && num_peers >= 1 && num_peers <= GDA_MAX_PEERS && rank >= 0 && rank < num_peers,
              "gda_peer_barrier_dev: bad arguments");
  FlagPtrs f;
  for (int i = 0; i < num_peers; ++i) {
    GDA_REQUIRE(peer_flags[i] != nullptr, "gda_peer_barrier_dev: NULL flag array");
    f.p[i] = peer_flags[i];
  }
  k_peer_barrier_dev<<<1, 32, 0, as_stream(stream)>>>(f, rank, num_peers, epoch_dev, error_flag);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // extern "C"
