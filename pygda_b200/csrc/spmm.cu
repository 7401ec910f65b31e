// Sparse neighbour aggregation  Y = A_hat X  /  Y = A_hat^T X   (HBM / L2 gather bound).
//
// Replaces MessagePassing.propagate (index_select -> mul -> scatter_add_) at
// pygda/nn/prop_gcn_conv.py:208-210,238 and pygda/nn/cached_gcn_conv.py:138,156,
// and its autograd backward (the same product with A_hat^T).
//
// Layout: CSR with int32 indices and fp32 weights (graph.cu); features row-major.
// A group of LPR lanes (a full warp at H = 128 fp32) owns a BLOCK of consecutive rows and
// streams the contiguous run of their non-zeros: (colidx, weight) pairs are read LPR at a
// time with one coalesced access and broadcast by shuffle; neighbour rows are gathered with
// 16-byte loads, U gathers in flight per lane, batches running straight across row
// boundaries so that the rowptr -> colidx -> gather latency chain is paid once per block,
// not once per row (round-1 profile: the row-per-warp form was latency bound at 135 us,
// DRAM 10 %, L2 18 %; see profiles/).  Accumulation is fp32, sequential in CSR (= COO) order,
// hence deterministic and equal to a sequential scatter_add.
// Rows longer than `seg` non-zeros are skipped by the row blocks and handled as `seg`-sized
// segments by separate groups scheduled FIRST; the last segment to finish reduces the partial
// sums in segment order (still deterministic) -- no hub row is walked by a single warp.
// Optional fused epilogue on the output row: + bias, ReLU, dropout (hash mask).
#include <cuda_bf16.h>

#include <cstdlib>

#include "graph.cuh"

#ifndef GDA_SPMM_MIN_CTAS
#define GDA_SPMM_MIN_CTAS 12
#endif
#ifndef GDA_SPMM_BLOCK
#define GDA_SPMM_BLOCK 64
#endif
#ifndef GDA_SPMM_U4
#define GDA_SPMM_U4 8
#endif
#ifndef GDA_ROWS_BLOCK
#define GDA_ROWS_BLOCK 128
#endif
#ifndef GDA_ROWS_MIN_CTAS
#define GDA_ROWS_MIN_CTAS 12
#endif
#ifndef GDA_TASKS_MIN_CTAS
#define GDA_TASKS_MIN_CTAS 10
#endif
#ifndef GDA_TASKS_MIN_CTAS_WIDE
#define GDA_TASKS_MIN_CTAS_WIDE 8
#endif

namespace gda {
namespace {

struct Epilogue {
  const float* bias;   // may be null
  int flags;           // GDA_EPI_*
  uint32_t thresh;     // dropout threshold
  float scale;         // 1/(1-p)
  uint64_t seed;
  const uint64_t* seed_offset;   // device, may be null
};

template <typename T, int VEC> struct VecIO;

// 16-byte gather of a neighbour-row slice.  GDA_GATHER_LOAD selects the cache policy at COMPILE time for A/B builds
// (`make VARIANT=ldcg|noalloc`, profiles/bench_spmm.py with GDA_LIB_PATH): 0 = ld.global.nc through L1 (the shipped
// default), 1 = ld.global.cg (L2 only), 2 = ld.global.nc.L1::no_allocate.  L1 hits on the gathers are 2 % at config 2
// while l1tex is the busiest unit of the kernel (profiles/r1_j_final), so not allocating may be the cheaper path.
#ifndef GDA_GATHER_LOAD
#define GDA_GATHER_LOAD 0
#endif
static __device__ __forceinline__ uint4 gather16(const void* p) {
#if GDA_GATHER_LOAD == 1
  return __ldcg(reinterpret_cast<const uint4*>(p));
#elif GDA_GATHER_LOAD == 2
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const uint4*>(p));
#endif
}

template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const uint4 v = gather16(p);
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[1]) { f[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&f)[1]) { *p = f[0]; }
};
template <> struct VecIO<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = gather16(p);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct VecIO<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[1]) {
    f[0] = __bfloat162float(*p);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[1]) {
    *p = __float2bfloat16_rn(f[0]);
  }
};

template <int VEC>
__device__ __forceinline__ void apply_epilogue(float (&acc)[VEC], const Epilogue& epi, int64_t row, int c0, int H) {
  uint64_t seed = epi.seed;
  if ((epi.flags & GDA_EPI_DROPOUT) && epi.seed_offset) seed += __ldg(epi.seed_offset);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float y = acc[v];
    if (epi.bias) y += __ldg(epi.bias + c0 + v);
    if (epi.flags & GDA_EPI_RELU) y = fmaxf(y, 0.f);
    if (epi.flags & GDA_EPI_DROPOUT)
      y = dropout_keep(seed, static_cast<uint64_t>(row) * H + c0 + v, epi.thresh) ? y * epi.scale : 0.f;
    acc[v] = y;
  }
}

inline bool fast_forced() {     // experiments: keep the row-block streaming kernel for the common width too
  static const bool on = [] { const char* e = std::getenv("GDA_SPMM_STREAM"); return e && e[0] == '1'; }();
  return on;
}

inline bool generic_forced() {
  static const bool on = [] { const char* e = std::getenv("GDA_SPMM_GENERIC"); return e && e[0] == '1'; }();
  return on;
}

// rows per group: the group keeps rowptr[row0 .. row0+RPG] in its lanes
template <int LPR> struct RowsPerGroup { static constexpr int value = (LPR - 1) < 8 ? (LPR - 1) : 8; };

template <typename T, int VEC, int LPR, int U>
__global__ void __launch_bounds__(256)
k_spmm(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ vals,
       int num_segs, const int* __restrict__ long_rows, const int* __restrict__ long_seg_ptr,
       const int* __restrict__ seg_long, int* __restrict__ counters, int seg,
       const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy, int N, int H,
       Epilogue epi, float* __restrict__ partial) {
  constexpr int GPW = 32 / LPR;                         // groups per warp
  constexpr int RPG = RowsPerGroup<LPR>::value;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int64_t warp = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t group = warp * GPW + sub;
  const int64_t row_groups = (static_cast<int64_t>(N) + RPG - 1) / RPG;
  if (group >= static_cast<int64_t>(num_segs) + row_groups) return;

  // ---- range of this group: lane i holds the boundary before its i-th row --------------------
  int row0, nrows, L = -1, rp;
  if (group < num_segs) {                               // one segment of a long row
    L = __ldg(seg_long + group);
    row0 = __ldg(long_rows + L);
    nrows = 1;
    const int j = static_cast<int>(group) - __ldg(long_seg_ptr + L);
    const int s = __ldg(rowptr + row0) + j * seg;
    const int e = min(s + seg, __ldg(rowptr + row0 + 1));
    rp = (l == 0) ? s : e;
  } else {
    row0 = static_cast<int>((group - num_segs) * RPG);
    nrows = min(RPG, N - row0);
    rp = __ldg(rowptr + row0 + min(l, nrows));
  }
  const int pos0 = __shfl_sync(gmask, rp, 0, LPR);
  const int end = __shfl_sync(gmask, rp, nrows, LPR);

  for (int cb = 0; cb < H; cb += LPR * VEC) {
    const int c0 = cb + l * VEC;
    const bool active = c0 < H;
    const T* __restrict__ Xc = X + c0;
    const unsigned ld32 = static_cast<unsigned>(ldx);   // N * ldx < 2^32 is checked on the host
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

    auto flush = [&](int r) {                           // r: row index inside the block
      if (!active) return;
      if (L < 0) {
        apply_epilogue<VEC>(acc, epi, row0 + r, c0, H);
        VecIO<T, VEC>::store(Y + static_cast<int64_t>(row0 + r) * ldy + c0, acc);
      } else {
        float* dst = partial + static_cast<int64_t>(group) * H + c0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) __stcg(dst + v, acc[v]);
      }
    };

    int r = 0, rs = pos0, re = __shfl_sync(gmask, rp, 1, LPR);
    int p = pos0;
    int cbase = pos0 - LPR;                             // chunk currently staged in (myc, myv): none
    int myc = 0;
    float myv = 0.f;

    while (true) {
      // ---- settle: finish rows that end at p, skip long rows (their segments cover them) ----
      bool moved = true;
      while (moved && r < nrows) {
        moved = false;
        const bool lng = (L < 0) && (re - rs > seg);
        if (lng) p = re;
        if (p >= re) {
          if (!lng) flush(r);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
          ++r;
          rs = re;
          if (r < nrows) re = __shfl_sync(gmask, rp, r + 1, LPR);
          moved = true;
        }
      }
      if (r >= nrows) break;

      // ---- stage the next LPR (colidx, weight) pairs when p left the staged chunk ----
      if (p >= cbase + LPR) {
        cbase = p;
        const int idx = cbase + l;
        myc = 0; myv = 0.f;
        if (idx < end) { myc = __ldg(colidx + idx); myv = __ldg(vals + idx); }
      }
      const int j = p - cbase;
      const int nvalid = min(U, min(cbase + LPR, end) - p);

      // ---- issue up to U gathers, regardless of row boundaries ----
      // (round-1 profile: the predicated zero-filling form of this loop cost ~320 warp
      // instructions per batch; full batches now take a predicate-free path with 32-bit
      // offset arithmetic)
      float xv[U][VEC];
      float wv[U];
      if (nvalid == U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned cj = static_cast<unsigned>(__shfl_sync(gmask, myc, j + u, LPR));
          wv[u] = __shfl_sync(gmask, myv, j + u, LPR);
          if (active) VecIO<T, VEC>::load(Xc + static_cast<size_t>(cj * ld32), xv[u]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned cj = static_cast<unsigned>(__shfl_sync(gmask, myc, j + u, LPR));
          wv[u] = __shfl_sync(gmask, myv, j + u, LPR);
          if (active && u < nvalid) VecIO<T, VEC>::load(Xc + static_cast<size_t>(cj * ld32), xv[u]);
        }
      }

      // ---- consume in order ----
      if (p + nvalid <= re) {                            // whole batch inside the current row
        if (active) {
          if (nvalid == U) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
          } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (u < nvalid) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
              }
          }
        }
        p += nvalid;
      } else {                                           // rows end (empty / long rows may begin) inside
        int consumed = nvalid;
        bool stop = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u < nvalid && !stop) {
            const int e = p + u;
            if (e >= re) {
              do {
                flush(r);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
                ++r;
                rs = re;
                re = (r < nrows) ? __shfl_sync(gmask, rp, r + 1, LPR) : 0x7fffffff;
              } while (r < nrows && e >= re);
              if (r >= nrows || ((L < 0) && (re - rs > seg))) { stop = true; consumed = u; }
            }
            if (!stop && active) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
            }
          }
        }
        p += consumed;
      }
    }
  }

  if (L >= 0) {                                         // last segment to arrive reduces, in order
    __threadfence();
    __syncwarp(gmask);
    int old = 0;
    const int first = __ldg(long_seg_ptr + L), nseg = __ldg(long_seg_ptr + L + 1) - first;
    if (l == 0) old = atomicAdd(counters + L, 1);
    old = __shfl_sync(gmask, old, 0, LPR);
    if (old == nseg - 1) {
      __threadfence();
      for (int cb = 0; cb < H; cb += LPR * VEC) {
        const int c0 = cb + l * VEC;
        if (c0 >= H) continue;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        for (int s = 0; s < nseg; ++s) {
          const float* src = partial + static_cast<int64_t>(first + s) * H + c0;
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += __ldcg(src + v);
        }
        apply_epilogue<VEC>(acc, epi, row0, c0, H);
        VecIO<T, VEC>::store(Y + static_cast<int64_t>(row0) * ldy + c0, acc);
      }
      if (l == 0) counters[L] = 0;                      // ready for the next launch
    }
  }
}

// ------------------------------------------------------------------------------------------
// Lean kernel for graphs in which every row has at least one non-zero (anything built with
// self loops, i.e. every GCN-normalised graph).  Same decomposition as k_spmm, but the row
// structure of each staged chunk of LPR non-zeros is turned into two bit masks up front
// (end-of-row, belongs-to-a-split-long-row), so the inner loop is, per non-zero:
// 2 SHFL + 2 address ops + 1 LDG.128 + VEC FFMA + 1 mask test -- no per-entry compares
// against row bounds, no predicated register merging (profiles/r1_b_*: the generic kernel
// executed 44 warp instructions per non-zero, this one ~12).
struct PeerTable { const void* p[GDA_MAX_PEERS]; };

template <typename T, int VEC, int LPR, int U, bool EPI, bool PEER>
__global__ void __launch_bounds__(GDA_SPMM_BLOCK, GDA_SPMM_MIN_CTAS)
k_spmm_fast(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ vals,
            int num_segs, const int* __restrict__ long_rows, const int* __restrict__ long_seg_ptr,
            const int* __restrict__ seg_long, int* __restrict__ counters, int seg,
            const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy, int N, int H,
            Epilogue epi, float* __restrict__ partial, PeerTable peers, int pad_col) {
  static_assert(LPR % U == 0, "batches must tile a chunk");
  constexpr int GPW = 32 / LPR;
  constexpr int RPG = RowsPerGroup<LPR>::value;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int64_t warp = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t group = warp * GPW + sub;
  const int64_t row_groups = (static_cast<int64_t>(N) + RPG - 1) / RPG;
  if (group >= static_cast<int64_t>(num_segs) + row_groups) return;

  int row0, nrows, L = -1, rp;
  if (group < num_segs) {
    L = __ldg(seg_long + group);
    row0 = __ldg(long_rows + L);
    nrows = 1;
    const int j = static_cast<int>(group) - __ldg(long_seg_ptr + L);
    const int s = __ldg(rowptr + row0) + j * seg;
    const int e = min(s + seg, __ldg(rowptr + row0 + 1));
    rp = (l == 0) ? s : e;
  } else {
    row0 = static_cast<int>((group - num_segs) * RPG);
    nrows = min(RPG, N - row0);
    rp = __ldg(rowptr + row0 + min(l, nrows));
  }
  // block boundaries, identical in every lane of the group (static shuffle indices)
  int b[RPG + 1];
#pragma unroll
  for (int i = 0; i <= RPG; ++i) b[i] = __shfl_sync(gmask, rp, i < LPR ? i : LPR - 1, LPR);
  const int end = b[RPG];
  const unsigned ldxb = static_cast<unsigned>(ldx) * static_cast<unsigned>(sizeof(T));   // row pitch in bytes
  const unsigned ldyb = static_cast<unsigned>(ldy) * static_cast<unsigned>(sizeof(T));

  for (int cb = 0; cb < H; cb += LPR * VEC) {
    const int c0 = cb + l * VEC;
    const bool active = c0 < H;
    const char* __restrict__ Xc = reinterpret_cast<const char*>(X + c0);
    char* __restrict__ Yc = reinterpret_cast<char*>(Y + c0);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    int r = 0;                                          // row (inside the block) being accumulated
    int p = b[0];

    while (p < end) {
      // ---- stage LPR non-zeros and classify them ----
      const int e = p + l;
      const bool valid = e < end;
      int myc = 0;
      float myv = 0.f;
      if (valid) { myc = __ldg(colidx + e); myv = __ldg(vals + e); }
      int rid = 0, rid1 = 0, mylen = 0, myend = end;
#pragma unroll
      for (int i = 1; i <= RPG; ++i) {
        rid += (e >= b[i]);
        rid1 += (e + 1 >= b[i]);
        if (e >= b[i - 1] && e < b[i]) { mylen = b[i] - b[i - 1]; myend = b[i]; }
      }
      const bool skip = valid && (L < 0) && (mylen > seg);
      const bool isend = valid && !skip && (rid1 != rid);
      const unsigned skipmask = (__ballot_sync(gmask, skip) >> (sub * LPR)) & ((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u));
      unsigned endmask = (__ballot_sync(gmask, isend) >> (sub * LPR)) & ((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u));
      if (skipmask & 1u) {                              // p sits in a split long row: jump over it
        p = __shfl_sync(gmask, myend, 0, LPR);
        r = __shfl_sync(gmask, rid, 0, LPR) + 1;
        continue;
      }
      const int cnt = skipmask ? (__ffs(skipmask) - 1) : min(LPR, end - p);
      if (l >= cnt) { myc = pad_col; myv = 0.f; }       // padding: weight 0, gathers a local row

      for (int j = 0; j < cnt; j += U) {
        float xv[U][VEC];
        float wv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned cj = static_cast<unsigned>(__shfl_sync(gmask, myc, j + u, LPR));
          wv[u] = __shfl_sync(gmask, myv, j + u, LPR);
          if (PEER) {
            // column = owner << 28 | row inside the owner's block: P2P load over NVLink when the
            // owner is another GPU (the owners' blocks are CUDA-IPC mapped, peer.cu)
            const char* base = static_cast<const char*>(peers.p[cj >> 28]) + c0 * sizeof(T);
            if (active)
              VecIO<T, VEC>::load(reinterpret_cast<const T*>(base + static_cast<uint64_t>(cj & 0x0FFFFFFFu) * ldxb), xv[u]);
          } else {
            // one IMAD.WIDE.U32: base + col * pitch
            if (active) VecIO<T, VEC>::load(reinterpret_cast<const T*>(Xc + static_cast<uint64_t>(cj) * ldxb), xv[u]);
          }
        }
        const unsigned em = endmask >> j;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (active) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
          }
          if (em & (1u << u)) {                         // last non-zero of row r
            if (active) {
              if (L < 0) {
                if (EPI) apply_epilogue<VEC>(acc, epi, row0 + r, c0, H);
                VecIO<T, VEC>::store(reinterpret_cast<T*>(Yc + static_cast<uint64_t>(static_cast<unsigned>(row0 + r)) * ldyb), acc);
              } else {
                float* dst = partial + static_cast<int64_t>(group) * H + c0;
#pragma unroll
                for (int v = 0; v < VEC; ++v) __stcg(dst + v, acc[v]);
              }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
            ++r;
          }
        }
      }
      p += cnt;
    }
  }

  if (L >= 0) {                                         // last segment to arrive reduces, in order
    __threadfence();
    __syncwarp(gmask);
    int old = 0;
    const int first = __ldg(long_seg_ptr + L), nseg = __ldg(long_seg_ptr + L + 1) - first;
    if (l == 0) old = atomicAdd(counters + L, 1);
    old = __shfl_sync(gmask, old, 0, LPR);
    if (old == nseg - 1) {
      __threadfence();
      for (int cb = 0; cb < H; cb += LPR * VEC) {
        const int c0 = cb + l * VEC;
        if (c0 >= H) continue;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        for (int s = 0; s < nseg; ++s) {
          const float* src = partial + static_cast<int64_t>(first + s) * H + c0;
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += __ldcg(src + v);
        }
        if (EPI) apply_epilogue<VEC>(acc, epi, row0, c0, H);
        VecIO<T, VEC>::store(Y + static_cast<int64_t>(row0) * ldy + c0, acc);
      }
      if (l == 0) counters[L] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-per-warp kernel for the common width (one 16-byte slice per lane: H = 128 fp32 / 256 bf16).
// The probes in profiles/probes/ show what bounds a gather on this GPU: with ~48-64 resident
// warps per SM and only 4 independent 512-byte row gathers in flight per warp, plain LDG.128
// reaches 16.5 TB/s of gathered rows under the aggregation's own read/write pattern (19.5 TB/s
// read-only) -- occupancy, not per-warp depth, hides the L2/HBM latency, and cp.async staging is
// slower (16.8 TB/s best).  So: minimal state per warp (~40 registers), (colidx, weight) read
// with warp-uniform loads (one L1 transaction, no shuffles), the next batch's indices fetched
// while the current batch's rows are in flight, grid-stride over rows.  Long rows are again
// handled as `seg`-sized segments (work items < num_segs) with an ordered reduction.
template <typename T, int VEC, bool EPI, bool PEER>
__global__ void __launch_bounds__(GDA_ROWS_BLOCK, GDA_ROWS_MIN_CTAS)
k_spmm_rows(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ vals,
            int num_segs, const int* __restrict__ long_rows, const int* __restrict__ long_seg_ptr,
            const int* __restrict__ seg_long, int* __restrict__ counters, int seg,
            const T* __restrict__ X, unsigned ldxb, T* __restrict__ Y, unsigned ldyb, int N, int H,
            Epilogue epi, float* __restrict__ partial, PeerTable peers, int pad_col) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int c0 = lane * VEC;
  const char* __restrict__ Xc = reinterpret_cast<const char*>(X + c0);
  const int total = num_segs + N;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
    int row, p, e, L = -1;
    if (item < num_segs) {
      L = __ldg(seg_long + item);
      row = __ldg(long_rows + L);
      p = __ldg(rowptr + row) + (item - __ldg(long_seg_ptr + L)) * seg;
      e = min(p + seg, __ldg(rowptr + row + 1));
    } else {
      row = item - num_segs;
      p = __ldg(rowptr + row);
      e = __ldg(rowptr + row + 1);
      if (e - p > seg) continue;                        // covered by its segments
    }
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    // full batches: no predicates, indices through one base pointer with immediate offsets
    // (profiles: the predicated / index-prefetching form of this loop executed 43 instructions per
    // non-zero and spilled; the pure gather probe needs 8.5)
    const int* __restrict__ ci = colidx + p;
    const float* __restrict__ cw = vals + p;
    int left = e - p;
    for (; left >= U; left -= U, ci += U, cw += U) {
      unsigned cj[U];
      float wv[U];
      float xv[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) { cj[u] = static_cast<unsigned>(__ldg(ci + u)); wv[u] = __ldg(cw + u); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const char* src = PEER ? static_cast<const char*>(peers.p[cj[u] >> 28]) + c0 * sizeof(T) +
                                     static_cast<uint64_t>(cj[u] & 0x0FFFFFFFu) * ldxb
                               : Xc + static_cast<uint64_t>(cj[u]) * ldxb;
        VecIO<T, VEC>::load(reinterpret_cast<const T*>(src), xv[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
    }
    if (left > 0) {                                     // 1..U-1 trailing non-zeros
      unsigned cj[U - 1];
      float wv[U - 1];
      float xv[U - 1][VEC];
#pragma unroll
      for (int u = 0; u < U - 1; ++u) {
        const bool ok = u < left;
        cj[u] = ok ? static_cast<unsigned>(__ldg(ci + u)) : static_cast<unsigned>(pad_col);
        wv[u] = ok ? __ldg(cw + u) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U - 1; ++u) {
        const char* src = PEER ? static_cast<const char*>(peers.p[cj[u] >> 28]) + c0 * sizeof(T) +
                                     static_cast<uint64_t>(cj[u] & 0x0FFFFFFFu) * ldxb
                               : Xc + static_cast<uint64_t>(cj[u]) * ldxb;
        VecIO<T, VEC>::load(reinterpret_cast<const T*>(src), xv[u]);    // padding: weight 0, a local row
      }
#pragma unroll
      for (int u = 0; u < U - 1; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fmaf(wv[u], xv[u][v], acc[v]);
    }
    if (L < 0) {
      if (EPI) apply_epilogue<VEC>(acc, epi, row, c0, H);
      VecIO<T, VEC>::store(reinterpret_cast<T*>(reinterpret_cast<char*>(Y + c0) + static_cast<uint64_t>(row) * ldyb), acc);
    } else {
      float* dst = partial + static_cast<int64_t>(item) * H + c0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) __stcg(dst + v, acc[v]);
      __threadfence();
      __syncwarp();
      int old = 0;
      const int first = __ldg(long_seg_ptr + L), nseg = __ldg(long_seg_ptr + L + 1) - first;
      if (lane == 0) old = atomicAdd(counters + L, 1);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == nseg - 1) {                            // last segment to arrive reduces, in order
        __threadfence();
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        for (int sgi = 0; sgi < nseg; ++sgi) {
          const float* srcp = partial + static_cast<int64_t>(first + sgi) * H + c0;
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += __ldcg(srcp + v);
        }
        if (EPI) apply_epilogue<VEC>(acc, epi, row, c0, H);
        VecIO<T, VEC>::store(reinterpret_cast<T*>(reinterpret_cast<char*>(Y + c0) + static_cast<uint64_t>(row) * ldyb), acc);
        if (lane == 0) counters[L] = 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Work-list kernel for the common width (one 16-byte slice per lane), the default when the graph
// carries a length-sorted task list (graph.cu: build_tasks).  What the ncu capture of k_spmm_rows
// showed (profiles/r1_d_rows_kernel): DRAM traffic = B_alg, but only 58 % of the resident warps
// active on average and 5 serial L2 latencies per average row (rowptr -> colidx -> gather, then
// colidx -> gather per further batch).  Two changes, both in the schedule, none in the arithmetic:
//  * static balance: tasks (short rows and 64-nnz segments of long rows) are sorted by descending
//    length at graph build; a grid-stride walk over that list gives every warp one task from each
//    length stratum, so all warps finish together (longest-processing-time-first scheduling);
//  * software pipeline over tasks: while the gathers of task t are in flight the warp has already
//    loaded the descriptor of task t+2 and the (colidx, weight) pairs of task t+1 (one coalesced
//    load, lane l holds pair l), so a task costs ceil(len / U) dependent L2 round trips, not
//    2 + ceil(len / U).  Indices and weights reach the lanes by shuffle.
// The summation order inside a row is unchanged (sequential in CSR = COO order, segments reduced
// in segment order), so results are bit-identical to k_spmm_rows.
// Task of warp w in round r of the work list.  The list is sorted by descending length; a plain grid-stride walk
// (r * W + w) hands warp 0 the longest task of EVERY round and warp W-1 the shortest, a telescoping imbalance of
// (longest - shortest task) ~ 60 non-zeros against a mean of ~115 per warp at config 2.  Reversing the direction in
// odd rounds (boustrophedon) pairs each long task with a short one: per-warp totals agree to within one task.
__device__ __forceinline__ int snake_task(int r, int w, int W) { return r * W + ((r & 1) ? (W - 1 - w) : w); }

template <typename T, int VEC, int K, bool PEER, int NB>
__device__ __forceinline__ void gather_batch(float (&acc)[NB][VEC], int myc, float myv, int j, const char* __restrict__ Xc,
                                             unsigned ldxb, uint64_t xbsb, int c0, const PeerTable& peers) {
  unsigned cj[K];
  float wv[K];
  float xv[K][NB][VEC];
#pragma unroll
  for (int u = 0; u < K; ++u) {
    cj[u] = static_cast<unsigned>(__shfl_sync(0xffffffffu, myc, j + u));
    wv[u] = __shfl_sync(0xffffffffu, myv, j + u);
  }
#pragma unroll
  for (int u = 0; u < K; ++u) {
    const char* src = PEER ? static_cast<const char*>(peers.p[cj[u] >> 28]) + c0 * sizeof(T) +
                                 static_cast<uint64_t>(cj[u] & 0x0FFFFFFFu) * ldxb
                           : Xc + static_cast<uint64_t>(cj[u]) * ldxb;
#pragma unroll
    for (int b = 0; b < NB; ++b) VecIO<T, VEC>::load(reinterpret_cast<const T*>(src + b * xbsb), xv[u][b]);
  }
#pragma unroll
  for (int u = 0; u < K; ++u)
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[b][v] = fmaf(wv[u], xv[u][b][v], acc[b][v]);
}

// Ordered sum of the segment partials of a split long row (the last segment to arrive reduces): same order as a plain
// loop -- bit-identical -- but the loads of eight segments are in flight together.  The biggest hub of the config-2
// graph has > 100 segments; summed four at a time with scalar loads this one warp was the tail of the launch
// (profiles/r2_d_structure/segment_length_ab.log: 50.7 us with long rows vs 41.3 us without).
template <int VEC>
__device__ __forceinline__ void sum_partials(float (&acc)[VEC], const float* __restrict__ base, int nseg, int64_t stride) {
  static_assert(VEC % 4 == 0, "partials are read 16 bytes at a time");
  int sgi = 0;
  for (; sgi + 8 <= nseg; sgi += 8) {
    float4 t[8][VEC / 4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q)
        t[u][q] = __ldcg(reinterpret_cast<const float4*>(base + static_cast<int64_t>(sgi + u) * stride) + q);
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q) {
        acc[4 * q] += t[u][q].x; acc[4 * q + 1] += t[u][q].y; acc[4 * q + 2] += t[u][q].z; acc[4 * q + 3] += t[u][q].w;
      }
  }
  for (; sgi < nseg; ++sgi)
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(base + static_cast<int64_t>(sgi) * stride) + q);
      acc[4 * q] += t.x; acc[4 * q + 1] += t.y; acc[4 * q + 2] += t.z; acc[4 * q + 3] += t.w;
    }
}

// NB > 1: NB feature matrices (xbsb / ybsb bytes apart, `nrows` rows each) share the graph -- one index read and one
// dependency chain per row for NB gathers (the two bottleneck evaluations of a domain, models/a2gnn.py).
// PUSH (partitioned graphs with little locality, dist.py): the output row is stored into the row block of this rank
// inside EVERY rank's gather buffer (outs.p[q], P2P stores over NVLink), so that the next propagation step gathers
// from local memory only -- the exchange is fused into the producer and overlaps its gathers row by row, and it moves
// each row once per peer in bulk-friendly 512-byte stores instead of once per referencing non-zero as a remote load.
// HALO (partitions with locality, dist.py): the row is stored locally AND into the halo slot of every rank that references
// it (halo_mask[row] = bit set of those ranks, halo_slot[row * nout + q] = its row in rank q's buffer): the halo exchange
// is fused into the producer, so the NVLink stores overlap the gathers of the rows still in flight instead of running
// as a separate pass between two barriers (measured at 2 GPUs, 25 MB of halo rows per step: separate push kernel
// 8.4 ms per training step, peer gathers 7.7 ms).
template <typename T, int VEC, int U, int CTAS, bool EPI, bool PEER, int NB, bool PUSH = false, bool HALO = false>
__global__ void __launch_bounds__(GDA_ROWS_BLOCK, CTAS)
k_spmm_tasks(const uint2* __restrict__ tasks, int num_tasks, const int* __restrict__ colidx,
             const float* __restrict__ vals, const int* __restrict__ long_rows, const int* __restrict__ long_seg_ptr,
             const int* __restrict__ seg_long, int* __restrict__ counters,
             const T* __restrict__ X, unsigned ldxb, T* __restrict__ Y, unsigned ldyb, int H,
             Epilogue epi, float* __restrict__ partial, PeerTable peers, uint64_t xbsb, uint64_t ybsb, int nrows,
             PeerTable outs = PeerTable{}, int nout = 0, const int* __restrict__ halo_mask = nullptr,
             const int* __restrict__ halo_slot = nullptr) {
  static_assert(NB == 1 || !PEER, "batched form is local only");
  static_assert(!PUSH || (PEER && NB == 1), "push mode is the partitioned single-matrix form");
  static_assert(!HALO || (!PEER && !PUSH && NB == 1), "halo mode runs on the local rectangular block");
  static_assert(U == 4 || U == 8, "batch depth");
  const int lane = threadIdx.x & 31;
  const int c0 = lane * VEC;
  const char* __restrict__ Xc = reinterpret_cast<const char*>(X + c0);
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int r = 0;
  if (wid >= num_tasks) return;

  uint2 cur = __ldg(tasks + wid);
  int inext = snake_task(1, wid, nwarps);
  bool has_next = inext < num_tasks;
  uint2 nxt = make_uint2(0u, 0u);
  if (has_next) nxt = __ldg(tasks + inext);
  int myc = 0;
  float myv = 0.f;
  if (lane <= static_cast<int>((cur.y >> 25) & 63u)) {
    myc = __ldg(colidx + cur.x + lane);
    myv = __ldg(vals + cur.x + lane);
  }

  while (true) {
    // ---- prefetch: pairs of the next task, descriptor of the one after ----
    int nc = 0;
    float nv = 0.f;
    uint2 nn = make_uint2(0u, 0u);
    const int inn = snake_task(r + 2, wid, nwarps);
    const bool has_nn = has_next && inn < num_tasks;
    if (has_next && lane <= static_cast<int>((nxt.y >> 25) & 63u)) {
      nc = __ldg(colidx + nxt.x + lane);
      nv = __ldg(vals + nxt.x + lane);
    }
    if (has_nn) nn = __ldg(tasks + inn);

    // ---- the task: len non-zeros starting at cur.x, first min(len, 32) pairs staged in (myc, myv) ----
    const int len = static_cast<int>((cur.y >> 25) & 63u) + 1;
    int hmask = 0;                                       // which ranks reference this row (issued now, used after the gathers)
    if (HALO && !(cur.y & 0x80000000u)) hmask = __ldg(halo_mask + (cur.y & 0x01FFFFFFu));
    float acc[NB][VEC];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[b][v] = 0.f;
    int base = 0;
    while (true) {
      const int n = min(len - base, 32);
      int j = 0;
      for (; j + U <= n; j += U) gather_batch<T, VEC, U, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers);
      switch (n - j) {                                   // warp-uniform: exact tails, no padding gathers
        case 1: gather_batch<T, VEC, 1, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 2: gather_batch<T, VEC, 2, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 3: gather_batch<T, VEC, 3, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 4: if (U > 4) gather_batch<T, VEC, 4, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 5: if (U > 4) gather_batch<T, VEC, 5, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 6: if (U > 4) gather_batch<T, VEC, 6, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        case 7: if (U > 4) gather_batch<T, VEC, 7, PEER, NB>(acc, myc, myv, j, Xc, ldxb, xbsb, c0, peers); break;
        default: break;
      }
      base += 32;
      if (base >= len) break;
      myc = 0; myv = 0.f;                                // tasks of 33..64 non-zeros: second chunk
      if (base + lane < len) {
        myc = __ldg(colidx + cur.x + base + lane);
        myv = __ldg(vals + cur.x + base + lane);
      }
    }

    if (!(cur.y & 0x80000000u)) {
      const unsigned row = cur.y & 0x01FFFFFFu;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (EPI) apply_epilogue<VEC>(acc[b], epi, static_cast<int64_t>(b) * nrows + row, c0, H);
        if (PUSH) {
          for (int q = 0; q < nout; ++q)
            VecIO<T, VEC>::store(reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(T) +
                                                      static_cast<uint64_t>(row) * ldyb), acc[b]);
        } else {
          VecIO<T, VEC>::store(reinterpret_cast<T*>(reinterpret_cast<char*>(Y + c0) + b * ybsb + static_cast<uint64_t>(row) * ldyb), acc[b]);
          if (HALO && hmask) {
            for (int q = 0; q < nout; ++q)
              if ((hmask >> q) & 1)
                VecIO<T, VEC>::store(reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(T) +
                                                          static_cast<uint64_t>(__ldg(halo_slot + row * nout + q)) * ldyb), acc[b]);
          }
        }
      }
    } else {                                             // segment of a long row: ordered reduction by the last arrival
      const int sgid = static_cast<int>(cur.y & 0x01FFFFFFu);
      const int L = __ldg(seg_long + sgid);
      float* dst = partial + static_cast<int64_t>(sgid) * (NB * H) + c0;
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int v = 0; v < VEC; ++v) __stcg(dst + b * H + v, acc[b][v]);
      __threadfence();
      __syncwarp();
      int old = 0;
      const int first = __ldg(long_seg_ptr + L), nseg = __ldg(long_seg_ptr + L + 1) - first;
      if (lane == 0) old = atomicAdd(counters + L, 1);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == nseg - 1) {
        __threadfence();
        const int row = __ldg(long_rows + L);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[b][v] = 0.f;
          if (VEC % 4 == 0) {
            sum_partials<(VEC % 4 == 0 ? VEC : 4)>(reinterpret_cast<float(&)[(VEC % 4 == 0 ? VEC : 4)]>(acc[b]),
                                                   partial + static_cast<int64_t>(first) * (NB * H) + b * H + c0, nseg, NB * H);
          } else {
            for (int sgi = 0; sgi < nseg; ++sgi) {
              const float* srcp = partial + static_cast<int64_t>(first + sgi) * (NB * H) + b * H + c0;
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[b][v] += __ldcg(srcp + v);
            }
          }
          if (EPI) apply_epilogue<VEC>(acc[b], epi, static_cast<int64_t>(b) * nrows + row, c0, H);
          if (PUSH) {
            for (int q = 0; q < nout; ++q)
              VecIO<T, VEC>::store(reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(T) +
                                                        static_cast<uint64_t>(row) * ldyb), acc[b]);
          } else {
            VecIO<T, VEC>::store(reinterpret_cast<T*>(reinterpret_cast<char*>(Y + c0) + b * ybsb + static_cast<uint64_t>(row) * ldyb), acc[b]);
            if (HALO) {
              const int hm = __ldg(halo_mask + row);
              for (int q = 0; q < nout; ++q)
                if ((hm >> q) & 1)
                  VecIO<T, VEC>::store(reinterpret_cast<T*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(T) +
                                                            static_cast<uint64_t>(__ldg(halo_slot + row * nout + q)) * ldyb), acc[b]);
            }
          }
        }
        if (lane == 0) counters[L] = 0;
      }
    }

    if (!has_next) break;
    cur = nxt; myc = nc; myv = nv;
    nxt = nn; has_next = has_nn;
    ++r;
  }
}

// ------------------------------------------------------------------------------------------
// Factored work-list kernel for UNIT-weight symmetric normalisation -- every GCN-normalised graph built from a plain
// edge_index (gcn_norm with edge_weight=None, prop_gcn_conv.py:67-81): w_e = dinv[src] * 1 * dinv[dst], so
//     A_hat = D S D,   S = A + I (0/1 entries, multi-edges repeated),  D = diag(dinv)
// and a chain  A_hat^k x = D S (D^2 S)^(k-1) D x :  one row scaling of the input, then k steps of
//     out[r] = scale[r] * sum_{c in row r} z[c],    scale = dinv^2 (intermediate steps) or dinv (last step).
// What this removes from the inner loop (profiles/probes/gather_probe6, B200): the per-edge weight -- its 4 bytes,
// its shuffle / load, the FFMA dependence on it (probe: +10 us of 51 at 40 warps/SM) -- and the lane-coalesced
// (colidx, weight) staging with two shuffles per non-zero (+5 us); indices are read with warp-uniform loads (one L1
// transaction, broadcast), the kernel keeps 4 + 16 + 4 live registers of payload and fits 16 CTAs (64 warps) per SM,
// the occupancy at which the regular gather probe peaks (41 us vs 51 us at 40 warps/SM).
// Summation order inside a row is unchanged (CSR = COO order, segments reduced in segment order); values differ from
// the weighted kernel by fp32 rounding only (scale applied once per row instead of once per edge).
// One batch: K row gathers issued from the indices in cj[], then -- while they are in flight -- the NEXT batch's four
// indices are fetched into the same registers (warp-uniform loads from `pn`; entries past the end of that task are
// never used), then the K rows are summed in order.  The ncu source view of the first version of this kernel put 15 %
// of the stall samples on the address computation that waits for the index load: a second serial L2 latency per batch.
template <int K, int NB>
__device__ __forceinline__ void gather_sum(float (&acc)[NB][4], unsigned (&cj)[4], const int* __restrict__ pn,
                                           const char* __restrict__ Xc, unsigned ldxb, uint64_t xbsb) {
  uint4 xv[K][NB];
#pragma unroll
  for (int u = 0; u < K; ++u) {
    const char* src = Xc + static_cast<uint64_t>(cj[u]) * ldxb;
#pragma unroll
    for (int b = 0; b < NB; ++b) xv[u][b] = gather16(src + b * xbsb);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) cj[u] = static_cast<unsigned>(__ldg(pn + u));
#pragma unroll
  for (int u = 0; u < K; ++u)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      acc[b][0] += __uint_as_float(xv[u][b].x); acc[b][1] += __uint_as_float(xv[u][b].y);
      acc[b][2] += __uint_as_float(xv[u][b].z); acc[b][3] += __uint_as_float(xv[u][b].w);
    }
}

// HALO (rectangular block of a partition, dist.py): the finished row is also stored into the halo slot of every rank
// that references it (halo_mask / halo_slot as in k_spmm_tasks<HALO>); matrix b of the stack lives hs.b[q] bytes into
// rank q's buffer.
struct HaloStride { uint64_t b[GDA_MAX_PEERS]; };

template <int U, int CTAS, bool EPI, int NB, bool HALO = false>
__global__ void __launch_bounds__(GDA_ROWS_BLOCK, CTAS)
k_spmm_unw(const uint2* __restrict__ tasks, int num_tasks, const int* __restrict__ colidx,
           const float* __restrict__ dinv, int last, const int* __restrict__ long_rows,
           const int* __restrict__ long_seg_ptr, const int* __restrict__ seg_long, int* __restrict__ counters,
           const float* __restrict__ X, unsigned ldxb, float* __restrict__ Y, unsigned ldyb, int H,
           Epilogue epi, float* __restrict__ partial, uint64_t xbsb, uint64_t ybsb, int nrows,
           const int* __restrict__ halo_mask = nullptr, const int* __restrict__ halo_slot = nullptr,
           PeerTable outs = PeerTable{}, int nout = 0, HaloStride hs = HaloStride{}) {
  static_assert(U == 4, "batches of four");
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 4;
  const char* __restrict__ Xc = reinterpret_cast<const char*>(X + c0);
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= num_tasks) return;
  uint2 cur = __ldg(tasks + wid);
  int inext = snake_task(1, wid, nwarps);
  bool has_next = inext < num_tasks;
  uint2 nxt = has_next ? __ldg(tasks + inext) : cur;
  unsigned cj[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) cj[u] = static_cast<unsigned>(__ldg(colidx + cur.x + u));
  for (int r = 0;; ++r) {
    // descriptor of the task after the next one: by the time the next task starts, its successor is known and the
    // first indices of that successor can be prefetched in its last batch
    const int inn = snake_task(r + 2, wid, nwarps);
    const bool has_nn = has_next && inn < num_tasks;
    uint2 nn = nxt;
    if (has_nn) nn = __ldg(tasks + inn);

    const int* __restrict__ ci = colidx + cur.x;
    const int* __restrict__ cn = colidx + nxt.x;          // first indices of the next task (== cur when none: unused)
    int left = static_cast<int>((cur.y >> 25) & 63u) + 1;
    float acc[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[b][v] = 0.f;
    for (; left > 4; left -= 4) { ci += 4; gather_sum<4, NB>(acc, cj, ci, Xc, ldxb, xbsb); }
    switch (left) {                                      // last batch of the task (warp-uniform): exact, no padding gathers
      case 1: gather_sum<1, NB>(acc, cj, cn, Xc, ldxb, xbsb); break;
      case 2: gather_sum<2, NB>(acc, cj, cn, Xc, ldxb, xbsb); break;
      case 3: gather_sum<3, NB>(acc, cj, cn, Xc, ldxb, xbsb); break;
      default: gather_sum<4, NB>(acc, cj, cn, Xc, ldxb, xbsb); break;
    }

    if (!(cur.y & 0x80000000u)) {
      const unsigned row = cur.y & 0x01FFFFFFu;
      const float d = __ldg(dinv + row);
      const float sc = last ? d : d * d;
      const int hm = HALO ? __ldg(halo_mask + row) : 0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[b][v] *= sc;
        if (EPI) apply_epilogue<4>(acc[b], epi, static_cast<int64_t>(b) * nrows + row, c0, H);
        VecIO<float, 4>::store(reinterpret_cast<float*>(reinterpret_cast<char*>(Y + c0) + b * ybsb + static_cast<uint64_t>(row) * ldyb), acc[b]);
        if (HALO && hm) {
          for (int q = 0; q < nout; ++q)
            if ((hm >> q) & 1)
              VecIO<float, 4>::store(reinterpret_cast<float*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(float) +
                                                              b * hs.b[q] + static_cast<uint64_t>(__ldg(halo_slot + row * nout + q)) * ldyb), acc[b]);
        }
      }
    } else {                                             // segment of a long row: ordered reduction by the last arrival
      const int sgid = static_cast<int>(cur.y & 0x01FFFFFFu);
      const int L = __ldg(seg_long + sgid);
      float* dst = partial + static_cast<int64_t>(sgid) * (NB * H) + c0;
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int v = 0; v < 4; ++v) __stcg(dst + b * H + v, acc[b][v]);
      __threadfence();
      __syncwarp();
      int old = 0;
      const int first = __ldg(long_seg_ptr + L), nseg = __ldg(long_seg_ptr + L + 1) - first;
      if (lane == 0) old = atomicAdd(counters + L, 1);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == nseg - 1) {
        __threadfence();
        const int row = __ldg(long_rows + L);
        const float d = __ldg(dinv + row);
        const float sc = last ? d : d * d;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[b][v] = 0.f;
          sum_partials<4>(acc[b], partial + static_cast<int64_t>(first) * (NB * H) + b * H + c0, nseg, NB * H);
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[b][v] *= sc;
          if (EPI) apply_epilogue<4>(acc[b], epi, static_cast<int64_t>(b) * nrows + row, c0, H);
          VecIO<float, 4>::store(reinterpret_cast<float*>(reinterpret_cast<char*>(Y + c0) + b * ybsb + static_cast<uint64_t>(row) * ldyb), acc[b]);
          if (HALO) {
            const int hm = __ldg(halo_mask + row);
            for (int q = 0; q < nout; ++q)
              if ((hm >> q) & 1)
                VecIO<float, 4>::store(reinterpret_cast<float*>(static_cast<char*>(const_cast<void*>(outs.p[q])) + c0 * sizeof(float) +
                                                                b * hs.b[q] + static_cast<uint64_t>(__ldg(halo_slot + row * nout + q)) * ldyb), acc[b]);
          }
        }
        if (lane == 0) counters[L] = 0;
      }
    }
    if (!has_next) break;
    cur = nxt;
    nxt = nn;
    has_next = has_nn;
  }
}

// z[b][r, :] = dinv[r] * x[b][r, :]   (the D x of the factored chain; one float4 per thread)
__global__ void __launch_bounds__(256)
k_row_scale(const float* __restrict__ X, int64_t ldx, int64_t xbs, float* __restrict__ Z, int64_t ldz, int64_t zbs,
            const float* __restrict__ dinv, int64_t N, int H4, int nb) {
  const int64_t total = static_cast<int64_t>(nb) * N * H4;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(t % H4);
    const int64_t rr = t / H4;
    const int64_t r = rr % N, b = rr / N;
    const float d = __ldg(dinv + r);
    float4 v = __ldg(reinterpret_cast<const float4*>(X + b * xbs + r * ldx) + c);
    v.x *= d; v.y *= d; v.z *= d; v.w *= d;
    reinterpret_cast<float4*>(Z + b * zbs + r * ldz)[c] = v;
  }
}

inline int unw_mode() {         // experiments: GDA_SPMM_UNW=0 disables the factored chain, =12 / =16 pick CTAs per SM
  static const int m = [] { const char* e = std::getenv("GDA_SPMM_UNW"); return e ? std::atoi(e) : 12; }();
  return m;
}

inline int tasks_mode() {       // experiments: GDA_SPMM_TASKS=0 keeps k_spmm_rows, =8 the 8-deep variant, =12 12 CTAs/SM
  static const int m = [] { const char* e = std::getenv("GDA_SPMM_TASKS"); return e ? std::atoi(e) : 4; }();
  return m;
}

template <typename T, int VEC, int LPR, int U>
int launch(const Csr& c, int seg, const T* X, int64_t ldx, T* Y, int64_t ldy, int64_t N, int H,
           const Epilogue& epi, float* partial, cudaStream_t st, const PeerTable* peers = nullptr, int pad_col = 0,
           int nb = 1, int64_t xbs = 0, int64_t ybs = 0) {
  constexpr int RPG = RowsPerGroup<LPR>::value;
  const int64_t groups = static_cast<int64_t>(c.num_segs) + ceil_div(N, RPG);
  const bool generic = c.may_have_empty_rows || generic_forced();
  const int kBlock = generic ? 256 : GDA_SPMM_BLOCK;
  const int groups_per_block = (kBlock / 32) * (32 / LPR);
  if (groups == 0) return GDA_OK;
  const unsigned grid = static_cast<unsigned>(ceil_div(groups, groups_per_block));
  constexpr int UF = U <= LPR ? U : LPR;               // batches must tile a chunk of LPR entries
  const bool has_epi = epi.bias != nullptr || epi.flags != 0;
  if (LPR == 32 && H == 32 * VEC && VEC * sizeof(T) == 16 && !generic && !fast_forced()) {
    // one 16-byte slice per lane: row-per-warp kernel, grid-stride over rows
    const unsigned ldxb = static_cast<unsigned>(ldx * sizeof(T)), ldyb = static_cast<unsigned>(ldy * sizeof(T));
    const int64_t items = static_cast<int64_t>(c.num_segs) + N;
    int64_t blocks = ceil_div(items, GDA_ROWS_BLOCK / 32);
    const int64_t cap = static_cast<int64_t>(kNumSMs) * GDA_ROWS_MIN_CTAS;
    if (blocks > cap) blocks = cap;
    const PeerTable pt = peers ? *peers : PeerTable{};
    const int tm = tasks_mode();
    if (c.tasks != nullptr && c.num_tasks > 0 && tm != 0) {
      const int per_sm = tm == 8 ? GDA_TASKS_MIN_CTAS_WIDE : (tm == 12 ? 12 : GDA_TASKS_MIN_CTAS);
      int64_t tb = ceil_div(c.num_tasks, GDA_ROWS_BLOCK / 32);
      if (tb > static_cast<int64_t>(kNumSMs) * per_sm) tb = static_cast<int64_t>(kNumSMs) * per_sm;
      if (nb == 2) {                                     // two matrices per index read (never on the peer path)
        int64_t tb2 = ceil_div(c.num_tasks, GDA_ROWS_BLOCK / 32);
        if (tb2 > static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS_WIDE) tb2 = static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS_WIDE;
        const uint64_t xb = static_cast<uint64_t>(xbs) * sizeof(T), yb = static_cast<uint64_t>(ybs) * sizeof(T);
        if (has_epi)
          k_spmm_tasks<T, VEC, 4, GDA_TASKS_MIN_CTAS_WIDE, true, false, 2><<<static_cast<unsigned>(tb2), GDA_ROWS_BLOCK, 0, st>>>(
              c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y,
              ldyb, H, epi, partial, pt, xb, yb, static_cast<int>(N));
        else
          k_spmm_tasks<T, VEC, 4, GDA_TASKS_MIN_CTAS_WIDE, false, false, 2><<<static_cast<unsigned>(tb2), GDA_ROWS_BLOCK, 0, st>>>(
              c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y,
              ldyb, H, epi, partial, pt, xb, yb, static_cast<int>(N));
        GDA_LAUNCH_CHECK();
        return GDA_OK;
      }
#define GDA_TASKS_LAUNCH(UU, CC, E, P)                                                                     \
      k_spmm_tasks<T, VEC, UU, CC, E, P, 1><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(         \
          c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters,      \
          X, ldxb, Y, ldyb, H, epi, partial, pt, 0, 0, static_cast<int>(N))
#define GDA_TASKS_LAUNCH_EP(UU, CC)                                                                        \
      do {                                                                                                 \
        if (peers) { if (has_epi) GDA_TASKS_LAUNCH(UU, CC, true, true); else GDA_TASKS_LAUNCH(UU, CC, false, true); } \
        else { if (has_epi) GDA_TASKS_LAUNCH(UU, CC, true, false); else GDA_TASKS_LAUNCH(UU, CC, false, false); }   \
      } while (0)
      if (tm == 8) GDA_TASKS_LAUNCH_EP(8, GDA_TASKS_MIN_CTAS_WIDE);
      else if (tm == 12) GDA_TASKS_LAUNCH_EP(4, 12);
      else GDA_TASKS_LAUNCH_EP(4, GDA_TASKS_MIN_CTAS);
#undef GDA_TASKS_LAUNCH_EP
#undef GDA_TASKS_LAUNCH
      GDA_LAUNCH_CHECK();
      return GDA_OK;
    }
#define GDA_ROWS_LAUNCH(E, P)                                                                              \
    k_spmm_rows<T, VEC, E, P><<<static_cast<unsigned>(blocks), GDA_ROWS_BLOCK, 0, st>>>(                  \
        c.rowptr, c.colidx, c.vals, c.num_segs, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, seg, \
        X, ldxb, Y, ldyb, static_cast<int>(N), H, epi, partial, pt, pad_col)
    if (peers) { if (has_epi) GDA_ROWS_LAUNCH(true, true); else GDA_ROWS_LAUNCH(false, true); }
    else { if (has_epi) GDA_ROWS_LAUNCH(true, false); else GDA_ROWS_LAUNCH(false, false); }
#undef GDA_ROWS_LAUNCH
    GDA_LAUNCH_CHECK();
    return GDA_OK;
  }
  if (generic) {
    k_spmm<T, VEC, LPR, U><<<grid, kBlock, 0, st>>>(c.rowptr, c.colidx, c.vals, c.num_segs, c.long_rows,
                                                   c.long_seg_ptr, c.seg_long, c.counters, seg, X, ldx, Y, ldy,
                                                   static_cast<int>(N), H, epi, partial);
  } else if (peers) {
    if (has_epi)
      k_spmm_fast<T, VEC, LPR, UF, true, true><<<grid, kBlock, 0, st>>>(c.rowptr, c.colidx, c.vals, c.num_segs,
          c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, seg, X, ldx, Y, ldy, static_cast<int>(N), H, epi, partial,
          *peers, pad_col);
    else
      k_spmm_fast<T, VEC, LPR, UF, false, true><<<grid, kBlock, 0, st>>>(c.rowptr, c.colidx, c.vals, c.num_segs,
          c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, seg, X, ldx, Y, ldy, static_cast<int>(N), H, epi, partial,
          *peers, pad_col);
  } else if (has_epi) {
    k_spmm_fast<T, VEC, LPR, UF, true, false><<<grid, kBlock, 0, st>>>(c.rowptr, c.colidx, c.vals, c.num_segs,
        c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, seg, X, ldx, Y, ldy, static_cast<int>(N), H, epi, partial,
        PeerTable{}, 0);
  } else {
    k_spmm_fast<T, VEC, LPR, UF, false, false><<<grid, kBlock, 0, st>>>(c.rowptr, c.colidx, c.vals, c.num_segs,
        c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, seg, X, ldx, Y, ldy, static_cast<int>(N), H, epi, partial,
        PeerTable{}, 0);
  }
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

inline int pow2_at_least(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <typename T, int VEC>
int dispatch_lpr(const Csr& c, int seg, const T* X, int64_t ldx, T* Y, int64_t ldy, int64_t N, int H,
                 const Epilogue& epi, float* partial, cudaStream_t st, const PeerTable* peers = nullptr, int pad_col = 0,
                 int nb = 1, int64_t xbs = 0, int64_t ybs = 0) {
  int lanes = pow2_at_least(static_cast<int>(ceil_div(H, VEC)));
  if (lanes > 32) lanes = 32;
  if (lanes < 4) lanes = 4;
  constexpr int U = (VEC == 4) ? GDA_SPMM_U4 : 4;               // 8 x float4 or 4 x (8 bf16) gathers in flight per lane
  switch (lanes) {
    case 4:  return launch<T, VEC, 4, U>(c, seg, X, ldx, Y, ldy, N, H, epi, partial, st, peers, pad_col);
    case 8:  return launch<T, VEC, 8, U>(c, seg, X, ldx, Y, ldy, N, H, epi, partial, st, peers, pad_col);
    case 16: return launch<T, VEC, 16, U>(c, seg, X, ldx, Y, ldy, N, H, epi, partial, st, peers, pad_col);
    default: return launch<T, VEC, 32, U>(c, seg, X, ldx, Y, ldy, N, H, epi, partial, st, peers, pad_col, nb, xbs, ybs);
  }
}

// true when launch<> will take the work-list kernel (the only one with a batched form)
template <typename T, int WIDE>
bool tasks_path(const Csr& c, int H, bool wide_ok) {
  return wide_ok && H == 32 * WIDE && WIDE * sizeof(T) == 16 && !c.may_have_empty_rows && !generic_forced() &&
         !fast_forced() && c.tasks != nullptr && c.num_tasks > 0 && tasks_mode() != 0;
}

// nb feature matrices X + b*xbs -> Y + b*ybs (elements) over the same graph.  The dropout mask of matrix b is the one of
// rows [b*N, (b+1)*N) of a stacked [nb*N, H] matrix.
template <typename T, int WIDE>
int spmm_any(const gda_graph* g, int transpose, const T* X, int64_t ldx, T* Y, int64_t ldy, int H,
             const float* bias, int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
             void* workspace, int64_t workspace_bytes, cudaStream_t st, const PeerTable* peers = nullptr,
             int my_rank = 0, int nb = 1, int64_t xbs = 0, int64_t ybs = 0) {
  GDA_REQUIRE(g != nullptr, "gda_spmm: graph is NULL");
  GDA_REQUIRE(g->peer_packed == (peers != nullptr), "gda_spmm: partitioned graphs need gda_spmm_peer_* (and only they)");
  GDA_REQUIRE(H > 0, "gda_spmm: H must be positive");
  GDA_REQUIRE(nb >= 1 && (nb == 1 || !peers), "gda_spmm: nb must be >= 1 (and 1 on the peer path)");
  if (g->N == 0) return GDA_OK;
  GDA_REQUIRE((X || peers) && Y, "gda_spmm: NULL feature pointer");
  GDA_REQUIRE(X != Y, "gda_spmm: X and Y must not alias");
  GDA_REQUIRE(ldx >= H && ldy >= H, "gda_spmm: leading dimension smaller than H");
  GDA_REQUIRE(g->N * ldx < (int64_t(1) << 32) && g->N * ldy < (int64_t(1) << 32),
              "gda_spmm: N * ld must be below 2^32 elements");
  GDA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gda_spmm: dropout_p outside [0,1)");
  GDA_REQUIRE(nb == 1 || (xbs >= g->N * ldx && ybs >= g->N * ldy), "gda_spmm: batch strides smaller than one matrix");
  const Csr& c = transpose ? g->csr_t : g->csr;
  Epilogue epi;
  epi.bias = bias;
  epi.flags = epi_flags;
  epi.thresh = dropout_threshold(dropout_p);
  epi.scale = 1.0f / (1.0f - dropout_p);
  epi.seed = seed;
  epi.seed_offset = seed_offset;
  float* partial = static_cast<float*>(workspace);
  bool wide_ok = (H % WIDE == 0) && (ldx % WIDE == 0) && (ldy % WIDE == 0) &&
                 (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (reinterpret_cast<uintptr_t>(Y) % 16 == 0);
  if (nb > 1) wide_ok = wide_ok && (xbs % WIDE == 0) && (ybs % WIDE == 0);
  if (c.skip_empty_rows)          // only the work-list kernel leaves rows without non-zeros untouched
    GDA_REQUIRE((tasks_path<T, WIDE>(c, H, wide_ok)), "gda_spmm: a GDA_SKIP_EMPTY_ROWS graph needs the H = 128 fp32 / 256 bf16 path");
  const int pad_col = peers ? (my_rank << 28) : 0;
  if (peers)
    for (int i = 0; i < GDA_MAX_PEERS; ++i) wide_ok = wide_ok && (reinterpret_cast<uintptr_t>(peers->p[i]) % 16 == 0);
  // pairs of matrices go through the batched work-list kernel; anything else one matrix at a time
  int b = 0;
  if (nb >= 2 && tasks_path<T, WIDE>(c, H, wide_ok)) {
    const int64_t need2 = static_cast<int64_t>(c.num_segs) * H * 2 * sizeof(float);
    if (need2 > 0 && (workspace == nullptr || workspace_bytes < need2))
      return fail(GDA_E_WORKSPACE, "gda_spmm: workspace smaller than gda_spmm_workspace_bytes(H * nb)");
    for (; b + 2 <= nb; b += 2) {
      Epilogue e2 = epi;
      e2.seed = seed + static_cast<uint64_t>(b) * static_cast<uint64_t>(g->N) * static_cast<uint64_t>(H) * 0x9E3779B97F4A7C15ull;
      int rc = dispatch_lpr<T, WIDE>(c, g->seg, X + b * xbs, ldx, Y + b * ybs, ldy, g->N, H, e2, partial, st, nullptr, 0,
                                     2, xbs, ybs);
      if (rc) return rc;
    }
  }
  const int64_t need = static_cast<int64_t>(c.num_segs) * H * sizeof(float);
  if (b < nb && need > 0 && (workspace == nullptr || workspace_bytes < need))
    return fail(GDA_E_WORKSPACE, "gda_spmm: workspace smaller than gda_spmm_workspace_bytes()");
  for (; b < nb; ++b) {
    Epilogue e1 = epi;        // mask index (b*N + row)*H + col == hash seed shifted by b*N*H golden-ratio steps (common.cuh)
    e1.seed = seed + static_cast<uint64_t>(b) * static_cast<uint64_t>(g->N) * static_cast<uint64_t>(H) * 0x9E3779B97F4A7C15ull;
    const T* Xb = X ? X + b * xbs : X;
    T* Yb = Y + b * ybs;
    int rc = wide_ok ? dispatch_lpr<T, WIDE>(c, g->seg, Xb, ldx, Yb, ldy, g->N, H, e1, partial, st, peers, pad_col)
                     : dispatch_lpr<T, 1>(c, g->seg, Xb, ldx, Yb, ldy, g->N, H, e1, partial, st, peers, pad_col);
    if (rc) return rc;
  }
  return GDA_OK;
}

// Push-mode step over a partition (fp32, one 16-byte slice per lane): gathers from the LOCAL copy of the whole matrix
// (`gather`: block q of every rank at gather.p[q]), stores every output row into outs.p[0 .. nout-1].
int spmm_push(const gda_graph* g, int transpose, const PeerTable& gather, const PeerTable& outs, int nout, int64_t ldx,
              int64_t ldy, int H, const Epilogue& epi, float* partial, int64_t workspace_bytes, cudaStream_t st) {
  const Csr& c = transpose ? g->csr_t : g->csr;
  bool wide_ok = H == 128 && ldx % 4 == 0 && ldy % 4 == 0;
  for (int i = 0; i < GDA_MAX_PEERS; ++i) wide_ok = wide_ok && (reinterpret_cast<uintptr_t>(gather.p[i]) % 16 == 0);
  for (int i = 0; i < nout; ++i) wide_ok = wide_ok && (reinterpret_cast<uintptr_t>(outs.p[i]) % 16 == 0);
  const bool on_tasks_path = tasks_path<float, 4>(c, H, wide_ok);
  GDA_REQUIRE(on_tasks_path, "gda_spmm_push: needs H = 128 fp32, aligned buffers and a task list");
  GDA_REQUIRE(g->rows_per_rank * ldx < (int64_t(1) << 32) && g->N * ldy < (int64_t(1) << 32), "gda_spmm_push: block too large");
  const int64_t need = static_cast<int64_t>(c.num_segs) * H * sizeof(float);
  if (need > 0 && (partial == nullptr || workspace_bytes < need))
    return fail(GDA_E_WORKSPACE, "gda_spmm_push: workspace smaller than gda_spmm_workspace_bytes()");
  if (g->N == 0) return GDA_OK;
  const unsigned ldxb = static_cast<unsigned>(ldx * sizeof(float)), ldyb = static_cast<unsigned>(ldy * sizeof(float));
  int64_t tb = ceil_div(c.num_tasks, GDA_ROWS_BLOCK / 32);
  if (tb > static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS) tb = static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS;
  const bool has_epi = epi.bias != nullptr || epi.flags != 0;
  const float* X = static_cast<const float*>(gather.p[0]);
  float* Y = static_cast<float*>(const_cast<void*>(outs.p[0]));
  if (has_epi)
    k_spmm_tasks<float, 4, 4, GDA_TASKS_MIN_CTAS, true, true, 1, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
        c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y, ldyb, H,
        epi, partial, gather, 0, 0, static_cast<int>(g->N), outs, nout);
  else
    k_spmm_tasks<float, 4, 4, GDA_TASKS_MIN_CTAS, false, true, 1, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
        c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y, ldyb, H,
        epi, partial, gather, 0, 0, static_cast<int>(g->N), outs, nout);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

// Halo-mode step on a local rectangular block (GDA_SKIP_EMPTY_ROWS graph): ordinary local gathers, every row stored to Y
// and to the halo slots of the ranks that reference it.
int spmm_halo(const gda_graph* g, const float* X, int64_t ldx, float* Y, int64_t ldy, int H, const int* halo_mask,
              const int* halo_slot, const PeerTable& outs, int nout, const Epilogue& epi, float* partial,
              int64_t workspace_bytes, cudaStream_t st) {
  const Csr& c = g->csr;
  bool wide_ok = H == 128 && ldx % 4 == 0 && ldy % 4 == 0 && reinterpret_cast<uintptr_t>(X) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(Y) % 16 == 0;
  for (int i = 0; i < nout; ++i) wide_ok = wide_ok && (reinterpret_cast<uintptr_t>(outs.p[i]) % 16 == 0);
  const bool on_tasks_path = tasks_path<float, 4>(c, H, wide_ok);
  GDA_REQUIRE(on_tasks_path, "gda_spmm_halo: needs H = 128 fp32, aligned buffers and a task list");
  GDA_REQUIRE(g->N * ldx < (int64_t(1) << 32) && g->N * ldy < (int64_t(1) << 32), "gda_spmm_halo: block too large");
  const int64_t need = static_cast<int64_t>(c.num_segs) * H * sizeof(float);
  if (need > 0 && (partial == nullptr || workspace_bytes < need))
    return fail(GDA_E_WORKSPACE, "gda_spmm_halo: workspace smaller than gda_spmm_workspace_bytes()");
  if (g->N == 0 || c.num_tasks == 0) return GDA_OK;
  const unsigned ldxb = static_cast<unsigned>(ldx * sizeof(float)), ldyb = static_cast<unsigned>(ldy * sizeof(float));
  int64_t tb = ceil_div(c.num_tasks, GDA_ROWS_BLOCK / 32);
  if (tb > static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS) tb = static_cast<int64_t>(kNumSMs) * GDA_TASKS_MIN_CTAS;
  const bool has_epi = epi.bias != nullptr || epi.flags != 0;
  if (has_epi)
    k_spmm_tasks<float, 4, 4, GDA_TASKS_MIN_CTAS, true, false, 1, false, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
        c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y, ldyb, H,
        epi, partial, PeerTable{}, 0, 0, static_cast<int>(g->N), outs, nout, halo_mask, halo_slot);
  else
    k_spmm_tasks<float, 4, 4, GDA_TASKS_MIN_CTAS, false, false, 1, false, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
        c.tasks, c.num_tasks, c.colidx, c.vals, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, X, ldxb, Y, ldyb, H,
        epi, partial, PeerTable{}, 0, 0, static_cast<int>(g->N), outs, nout, halo_mask, halo_slot);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

// true when the factored (weight-free) kernel can take a step over this graph at this width
bool unw_path(const gda_graph* g, const Csr& c, int H, int64_t ldx, int64_t ldy, const float* X, const float* Y, int nb,
              int64_t xbs, int64_t ybs) {
  bool wide_ok = (H == 128) && (ldx % 4 == 0) && (ldy % 4 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(Y) % 16 == 0);
  if (nb > 1) wide_ok = wide_ok && (xbs % 4 == 0) && (ybs % 4 == 0);
  return g->unit_weights && !g->peer_packed && g->dinv != nullptr && (nb == 1 || nb == 2) && unw_mode() != 0 &&
         tasks_path<float, 4>(c, H, wide_ok);
}

int spmm_unw(const gda_graph* g, int transpose, int nb, const float* Z, int64_t ldz, int64_t zbs, float* Y, int64_t ldy,
             int64_t ybs, int H, int last, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
             const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, cudaStream_t st,
             const int* halo_mask = nullptr, const int* halo_slot = nullptr, const PeerTable* outs = nullptr, int nout = 0,
             const HaloStride* hs = nullptr) {
  const Csr& c = transpose ? g->csr_t : g->csr;
  GDA_REQUIRE(Z && Y && Z != Y, "gda_spmm_unw: bad feature pointers");
  GDA_REQUIRE(g->N * ldz < (int64_t(1) << 32) && g->N * ldy < (int64_t(1) << 32), "gda_spmm_unw: N * ld must be below 2^32");
  GDA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gda_spmm_unw: dropout_p outside [0,1)");
  GDA_REQUIRE(unw_path(g, c, H, ldz, ldy, Z, Y, nb, zbs, ybs), "gda_spmm_unw: graph / shape not eligible (gda_graph_unit_weights)");
  const int64_t need = static_cast<int64_t>(c.num_segs) * H * nb * sizeof(float);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need))
    return fail(GDA_E_WORKSPACE, "gda_spmm_unw: workspace smaller than gda_spmm_workspace_bytes(H * nb)");
  Epilogue epi;
  epi.bias = bias; epi.flags = epi_flags; epi.thresh = dropout_threshold(dropout_p);
  epi.scale = 1.0f / (1.0f - dropout_p); epi.seed = seed; epi.seed_offset = seed_offset;
  const bool has_epi = bias != nullptr || epi_flags != 0;
  float* partial = static_cast<float*>(workspace);
  const unsigned ldzb = static_cast<unsigned>(ldz * sizeof(float)), ldyb = static_cast<unsigned>(ldy * sizeof(float));
  const uint64_t zb = static_cast<uint64_t>(zbs) * sizeof(float), yb = static_cast<uint64_t>(ybs) * sizeof(float);
  const int per_sm = nb == 2 ? 10 : (unw_mode() == 16 ? 16 : 12);
  int64_t tb = ceil_div(c.num_tasks, GDA_ROWS_BLOCK / 32);
  if (tb > static_cast<int64_t>(kNumSMs) * per_sm) tb = static_cast<int64_t>(kNumSMs) * per_sm;
#define GDA_UNW_LAUNCH(CC, E, B)                                                                              \
  k_spmm_unw<4, CC, E, B><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(                              \
      c.tasks, c.num_tasks, c.colidx, g->dinv, last, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, Z, ldzb, Y, \
      ldyb, H, epi, partial, zb, yb, static_cast<int>(g->N))
  if (halo_mask) {                                       // intermediate step of a halo-mode chain: no epilogue
    GDA_REQUIRE(!has_epi && !last && halo_slot && outs && hs && nout >= 1, "gda_spmm_unw_halo: bad halo arguments");
    if (nb == 2)
      k_spmm_unw<4, 10, false, 2, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
          c.tasks, c.num_tasks, c.colidx, g->dinv, 0, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, Z, ldzb, Y, ldyb,
          H, epi, partial, zb, yb, static_cast<int>(g->N), halo_mask, halo_slot, *outs, nout, *hs);
    else
      k_spmm_unw<4, 12, false, 1, true><<<static_cast<unsigned>(tb), GDA_ROWS_BLOCK, 0, st>>>(
          c.tasks, c.num_tasks, c.colidx, g->dinv, 0, c.long_rows, c.long_seg_ptr, c.seg_long, c.counters, Z, ldzb, Y, ldyb,
          H, epi, partial, zb, yb, static_cast<int>(g->N), halo_mask, halo_slot, *outs, nout, *hs);
    GDA_LAUNCH_CHECK();
    return GDA_OK;
  }
  if (nb == 2) { if (has_epi) GDA_UNW_LAUNCH(10, true, 2); else GDA_UNW_LAUNCH(10, false, 2); }
  else if (per_sm == 12) { if (has_epi) GDA_UNW_LAUNCH(12, true, 1); else GDA_UNW_LAUNCH(12, false, 1); }
  else { if (has_epi) GDA_UNW_LAUNCH(16, true, 1); else GDA_UNW_LAUNCH(16, false, 1); }
#undef GDA_UNW_LAUNCH
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // namespace
}  // namespace gda

extern "C" {

int gda_graph_unit_weights(const gda_graph_t* g, int transpose, int H, int nb) {
  if (!g || H != 128) return 0;
  const gda::Csr& c = transpose ? g->csr_t : g->csr;
  return gda::unw_path(g, c, H, H, H, nullptr, nullptr, nb, static_cast<int64_t>(g->N) * H,
                       static_cast<int64_t>(g->N) * H) ? 1 : 0;
}

int gda_row_scale_f32(const gda_graph_t* g, int nb, const float* X, int64_t ldx, int64_t x_batch_stride, float* Z,
                      int64_t ldz, int64_t z_batch_stride, int H, gda_stream_t stream) {
  GDA_REQUIRE(g && g->dinv, "gda_row_scale_f32: graph without a normalisation vector");
  GDA_REQUIRE(nb >= 1 && H > 0 && H % 4 == 0 && ldx % 4 == 0 && ldz % 4 == 0 && x_batch_stride % 4 == 0 &&
                  z_batch_stride % 4 == 0,
              "gda_row_scale_f32: H, leading dimensions and batch strides must be multiples of 4");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(Z) % 16 == 0,
              "gda_row_scale_f32: pointers must be 16-byte aligned");
  if (g->N == 0) return GDA_OK;
  GDA_REQUIRE(X && Z, "gda_row_scale_f32: NULL pointer");
  const int64_t total = static_cast<int64_t>(nb) * g->N * (H / 4);
  int64_t blocks = gda::ceil_div(total, 256);
  if (blocks > gda::kNumSMs * 16) blocks = gda::kNumSMs * 16;
  gda::k_row_scale<<<static_cast<unsigned>(blocks), 256, 0, gda::as_stream(stream)>>>(
      X, ldx, x_batch_stride, Z, ldz, z_batch_stride, g->dinv, g->N, H / 4, nb);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_row_scale_rows_f32(const gda_graph_t* g, int64_t rows, int nb, const float* X, int64_t ldx, int64_t x_batch_stride,
                           float* Z, int64_t ldz, int64_t z_batch_stride, int H, gda_stream_t stream) {
  GDA_REQUIRE(g && g->dinv, "gda_row_scale_rows_f32: graph without a normalisation vector");
  GDA_REQUIRE(rows >= 0 && rows <= g->N, "gda_row_scale_rows_f32: rows outside [0, N]");
  GDA_REQUIRE(nb >= 1 && H > 0 && H % 4 == 0 && ldx % 4 == 0 && ldz % 4 == 0 && x_batch_stride % 4 == 0 &&
                  z_batch_stride % 4 == 0,
              "gda_row_scale_rows_f32: H, leading dimensions and batch strides must be multiples of 4");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(Z) % 16 == 0,
              "gda_row_scale_rows_f32: pointers must be 16-byte aligned");
  if (rows == 0) return GDA_OK;
  GDA_REQUIRE(X && Z, "gda_row_scale_rows_f32: NULL pointer");
  const int64_t total = static_cast<int64_t>(nb) * rows * (H / 4);
  int64_t blocks = gda::ceil_div(total, 256);
  if (blocks > gda::kNumSMs * 16) blocks = gda::kNumSMs * 16;
  gda::k_row_scale<<<static_cast<unsigned>(blocks), 256, 0, gda::as_stream(stream)>>>(
      X, ldx, x_batch_stride, Z, ldz, z_batch_stride, g->dinv, rows, H / 4, nb);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

int gda_graph_export_dinv(const gda_graph_t* g, float* dinv_out, gda_stream_t stream) {
  GDA_REQUIRE(g && g->dinv && dinv_out, "gda_graph_export_dinv: graph without a normalisation vector, or NULL output");
  GDA_CUDA(cudaMemcpyAsync(dinv_out, g->dinv, sizeof(float) * g->N, cudaMemcpyDeviceToDevice, gda::as_stream(stream)));
  return GDA_OK;
}

int gda_graph_set_unit_dinv(gda_graph_t* g, const float* dinv, gda_stream_t stream) {
  GDA_REQUIRE(g && dinv, "gda_graph_set_unit_dinv: NULL argument");
  GDA_REQUIRE(!g->peer_packed, "gda_graph_set_unit_dinv: not for partitioned (peer-packed) graphs");
  if (g->N == 0) return GDA_OK;
  if (!g->dinv) GDA_CUDA(cudaMalloc(&g->dinv, sizeof(float) * g->N));
  GDA_CUDA(cudaMemcpyAsync(g->dinv, dinv, sizeof(float) * g->N, cudaMemcpyDeviceToDevice, gda::as_stream(stream)));
  g->unit_weights = true;
  return GDA_OK;
}

int gda_spmm_unw_nb_f32(const gda_graph_t* g, int transpose, int nb, const float* Z, int64_t ldz, int64_t z_batch_stride,
                        float* Y, int64_t ldy, int64_t y_batch_stride, int H, int last, const float* bias, int epi_flags,
                        float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                        int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(g != nullptr, "gda_spmm_unw: graph is NULL");
  if (g->N == 0) return GDA_OK;
  return gda::spmm_unw(g, transpose, nb, Z, ldz, z_batch_stride, Y, ldy, y_batch_stride, H, last, bias, epi_flags,
                       dropout_p, seed, seed_offset, workspace, workspace_bytes, gda::as_stream(stream));
}

int gda_spmm_unw_halo_f32(const gda_graph_t* block, int nb, const float* Z, int64_t ldz, int64_t z_batch_stride, float* Y,
                          int64_t ldy, int64_t y_batch_stride, int H, const int32_t* halo_mask, const int32_t* halo_slot,
                          void* const* peer_base, const int64_t* peer_batch_stride, int num_peers, void* workspace,
                          int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(block && halo_mask && halo_slot && peer_base && peer_batch_stride, "gda_spmm_unw_halo_f32: NULL argument");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS, "gda_spmm_unw_halo_f32: bad peer count");
  if (block->N == 0) return GDA_OK;
  gda::PeerTable ot;
  gda::HaloStride hs;
  for (int i = 0; i < GDA_MAX_PEERS; ++i) {
    ot.p[i] = peer_base[i < num_peers ? i : 0];
    hs.b[i] = static_cast<uint64_t>(peer_batch_stride[i < num_peers ? i : 0]) * sizeof(float);
    GDA_REQUIRE(reinterpret_cast<uintptr_t>(ot.p[i]) % 16 == 0 && hs.b[i] % 16 == 0, "gda_spmm_unw_halo_f32: peer buffers must be 16-byte aligned");
  }
  return gda::spmm_unw(block, 0, nb, Z, ldz, z_batch_stride, Y, ldy, y_batch_stride, H, 0, nullptr, 0, 0.f, 0, nullptr,
                       workspace, workspace_bytes, gda::as_stream(stream), halo_mask, halo_slot, &ot, num_peers, &hs);
}

int64_t gda_spmm_workspace_bytes(const gda_graph_t* g, int transpose, int H) {
  if (!g || H <= 0) return 0;
  const gda::Csr& c = transpose ? g->csr_t : g->csr;
  return static_cast<int64_t>(c.num_segs) * H * static_cast<int64_t>(sizeof(float));
}

int gda_spmm_f32(const gda_graph_t* g, int transpose, const float* X, int64_t ldx, float* Y, int64_t ldy,
                 int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                 const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  return gda::spmm_any<float, 4>(g, transpose, X, ldx, Y, ldy, H, bias, epi_flags, dropout_p, seed, seed_offset,
                                 workspace, workspace_bytes, gda::as_stream(stream));
}

int gda_spmm_bf16(const gda_graph_t* g, int transpose, const void* X, int64_t ldx, void* Y, int64_t ldy,
                  int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                  const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  return gda::spmm_any<__nv_bfloat16, 8>(g, transpose, static_cast<const __nv_bfloat16*>(X), ldx,
                                         static_cast<__nv_bfloat16*>(Y), ldy, H, bias, epi_flags, dropout_p,
                                         seed, seed_offset, workspace, workspace_bytes, gda::as_stream(stream));
}

int gda_spmm_peer_f32(const gda_graph_t* part, int transpose, const void* const* peer_x, int num_peers, int my_rank,
                      int64_t ldx, float* Y, int64_t ldy, int H, const float* bias, int epi_flags, float dropout_p,
                      uint64_t seed, const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes,
                      gda_stream_t stream) {
  GDA_REQUIRE(part && part->peer_packed, "gda_spmm_peer_f32: graph is not a partition (gda_graph_partition)");
  GDA_REQUIRE(peer_x && num_peers >= 1 && num_peers <= GDA_MAX_PEERS && my_rank >= 0 && my_rank < num_peers,
              "gda_spmm_peer_f32: bad peer arguments");
  GDA_REQUIRE(gda::ceil_div(part->global_N, part->rows_per_rank) <= num_peers, "gda_spmm_peer_f32: too few peers");
  gda::PeerTable t;
  for (int i = 0; i < GDA_MAX_PEERS; ++i) t.p[i] = peer_x[i < num_peers ? i : my_rank];
  for (int i = 0; i < num_peers; ++i) GDA_REQUIRE(peer_x[i] != nullptr, "gda_spmm_peer_f32: NULL peer block");
  GDA_REQUIRE(part->rows_per_rank * ldx < (int64_t(1) << 32), "gda_spmm_peer_f32: block too large");
  return gda::spmm_any<float, 4>(part, transpose, static_cast<const float*>(peer_x[my_rank]), ldx, Y, ldy, H, bias,
                                 epi_flags, dropout_p, seed, seed_offset, workspace, workspace_bytes,
                                 gda::as_stream(stream), &t, my_rank);
}

int gda_spmm_nb_f32(const gda_graph_t* g, int transpose, int nb, const float* X, int64_t ldx, int64_t x_batch_stride,
                    float* Y, int64_t ldy, int64_t y_batch_stride, int H, const float* bias, int epi_flags,
                    float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                    int64_t workspace_bytes, gda_stream_t stream) {
  return gda::spmm_any<float, 4>(g, transpose, X, ldx, Y, ldy, H, bias, epi_flags, dropout_p, seed, seed_offset,
                                 workspace, workspace_bytes, gda::as_stream(stream), nullptr, 0, nb, x_batch_stride,
                                 y_batch_stride);
}

// k chained steps in one call (host-side loop: k launches, one crossing of the ABI).
// T0 / T1: ping-pong scratch [nb * N, H] (needed for k >= 2 / k >= 3); the epilogue is applied on the last step.
int gda_spmm_k_nb_f32(const gda_graph_t* g, int transpose, int k, int nb, const float* X, int64_t ldx,
                      int64_t x_batch_stride, float* Y, int64_t ldy, int64_t y_batch_stride, float* T0, float* T1, int H,
                      const float* bias, int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                      void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(g != nullptr, "gda_spmm_k: graph is NULL");
  GDA_REQUIRE(k >= 1, "gda_spmm_k: k must be >= 1");
  GDA_REQUIRE(k < 2 || T0, "gda_spmm_k: T0 scratch needed for k >= 2");
  GDA_REQUIRE(k < 3 || T1, "gda_spmm_k: T1 scratch needed for k >= 3");
  {
    // unit-weight graphs, k >= 3: D S (D^2 S)^(k-1) D x on the weight-free kernel (one row scaling + k lean steps)
    const gda::Csr& c = transpose ? g->csr_t : g->csr;
    const int64_t nh = g->N * static_cast<int64_t>(H);
    if (k >= 3 && g->N > 0 && gda::unw_path(g, c, H, H, ldy, T0, Y, nb, nh, y_batch_stride) &&
        reinterpret_cast<uintptr_t>(T1) % 16 == 0 && ldx % 4 == 0 && x_batch_stride % 4 == 0 &&
        reinterpret_cast<uintptr_t>(X) % 16 == 0) {
      int rc = gda_row_scale_f32(g, nb, X, ldx, x_batch_stride, T0, H, nh, H, stream);
      if (rc) return rc;
      const float* zsrc = T0;
      for (int i = 0; i < k; ++i) {
        const bool last = i == k - 1;
        float* dst = last ? Y : ((i & 1) ? T0 : T1);
        rc = gda_spmm_unw_nb_f32(g, transpose, nb, zsrc, H, nh, dst, last ? ldy : H, last ? y_batch_stride : nh, H,
                                 last ? 1 : 0, last ? bias : nullptr, last ? epi_flags : 0, last ? dropout_p : 0.f, seed,
                                 seed_offset, workspace, workspace_bytes, stream);
        if (rc) return rc;
        zsrc = dst;
      }
      return GDA_OK;
    }
  }
  const float* src = X;
  int64_t ld_src = ldx, bs_src = x_batch_stride;
  for (int i = 0; i < k; ++i) {
    const bool last = i == k - 1;
    float* dst = last ? Y : ((i & 1) ? T1 : T0);
    const int64_t ld_dst = last ? ldy : H;
    const int64_t bs_dst = last ? y_batch_stride : g->N * static_cast<int64_t>(H);
    int rc = gda_spmm_nb_f32(g, transpose, nb, src, ld_src, bs_src, dst, ld_dst, bs_dst, H, last ? bias : nullptr,
                             last ? epi_flags : 0, last ? dropout_p : 0.f, seed, seed_offset, workspace,
                             workspace_bytes, stream);
    if (rc) return rc;
    src = dst;
    ld_src = ld_dst;
    bs_src = bs_dst;
  }
  return GDA_OK;
}

int gda_spmm_k_f32(const gda_graph_t* g, int transpose, int k, const float* X, int64_t ldx, float* Y, int64_t ldy,
                   float* T0, float* T1, int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                   const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  return gda_spmm_k_nb_f32(g, transpose, k, 1, X, ldx, 0, Y, ldy, 0, T0, T1, H, bias, epi_flags, dropout_p, seed,
                           seed_offset, workspace, workspace_bytes, stream);
}

int gda_peer_barrier(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t epoch, int* error_flag,
                     gda_stream_t stream);   // peer.cu

// The partitioned form: barrier, copy x into the rank's symmetric buffer 0, then k x (barrier, peer
// aggregation) ping-ponging between the symmetric buffers; the last step writes Y (plain memory).
// sym0 / sym1: per-rank pointers of the two CUDA-IPC buffers ([rows_per_rank, H] each); uses barrier
// epochs epoch0+1 .. epoch0+k+1 (the caller advances its counter by k+1).
int gda_spmm_peer_k_f32(const gda_graph_t* part, int transpose, int k, const float* x_local,
                        const void* const* sym0, const void* const* sym1, int num_peers, int my_rank, float* Y,
                        int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                        const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes,
                        uint64_t* const* peer_flags, uint64_t epoch0, int* error_flag, gda_stream_t stream) {
  GDA_REQUIRE(part && part->peer_packed && k >= 1 && x_local && sym0 && sym1 && Y, "gda_spmm_peer_k_f32: bad arguments");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS && my_rank >= 0 && my_rank < num_peers,
              "gda_spmm_peer_k_f32: bad peer arguments");
  cudaStream_t st = gda::as_stream(stream);
  int rc = gda_peer_barrier(peer_flags, my_rank, num_peers, ++epoch0, error_flag, stream);
  if (rc) return rc;
  GDA_CUDA(cudaMemcpyAsync(const_cast<void*>(sym0[my_rank]), x_local, sizeof(float) * part->N * H,
                           cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < k; ++i) {
    const bool last = i == k - 1;
    if ((rc = gda_peer_barrier(peer_flags, my_rank, num_peers, ++epoch0, error_flag, stream))) return rc;
    const void* const* src = (i & 1) ? sym1 : sym0;
    float* dst = last ? Y : static_cast<float*>(const_cast<void*>(((i & 1) ? sym0 : sym1)[my_rank]));
    rc = gda_spmm_peer_f32(part, transpose, src, num_peers, my_rank, H, dst, H, H, last ? bias : nullptr,
                           last ? epi_flags : 0, last ? dropout_p : 0.f, seed, seed_offset, workspace,
                           workspace_bytes, stream);
    if (rc) return rc;
  }
  return GDA_OK;
}

int gda_peer_barrier_dev(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t* epoch_dev, int* error_flag,
                         gda_stream_t stream);   // peer.cu

// gda_spmm_peer_k_f32 with the barrier epochs kept in device memory (*epoch_dev advances by k + 1): no launch
// argument depends on host state, so the whole sequence can be captured in a CUDA graph and replayed.
int gda_spmm_peer_k_dev_f32(const gda_graph_t* part, int transpose, int k, const float* x_local,
                            const void* const* sym0, const void* const* sym1, int num_peers, int my_rank, float* Y,
                            int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                            const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes,
                            uint64_t* const* peer_flags, uint64_t* epoch_dev, int* error_flag, gda_stream_t stream) {
  GDA_REQUIRE(part && part->peer_packed && k >= 1 && x_local && sym0 && sym1 && Y && epoch_dev,
              "gda_spmm_peer_k_dev_f32: bad arguments");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS && my_rank >= 0 && my_rank < num_peers,
              "gda_spmm_peer_k_dev_f32: bad peer arguments");
  cudaStream_t st = gda::as_stream(stream);
  int rc = gda_peer_barrier_dev(peer_flags, my_rank, num_peers, epoch_dev, error_flag, stream);
  if (rc) return rc;
  GDA_CUDA(cudaMemcpyAsync(const_cast<void*>(sym0[my_rank]), x_local, sizeof(float) * part->N * H,
                           cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < k; ++i) {
    const bool last = i == k - 1;
    if ((rc = gda_peer_barrier_dev(peer_flags, my_rank, num_peers, epoch_dev, error_flag, stream))) return rc;
    const void* const* src = (i & 1) ? sym1 : sym0;
    float* dst = last ? Y : static_cast<float*>(const_cast<void*>(((i & 1) ? sym0 : sym1)[my_rank]));
    rc = gda_spmm_peer_f32(part, transpose, src, num_peers, my_rank, H, dst, H, H, last ? bias : nullptr,
                           last ? epi_flags : 0, last ? dropout_p : 0.f, seed, seed_offset, workspace,
                           workspace_bytes, stream);
    if (rc) return rc;
  }
  return GDA_OK;
}

// ---- push mode: A_hat^k over a partition with the exchange fused into the producer (dist.py) ------------------------
// gbuf0 / gbuf1: per-rank base pointers of two symmetric GATHER buffers, each [num_peers * rows_per_rank, H] fp32 -- a
// full copy of the (padded) global matrix on every rank, block q = rank q's rows.  Step i gathers from the local copy
// and stores each output row into block `my_rank` of the other buffer on EVERY rank (P2P stores); the last step writes
// Y (local, plain).  Barriers (device epoch) order the steps.  One aggregation launch moves (num_peers - 1) x block
// bytes over NVLink as coalesced 512-byte row stores, instead of one remote 512-byte load per referencing non-zero.
int gda_spmm_push_f32(const gda_graph_t* part, int transpose, const void* gather_local, void* const* out_blocks,
                      int num_out, int num_peers, int64_t ldx, int64_t ldy, int H, const float* bias, int epi_flags,
                      float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                      int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(part && part->peer_packed && gather_local && out_blocks, "gda_spmm_push_f32: bad arguments");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS && num_out >= 1 && num_out <= GDA_MAX_PEERS,
              "gda_spmm_push_f32: bad peer counts");
  GDA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gda_spmm_push_f32: dropout_p outside [0,1)");
  gda::PeerTable gt, ot;
  const char* base = static_cast<const char*>(gather_local);
  for (int i = 0; i < GDA_MAX_PEERS; ++i)
    gt.p[i] = base + static_cast<int64_t>(i < num_peers ? i : 0) * part->rows_per_rank * ldx * static_cast<int64_t>(sizeof(float));
  for (int i = 0; i < GDA_MAX_PEERS; ++i) {
    ot.p[i] = out_blocks[i < num_out ? i : 0];
    GDA_REQUIRE(ot.p[i] != nullptr, "gda_spmm_push_f32: NULL output block");
  }
  gda::Epilogue epi;
  epi.bias = bias; epi.flags = epi_flags; epi.thresh = gda::dropout_threshold(dropout_p);
  epi.scale = 1.0f / (1.0f - dropout_p); epi.seed = seed; epi.seed_offset = seed_offset;
  return gda::spmm_push(part, transpose, gt, ot, num_out, ldx, ldy, H, epi, static_cast<float*>(workspace),
                        workspace_bytes, gda::as_stream(stream));
}

int gda_spmm_halo_f32(const gda_graph_t* block, const float* X, int64_t ldx, float* Y, int64_t ldy, int H,
                      const int32_t* halo_mask, const int32_t* halo_slot, void* const* peer_base, int num_peers,
                      const float* bias, int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                      void* workspace, int64_t workspace_bytes, gda_stream_t stream) {
  GDA_REQUIRE(block && X && Y && X != Y && halo_mask && halo_slot && peer_base, "gda_spmm_halo_f32: bad arguments");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS, "gda_spmm_halo_f32: bad peer count");
  GDA_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gda_spmm_halo_f32: dropout_p outside [0,1)");
  gda::PeerTable ot;
  for (int i = 0; i < GDA_MAX_PEERS; ++i) ot.p[i] = peer_base[i < num_peers ? i : 0];
  gda::Epilogue epi;
  epi.bias = bias; epi.flags = epi_flags; epi.thresh = gda::dropout_threshold(dropout_p);
  epi.scale = 1.0f / (1.0f - dropout_p); epi.seed = seed; epi.seed_offset = seed_offset;
  return gda::spmm_halo(block, X, ldx, Y, ldy, H, halo_mask, halo_slot, ot, num_peers, epi,
                        static_cast<float*>(workspace), workspace_bytes, gda::as_stream(stream));
}

int gda_spmm_push_k_f32(const gda_graph_t* part, int transpose, int k, const float* x_local, void* const* gbuf0,
                        void* const* gbuf1, int num_peers, int my_rank, float* Y, int H, const float* bias,
                        int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                        int64_t workspace_bytes, uint64_t* const* peer_flags, uint64_t* epoch_dev, int* error_flag,
                        gda_stream_t stream) {
  GDA_REQUIRE(part && part->peer_packed && k >= 1 && x_local && gbuf0 && gbuf1 && Y && epoch_dev,
              "gda_spmm_push_k_f32: bad arguments");
  GDA_REQUIRE(num_peers >= 1 && num_peers <= GDA_MAX_PEERS && my_rank >= 0 && my_rank < num_peers,
              "gda_spmm_push_k_f32: bad peer arguments");
  cudaStream_t st = gda::as_stream(stream);
  const int64_t block_bytes = part->rows_per_rank * static_cast<int64_t>(H) * sizeof(float);
  int rc = gda_peer_barrier_dev(peer_flags, my_rank, num_peers, epoch_dev, error_flag, stream);
  if (rc) return rc;
  for (int q = 0; q < num_peers; ++q)               // copy-in: this rank's rows into block my_rank of every gather buffer 0
    GDA_CUDA(cudaMemcpyAsync(static_cast<char*>(gbuf0[q]) + my_rank * block_bytes, x_local, sizeof(float) * part->N * H,
                             cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < k; ++i) {
    const bool last = i == k - 1;
    if ((rc = gda_peer_barrier_dev(peer_flags, my_rank, num_peers, epoch_dev, error_flag, stream))) return rc;
    void* const* src = (i & 1) ? gbuf1 : gbuf0;
    void* const* dstb = (i & 1) ? gbuf0 : gbuf1;
    void* outs[GDA_MAX_PEERS];
    int nout = 1;
    if (last) {
      outs[0] = Y;
    } else {
      nout = num_peers;
      for (int q = 0; q < num_peers; ++q) outs[q] = static_cast<char*>(dstb[q]) + my_rank * block_bytes;
    }
    rc = gda_spmm_push_f32(part, transpose, src[my_rank], outs, nout, num_peers, H, H, H, last ? bias : nullptr,
                           last ? epi_flags : 0, last ? dropout_p : 0.f, seed, seed_offset, workspace, workspace_bytes,
                           stream);
    if (rc) return rc;
  }
  return GDA_OK;
}

}  // extern "C"
