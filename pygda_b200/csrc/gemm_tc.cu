// Dense GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
// C[M,N] (fp32) = op(A) * op(B) with each fp32 operand given as a SPLIT-BF16 pair
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi), |x - hi - lo| <= 2^-17 |x|).  The product is
// evaluated as  Ah*Bh + Ah*Bl + Al*Bh  -- three kind::f16 UMMAs per K step accumulating in one
// fp32 TMEM tile -- which keeps the relative error of every product term below ~2^-16, i.e.
// inside the 1e-4 fp32 parity bar of the reference's `x @ W^T` (pygda/nn/prop_gcn_conv.py:205),
// while streaming the same 4 bytes per element as fp32 would.  Plain bf16 or TF32 alone miss
// that bar (2^-9 / 2^-11 per term).
//
// One CTA = one 128x128 output tile (optionally one K split of it), 192 threads:
//   warp 0    TMA producer : cp.async.bulk.tensor (SWIZZLE_128B boxes) into a 3-stage smem ring
//   warp 1    MMA issuer   : one elected lane issues tcgen05.mma (M=128, N=128, K=16), commits to
//                            the stage's `empty` mbarrier; owns the TMEM allocation (128 columns)
//   warps 2-5 epilogue     : tcgen05.ld 32x32b.x32 -> registers -> global (row = TMEM lane)
// Operand majors: row-major A[M,K] / B[N,K] are K-major (box 64 x 128, SBO = 1024 B); the
// transposed forms A[K,M] / B[K,N] are MN-major (two boxes 64 x 64, LBO = 8192 B, SBO = 1024 B),
// so weight gradients  dW = G^T X  (reduction over the node dimension) need no transposes.
// Out-of-bounds parts of a box are zero-filled by TMA; the epilogue masks rows/columns.
#include <cstdlib>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace gda {
namespace {

// Tile shapes: MT x 128 rows by BN columns per CTA.  (MT, BN) = (2, 128) halves the L2 re-reads of
// the B operand for tall outputs (round-1 profile: with 128 x 128 tiles every one of 782 CTAs
// re-read all of W through L2 -- L2 at 70 %, DRAM only 60 %); (1, 256) does the same for the A
// operand of the weight-gradient GEMM; (1, 128) is the small-shape default.
// TERMS = 3: split-bf16 operands (hi, lo), three UMMAs per K step (fp32-accurate).  TERMS = 1: plain bf16
// operands, one UMMA per K step (the bf16 feature path of BASELINE config 3) -- half the bytes per stage.
template <int MT, int BN_, int TERMS = 3> struct TileCfg {
  static constexpr int A_BYTES = MT * BM * BK * 2;      // one of (hi, lo)
  static constexpr int B_BYTES = BN_ * BK * 2;
  static constexpr int PARTS = TERMS == 3 ? 2 : 1;
  static constexpr int STAGE_BYTES = PARTS * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 4 ? 4 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = MT * BN_;
};


template <bool A_MN, bool B_MN, int MT, int BN, int TERMS = 3, bool OUT_BF16 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16x3(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
              const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
              float* __restrict__ C, int64_t ldc, int M, int N, int K, int kb_per_split, int64_t split_stride) {
  using Cfg = TileCfg<MT, BN, TERMS>;
  static_assert(TERMS == 3 || TERMS == 1, "three split terms or one plain bf16 term");
  constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES;
  static_assert(!(A_MN && MT != 1), "MN-major A uses one 128-row tile");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;                    // SWIZZLE_128B wants 1024 B alignment
  const uint32_t bars = tiles + STAGES * STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tmem_full = bars + 16 * STAGES;
  const uint32_t tmem_slot = bars + 16 * STAGES + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * (MT * BM), n0 = blockIdx.x * BN;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapAl); tma_prefetch_desc(&mapBh); tma_prefetch_desc(&mapBl);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
        const uint32_t st = tiles + stage * STAGE_BYTES;
        const uint32_t bar = full0 + 8 * stage;
        mbar_expect_tx(bar, STAGE_BYTES);
        const int k0 = (kb_begin + kb) * BK;
        constexpr int PARTS = Cfg::PARTS;              // stage layout: Ah [Al] Bh [Bl]
        const uint32_t sAh = st, sAl = st + A_BYTES, sBh = st + PARTS * A_BYTES, sBl = sBh + B_BYTES;
        if (!A_MN) {                                   // one box 64 x (MT*128)
          tma_load_2d(sAh, &mapAh, bar, k0, m0);
          if (TERMS == 3) tma_load_2d(sAl, &mapAl, bar, k0, m0);
        } else {                                       // two boxes 64(m) x 64(k)
          tma_load_2d(sAh, &mapAh, bar, m0, k0);
          tma_load_2d(sAh + 8192, &mapAh, bar, m0 + 64, k0);
          if (TERMS == 3) {
            tma_load_2d(sAl, &mapAl, bar, m0, k0);
            tma_load_2d(sAl + 8192, &mapAl, bar, m0 + 64, k0);
          }
        }
        if (!B_MN) {                                   // one box 64 x BN
          tma_load_2d(sBh, &mapBh, bar, k0, n0);
          if (TERMS == 3) tma_load_2d(sBl, &mapBl, bar, k0, n0);
        } else {                                       // BN/64 boxes 64(n) x 64(k)
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) {
            tma_load_2d(sBh + i * 8192, &mapBh, bar, n0 + 64 * i, k0);
            if (TERMS == 3) tma_load_2d(sBl + i * 8192, &mapBl, bar, n0 + 64 * i, k0);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN, B_MN, BN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t st = tiles + stage * STAGE_BYTES;
        const uint32_t sAh = st, sAl = st + A_BYTES, sBh = st + Cfg::PARTS * A_BYTES, sBl = sBh + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          // K-major: +32 B per K=16 step inside the 128 B swizzle row; MN-major: +16 rows * 128 B
          const uint32_t a_off = A_MN ? ks * 2048u : ks * 32u;
          const uint32_t b_off = B_MN ? ks * 2048u : ks * 32u;
          const uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
          const uint64_t bh = make_desc(sBh + b_off, b_lbo, 1024u);
          const uint64_t bl = make_desc(sBl + b_off, b_lbo, 1024u);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t ah = make_desc(sAh + mt * (BM * 128) + a_off, a_lbo, 1024u);   // 128 rows x 128 B
            const uint64_t al = make_desc(sAl + mt * (BM * 128) + a_off, a_lbo, 1024u);
            const uint32_t d = tmem_base + mt * BN;
            umma_bf16(d, ah, bh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            if (TERMS == 3) {
              umma_bf16(d, ah, bl, idesc, 1u);
              umma_bf16(d, al, bh, idesc, 1u);
            }
          }
        }
        umma_commit(empty0 + 8 * stage);          // frees the smem stage once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full);                     // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int row = m0 + mt * BM + q * 32 + lane;
      float* crow = C + blockIdx.z * split_stride + static_cast<int64_t>(row) * ldc + n0;
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(crow) & 15u) == 0) && (ldc % 4 == 0);
      // bf16 output (TERMS == 1 feature path): C is a __nv_bfloat16 matrix, ldc in bf16 elements, no K splits
      __nv_bfloat16* hrow = reinterpret_cast<__nv_bfloat16*>(C) + static_cast<int64_t>(row) * ldc + n0;
      const bool hvec_ok = ((reinterpret_cast<uintptr_t>(hrow) & 15u) == 0) && (ldc % 8 == 0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mt * BN + c * 32, r);
        if (OUT_BF16) {
          if (row < M) {
            const int col0 = n0 + c * 32;
            if (hvec_ok && col0 + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[8 * j + 2 * i]),
                                                                  __uint_as_float(r[8 * j + 2 * i + 1]));
                  w[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(hrow + c * 32 + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < N) hrow[c * 32 + j] = __float2bfloat16_rn(__uint_as_float(r[j]));
            }
          }
          continue;
        }
        if (row < M) {
          const int col0 = n0 + c * 32;
          if (vec_ok && col0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(crow + c * 32 + 4 * j) =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                              __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) crow[c * 32 + j] = __uint_as_float(r[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

// ------------------------------------------------------------------ fp32 -> split bf16
__global__ void k_split_bf16(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ldx,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ldo) {
  const int64_t groups = ldo / 8;                               // 8 output columns per thread
  const int64_t total = rows * groups;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups, c0 = (i % groups) * 8;
    float v[8];
    const float* src = x + r * ldx + c0;
    if (c0 + 8 <= cols && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? __ldg(src + j) : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
      h[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      l[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    *reinterpret_cast<uint4*>(hi + r * ldo + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + r * ldo + c0) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}


struct TcPlan { int mt, bn, splits; };

TcPlan plan_for(bool a_mn, bool b_mn, int64_t M, int64_t N, int64_t K) {
  TcPlan p{1, 128, 1};
  if (!a_mn && ceil_div(M, 256) * ceil_div(N, 128) >= kNumSMs) p.mt = 2;       // tall output: share B
  // (1, 256) tiles (share A across a wide output) measured SLOWER on the B200 for dW = G^T X
  // (581 vs 474 us): only 2 pipeline stages fit; kept compiled for experiments, not selected.
  else if (b_mn && N >= 256 && std::getenv("GDA_TC_BN256")) p.bn = 256;
  const int64_t tiles = ceil_div(M, p.mt * BM) * ceil_div(N, p.bn);
  const int64_t nkb = ceil_div(K, BK);
  if (tiles < kNumSMs) {
    int64_t s = (4 * kNumSMs + tiles / 2) / tiles;
    const int64_t cap = nkb / 8 > 0 ? nkb / 8 : 1;
    if (s > cap) s = cap;
    p.splits = s < 1 ? 1 : static_cast<int>(s);
  }
  return p;
}

template <bool A_MN, bool B_MN, int MT, int BN_, int TERMS = 3, bool OUT_BF16 = false>
int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, int splits, int64_t split_stride, cudaStream_t st) {
  using Cfg = TileCfg<MT, BN_, TERMS>;
  static bool attr_set = false;
  if (!attr_set) {
    GDA_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3<A_MN, B_MN, MT, BN_, TERMS, OUT_BF16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int nkb = static_cast<int>(ceil_div(K, BK));
  const int kbps = static_cast<int>(ceil_div(nkb, splits));
  dim3 grid(static_cast<unsigned>(ceil_div(N, BN_)), static_cast<unsigned>(ceil_div(M, MT * BM)),
            static_cast<unsigned>(ceil_div(nkb, kbps)));
  k_gemm_bf16x3<A_MN, B_MN, MT, BN_, TERMS, OUT_BF16><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(
      ah, al, bh, bl, C, ldc, (int)M, (int)N, (int)K, kbps, split_stride);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

}  // namespace

int splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float alpha, float beta, float* C, int64_t ldc,
                  cudaStream_t st);   // gemm_simt.cu

int64_t round_up8(int64_t x) { return (x + 7) / 8 * 8; }

int split_bf16(const float* x, int64_t rows, int64_t cols, int64_t ldx, void* hi, void* lo, int64_t ldo,
               cudaStream_t st) {
  GDA_REQUIRE(rows >= 0 && cols >= 0 && ldx >= cols, "gda_split_bf16: bad shape");
  GDA_REQUIRE(ldo >= cols && ldo % 8 == 0, "gda_split_bf16: ld_out must be >= cols and a multiple of 8");
  if (rows == 0 || cols == 0) return GDA_OK;
  GDA_REQUIRE(x && hi && lo, "gda_split_bf16: NULL pointer");
  GDA_REQUIRE((reinterpret_cast<uintptr_t>(hi) % 16 == 0) && (reinterpret_cast<uintptr_t>(lo) % 16 == 0),
              "gda_split_bf16: outputs must be 16-byte aligned");
  const int64_t total = rows * (ldo / 8);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  k_split_bf16<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, rows, cols, ldx, static_cast<__nv_bfloat16*>(hi),
                                                             static_cast<__nv_bfloat16*>(lo), ldo);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

bool bf16x3_shape_ok(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb) {
  return M >= 1 && N >= 64 && K >= 64 && lda % 8 == 0 && ldb % 8 == 0 && M < (int64_t(1) << 31) &&
         N < (int64_t(1) << 31) && K < (int64_t(1) << 31) && ceil_div(M, BM) <= 65535;
}

int64_t bf16x3_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  // upper bound over the operand majors (the plan depends on them only through the tile shape)
  int s = 1;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) { const int t = plan_for(a, b, M, N, K).splits; if (t > s) s = t; }
  return s > 1 ? static_cast<int64_t>(s) * M * N * sizeof(float) : 0;
}

int gemm_bf16x3(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_lo,
                int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb, float* C, int64_t ldc, void* ws,
                int64_t ws_bytes, cudaStream_t st) {
  GDA_REQUIRE(a_hi && a_lo && b_hi && b_lo && C, "gda_gemm_bf16x3: NULL pointer");
  GDA_REQUIRE(bf16x3_shape_ok(M, N, K, lda, ldb), "gda_gemm_bf16x3: needs N >= 64, K >= 64 and lda, ldb multiples of 8");
  GDA_REQUIRE(ldc >= N, "gda_gemm_bf16x3: ldc < N");
  for (const void* p : {a_hi, a_lo, b_hi, b_lo})
    GDA_REQUIRE(reinterpret_cast<uintptr_t>(p) % 16 == 0, "gda_gemm_bf16x3: operands must be 16-byte aligned");
  CUtensorMap ah, al, bh, bl;
  int rc;
  // A row-major [M,K] (K-major): inner = K, outer = M, box 64 x 128.  A^T stored [K,M] (MN-major):
  // inner = M, outer = K, box 64 x 64.  Same for B with N.
  const int64_t a_in = transA ? M : K, a_out = transA ? K : M;
  const int64_t b_in = transB ? K : N, b_out = transB ? N : K;
  const bool a_mn = transA != 0, b_mn = transB == 0;
  const TcPlan plan = plan_for(a_mn, b_mn, M, N, K);
  const int a_box = a_mn ? BK : plan.mt * BM, b_box = b_mn ? BK : plan.bn;
  if ((rc = make_map(&ah, a_hi, a_in, a_out, lda, a_box)) || (rc = make_map(&al, a_lo, a_in, a_out, lda, a_box)) ||
      (rc = make_map(&bh, b_hi, b_in, b_out, ldb, b_box)) || (rc = make_map(&bl, b_lo, b_in, b_out, ldb, b_box)))
    return rc;
  const int splits = plan.splits;
  float* out = C;
  int64_t out_ld = ldc, stride = 0;
  if (splits > 1) {
    const int64_t need = static_cast<int64_t>(splits) * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return fail(GDA_E_WORKSPACE, "gda_gemm_bf16x3: workspace too small");
    out = static_cast<float*>(ws); out_ld = N; stride = M * N;
  }
#define GDA_TC_LAUNCH(AM, BM_, MTV, BNV) \
  rc = launch_tc<AM, BM_, MTV, BNV>(ah, al, bh, bl, out, out_ld, M, N, K, splits, stride, st)
  if (plan.mt == 2) {
    if (!b_mn) GDA_TC_LAUNCH(false, false, 2, 128); else GDA_TC_LAUNCH(false, true, 2, 128);
  } else if (plan.bn == 256) {
    if (!a_mn) GDA_TC_LAUNCH(false, true, 1, 256); else GDA_TC_LAUNCH(true, true, 1, 256);
  } else {
    if (!a_mn && !b_mn) GDA_TC_LAUNCH(false, false, 1, 128);
    else if (!a_mn && b_mn) GDA_TC_LAUNCH(false, true, 1, 128);
    else if (a_mn && !b_mn) GDA_TC_LAUNCH(true, false, 1, 128);
    else GDA_TC_LAUNCH(true, true, 1, 128);
  }
#undef GDA_TC_LAUNCH
  if (rc) return rc;
  if (splits > 1) {
    const int nkb = static_cast<int>(ceil_div(K, BK));
    const int kbps = static_cast<int>(ceil_div(nkb, splits));
    const int used = static_cast<int>(ceil_div(nkb, kbps));
    return splitk_reduce(static_cast<float*>(ws), used, M, N, 1.f, 0.f, C, ldc, st);
  }
  return GDA_OK;
}

// ---- plain bf16 operands, fp32 accumulation, fp32 or bf16 output (BASELINE config 3 feature path) ----
int64_t bf16_workspace_bytes(int64_t M, int64_t N, int64_t K, int out_bf16) {
  return out_bf16 ? 0 : bf16x3_workspace_bytes(M, N, K);
}

int gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
              int64_t ldb, void* C, int64_t ldc, int out_bf16, void* ws, int64_t ws_bytes, cudaStream_t st) {
  GDA_REQUIRE(A && B && C, "gda_gemm_bf16: NULL pointer");
  GDA_REQUIRE(bf16x3_shape_ok(M, N, K, lda, ldb), "gda_gemm_bf16: needs N >= 64, K >= 64 and lda, ldb multiples of 8");
  GDA_REQUIRE(ldc >= N, "gda_gemm_bf16: ldc < N");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0,
              "gda_gemm_bf16: operands must be 16-byte aligned");
  CUtensorMap am, bm;
  int rc;
  const int64_t a_in = transA ? M : K, a_out = transA ? K : M;
  const int64_t b_in = transB ? K : N, b_out = transB ? N : K;
  const bool a_mn = transA != 0, b_mn = transB == 0;
  TcPlan plan = plan_for(a_mn, b_mn, M, N, K);
  plan.bn = 128;
  if (out_bf16) plan.splits = 1;                       // partial sums would need an fp32 detour
  const int a_box = a_mn ? BK : plan.mt * BM, b_box = b_mn ? BK : plan.bn;
  if ((rc = make_map(&am, A, a_in, a_out, lda, a_box)) || (rc = make_map(&bm, B, b_in, b_out, ldb, b_box))) return rc;
  const int splits = plan.splits;
  float* out = static_cast<float*>(C);
  int64_t out_ld = ldc, stride = 0;
  if (splits > 1) {
    const int64_t need = static_cast<int64_t>(splits) * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return fail(GDA_E_WORKSPACE, "gda_gemm_bf16: workspace too small");
    out = static_cast<float*>(ws); out_ld = N; stride = M * N;
  }
#define GDA_TC1_LAUNCH(AM, BM_, MTV, OB) \
  rc = launch_tc<AM, BM_, MTV, 128, 1, OB>(am, am, bm, bm, out, out_ld, M, N, K, splits, stride, st)
#define GDA_TC1_PICK(OB)                                                                   \
  do {                                                                                     \
    if (plan.mt == 2) {                                                                    \
      if (!b_mn) GDA_TC1_LAUNCH(false, false, 2, OB); else GDA_TC1_LAUNCH(false, true, 2, OB); \
    } else if (!a_mn && !b_mn) GDA_TC1_LAUNCH(false, false, 1, OB);                         \
    else if (!a_mn && b_mn) GDA_TC1_LAUNCH(false, true, 1, OB);                             \
    else if (a_mn && !b_mn) GDA_TC1_LAUNCH(true, false, 1, OB);                             \
    else GDA_TC1_LAUNCH(true, true, 1, OB);                                                 \
  } while (0)
  if (out_bf16) GDA_TC1_PICK(true); else GDA_TC1_PICK(false);
#undef GDA_TC1_PICK
#undef GDA_TC1_LAUNCH
  if (rc) return rc;
  if (splits > 1) {
    const int nkb = static_cast<int>(ceil_div(K, BK));
    const int kbps = static_cast<int>(ceil_div(nkb, splits));
    const int used = static_cast<int>(ceil_div(nkb, kbps));
    return splitk_reduce(static_cast<float*>(ws), used, M, N, 1.f, 0.f, static_cast<float*>(C), ldc, st);
  }
  return GDA_OK;
}

// ---- fp32-operand entry used by gda_gemm_f32: split both operands into the workspace first ----
bool tc_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t, int64_t, int64_t, const float*,
                  const float*, const float*) {
  (void)transA; (void)transB;
  // worth it only when the SIMT kernel would be compute-bound: >= ~0.5 GFLOP and tensor-core friendly
  return N >= 64 && K >= 64 && M >= 128 && 2.0 * M * N * K >= 5e8 && encode_fn() != nullptr;
}

static void f32_layout(int trans, int64_t rows_logical, int64_t cols_logical, int64_t* r, int64_t* c) {
  // stored shape of op(X): not transposed -> [rows, cols]; transposed -> [cols, rows]
  *r = trans ? cols_logical : rows_logical;
  *c = trans ? rows_logical : cols_logical;
}

int64_t tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K) {
  if (!(N >= 64 && K >= 64 && M >= 128 && 2.0 * M * N * K >= 5e8)) return 0;
  int64_t ar, ac, br, bc;
  f32_layout(transA, M, K, &ar, &ac);
  f32_layout(!transB, N, K, &br, &bc);      // B stored [K,N] unless transB ([N,K])
  const int64_t a_bytes = 2 * ar * round_up8(ac) * 2, b_bytes = 2 * br * round_up8(bc) * 2;
  return a_bytes + b_bytes + bf16x3_workspace_bytes(M, N, K) + 1024;
}

int gemm_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
            const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* ws, int64_t ws_bytes,
            cudaStream_t st) {
  if (alpha != 1.f || beta != 0.f)
    return gemm_simt(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, ws, ws_bytes, st);
  if (!ws || ws_bytes < tc_workspace_bytes(transA, transB, M, N, K))
    return fail(GDA_E_WORKSPACE, "gda_gemm_f32: workspace smaller than gda_gemm_workspace_bytes()");
  int64_t ar, ac, br, bc;
  f32_layout(transA, M, K, &ar, &ac);
  f32_layout(!transB, N, K, &br, &bc);
  const int64_t alo = round_up8(ac), blo = round_up8(bc);
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  void* a_hi = p; p += ar * alo * 2;
  void* a_lo = p; p += ar * alo * 2;
  void* b_hi = p; p += br * blo * 2;
  void* b_lo = p; p += br * blo * 2;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 255) / 256 * 256);
  int rc;
  if ((rc = split_bf16(A, ar, ac, lda, a_hi, a_lo, alo, st))) return rc;
  if ((rc = split_bf16(B, br, bc, ldb, b_hi, b_lo, blo, st))) return rc;
  const int64_t rest = ws_bytes - (p - static_cast<char*>(ws));
  return gemm_bf16x3(transA, transB, M, N, K, a_hi, a_lo, alo, b_hi, b_lo, blo, C, ldc, p, rest, st);
}

}  // namespace gda
