// First-layer GEMMs on a SPARSE input matrix X (bag-of-words features: a few per cent non-zero), fp32-accurate,
// on the tcgen05 tensor cores:  lin(x) = X W^T (pygda/nn/prop_gcn_conv.py:205; x @ W, cached_gcn_conv.py:130)
// and its weight gradient dW = G^T X.
//
// The dense kernel (gemm_tc.cu) streams X as a split-bf16 pair -- 4 bytes per element, 2.7 GB per pass at
// BASELINE config 2 (100 k x 6775), HBM-bound at ~0.5 ms while the tensor pipe idles half of the time.  Here X
// stays TILE-PACKED in HBM (5 bytes per NON-ZERO: fp32 value + one position byte, ~0.24 GB) and the
// dense 128 x 64 operand tiles exist only in shared memory: four "expander" warps zero a stage, decode their
// 32-row x 64-column sub-tile (every entry on its own: see the format below), split every value into
// hi = bf16(v), lo = bf16(v - hi) and scatter both into the SWIZZLE_128B layout the UMMA descriptors expect;
// `fence.proxy.async` + an mbarrier arrive hand the stage to the MMA warp.  W / G (dense, small or streamed once)
// still arrive by TMA.  Same three product terms per K step as the dense kernel (Ah*Bh + Ah*Bl + Al*Bh) in the same
// K order; Ah*Bl is kept in a second accumulator and added in the epilogue (one N = 256 UMMA on [Bh | Bl] reads Ah
// from shared memory once).  What bounds the kernel is the L2 -> SM stream of the dense operand and the expanders'
// pace, not HBM: it runs at 0.82 of the measured sustained bf16 tensor rate.
//
// Tile-packed X ("XT", built by pygda_b200.data.PackedTiles; the pinned host staging form is this form with the
// values exponent-packed, so the per-step host->device copy lands in the buffers this kernel reads after one
// streaming kernel that rebuilds the fp32 values, gda_unpack_values_f32):
//   sub-tile t = strip * nkb + kb, strip = row / 32, kb = col / 64 (nkb = ceil(cols / 64));
//   position of an entry inside its sub-tile p = (row % 32) * 64 + (col % 64) in [0, 2048), entries sorted by p;
//   entries of sub-tile t: [ptr[t], ptr[t+1]) in vals (fp32) and codes (uint8 = p & 255);
//   seg[t][k], k = 0..7 (uint16) = number of entries of the sub-tile with p < 256 (k + 1)   (seg[t][7] = count),
//   so the high bits of entry i's position are #{k < 7 : seg[t][k] <= i}: decoding an entry needs no information
//   from its neighbours -- no prefix sums, no warp shuffles in the expander (a delta code with warp scans was
//   measured first: the scans' shuffles share the shared-memory data pipe with the zero fill, the scattered
//   stores and the UMMA operand reads, and cost 140 us of a 490 us launch).  5 bytes per non-zero + 20 bytes per
//   sub-tile.  The strip count is padded to a multiple of 4.
//
// One CTA = MT (1 or 2) 128 x 128 output tiles that share the dense operand (x one K split):
//   warp 0    TMA producer of the dense operand (B) tiles
//   warp 1    MMA issuer (one elected lane), owns the 256 MT TMEM columns
//   warps 2.. expanders of the sparse operand (A) tiles -- one group of 4 MT warps per ring slot (3 slots for
//             MT = 1, 2 for MT = 2), group g always fills slot g -- then the epilogue (TMEM -> registers -> global)
// DW = false: C[rows, N] = X B      A tile = 128 rows x 64 cols, K-major      (warp e: strip 4*tile + e)
// DW = true : C = X^T G             A tile = 64 rows(k) x 128 cols(m), MN-major (warp e: strip 2*kb + e/2, kb' e%2)
#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace gda {
namespace {

constexpr int XT_TILE = BM * BK * 2;                       // 16 KB: one of (Ah, Al, Bh, Bl)
// MT row tiles of the sparse operand per CTA share one dense-operand tile per stage: every CTA re-reads the
// whole dense operand through L2 (the 3.5 MB weight matrix 782 times per launch at config 2 -- 8.6 TB/s of
// L2 -> SM traffic with MT = 1, which is what bounded the kernel), MT = 2 halves that.
template <int MT> struct XtCfg {
  static constexpr int STAGES = MT == 1 ? 3 : 2;           // ring slots = expander groups (group g owns slot g)
  static constexpr int STAGE_BYTES = (2 * MT + 2) * XT_TILE;   // A0h A0l [A1h A1l] Bh Bl
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int EXP_WARPS = 4 * MT;                 // expander warps per group: one per 32-row sub-tile
  static constexpr int THREADS = 32 * (2 + EXP_WARPS * STAGES);
  static constexpr int TMEM_COLS = 256 * MT;
};
constexpr int XT_MAXC = 8;                                 // register-prefetched chunks of 32 entries per sub-tile

struct XtView {
  const float* vals;
  const uint8_t* codes;
  const int32_t* ptr;       // [nstrips * nkb + 1] entry offsets
  const uint4* seg;         // [nstrips * nkb] eight uint16 cumulative segment counts per sub-tile
  int nkb;                  // 64-column blocks per strip
  int nstrips;              // 32-row strips (multiple of 4)
};

struct SubPtr { int e0, n; };                               // first entry, entry count
struct SubData { uint32_t code[XT_MAXC]; float val[XT_MAXC]; uint4 seg; };

__device__ __forceinline__ void xt_load(const XtView& x, const SubPtr& p, int t, int lane, SubData& d) {
  d.seg = make_uint4(0u, 0u, 0u, 0u);
  if (t >= 0) d.seg = __ldg(x.seg + t);
#pragma unroll
  for (int j = 0; j < XT_MAXC; ++j) {
    const int i = j * 32 + lane;
    const bool on = i < p.n;
    d.code[j] = on ? static_cast<uint32_t>(__ldg(x.codes + p.e0 + i)) : 0u;
    d.val[j] = on ? __ldg(x.vals + p.e0 + i) : 0.f;
  }
}

// high bits of the position of entry i: the number of segment boundaries at or below i
__device__ __forceinline__ uint32_t xt_high(const uint4& sg, uint32_t i) {
  uint32_t h = 0;
  h += i >= (sg.x & 0xffffu); h += i >= (sg.x >> 16);
  h += i >= (sg.y & 0xffffu); h += i >= (sg.y >> 16);
  h += i >= (sg.z & 0xffffu); h += i >= (sg.z >> 16);
  h += i >= (sg.w & 0xffffu);
  return h;
}

// one entry -> the swizzled (hi, lo) tiles; `base_hi` / `base_lo` address this warp's 32 x 128-byte rows
__device__ __forceinline__ void xt_put(uint32_t base_hi, uint32_t base_lo, uint32_t pos, float v) {
  const uint32_t r = pos >> 6, c = pos & 63u;
  const uint32_t off = (r * 128u + c * 2u) ^ ((r & 7u) << 4);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(base_hi + off), "h"(__bfloat16_as_ushort(h)) : "memory");
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(base_lo + off), "h"(__bfloat16_as_ushort(l)) : "memory");
}

template <bool DW, bool B_MN, bool OUT_T, int MT>
__global__ void __launch_bounds__(XtCfg<MT>::THREADS, 1)
k_gemm_xt(const XtView x, const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
          float* __restrict__ C, int64_t ldc, int M, int N, int K, int kb_per_split, int64_t split_stride) {
  static_assert(!DW || B_MN, "the weight-gradient form reads G row-major (MN-major B)");
  using Cfg = XtCfg<MT>;
  constexpr int XT_STAGES = Cfg::STAGES, XT_STAGE_BYTES = Cfg::STAGE_BYTES, XT_GROUPS = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t bars = tiles + XT_STAGES * XT_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * XT_STAGES, tmem_full = bars + 16 * XT_STAGES;
  const uint32_t tmem_slot = bars + 16 * XT_STAGES + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * (MT * BM), n0 = blockIdx.x * BM;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapBh); tma_prefetch_desc(&mapBl);
    // a stage is full when the TMA bytes of B have landed (1 arrival + tx) and all its expander threads arrived
    for (int s = 0; s < XT_STAGES; ++s) { mbar_init(full0 + 8 * s, 1 + 32 * Cfg::EXP_WARPS); mbar_init(empty0 + 8 * s, 1); }
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
    // ===================== TMA producer: dense operand B (hi, lo) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
        const uint32_t st = tiles + stage * XT_STAGE_BYTES;
        const uint32_t bar = full0 + 8 * stage;
        mbar_expect_tx(bar, 2 * XT_TILE);
        const int k0 = (kb_begin + kb) * BK;
        const uint32_t sBh = st + 2 * MT * XT_TILE, sBl = sBh + XT_TILE;
        if (!B_MN) {                                   // B stored [N, K]: one box 64(k) x 128(n)
          tma_load_2d(sBh, &mapBh, bar, k0, n0);
          tma_load_2d(sBl, &mapBl, bar, k0, n0);
        } else {                                       // B stored [K, N]: two boxes 64(n) x 64(k)
          tma_load_2d(sBh, &mapBh, bar, n0, k0);
          tma_load_2d(sBh + 8192, &mapBh, bar, n0 + 64, k0);
          tma_load_2d(sBl, &mapBl, bar, n0, k0);
          tma_load_2d(sBl + 8192, &mapBl, bar, n0 + 64, k0);
        }
        if (++stage == XT_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // Bh and Bl are adjacent in the stage, so ONE N = 256 UMMA multiplies Ah by both (columns [0, 128) of the
      // accumulator collect Ah*Bh, [128, 256) Ah*Bl: Ah is read from shared memory once instead of twice); a second
      // N = 128 UMMA adds Al*Bh to the first half.  The epilogue adds the halves.
      constexpr uint32_t idesc = make_idesc(DW, B_MN, BM), idesc2 = make_idesc(DW, B_MN, 2 * BM);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t st = tiles + stage * XT_STAGE_BYTES;
        const uint32_t sBh = st + 2 * MT * XT_TILE;                         // Bl follows Bh
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint32_t a_off = DW ? ks * 2048u : ks * 32u;
          const uint32_t b_off = B_MN ? ks * 2048u : ks * 32u;
          const uint32_t a_lbo = DW ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
          const uint64_t bh = make_desc(sBh + b_off, b_lbo, 1024u);          // also the base of [Bh | Bl]
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sAh = st + mt * 2 * XT_TILE, sAl = sAh + XT_TILE;
            const uint64_t ah = make_desc(sAh + a_off, a_lbo, 1024u);
            const uint64_t al = make_desc(sAl + a_off, a_lbo, 1024u);
            umma_bf16(tmem_base + mt * 256, ah, bh, idesc2, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_bf16(tmem_base + mt * 256, al, bh, idesc, 1u);
          }
        }
        umma_commit(empty0 + 8 * stage);
        if (++stage == XT_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===================== expanders: tile-packed X -> dense swizzled (hi, lo) tiles =====================
    // Three groups of four warps; group g fills the stages kb = g, g + 3, ... (always ring slot g) so that a warp
    // has three MMA stage times per sub-tile.  Everything a stage needs is in registers before the stage is touched: the pointer
    // quadruples of 64 upcoming sub-tiles are lane-cached (one coalesced load per 32 stages, broadcast by
    // shuffle), the entries of the NEXT sub-tile are loaded while the current one is expanded (two register
    // buffers, loop unrolled by two so that no register waits on a load just to be moved).
    const int e8 = warp - 2, grp = e8 / Cfg::EXP_WARPS, eg = e8 % Cfg::EXP_WARPS;
    const int emt = eg >> 2, e = eg & 3;                    // row tile of the CTA, 32-row strip (or sub-tile) inside it
    const int mt0 = m0 + emt * BM;
    const int n_i = (nkb - grp + XT_GROUPS - 1) / XT_GROUPS;   // stages of this group
    // sub-tile of this warp in its i-th stage, or -1
    auto sub_tile = [&](int i) -> int {
      const int kb_abs = kb_begin + grp + XT_GROUPS * i;
      if (i >= n_i || kb_abs >= kb_end) return -1;
      int strip, kbx;
      if (!DW) { strip = (mt0 >> 5) + e; kbx = kb_abs; }
      else     { strip = 2 * kb_abs + (e >> 1); kbx = (mt0 >> 6) + (e & 1); }
      return (strip < x.nstrips && kbx < x.nkb) ? strip * x.nkb + kbx : -1;
    };
    // this warp's 32 rows of 128 bytes inside the stage's Ah tile (Al follows at + XT_TILE)
    const uint32_t my_rows = static_cast<uint32_t>(emt * 2 * XT_TILE) +
                             (DW ? static_cast<uint32_t>((e & 1) * 8192 + (e >> 1) * 4096) : static_cast<uint32_t>(e * 4096));

    struct Win { int e0, e1, t; };
    auto load_win = [&](int base_i) -> Win {
      Win w{0, 0, -1};
      w.t = sub_tile(base_i + lane);
      if (w.t >= 0) { w.e0 = __ldg(x.ptr + w.t); w.e1 = __ldg(x.ptr + w.t + 1); }
      return w;
    };
    int win_base = 0;
    Win wa = load_win(0), wb = load_win(32);
    auto get_ptr = [&](int i, int& t) -> SubPtr {           // i in [win_base, win_base + 64), warp-uniform
      const int j = i - win_base;
      const bool hi = j >= 32;
      const int e0 = __shfl_sync(0xffffffffu, hi ? wb.e0 : wa.e0, j & 31);
      const int e1 = __shfl_sync(0xffffffffu, hi ? wb.e1 : wa.e1, j & 31);
      t = __shfl_sync(0xffffffffu, hi ? wb.t : wa.t, j & 31);
      if (hi) {                                             // the first window is used up: slide
        wa = wb; win_base += 32;
        wb = load_win(win_base + 32);
      }
      return SubPtr{e0, e1 - e0};
    };

    auto do_stage = [&](int i, const SubPtr& p, const SubData& d) {
      const int stage = grp;                               // kb = grp + XT_GROUPS * i
      const uint32_t phase = static_cast<uint32_t>(i) & 1u;
      mbar_wait(empty0 + 8 * stage, phase ^ 1u);
      const uint32_t base_hi = tiles + stage * XT_STAGE_BYTES + my_rows, base_lo = base_hi + XT_TILE;
#pragma unroll
      for (int z = 0; z < 8; ++z) {
        const uint32_t a = (z * 32 + lane) * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base_hi + a), "r"(0u) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base_lo + a), "r"(0u) : "memory");
      }
      __syncwarp();
      // the entries were requested one stage ago; nothing below may be scheduled above this point (the compiler
      // would otherwise start combining the registers of the loads it has just issued for the NEXT stage)
      uint32_t code[XT_MAXC];
      float val[XT_MAXC];
      uint4 sg = d.seg;
      asm volatile("" : "+r"(sg.x), "+r"(sg.y), "+r"(sg.z), "+r"(sg.w));
#pragma unroll
      for (int j = 0; j < XT_MAXC; ++j) {
        code[j] = d.code[j]; val[j] = d.val[j];
        asm volatile("" : "+r"(code[j]), "+f"(val[j]));
      }
#pragma unroll
      for (int j = 0; j < XT_MAXC; ++j) {
        const uint32_t k = j * 32 + lane;
        if (k < static_cast<uint32_t>(p.n)) xt_put(base_hi, base_lo, (xt_high(sg, k) << 8) | code[j], val[j]);
      }
      // a sub-tile with more than 32 * XT_MAXC entries (> 12.5 % non-zero): the rest straight from memory
      for (int k = 32 * XT_MAXC + lane; k < p.n; k += 32)
        xt_put(base_hi, base_lo, (xt_high(sg, k) << 8) | __ldg(x.codes + p.e0 + k), __ldg(x.vals + p.e0 + k));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the UMMA
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * stage) : "memory");
    };

    int ta, tb;
    SubPtr pa = get_ptr(0, ta), pb;
    SubData da, db;
    xt_load(x, pa, ta, lane, da);
    for (int i = 0; i < n_i; i += 2) {
      pb = get_ptr(i + 1, tb);
      xt_load(x, pb, tb, lane, db);                         // entries of stage i + 1 while stage i is expanded
      do_stage(i, pa, da);
      if (i + 1 >= n_i) break;
      pa = get_ptr(i + 2, ta);
      xt_load(x, pa, ta, lane, da);
      do_stage(i + 1, pb, db);
    }

    // ===================== epilogue: TMEM -> registers -> global =====================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    float* cbase = C + blockIdx.z * split_stride;
    // the warps of a quarter share its MT x four 32-column chunks
#pragma unroll 1
    for (int task = e8 >> 2; task < 4 * MT; task += Cfg::EXP_WARPS * XT_GROUPS / 4) {
      const int mt = task >> 2, c = task & 3;
      const int row = m0 + mt * BM + q * 32 + lane;
      uint32_t r[32], r2[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mt * 256 + c * 32, r);
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mt * 256 + BM + c * 32, r2);
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      const int col0 = n0 + c * 32;
      if (OUT_T) {                                // C[n][m]: consecutive lanes write consecutive addresses
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) cbase[static_cast<int64_t>(col0 + j) * ldc + row] = __uint_as_float(r[j]);
        }
      } else if (row < M) {
        float* crow = cbase + static_cast<int64_t>(row) * ldc + n0;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(crow) & 15u) == 0) && (ldc % 4 == 0);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

template <bool DW, bool B_MN, bool OUT_T, int MT>
int launch_xt(const XtView& x, const CUtensorMap& bh, const CUtensorMap& bl, float* C, int64_t ldc, int64_t M,
              int64_t N, int64_t K, int splits, int64_t split_stride, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    GDA_CUDA(cudaFuncSetAttribute(k_gemm_xt<DW, B_MN, OUT_T, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  XtCfg<MT>::SMEM));
    attr_set = true;
  }
  const int nkb = static_cast<int>(ceil_div(K, BK));
  const int kbps = static_cast<int>(ceil_div(nkb, splits));
  dim3 grid(static_cast<unsigned>(ceil_div(N, BM)), static_cast<unsigned>(ceil_div(M, MT * BM)),
            static_cast<unsigned>(ceil_div(nkb, kbps)));
  k_gemm_xt<DW, B_MN, OUT_T, MT><<<grid, XtCfg<MT>::THREADS, XtCfg<MT>::SMEM, st>>>(x, bh, bl, C, ldc, (int)M, (int)N, (int)K, kbps,
                                                               split_stride);
  GDA_LAUNCH_CHECK();
  return GDA_OK;
}

// K splits of the weight-gradient form: fill the 148 SMs in whole waves (every CTA does the same work)
bool xt_mt2() {                                             // GDA_XT_MT1=1: one row tile per CTA (A/B experiments)
  static const bool on = [] { const char* e = std::getenv("GDA_XT_MT1"); return !(e && e[0] == '1'); }();
  return on;
}

int xt_dw_splits(int64_t tiles, int64_t nkb) {
  int best = 1;
  double best_eff = 0.0;
  const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(64, nkb / 16));
  for (int64_t s = 1; s <= cap; ++s) {
    const int64_t ctas = tiles * s;
    const double eff = static_cast<double>(ctas) / static_cast<double>(ceil_div(ctas, kNumSMs) * kNumSMs);
    if (eff > best_eff + 0.02) { best_eff = eff; best = static_cast<int>(s); }
  }
  return best;
}

int xt_view(XtView* v, const float* vals, const void* codes, const int32_t* ptr, const void* seg,
            int64_t rows, int64_t cols) {
  GDA_REQUIRE(rows >= 1 && cols >= 1 && rows < (int64_t(1) << 31) && cols < (int64_t(1) << 31), "gda_gemm_xt: bad shape");
  GDA_REQUIRE(vals && codes && ptr && seg, "gda_gemm_xt: NULL pointer in the tile-packed matrix");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(seg) % 16 == 0, "gda_gemm_xt: seg must be 16-byte aligned");
  v->vals = vals; v->codes = static_cast<const uint8_t*>(codes); v->ptr = ptr; v->seg = static_cast<const uint4*>(seg);
  v->nkb = static_cast<int>(ceil_div(cols, 64));
  v->nstrips = static_cast<int>(ceil_div(ceil_div(rows, 32), 4) * 4);
  GDA_REQUIRE(static_cast<int64_t>(v->nstrips) * v->nkb < (int64_t(1) << 31), "gda_gemm_xt: too many sub-tiles");
  return GDA_OK;
}

}  // namespace

int64_t xt_ptr_entries(int64_t rows, int64_t cols) {
  return ceil_div(ceil_div(rows, 32), 4) * 4 * ceil_div(cols, 64) + 1;
}

int gemm_xt_fwd(const float* vals, const void* codes, const int32_t* ptr, const void* seg, int64_t rows,
                int64_t cols, int transB, int64_t N, const void* b_hi, const void* b_lo, int64_t ldb, float* C,
                int64_t ldc, cudaStream_t st) {
  XtView x;
  int rc;
  if ((rc = xt_view(&x, vals, codes, ptr, seg, rows, cols))) return rc;
  GDA_REQUIRE(N >= 1 && b_hi && b_lo && C && ldc >= N, "gda_gemm_xt_fwd: bad dense operand / output");
  GDA_REQUIRE(ldb % 8 == 0 && ldb >= (transB ? cols : N), "gda_gemm_xt_fwd: ldb must be a multiple of 8 and cover a row");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(b_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(b_lo) % 16 == 0,
              "gda_gemm_xt_fwd: operands must be 16-byte aligned");
  GDA_REQUIRE(ceil_div(rows, BM) <= 65535, "gda_gemm_xt_fwd: too many row tiles");
  CUtensorMap bh, bl;
  // transB: B stored [N, cols] (K-major, box 64 x 128); else [cols, N] (MN-major, boxes 64 x 64)
  const int64_t b_in = transB ? cols : N, b_out = transB ? N : cols;
  const int b_box = transB ? BM : BK;
  if ((rc = make_map(&bh, b_hi, b_in, b_out, ldb, b_box)) || (rc = make_map(&bl, b_lo, b_in, b_out, ldb, b_box)))
    return rc;
  // two row tiles per CTA (one pass over the dense operand per 256 rows) once that still fills the SMs
  const bool mt2 = xt_mt2() && ceil_div(rows, 2 * BM) * ceil_div(N, BM) >= kNumSMs;
  if (transB) return mt2 ? launch_xt<false, false, false, 2>(x, bh, bl, C, ldc, rows, N, cols, 1, 0, st)
                         : launch_xt<false, false, false, 1>(x, bh, bl, C, ldc, rows, N, cols, 1, 0, st);
  return mt2 ? launch_xt<false, true, false, 2>(x, bh, bl, C, ldc, rows, N, cols, 1, 0, st)
             : launch_xt<false, true, false, 1>(x, bh, bl, C, ldc, rows, N, cols, 1, 0, st);
}

// weight-gradient plan: row tiles per CTA and K splits
struct XtDwPlan { int mt, splits; };
XtDwPlan xt_dw_plan(int64_t rows, int64_t cols, int64_t N) {
  XtDwPlan p;
  // two feature tiles per CTA measured SLOWER here (503 vs 460 us at config 2: 27 tiles x splits leave only two
  // ring slots per CTA and the node-range splits already keep G's re-reads low); GDA_XT_DW_MT2=1 selects it
  static const bool dw2 = [] { const char* e = std::getenv("GDA_XT_DW_MT2"); return e && e[0] == '1'; }();
  p.mt = (dw2 && xt_mt2() && cols >= 4 * BM) ? 2 : 1;
  p.splits = xt_dw_splits(ceil_div(cols, p.mt * BM) * ceil_div(N, BM), ceil_div(rows, BK));
  return p;
}

int64_t gemm_xt_dw_workspace_bytes(int64_t rows, int64_t cols, int64_t N) {
  const int s = xt_dw_plan(rows, cols, N).splits;
  return s > 1 ? static_cast<int64_t>(s) * cols * N * sizeof(float) : 0;
}

int splitk_reduce(const float* part, int splits, int64_t M, int64_t N, float alpha, float beta, float* C, int64_t ldc,
                  cudaStream_t st);   // gemm_simt.cu

int gemm_xt_dw(const float* vals, const void* codes, const int32_t* ptr, const void* seg, int64_t rows,
               int64_t cols, int64_t N, const void* g_hi, const void* g_lo, int64_t ldg, int out_t, float* C,
               int64_t ldc, void* ws, int64_t ws_bytes, cudaStream_t st) {
  XtView x;
  int rc;
  if ((rc = xt_view(&x, vals, codes, ptr, seg, rows, cols))) return rc;
  GDA_REQUIRE(N >= 1 && g_hi && g_lo && C, "gda_gemm_xt_dw: bad dense operand / output");
  GDA_REQUIRE(ldc >= (out_t ? cols : N), "gda_gemm_xt_dw: ldc too small");
  GDA_REQUIRE(ldg % 8 == 0 && ldg >= N, "gda_gemm_xt_dw: ldg must be a multiple of 8 and >= N");
  GDA_REQUIRE(reinterpret_cast<uintptr_t>(g_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(g_lo) % 16 == 0,
              "gda_gemm_xt_dw: operands must be 16-byte aligned");
  CUtensorMap gh, gl;
  if ((rc = make_map(&gh, g_hi, N, rows, ldg, BK)) || (rc = make_map(&gl, g_lo, N, rows, ldg, BK))) return rc;
  const int64_t nkb = ceil_div(rows, BK);
  const XtDwPlan plan = xt_dw_plan(rows, cols, N);
  const int splits = plan.splits;
  float* out = C;
  int64_t out_ld = ldc, stride = 0;
  // partial results in the orientation of the output: [N, cols] when out_t, else [cols, N]
  const int64_t pr = out_t ? N : cols, pc = out_t ? cols : N;
  if (splits > 1) {
    const int64_t need = static_cast<int64_t>(splits) * cols * N * sizeof(float);
    if (!ws || ws_bytes < need) return fail(GDA_E_WORKSPACE, "gda_gemm_xt_dw: workspace too small");
    out = static_cast<float*>(ws); out_ld = pc; stride = pr * pc;
  }
  if (plan.mt == 2) {
    if (out_t) rc = launch_xt<true, true, true, 2>(x, gh, gl, out, out_ld, cols, N, rows, splits, stride, st);
    else rc = launch_xt<true, true, false, 2>(x, gh, gl, out, out_ld, cols, N, rows, splits, stride, st);
  } else {
    if (out_t) rc = launch_xt<true, true, true, 1>(x, gh, gl, out, out_ld, cols, N, rows, splits, stride, st);
    else rc = launch_xt<true, true, false, 1>(x, gh, gl, out, out_ld, cols, N, rows, splits, stride, st);
  }
  if (rc) return rc;
  if (splits > 1) {
    const int kbps = static_cast<int>(ceil_div(nkb, splits));
    const int used = static_cast<int>(ceil_div(nkb, kbps));
    return splitk_reduce(static_cast<float*>(ws), used, pr, pc, 1.f, 0.f, C, ldc, st);
  }
  return GDA_OK;
}

}  // namespace gda

extern "C" {

int64_t gda_xt_ptr_entries(int64_t rows, int64_t cols) { return gda::xt_ptr_entries(rows, cols); }

int gda_gemm_xt_fwd(const float* vals, const void* codes, const int32_t* ptr, const void* seg,
                    int64_t rows, int64_t cols, int transB, int64_t N, const void* b_hi, const void* b_lo,
                    int64_t ldb, float* C, int64_t ldc, gda_stream_t stream) {
  if (rows == 0 || N == 0) return GDA_OK;
  return gda::gemm_xt_fwd(vals, codes, ptr, seg, rows, cols, transB, N, b_hi, b_lo, ldb, C, ldc,
                          gda::as_stream(stream));
}

int64_t gda_gemm_xt_dw_workspace_bytes(int64_t rows, int64_t cols, int64_t N) {
  return gda::gemm_xt_dw_workspace_bytes(rows, cols, N);
}

int gda_gemm_xt_dw(const float* vals, const void* codes, const int32_t* ptr, const void* seg,
                   int64_t rows, int64_t cols, int64_t N, const void* g_hi, const void* g_lo, int64_t ldg,
                   int out_transposed, float* C, int64_t ldc, void* workspace, int64_t workspace_bytes,
                   gda_stream_t stream) {
  return gda::gemm_xt_dw(vals, codes, ptr, seg, rows, cols, N, g_hi, g_lo, ldg, out_transposed, C, ldc,
                         workspace, workspace_bytes, gda::as_stream(stream));
}

}  // extern "C"
