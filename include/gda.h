/* libgda -- C ABI of the B200-native PyGDA hot path (sm_100a).
 *
 * The reference (pygda-team/pygda) is pure Python and has NO FFI/plugin
 * interface; its hot path reaches native code only through torch / PyG /
 * torch_scatter operators.  Each entry point below therefore names the
 * reference operator call site (file:line under /root/reference) it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the
 * reference side.
 *
 * Conventions
 *  - plain C symbols, `int` return: 0 = ok, negative = GDA_E_* ; never throws or
 *    aborts across the ABI.  `gda_last_error()` returns a thread-local message.
 *  - every buffer is a CALLER-ALLOCATED DEVICE pointer (the caller's allocator --
 *    torch in this repo -- owns all memory).  The library retains nothing beyond
 *    the call except the opaque `gda_graph_t` (CSR/CSC + weights), which is
 *    explicitly created and destroyed.
 *  - every call takes the CUDA stream to run on (`void*` == cudaStream_t) and is
 *    asynchronous with respect to the host unless stated otherwise.
 *  - re-entrant; no hidden global state.  A `gda_graph_t` may be used from
 *    several host threads as long as calls on it are ordered on one stream at a
 *    time (PyTorch's autograd thread uses the forward's stream).
 *  - row-major matrices with explicit leading dimensions in ELEMENTS.
 */
#ifndef GDA_H_
#define GDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDA_VERSION 100          /* 0.1.0 */

/* error codes */
#define GDA_OK              0
#define GDA_E_INVALID      -1    /* bad argument (null pointer, negative size, bad flag) */
#define GDA_E_CUDA         -2    /* CUDA runtime error; see gda_last_error() */
#define GDA_E_INDEX        -3    /* edge_index entry outside [0, N) */
#define GDA_E_UNSUPPORTED  -4    /* shape/dtype combination not built */
#define GDA_E_WORKSPACE    -5    /* workspace too small */

typedef void* gda_stream_t;      /* cudaStream_t */
typedef struct gda_graph gda_graph_t;
typedef struct gda_edges gda_edges_t;   /* device edge list produced by the library (TDSS smoothing graph) */

int         gda_version(void);
int         gda_sm_arch(void);   /* 100: built for sm_100a only */
const char* gda_last_error(void);
/* number of kernels this library has launched in this process (statistics for bench.py) */
uint64_t    gda_launch_count(void);

/* ------------------------------------------------------------------ graph --
 * gda_graph_create replaces, in one device pass, what the reference recomputes
 * on every conv call:
 *   gcn_norm            pygda/nn/prop_gcn_conv.py:24-81   (GDA_NORM_SYM_COL)
 *   CachedGCNConv.norm  pygda/nn/cached_gcn_conv.py:63-103 (GDA_NORM_SYM_ROW)
 *   add_remaining_self_loops / scatter_add call sites therein (:71-78 / :95-99)
 * plus the COO -> CSR (by destination) and CSC (by source) builds that the
 * aggregation kernels need.  Index results are bit-exact with the reference
 * (kept edges in order, then loops 0..N-1); within a CSR row entries keep COO
 * order, so the fp32 summation order equals a sequential scatter_add.
 *
 * edge_index: device int64 [2, E] (row 0 = source, row 1 = target), edge_weight:
 * device float [E] or NULL (= ones).  Synchronises `stream` once (it has to read
 * the surviving edge count). */
#define GDA_SELF_LOOPS    1      /* add_remaining_self_loops */
#define GDA_IMPROVED      2      /* fill value 2 instead of 1 */
#define GDA_NORM_SYM_COL  4      /* D^-1/2 A D^-1/2, degree summed at the target index */
#define GDA_NORM_SYM_ROW  8      /* same, degree summed at the source index */
#define GDA_SKIP_EMPTY_ROWS 16   /* rows without non-zeros are NOT written by the aggregation (their output keeps its
                                    content): rectangular local blocks of a partitioned graph, whose trailing "rows" are
                                    halo slots filled by other GPUs (pygda_b200/dist.py, halo mode); H = 128 fp32 only */
int gda_graph_create(const int64_t* edge_index, int64_t E, int64_t N,
                     const float* edge_weight, int flags, gda_stream_t stream,
                     gda_graph_t** out);
int gda_graph_destroy(gda_graph_t* g);
int gda_graph_info(const gda_graph_t* g, int64_t* N, int64_t* nnz,
                   int64_t* num_long_rows, int64_t* num_long_rows_t);
/* self-looped, normalised COO in the reference's order: edge_index_out int64 [2,nnz],
 * weight_out float [nnz]  (what gcn_norm returns, prop_gcn_conv.py:81) */
int gda_graph_export_coo(const gda_graph_t* g, int64_t* edge_index_out, float* weight_out,
                         gda_stream_t stream);
/* CSR of A_hat (transpose=0: rows = targets, cols = sources) or of A_hat^T
 * (transpose=1): rowptr int32 [N+1], colidx int32 [nnz], vals float [nnz] */
int gda_graph_export_csr(const gda_graph_t* g, int transpose, int32_t* rowptr,
                         int32_t* colidx, float* vals, gda_stream_t stream);

/* ------------------------------------------------------------ aggregation --
 * Y = A_hat * X (transpose=0) or A_hat^T * X (transpose=1, the backward of the
 * former).  Replaces MessagePassing.propagate = index_select -> mul ->
 * scatter_add_ at pygda/nn/prop_gcn_conv.py:208-210 (+message :238) and
 * pygda/nn/cached_gcn_conv.py:138 (+message :156), and its autograd backward.
 * X [N,H] (ldx), Y [N,H] (ldy), fp32; X and Y must not alias.
 * Optional fused epilogue on Y: + bias[H] (prop_gcn_conv.py:212-213 /
 * cached_gcn_conv.py:171-172), then ReLU, then inverted dropout with keep mask
 * hash(seed + *seed_offset, row*H+col) (a2gnn_base.py:136-138).  bias may be NULL;
 * seed_offset is a device uint64 (or NULL = 0) so that a captured CUDA graph draws
 * a fresh mask on every replay.
 * workspace: device scratch of gda_spmm_workspace_bytes(); may be NULL when that
 * is 0 (no rows longer than the split threshold). */
#define GDA_EPI_RELU      1
#define GDA_EPI_DROPOUT   2
int64_t gda_spmm_workspace_bytes(const gda_graph_t* g, int transpose, int H);
int gda_spmm_f32(const gda_graph_t* g, int transpose, const float* X, int64_t ldx,
                 float* Y, int64_t ldy, int H, const float* bias, int epi_flags,
                 float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                 void* workspace, int64_t workspace_bytes, gda_stream_t stream);
/* A_hat^k X in one call (k launches): what `for i in range(prop_nums): out = self.propagate(...)`
 * does at pygda/nn/prop_gcn_conv.py:208-210.  T0/T1: ping-pong scratch [N,H] (k>=2 / k>=3);
 * the epilogue applies to the last step only. */
int gda_spmm_k_f32(const gda_graph_t* g, int transpose, int k, const float* X, int64_t ldx, float* Y,
                   int64_t ldy, float* T0, float* T1, int H, const float* bias, int epi_flags,
                   float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                   int64_t workspace_bytes, gda_stream_t stream);
/* The same over nb feature matrices that share the graph: matrix b is X + b*x_batch_stride ->
 * Y + b*y_batch_stride (elements).  Used for the repeated bottleneck evaluations of one domain per
 * step (pygda/models/a2gnn.py:192-193 and :181/:211 run feat_bottleneck twice on the same graph with
 * fresh dropout masks): one (colidx, weight) read and one dependency chain per row serve nb gathers.
 * The dropout mask of matrix b is that of rows [b*N, (b+1)*N) of the stacked [nb*N, H] matrix;
 * workspace: gda_spmm_workspace_bytes(g, transpose, H * nb). */
int gda_spmm_nb_f32(const gda_graph_t* g, int transpose, int nb, const float* X, int64_t ldx,
                    int64_t x_batch_stride, float* Y, int64_t ldy, int64_t y_batch_stride, int H,
                    const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                    const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, gda_stream_t stream);
/* The factored form for UNIT-weight symmetric normalisation (gcn_norm with edge_weight=None,
 * pygda/nn/prop_gcn_conv.py:67-81: w_e = dinv[src] * 1 * dinv[dst]):  A_hat = D (A + I) D, so
 * A_hat^k x = D S (D^2 S)^(k-1) D x with S = A + I.  gda_row_scale_f32 computes z = D x once,
 * gda_spmm_unw_nb_f32 one step out[r] = scale[r] * sum_{c in row r} z[c] (scale = dinv^2, or dinv when
 * `last`), with the epilogue of gda_spmm_f32 -- no per-edge weight is read or multiplied.  Values equal
 * the weighted kernel's up to fp32 rounding.  gda_graph_unit_weights: 1 when the graph / width / nb
 * qualify (H = 128 fp32, nb 1 or 2).  gda_spmm_k_nb_f32 takes this route by itself for k >= 3. */
int gda_graph_unit_weights(const gda_graph_t* g, int transpose, int H, int nb);
int gda_row_scale_f32(const gda_graph_t* g, int nb, const float* X, int64_t ldx, int64_t x_batch_stride, float* Z,
                      int64_t ldz, int64_t z_batch_stride, int H, gda_stream_t stream);
/* Rectangular LOCAL blocks of a partitioned graph (halo mode of the multi-GPU path, pygda_b200/dist.py): the block of
 * a rank keeps the normalised weights of the WHOLE graph, whose degrees the library cannot recompute from the block.
 * gda_graph_export_dinv copies deg^-1/2 of a normalised graph ([N] floats); gda_graph_set_unit_dinv installs such a
 * vector (the rank's own rows; halo rows unused) on a block whose weights are dinv[row] * dinv[col] of a unit-weight
 * graph, which makes the block eligible for the factored chain (gda_graph_unit_weights / gda_spmm_unw_nb_f32): halo
 * rows arrive already scaled by their owners.  gda_row_scale_rows_f32 = gda_row_scale_f32 over the first `rows` rows
 * only (the rank's own rows; the halo slots are written by the peers). */
int gda_graph_export_dinv(const gda_graph_t* g, float* dinv_out, gda_stream_t stream);
int gda_graph_set_unit_dinv(gda_graph_t* g, const float* dinv, gda_stream_t stream);
int gda_row_scale_rows_f32(const gda_graph_t* g, int64_t rows, int nb, const float* X, int64_t ldx,
                           int64_t x_batch_stride, float* Z, int64_t ldz, int64_t z_batch_stride, int H,
                           gda_stream_t stream);
int gda_spmm_unw_nb_f32(const gda_graph_t* g, int transpose, int nb, const float* Z, int64_t ldz, int64_t z_batch_stride,
                        float* Y, int64_t ldy, int64_t y_batch_stride, int H, int last, const float* bias, int epi_flags,
                        float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                        int64_t workspace_bytes, gda_stream_t stream);
int gda_spmm_k_nb_f32(const gda_graph_t* g, int transpose, int k, int nb, const float* X, int64_t ldx,
                      int64_t x_batch_stride, float* Y, int64_t ldy, int64_t y_batch_stride, float* T0, float* T1,
                      int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                      const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes, gda_stream_t stream);
/* bf16 features, fp32 edge weights and accumulation (BASELINE config 3) */
int gda_spmm_bf16(const gda_graph_t* g, int transpose, const void* X, int64_t ldx,
                  void* Y, int64_t ldy, int H, const float* bias, int epi_flags,
                  float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                  void* workspace, int64_t workspace_bytes, gda_stream_t stream);

/* ---------------------------------------------------- multi-GPU (peer path) --
 * 1-D node (row-block) partition over the GPUs of one NVSwitch box (the
 * reference is single-device).  gda_graph_partition keeps rows [row_lo, row_hi) of a graph built
 * with self loops and re-encodes columns as (owner rank << 28 | row inside the owner's block),
 * owner = col / rows_per_rank.  gda_spmm_peer_f32 is gda_spmm_f32 on such a partition with the
 * input matrix given as one pointer PER RANK (peer_x[q] = rank q's [rows_per_rank, H] block,
 * CUDA-IPC mapped): neighbour rows owned by other GPUs are read over NVLink inside the kernel,
 * so the exchange overlaps the gather/FMA work and only referenced rows cross the fabric.
 * gda_sym_alloc / gda_sym_open give the IPC-shareable buffers; gda_peer_barrier is a device-side
 * all-ranks barrier (flags in peer memory) that orders the k propagation steps across GPUs:
 * peer_flags[q] = rank q's uint64[num_peers] flag array (zero-initialised), epoch strictly
 * increasing; *error_flag (device int, may be NULL) is set if a peer does not arrive in ~10 s. */
#define GDA_MAX_PEERS 8
int gda_graph_partition(const gda_graph_t* g, int64_t row_lo, int64_t row_hi, int64_t rows_per_rank,
                        gda_stream_t stream, gda_graph_t** out);
int gda_spmm_peer_f32(const gda_graph_t* part, int transpose, const void* const* peer_x, int num_peers,
                      int my_rank, int64_t ldx, float* Y, int64_t ldy, int H, const float* bias,
                      int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                      void* workspace, int64_t workspace_bytes, gda_stream_t stream);
/* barrier, copy-in, then k x (barrier, gda_spmm_peer_f32) ping-ponging between the two symmetric
 * buffers sym0/sym1 (per-rank pointer arrays); barrier epochs epoch0+1 .. epoch0+k+1 are used. */
int gda_spmm_peer_k_f32(const gda_graph_t* part, int transpose, int k, const float* x_local,
                        const void* const* sym0, const void* const* sym1, int num_peers, int my_rank,
                        float* Y, int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                        const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes,
                        uint64_t* const* peer_flags, uint64_t epoch0, int* error_flag, gda_stream_t stream);
/* The same with the barrier epoch kept in DEVICE memory (*epoch_dev, a uint64 owned by the rank, advanced by
 * the barrier kernel itself): no launch argument depends on host state, so the partitioned training step can be
 * captured in a CUDA graph and replayed.  gda_peer_barrier_dev is the barrier alone. */
int gda_spmm_peer_k_dev_f32(const gda_graph_t* part, int transpose, int k, const float* x_local,
                            const void* const* sym0, const void* const* sym1, int num_peers, int my_rank,
                            float* Y, int H, const float* bias, int epi_flags, float dropout_p, uint64_t seed,
                            const uint64_t* seed_offset, void* workspace, int64_t workspace_bytes,
                            uint64_t* const* peer_flags, uint64_t* epoch_dev, int* error_flag, gda_stream_t stream);
int gda_peer_barrier_dev(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t* epoch_dev,
                         int* error_flag, gda_stream_t stream);
/* PUSH mode for partitions with little locality (a random graph over P GPUs references (P-1)/P of its columns on other
 * GPUs): every rank keeps a full local copy of the (padded) global matrix in a symmetric GATHER buffer
 * [num_peers * rows_per_rank, H]; an aggregation gathers from that local copy and stores each output row into block
 * `my_rank` of the next gather buffer on EVERY rank (P2P stores over NVLink inside the kernel), so the exchange is
 * fused into the producer, overlaps its gathers and moves each row once per peer as coalesced 512-byte stores --
 * instead of one remote 512-byte load per referencing non-zero (gda_spmm_peer_f32).  gda_spmm_push_f32: one step
 * (gather_local = this rank's copy; out_blocks[0..num_out) = where each output row goes, pre-offset to this rank's
 * block); gda_spmm_push_k_f32: barrier, copy-in to every rank, k x (barrier, step), the last step writing Y. */
int gda_spmm_push_f32(const gda_graph_t* part, int transpose, const void* gather_local, void* const* out_blocks,
                      int num_out, int num_peers, int64_t ldx, int64_t ldy, int H, const float* bias, int epi_flags,
                      float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                      int64_t workspace_bytes, gda_stream_t stream);
int gda_spmm_push_k_f32(const gda_graph_t* part, int transpose, int k, const float* x_local, void* const* gbuf0,
                        void* const* gbuf1, int num_peers, int my_rank, float* Y, int H, const float* bias,
                        int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset, void* workspace,
                        int64_t workspace_bytes, uint64_t* const* peer_flags, uint64_t* epoch_dev, int* error_flag,
                        gda_stream_t stream);
/* HALO mode for partitions WITH locality (few remote columns): the local block is a rectangular graph whose trailing
 * columns are halo slots; after every step each rank copies the rows other ranks reference into their halo slots
 * (gda_push_rows_f32: entry i = row rows[i] of src -> row slots[i] of rank peer[i]'s buffer, P2P stores), so that the
 * aggregation itself runs on local memory with the ordinary kernels.  Each referenced row crosses NVLink once per
 * consumer and step instead of once per referencing non-zero. */
/* gda_spmm_halo_f32: gda_spmm_f32 on such a local block with the halo exchange FUSED into the producer: row r is
 * stored to Y and, for every rank q with bit q set in halo_mask[r], into row halo_slot[r * num_peers + q] of
 * peer_base[q] (P2P stores issued as soon as the row is reduced, overlapping the gathers of the rows in flight). */
int gda_spmm_halo_f32(const gda_graph_t* block, const float* X, int64_t ldx, float* Y, int64_t ldy, int H,
                      const int32_t* halo_mask, const int32_t* halo_slot, void* const* peer_base, int num_peers,
                      const float* bias, int epi_flags, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                      void* workspace, int64_t workspace_bytes, gda_stream_t stream);
/* gda_spmm_unw_halo_f32: one INTERMEDIATE step z' = D^2 S z of the factored chain on such a block (made eligible by
 * gda_graph_set_unit_dinv), nb = 1 or 2 stacked matrices, with the same fused halo exchange; matrix b of the stack
 * lives peer_batch_stride[q] * b elements into rank q's buffer (the ranks' blocks differ in size). */
int gda_spmm_unw_halo_f32(const gda_graph_t* block, int nb, const float* Z, int64_t ldz, int64_t z_batch_stride, float* Y,
                          int64_t ldy, int64_t y_batch_stride, int H, const int32_t* halo_mask, const int32_t* halo_slot,
                          void* const* peer_base, const int64_t* peer_batch_stride, int num_peers, void* workspace,
                          int64_t workspace_bytes, gda_stream_t stream);
int gda_push_rows_f32(const float* src, int64_t ld, const int32_t* rows, const int32_t* slots, const int32_t* peer,
                      int64_t count, void* const* peer_base, int num_peers, int64_t peer_ld, int H,
                      gda_stream_t stream);
int gda_sym_alloc(int64_t bytes, void** ptr, unsigned char* handle_out /* 64 bytes */);
int gda_sym_open(const unsigned char* handle /* 64 bytes */, void** ptr);
int gda_sym_close(void* ptr);
int gda_sym_free(void* ptr);
int gda_peer_barrier(uint64_t* const* peer_flags, int rank, int num_peers, uint64_t epoch, int* error_flag,
                     gda_stream_t stream);

/* ------------------------------------------------------------ dense GEMM --
 * C[M,N] = alpha * op(A) * op(B) + beta * C, fp32 row-major, op = transpose when
 * the flag is set (A is [M,K] or [K,M]; B is [K,N] or [N,K]).
 * Replaces `self.lin(x)` (prop_gcn_conv.py:205, PyG Linear = x @ W^T),
 * `torch.matmul(x, self.weight)` (cached_gcn_conv.py:130), the nn.Linear heads
 * (a2gnn_base.py:67,70; udagcn_base.py:155-162; grade_base.py:63-74) and their
 * autograd backward GEMMs.  Large aligned shapes run on tcgen05 tensor cores
 * with a split-bf16 (3-term) decomposition that keeps fp32-level accuracy
 * (DESIGN.md section 4); everything else on a SIMT fp32 kernel. */
int64_t gda_gemm_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K);
int gda_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                 float* C, int64_t ldc, void* workspace, int64_t workspace_bytes,
                 gda_stream_t stream);

/* Split-bf16 operands for the tensor-core path.  gda_split_bf16 writes hi = bf16(x) and
 * lo = bf16(x - hi) as [rows, ld_out] bf16 (ld_out >= cols, multiple of 8, pad columns zeroed;
 * outputs 16-byte aligned).  gda_gemm_bf16x3 computes C = op(A) op(B) in fp32 from such pairs on
 * tcgen05 (Ah*Bh + Ah*Bl + Al*Bh, fp32 accumulation in TMEM); lda/ldb are the bf16 leading
 * dimensions (multiples of 8).  Requirements: N >= 64, K >= 64 (gda_gemm_bf16x3_supported).
 * A constant operand (the input features x, models/a2gnn.py:181-211) is split once and re-used
 * by every forward (x W^T) and backward (G^T x) pass. */
int     gda_split_bf16(const float* x, int64_t rows, int64_t cols, int64_t ldx, void* hi, void* lo,
                       int64_t ld_out, gda_stream_t stream);
int     gda_gemm_bf16x3_supported(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb);
int64_t gda_gemm_bf16x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
int     gda_gemm_bf16x3(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* a_hi,
                        const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                        float* C, int64_t ldc, void* workspace, int64_t workspace_bytes,
                        gda_stream_t stream);

/* ------------------------------ first-layer GEMMs on a SPARSE input matrix X --
 * The input features of the citation benchmarks are bag-of-words rows, a few per cent non-zero (config 2: 7 %).
 * `self.lin(x)` (prop_gcn_conv.py:205) / `torch.matmul(x, self.weight)` (cached_gcn_conv.py:130) and the weight
 * gradient of that product are evaluated from a TILE-PACKED copy of X -- 5 bytes per NON-ZERO instead of the 4 bytes
 * per element of the dense split-bf16 pair -- that is expanded into dense swizzled (hi, lo) bf16 operand tiles in
 * shared memory and multiplied on the tcgen05 tensor cores with the same three-term split as gda_gemm_bf16x3
 * (fp32-accurate: same product terms, same K order).
 *
 * Tile-packed X [rows, cols] (pygda_b200.data.PackedTiles; also the pinned host staging form of the per-step
 * `.to(device)`, a2gnn.py:311-312): sub-tile t = (row / 32) * nkb + col / 64, nkb = ceil(cols / 64); inside a
 * sub-tile the entries are sorted by p = (row % 32) * 64 + col % 64; the entries of sub-tile t are
 * [ptr[t], ptr[t+1]) of vals (fp32) and codes (uint8 = p & 255); seg[t][k], k = 0..7 (uint16, 16-byte aligned
 * rows) = number of entries of the sub-tile with p < 256 (k + 1), so that entry i has p >> 8 = #{k < 7 :
 * seg[t][k] <= i} -- an entry is decoded without looking at its neighbours.  The number of 32-row strips is padded
 * to a multiple of 4: ptr has gda_xt_ptr_entries(rows, cols) int32 entries, seg one row fewer.
 *
 * gda_gemm_xt_fwd: C[rows, N] = X B^T (transB = 1, B stored [N, cols]) or X B (transB = 0, B stored [cols, N]);
 *   B as a split-bf16 pair (gda_split_bf16, ldb in bf16 elements, multiple of 8).
 * gda_gemm_xt_dw:  C = X^T G [cols, N] (out_transposed = 0) or its transpose G^T X [N, cols] (out_transposed = 1,
 *   the layout of PyG's Linear weight); G [rows, N] as a split-bf16 pair (ldg multiple of 8).  The reduction over
 *   the rows is split into whole waves of CTAs and reduced in fixed order (deterministic); workspace of
 *   gda_gemm_xt_dw_workspace_bytes. */
int64_t gda_xt_ptr_entries(int64_t rows, int64_t cols);
int gda_gemm_xt_fwd(const float* vals, const void* codes, const int32_t* ptr, const void* seg,
                    int64_t rows, int64_t cols, int transB, int64_t N, const void* b_hi, const void* b_lo,
                    int64_t ldb, float* C, int64_t ldc, gda_stream_t stream);
int64_t gda_gemm_xt_dw_workspace_bytes(int64_t rows, int64_t cols, int64_t N);
int gda_gemm_xt_dw(const float* vals, const void* codes, const int32_t* ptr, const void* seg,
                   int64_t rows, int64_t cols, int64_t N, const void* g_hi, const void* g_lo, int64_t ldg,
                   int out_transposed, float* C, int64_t ldc, void* workspace, int64_t workspace_bytes,
                   gda_stream_t stream);

/* --------------------------------------- bf16 feature path (BASELINE config 3) --
 * "UDAGCN ... hid=256, bf16": activations and input features stored in bf16, parameters and every accumulation
 * in fp32.  gda_gemm_bf16: C = op(A) op(B) from plain bf16 operands on the tcgen05 kernel (ONE UMMA per K step
 * instead of the three of gda_gemm_bf16x3, half the operand bytes), C fp32 (weight gradients; split-K workspace
 * of gda_gemm_bf16_workspace_bytes) or bf16 (activations; out_bf16 = 1, ldc in bf16 elements).  Same shape rules
 * as gda_gemm_bf16x3.  The aggregation on bf16 rows is gda_spmm_bf16.  Casts, act/dropout (same counter-hash
 * masks as the fp32 kernels) and bias-gradient column sums on bf16 follow. */
int64_t gda_gemm_bf16_workspace_bytes(int64_t M, int64_t N, int64_t K, int out_bf16);
int gda_gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                  const void* B, int64_t ldb, void* C, int64_t ldc, int out_bf16, void* workspace,
                  int64_t workspace_bytes, gda_stream_t stream);
int gda_cast_f32_bf16(const float* x, void* y, int64_t n, gda_stream_t stream);
int gda_cast_bf16_f32(const void* x, float* y, int64_t n, gda_stream_t stream);
int gda_act_dropout_bf16_fwd(const void* x, void* y, int64_t n, int act, float dropout_p, uint64_t seed,
                             const uint64_t* seed_offset, gda_stream_t stream);
int gda_act_dropout_bf16_bwd(const void* gy, const void* y, void* gx, int64_t n, int act, float dropout_p,
                             uint64_t seed, const uint64_t* seed_offset, gda_stream_t stream);
int gda_colsum_bf16(const void* x, int64_t rows, int64_t cols, int64_t ldx, float* out, gda_stream_t stream);

/* -------------------------------------- Bernstein propagation --
 * BernProp.forward (pygda/nn/dgsda_base.py:101-153): out = sum_k C(K,k)/2^K relu(temp_k) L^k (2I-L)^(K-k) x.
 * The propagations are gda_spmm_f32 on two graphs (L and 2I-L, built with explicit weights); these two are the
 * glue between them: y = beta*y + coef*relu(*temp_k)*x with the temperature read on the device, and
 * *out = coef * [temp_k > 0] * <g, x> (fp64 accumulation; scratch: one device double). */
int gda_bern_axpy_f32(float* y, const float* x, int64_t n, float beta, float coef, const float* temp_k,
                      gda_stream_t stream);
int gda_bern_dtemp_f32(const float* g, const float* x, int64_t n, float coef, const float* temp_k, float* out,
                       double* scratch, gda_stream_t stream);

/* ------------------------------------------------------------ elementwise --
 * y = dropout(act(x + bias)) and its backward; act: 0 none, 1 relu.  The keep
 * mask is regenerated from (seed, element index): nothing is stored.
 * Replaces `self.act(x)` + `F.dropout` (a2gnn_base.py:136-138, udagcn_base.py:86-88).
 * x,y [rows, cols] contiguous; y may alias x.  bias may be NULL. */
int gda_bias_act_dropout_fwd(const float* x, const float* bias, float* y, int64_t rows,
                             int64_t cols, int act, float dropout_p, uint64_t seed,
                             const uint64_t* seed_offset, gda_stream_t stream);
/* gx = gy * act'(y) * keep/(1-p) ; `y` is the forward OUTPUT.  gbias [cols] (may be
 * NULL) receives the column sums of gx (it is overwritten, not accumulated). */
int gda_bias_act_dropout_bwd(const float* gy, const float* y, float* gx, float* gbias,
                             int64_t rows, int64_t cols, int act, float dropout_p,
                             uint64_t seed, const uint64_t* seed_offset, gda_stream_t stream);
/* `rep` independent dropout masks over ONE input: y [rep*rows, cols] stacked, copy r = dropout_r(act(x + bias)),
 * mask index r*rows*cols + element.  The reference evaluates feat_bottleneck twice per domain and step on the same
 * graph (a2gnn.py:181/192, 193/211); layer 1 is shared (dropout acts after it, a2gnn_base.py:136-138) and fans
 * out here.  Backward: gy0 / gy1 = gradient of copy 0 / 1 (NULL: none arrived), y = the stacked forward output;
 * sum = 1: gx [rows, cols] = sum of the masked copies; sum = 0: gx [rep*rows, cols] stacked (gbias over all rows). */
int gda_bias_act_dropout_rep_fwd(const float* x, const float* bias, float* y, int64_t rows, int64_t cols, int rep,
                                 int act, float dropout_p, uint64_t seed, const uint64_t* seed_offset,
                                 gda_stream_t stream);
int gda_bias_act_dropout_rep_bwd(const float* gy0, const float* gy1, const float* y, float* gx, float* gbias,
                                 int64_t rows, int64_t cols, int rep, int sum, int act, float dropout_p,
                                 uint64_t seed, const uint64_t* seed_offset, gda_stream_t stream);
/* column sums: out[cols] = sum_r x[r, :]   (bias gradients) */
int gda_colsum_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* out,
                   gda_stream_t stream);

/* ------------------------------------------------------------------ losses --
 * Mean cross-entropy of log_softmax(logits) against integer labels, forward and
 * backward in one pass.  Replaces F.nll_loss(F.log_softmax(..)) (models/a2gnn.py:182),
 * F.cross_entropy on the domain logits (a2gnn.py:204, udagcn.py:180-187, grade.py:176).
 * labels: int64 [rows] or NULL; when NULL the label of row r is (r >= split) --
 * the reference's `[0]*N_s + [1]*N_t` host list (a2gnn.py:200-202) folded into the
 * kernel.  loss_out: device float[1] (overwritten).  dlogits (may be NULL):
 * d(mean CE)/dlogits, [rows, C].  C <= 64. */
int gda_softmax_ce_fwd_bwd(const float* logits, int64_t rows, int C, int64_t ld,
                           const int64_t* labels, int64_t split, float* loss_out,
                           float* dlogits, gda_stream_t stream);
/* Target-entropy term of UDAGCN (models/udagcn.py:193-199): mean_r sum_c -p log p with
 * p = clamp(softmax(logits), 1e-9, 1); backward into dlogits. */
int gda_softmax_entropy_fwd_bwd(const float* logits, int64_t rows, int C, int64_t ld,
                                float* loss_out, float* dlogits, gda_stream_t stream);

/* --------------------------------------------------------------------- MMD --
 * Multi-bandwidth Gaussian MMD between sampled source/target rows.
 * Replaces pygda/utils/mmd.py:4-158 (guassian_kernel :4-55, get_MMD :57-107, MMD
 * :109-158) without the n x n x d temporary.  The sample indices are INPUTS
 * (drawn by the host exactly as mmd.py:148-149 draws them).
 *  src [Ns, d] (lds), tgt [Nt, d] (ldt) fp32; src_idx / tgt_idx int64 [times, b].
 *  loss_out: device float[1] = mean over `times` of get_MMD.
 *  workspace holds, per sample, the n x n matrix (n = 2b) that the backward
 *  re-uses; keep it alive until gda_mmd_bwd has run.
 * gda_mmd_bwd ACCUMULATES grad_scale * dloss/dsrc into gsrc [Ns,d] and
 * dloss/dtgt into gtgt [Nt,d] (atomic adds: sampled rows repeat). */
int64_t gda_mmd_workspace_bytes(int times, int b, int d);
int gda_mmd_fwd(const float* src, int64_t lds, const float* tgt, int64_t ldt, int d,
                const int64_t* src_idx, const int64_t* tgt_idx, int times, int b,
                float kernel_mul, int kernel_num, float* loss_out, void* workspace,
                int64_t workspace_bytes, gda_stream_t stream);
int gda_mmd_bwd(const float* src, int64_t lds, const float* tgt, int64_t ldt, int d,
                const int64_t* src_idx, const int64_t* tgt_idx, int times, int b,
                const float* grad_scale /* device float[1] */, float* gsrc, int64_t ldgs,
                float* gtgt, int64_t ldgt, void* workspace, int64_t workspace_bytes,
                gda_stream_t stream);

/* ------------------------------------------------------------------- GAT --
 * Edge-softmax aggregation of the stock PyG GATConv(heads=1, concat=False) used by
 * pygda/nn/gnn_base.py:81-87: e_ij = leaky_relu(a_src[j] + a_dst[i]), alpha = softmax over the
 * in-edges of i (+1e-16), out_i = sum_j alpha_ij h_j.  g: graph built with GDA_SELF_LOOPS and no
 * normalisation.  alpha: float [nnz] (CSR order) written by the forward and read by the backward;
 * scratch: float [nnz].  The backward OVERWRITES dh [N,C], da_src [N], da_dst [N]. */
int gda_gat_fwd(const gda_graph_t* g, const float* h, int C, const float* a_src, const float* a_dst,
                float negative_slope, float* out, float* alpha, gda_stream_t stream);
int gda_gat_bwd(const gda_graph_t* g, const float* h, int C, const float* a_src, const float* a_dst,
                float negative_slope, const float* alpha, const float* gout, float* dh, float* da_src,
                float* da_dst, float* scratch, gda_stream_t stream);

/* ------------------------------------------------- TDSS smoothness --
 * The smoothing graph `edge_index_smooth` that TDSS.fit builds once per target graph
 * (pygda/models/tdss.py:503) and the Laplacian regulariser evaluated on it every step.
 *
 * gda_khop_create: TDSS.smoothness(smooth_mode='K-hop') (tdss.py:374-385) = (k-1) applications
 *   of TwoHopNeighbor (tdss.py:66-90: spspmm(A, A) pattern, remove_self_loops, cat with the
 *   input, coalesce) then add_remaining_self_loops.  Bit-exact: sorted distinct non-loop pairs,
 *   then loops 0..N-1 (k = 1: the input's non-loop edges in their order, then the loops).
 * gda_rw_create: smooth_mode='RW' (tdss.py:367-373): edge (v, i) for every node v on a uniform
 *   random walk of `walk_length` steps from i (walks stay put at nodes without out-edges), in
 *   row-major order, duplicates removed.  Own counter-based random stream (`seed`).
 * The result is an opaque device edge list; its size is only known after the build, so the
 * caller asks for it, allocates int64 [2, E] and exports.  Errors: GDA_E_INDEX for ids outside
 * [0, N); GDA_E_INVALID when the number of 2-paths exceeds 2^31. */
int gda_khop_create(const int64_t* edge_index, int64_t E, int64_t N, int k, gda_stream_t stream, gda_edges_t** out);
int gda_rw_create(const int64_t* edge_index, int64_t E, int64_t N, int walk_length, uint64_t seed,
                  gda_stream_t stream, gda_edges_t** out);
int64_t gda_edges_size(const gda_edges_t* edges);
int gda_edges_export(const gda_edges_t* edges, int64_t* edge_index_out /* [2, E] */, gda_stream_t stream);
int gda_edges_destroy(gda_edges_t* edges);

/* compute_laplacian_loss (tdss.py:385-449):  loss = 1/2 sum_e |f[row] d[row]^-1/2 - f[col] d[col]^-1/2|^2,
 * d = scatter_add(ones, row).  Evaluated without the [E, H] edge tensors as
 *   g = D^-1/2 f (gda_row_scale_rsqrt_f32),  u_in = A g, u_out = A^T g (gda_spmm_f32 on a graph built
 *   with flags = 0 from edge_index_smooth),  r = (d_out + d_in) g - u_in - u_out,
 *   loss = 1/2 sum_i g_i . r_i,  d loss / d f = D^-1/2 r   (gda_laplacian_finish_f32; fp64 reduction,
 *   fixed order).  gda_graph_degrees: edge counts per node as floats (out = by source = the
 *   reference's `deg`, in = by target). */
int gda_graph_degrees(const gda_graph_t* g, float* out_deg, float* in_deg, gda_stream_t stream);
int gda_row_scale_rsqrt_f32(const float* x, int64_t ldx, const float* deg, float* y, int64_t N, int H,
                            gda_stream_t stream);
int64_t gda_laplacian_workspace_bytes(void);
int gda_laplacian_finish_f32(const float* g, const float* u_in, const float* u_out, const float* out_deg,
                             const float* in_deg, int64_t N, int H, float* loss, float* df, void* workspace,
                             int64_t workspace_bytes, gda_stream_t stream);

/* ------------------------------------------------ two-view attention fusion --
 * Attention.forward (pygda/nn/attention.py:52-55) for the two views UDAGCN(ppmi=True) fuses
 * (pygda/nn/udagcn_base.py:262-264): out = a0 x0 + a1 x1, (a0, a1) = softmax(w.x0 + b, w.x1 + b).
 * x0, x1, out, gout, g0, g1: [N, H] contiguous; w [H]; b device scalar (may be NULL); a0_out / a0 [N] saved
 * for the backward.  The backward OVERWRITES g0, g1 (either may be NULL) and dw [H] (may be NULL); db = 0. */
int gda_attention2_fwd(const float* x0, const float* x1, int64_t N, int H, const float* w, const float* b,
                       float* out, float* a0_out, gda_stream_t stream);
int gda_attention2_bwd(const float* x0, const float* x1, int64_t N, int H, const float* w, const float* a0,
                       const float* gout, float* g0, float* g1, float* dw, gda_stream_t stream);

/* ------------------------------------------------- PPMI graph --
 * gda_ppmi_create replaces PPMIConv.norm up to the PPMI scores (pygda/nn/ppmi_conv.py:98-172): undirected
 * de-duplicated adjacency, `rounds` (reference: 40, :134) random walks of random length in [1, path_len] from
 * every node that has an edge, visit counters -> probabilities -> column sums ->
 * ppmi = max(log(p / colsum * #visited / path_len), 0), one weighted edge (start, visited) per visited pair
 * (zero scores kept, as the reference keeps them).  Edges come out sorted by (start, visited).  The symmetric
 * normalisation with remaining self loops (:174-184) is gda_graph_create(flags = SELF_LOOPS | NORM_SYM_ROW,
 * edge_weight = these scores).  The walks use the library's counter hash (np.random cannot be reproduced):
 * same seed, same graph.  gda_ppmi_walks exports the walks themselves, int32 [rounds, N, path_len], -1 past a
 * walk's length or for nodes without an edge (tests).  counts_out (may be NULL): visit count per edge.
 * Errors: GDA_E_INDEX for ids outside [0, N). */
typedef struct gda_wedges gda_wedges_t;
int gda_ppmi_create(const int64_t* edge_index, int64_t E, int64_t N, int path_len, int rounds, uint64_t seed,
                    gda_stream_t stream, gda_wedges_t** out);
int gda_ppmi_walks(const int64_t* edge_index, int64_t E, int64_t N, int path_len, int rounds, uint64_t seed,
                   int32_t* walks_out /* [rounds, N, path_len] */, gda_stream_t stream);
int64_t gda_wedges_size(const gda_wedges_t* edges);
int gda_wedges_export(const gda_wedges_t* edges, int64_t* edge_index_out /* [2, M] */, float* weight_out /* [M] */,
                      int32_t* counts_out /* [M] or NULL */, gda_stream_t stream);
int gda_wedges_destroy(gda_wedges_t* edges);

/* ------------------------------------------- either side of the path --
 * gda_collate_graphs: graph-mode mini-batch collation on the device -- what PyG's
 * DataLoader / Batch.from_data_list yields for `DataLoader(dataset, batch_size, shuffle=True)`
 * (pygda/models/a2gnn.py:276-286, adagcn.py:240-251): x rows of the chosen graphs concatenated,
 * edge_index re-based to the running node count, batch[i] = position of node i's graph in the batch.
 * The dataset lives in HBM as one concatenation: x_all [N_all, F], edge_index_all int64 [2, E_all] (node ids
 * global in x_all, every edge inside its graph, edges of one graph contiguous), node_ptr / edge_ptr
 * int64 [G_all + 1].  graph_ids int64 [B] = the batch, in order; out_node_ptr / out_edge_ptr int64 [B + 1] =
 * exclusive scans of the chosen graphs' node / edge counts (device), num_*_out their last entries.
 * Outputs: x_out [num_nodes_out, F], edge_index_out int64 [2, num_edges_out], batch_out int64 [num_nodes_out]. */
int gda_collate_graphs(const float* x_all, int F, const int64_t* edge_index_all, int64_t E_all,
                       const int64_t* node_ptr, const int64_t* edge_ptr, const int64_t* graph_ids, int64_t B,
                       const int64_t* out_node_ptr, const int64_t* out_edge_ptr, int64_t num_nodes_out,
                       int64_t num_edges_out, float* x_out, int64_t* edge_index_out, int64_t* batch_out,
                       gda_stream_t stream);
/* gda_unpack_rows_f32: dense out [N, F] (row pitch ldo) from a row-compressed copy -- vals fp32 [nnz], cols uint16
 * (col_bytes = 2, F <= 65536) or int32 (col_bytes = 4) [nnz], rowptr int64 [N + 1]; entries not listed are +0.0f.
 * The reference moves every batch host->device on every step (`.to(self.device)`, pygda/models/a2gnn.py:311-312);
 * for bag-of-words features (a few per cent non-zero) the pinned staging copy is kept compressed
 * (pygda_b200.data.Data.pin_memory) and only the non-zeros cross PCIe. */
/* gda_unpack_rows_delta_f32: the same from DELTA-coded column ids, one byte per entry (half the index bytes of the
 * uint16 form): per row the running column starts at 0, every byte adds its value, a byte of 255 only advances
 * (escape), any other byte emits the row's next value at the running column.  val_ptr / byte_ptr int32 [N + 1] are
 * the row offsets into vals / deltas.  *error_flag (device int, may be NULL) is set if a column lands outside F. */
int gda_unpack_rows_delta_f32(const float* vals, const uint8_t* deltas, const int32_t* val_ptr,
                              const int32_t* byte_ptr, int64_t N, int64_t F, float* out, int64_t ldo,
                              int* error_flag, gda_stream_t stream);
/* gda_unpack_tiles_f32: the same from the TILE-PACKED form of gda_gemm_xt_fwd / gda_gemm_xt_dw (sub-tiles of
 * 32 rows x 64 columns): for consumers that need the dense matrix. */
int gda_unpack_tiles_f32(const float* vals, const uint8_t* codes, const int32_t* ptr, const void* seg,
                         int64_t N, int64_t F, float* out, int64_t ldo, int* error_flag, gda_stream_t stream);
/* gda_unpack_values_f32: fp32 values from the exponent-packed staging form (pygda_b200.data.PackedTiles): value i =
 * 3 bytes of m24 (little endian: sign << 23 | mantissa) + the 4-bit code i of ecode (low nibble first): exponent =
 * meta[0] + code; code 15 marks an escape -- the first min(meta[1], esc_capacity) entries of (esc_idx, esc_val) hold
 * those values verbatim (also every zero / denormal / inf / nan).  Bit-exact for any input.  m24 holds
 * 3 * ceil(n / 4) * 4 bytes, ecode ceil(n / 4) * 2 bytes (padded with zeros); out must be 16-byte aligned. */
int gda_unpack_values_f32(const void* m24, const void* ecode, const int32_t* meta, const int32_t* esc_idx,
                          const float* esc_val, int64_t esc_capacity, int64_t n, float* out, gda_stream_t stream);
int gda_unpack_rows_f32(const float* vals, const void* cols, int col_bytes, const int64_t* rowptr, int64_t N,
                        int64_t F, float* out, int64_t ldo, gda_stream_t stream);
/* gda_argmax_confusion: pred[r] = argmax_c logits[r, c] (first maximal index) and counts[y * C + p] += 1
 * (int64 [C, C], zeroed by the call): the per-epoch training score
 * eval_micro_f1(labels, logits.argmax(dim=1)) (pygda/models/a2gnn.py:328-329, pygda/metrics/metrics.py)
 * without moving N labels and N predictions to the host.  pred_out may be NULL; *bad_label (device int) is set
 * when a label lies outside [0, C).  C <= 64. */
int gda_argmax_confusion(const float* logits, int64_t rows, int C, int64_t ld, const int64_t* labels,
                         int64_t* pred_out, int64_t* counts, int* bad_label, gda_stream_t stream);

/* ---------------------------------------------------------------- pooling --
 * global_mean_pool over a sorted `batch` vector given as CSR-style ptr
 * (int64 [G+1]).  Call sites: a2gnn_base.py:141, adagcn_base.py:94,
 * grade_base.py:154,157, models/udagcn.py:169-170. */
int gda_segment_mean_fwd(const float* x, int64_t ldx, const int64_t* ptr, int64_t G, int H,
                         float* out, gda_stream_t stream);
int gda_segment_mean_bwd(const float* gout, const int64_t* ptr, int64_t G, int H,
                         float* gx, int64_t ldgx, gda_stream_t stream);

/* --------------------------------------------------------------- optimiser --
 * torch.optim.Adam (L2 weight decay, no amsgrad) over up to GDA_ADAM_MAX_TENSORS
 * tensors in ONE launch (models/a2gnn.py:292-296,317-319).  `state` is a device
 * float[4] owned by the caller, zero-initialised: {step, bias_corr1, bias_corr2, -}.
 * gda_adam_step first advances state (so it is CUDA-graph replayable), then
 * updates p, m, v in place. */
#define GDA_ADAM_MAX_TENSORS 48
int gda_adam_step(int num_tensors, float* const* params, const float* const* grads,
                  float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                  float lr, float beta1, float beta2, float eps, float weight_decay,
                  float* state, gda_stream_t stream);

/* ------------------------------------------------------------------ misc ---
 * fill / axpy helpers so the hot path never falls back to framework kernels */
int gda_fill_f32(float* x, int64_t n, float value, gda_stream_t stream);
int gda_axpy_f32(float* y, const float* x, int64_t n, float alpha, gda_stream_t stream); /* y += alpha*x */
/* y = alpha * x : backward of GradReverse (pygda/nn/reverse_layer.py:65-66) with alpha := -alpha */
int gda_scale_f32(float* y, const float* x, int64_t n, float alpha, gda_stream_t stream);
/* y = alpha * (*alpha_dev) * x with a DEVICE scalar (loss-gradient scaling without a host sync) */
int gda_scale_dev_f32(float* y, const float* x, int64_t n, float alpha, const float* alpha_dev,
                      gda_stream_t stream);
/* *counter += 1 (device uint64): advances the dropout seed offset inside a CUDA graph */
int gda_counter_inc(uint64_t* counter, gda_stream_t stream);
/* out[0] = sum_i w[i] * (*terms[i]) for up to 8 device scalars: the loss combination
 * `loss = cls + weight * mmd` (models/a2gnn.py:183-209) without framework kernels */
int gda_combine_scalars(int n, const float* const* terms, const float* weights, float* out,
                        gda_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* GDA_H_ */
