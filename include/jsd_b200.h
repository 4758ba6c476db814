/* C ABI of the B200-native JSD contrastive-loss hot path (libjsd_b200.so).
 *
 * The reference (4m4n5/CLIP-Lite) has no FFI: its boundary for this path is the
 * Python class JSDInfoMaxLoss (loss.py:110-314), registered as
 * LossFactory.PRODUCTS["jsd"] (factories.py:374-376) and called from
 * model.py:94-101.  The entry points below are what a ctypes binding inside that
 * class calls in place of the PyTorch ops of loss.py:94-105 (normalise + row dot
 * * exp(t)), loss.py:204-254 (softplus / mean / roll-by-one negatives) and their
 * autograd backward (entered from train.py:218).  See INTEGRATION.md for the stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch-allocated);
 *    matrices are row-major and contiguous unless a pitch ("ld", in elements) is given;
 *  - `stream` is a cudaStream_t; all work is enqueued asynchronously on it, the
 *    library never synchronises and keeps no mutable global state besides the
 *    last error string (thread-local);
 *  - scalars that live on the device in the reference (the `temperature`
 *    parameter, the upstream gradient) are passed as device pointers so that no
 *    call forces a host round-trip;
 *  - return value 0 = success; non-zero = error, message from jsd_last_error().
 *    The library never throws and never falls back to the CPU.
 */
#ifndef JSD_B200_H_
#define JSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* jsd_stream_t; /* cudaStream_t */

enum jsd_dtype { JSD_F32 = 0, JSD_BF16 = 1, JSD_F16 = 2 };

/* ABI version of this header (bumped on any signature change). */
int jsd_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* jsd_last_error(void);
/* Number of SMs of the current device (the persistent kernels launch one CTA each). */
int jsd_sm_count(void);

/* ------------------------------------------------------------------ index mode
 * The reference's estimator, fused forward + backward.
 *   replaces: GlobalDiscriminatorDot.forward loss.py:94-105 (called twice, :206-222),
 *             JSDInfoMaxLoss.forward normal mode loss.py:204-222,254, cluster mode
 *             loss.py:225-252 (through neg_index), SSL call sites loss.py:257-300,
 *             and the autograd backward of all of it.
 * F, G       [B, D] features AFTER the projection heads, dtype `dtype`
 * neg_index  [B] int32 column of each row's negative, or NULL for (i+1) mod B
 * inv_ptr/inv_idx  CSR inverse of neg_index ([B+1] / [B]); both NULL iff neg_index is NULL
 * t_dev      device scalar: the `temperature` parameter (tau = exp(t))
 * workspace  >= jsd_index_workspace_bytes(B) bytes
 * out4       device float[4]: {mean softplus(-s_pos), mean softplus(s_neg), their sum, dL/dt}
 * loss_out   optional (may be NULL) separate device float receiving the loss (out4[2])
 * dF, dG     [B, D] grad_scale * gamma * dL/dF and ... dL/dG (same dtype as F, G); both NULL = forward only (eval /
 *            no_grad, or the forward half of an autograd step: the write-back pass is skipped)
 * grad_scale host scalar > 0 folded into the stored gradients (1 unless the caller wants to rescale later)
 * gamma_dev  optional (may be NULL = 1) device scalar: the upstream gradient dTotal/dCROSS (GradScaler's factor
 *            included).  It is applied in fp32 inside the kernel, so fp16 gradients are rounded ONCE and never
 *            pass through the subnormal range (the per-row coefficients are sigma / B).  The autograd binding calls
 *            this entry point twice per training step: forward with dF = dG = NULL (reads F, G), backward with
 *            gamma_dev (reads F, G again, writes dF, dG): 6 B D elements of traffic instead of the 8 of "fused
 *            call + one scaling pass over dF, dG".
 */
size_t jsd_index_workspace_bytes(int64_t B);
int jsd_index_fwd_bwd(const void* F, const void* G, int dtype, int64_t B, int64_t D, const int32_t* neg_index,
                      const int32_t* inv_ptr, const int32_t* inv_idx, const float* t_dev, void* workspace,
                      float* out4, float* loss_out, void* dF, void* dG, float grad_scale, const float* gamma_dev,
                      jsd_stream_t stream);

/* ------------------------------------------------------------------ dense mode
 * All off-diagonal pairs as negatives (BASELINE.json north star); a row slab
 * [M rows] x [N columns] of the global score matrix, positives on column
 * row_offset + i.  Forward and backward are separate calls so that they sit in
 * autograd's forward / backward.
 */

/* F.normalize (loss.py:94-95) + cast: Xn = bf16(x / max(||x||, 1e-12)), inv_norm = 1 / max(||x||, 1e-12). */
int jsd_normalize_cast(const void* X, int dtype, int64_t rows, int64_t D, void* Xn_bf16, float* inv_norm,
                       jsd_stream_t stream);

/* Scratch of the dense path: 16 bytes of "last block" tickets (word 0: loss finalisation of jsd_dense_fwd,
 * word 1: dL/dt reduction of jsd_normalize_bwd) followed by the per-warp loss partials.  The tickets must be
 * ZERO when the buffer is first used; every launch leaves them zero again (the last CTA to finish does the
 * reduction -- no separate reduction launch).  One buffer per concurrently used stream. */
size_t jsd_dense_workspace_bytes(void);

/* Forward: S = tau U V^T on tcgen05 tensor cores, softplus/sigmoid epilogue, S never stored.
 * U [M, D], V [N, D] bf16 unit rows (D % 8 == 0).  Gmat (optional, NULL => loss only)
 * [M, ldg] bf16 receives sigma(S_ij) with 0 on the positives (ldg % 64 == 0, ldg >= N);
 * gdiag [M] receives -sigma(-S_ii').  out4 = {pos, neg, pos + neg, 0} of
 *   L = mean_i softplus(-S_ii') + (1 / (M (N - 1))) sum_{j != i'} softplus(S_ij);
 * loss_out (optional, may be NULL) receives a separate copy of L.  dL/dt of the dense mode comes out of the
 * backward (jsd_normalize_bwd rowdot): it is the sum over rows of <u_i, dU_i>. */
int jsd_dense_fwd(const void* U_bf16, const void* V_bf16, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                  const float* t_dev, void* Gmat_bf16, int64_t ldg, float* gdiag, void* workspace, float* out4,
                  float* loss_out, jsd_stream_t stream);

/* Stream-K: the backward contractions can cut tiles along the contraction so that every CTA pair has work.
 * Passing `sk_workspace` (NULL = never) allows it; the single-call backward entry points below use it only when the
 * tiles fill at most half of the GPU (e.g. dU of a 1024-row rank: 16 tiles for 74 CTA pairs, 4.5x faster split),
 * jsd_dense_bwd_du/dv and jsd_gemm_bf16 additionally to balance a ragged last wave.
 * Workspace of the backward contractions (stream-K partial tiles + hand-off flags).  The first
 * jsd_streamk_flag_bytes() bytes must be ZERO when the buffer is first used; every launch leaves
 * them zero again.  One buffer per concurrently used stream. */
size_t jsd_streamk_workspace_bytes(void);
size_t jsd_streamk_flag_bytes(void);

/* Backward contractions (tensor cores), off-diagonal part, fp32 out; operands are read as they
 * lie in memory (MN-major UMMA descriptors), nothing is transposed:
 *   dUacc [M, D] = gamma tau / (M (N-1)) * Gmat   . V     (V [N, D] bf16 unit rows)
 *   dVacc [N, D] = gamma tau / (M (N-1)) * Gmat^T . U     (U [M, D] bf16 unit rows)
 * gamma_dev: device scalar upstream gradient (NULL => 1).  sk_workspace may be NULL (no stream-K). */
int jsd_dense_bwd_du(const void* Gmat_bf16, int64_t ldg, const void* V_bf16, int64_t M, int64_t N, int64_t D,
                     const float* t_dev, const float* gamma_dev, void* sk_workspace, float* dUacc,
                     jsd_stream_t stream);
int jsd_dense_bwd_dv(const void* Gmat_bf16, int64_t ldg, const void* U_bf16, int64_t M, int64_t N, int64_t D,
                     const float* t_dev, const float* gamma_dev, void* sk_workspace, float* dVacc,
                     jsd_stream_t stream);

/* Positive-pair term + Jacobian of F.normalize:
 *   d_row = acc_row + gamma tau / M_rows * gdiag[row] * partner[row + partner_offset]
 *   dX_row = (d_row - u_row <u_row, d_row>) * inv_norm[row],   u_row = X_row * inv_norm[row]
 * X, dX [rows, D] in `dtype`; acc fp32 [rows, D]; partner bf16 [*, D]; gdiag may be NULL (no positive term).
 * rowdot (optional, may be NULL) [rows] receives <u_row, d_row>.  dt_out (optional; needs rowdot and the forward's
 * `workspace`, whose second ticket word serialises it) receives sum_rows rowdot = gamma * dL/dt when X are the
 * image rows: the last block of the same launch reduces the row dots in a fixed order (fp64, deterministic). */
int jsd_normalize_bwd(const void* X, int dtype, int64_t rows, int64_t D, const float* inv_norm, const float* acc,
                      const void* partner_bf16, int64_t partner_offset, const float* gdiag, const float* t_dev,
                      const float* gamma_dev, int64_t M_rows, void* dX, float* rowdot, void* workspace, float* dt_out,
                      jsd_stream_t stream);

/* Single-GPU convenience (M == N == B, row_offset 0): the whole forward, resp. the whole backward, in
 * one call, so that the host pays one FFI crossing per autograd direction (2 + 3 kernel launches per step).
 *   forward : jsd_normalize_cast_pair(F, G), jsd_dense_fwd (loss reduced by its last CTA)
 *   backward: jsd_dense_bwd_du, then the image-side jsd_normalize_bwd (+ dt_out) on a library-owned helper
 *             stream next to jsd_dense_bwd_dv (forked / joined with events: legal under CUDA-graph capture), then
 *             the text-side jsd_normalize_bwd; with JSD_OVERLAP=0: both Jacobians in one launch after the GEMMs
 * F, G [B, D] in `dtype`; U, V bf16 [B, D]; acc_u, acc_v fp32 [B, D] and rowdot fp32 [B] scratch;
 * workspace = the forward's; dF, dG in `dtype`; dt_out = gamma * dL/dt. */
int jsd_dense_forward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const float* t_dev, void* U_bf16,
                      void* V_bf16, float* inv_f, float* inv_g, void* Gmat_bf16, int64_t ldg, float* gdiag,
                      void* workspace, float* out4, float* loss_out, jsd_stream_t stream);
int jsd_dense_backward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const void* U_bf16,
                       const void* V_bf16, const float* inv_f, const float* inv_g, const void* Gmat_bf16, int64_t ldg,
                       const float* gdiag, const float* t_dev, const float* gamma_dev, float* acc_u, float* acc_v,
                       float* rowdot, void* workspace, void* sk_workspace, void* dF, void* dG, float* dt_out,
                       jsd_stream_t stream);

/* Fused single-pass forward + backward for D in {64, 128, 192, 256} (single GPU, M == N == B): ONE tensor-core
 * launch in which each score tile stays in TMEM, sigma(S) goes through shared memory straight back into the tensor
 * cores and the gradient accumulators stay in TMEM (2 x 128 + D <= 512 columns) -- neither S nor sigma(S) reaches
 * HBM and no B x B buffer exists.  The text side runs as a second set of CTAs of the same launch with the
 * modalities swapped (the score tiles are recomputed: 8 B^2 D executed FLOPs for 6 B^2 D algorithmic ones).
 *   jsd_dense_fused_supported  1 iff (B, D) can take this path
 *   jsd_dense_fused_splits     S = CTAs per 128-row block (1..8): small batches are cut along the other modality so
 *                              that the launch fills the GPU; every accumulator below then has S slices
 *   jsd_dense_fused_fwd_bwd    U, V [B, D] bf16 unit rows -> loss (out4 / loss_out as jsd_dense_fwd), gdiag [B],
 *                              acc_u / acc_v [S, B, D] fp32 whose sum over S is the sum over the negatives of
 *                              sigma(S_ij) v_j resp. u_i, NOT yet scaled (autograd has no upstream gradient when
 *                              its forward runs)
 *   jsd_dense_fused_forward    jsd_normalize_cast_pair + jsd_dense_fused_fwd_bwd in one call
 *   jsd_dense_fused_backward   both Jacobians of F.normalize in one launch: applies gamma tau / (B (B - 1)) to the
 *                              accumulators and adds the positive-pair term in fp32, dt_out = gamma * dL/dt
 * workspace = jsd_dense_workspace_bytes() bytes, as for jsd_dense_fwd. */
int jsd_dense_fused_supported(int64_t B, int64_t D);
int jsd_dense_fused_splits(int64_t B, int64_t D);
int jsd_dense_fused_fwd_bwd(const void* U_bf16, const void* V_bf16, int64_t B, int64_t D, const float* t_dev,
                            float* acc_u, float* acc_v, float* gdiag, void* workspace, float* out4, float* loss_out,
                            jsd_stream_t stream);
int jsd_dense_fused_forward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const float* t_dev,
                            void* U_bf16, void* V_bf16, float* inv_f, float* inv_g, float* acc_u, float* acc_v,
                            float* gdiag, void* workspace, float* out4, float* loss_out, jsd_stream_t stream);
int jsd_dense_fused_backward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const void* U_bf16,
                             const void* V_bf16, const float* inv_f, const float* inv_g, const float* gdiag,
                             const float* t_dev, const float* gamma_dev, const float* acc_u, const float* acc_v,
                             float* rowdot, void* workspace, void* dF, void* dG, float* dt_out, jsd_stream_t stream);

/* Row-slab (multi-GPU) convenience calls: one FFI crossing each.
 *   jsd_normalize_cast_pair       = jsd_normalize_cast of F and of G in one launch
 *   jsd_dense_backward_image_side = jsd_dense_bwd_du, jsd_normalize_bwd (positives at column row_offset + i of
 *                                   V_all, row dots, dt_out = gamma * dL_r/dt) */
int jsd_normalize_cast_pair(const void* F, const void* G, int dtype, int64_t rows, int64_t D, void* U_bf16,
                            void* V_bf16, float* inv_f, float* inv_g, jsd_stream_t stream);
int jsd_dense_backward_image_side(const void* F, int dtype, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                                  const void* V_all_bf16, const float* inv_f, const void* Gmat_bf16, int64_t ldg,
                                  const float* gdiag, const float* t_dev, const float* gamma_dev, float* acc_u,
                                  float* rowdot, void* workspace, void* sk_workspace, void* dF, float* dt_out,
                                  jsd_stream_t stream);

/* ------------------------------------------------------------------ projection-head tail (LayerNorm + normalise)
 * replaces: the nn.LayerNorm that ends MILinearBlock.forward (loss.py:36-38) fused with the F.normalize of
 *           GlobalDiscriminatorDot.forward (loss.py:94-95), forward and backward.
 * X0 / X1   [rows, D] head outputs BEFORE LayerNorm (`dtype`): image head / text head; X1 (with out1, stats1, ...)
 *           may be NULL = one row set only
 * w, b      LayerNorm weight / bias, fp32 [D]; NULL = 1 / 0 (elementwise_affine=False).  eps = LayerNorm's eps.
 *
 * jsd_ln_normalize_pair      u = LN(x) / max(||LN(x)||, 1e-12), one pass over each row:
 *     out_bf16 = 0: out fp32 [rows, D]  (input of jsd_index_fwd_bwd or of any other consumer of unit rows)
 *     out_bf16 = 1: out bf16 [rows, D]  (the U / V operands of jsd_dense_fwd, jsd_dense_fused_fwd_bwd)
 *     stats [3, rows] fp32: mean | rstd | 1 / max(||LN(x)||, 1e-12)  -- all the backward keeps of LayerNorm's output
 * jsd_ln_normalize_bwd_pair  one pass: dU = acc_scale' * sum_s acc[s] + gamma tau / M_rows * gdiag[row] * partner[row
 *     + partner_offset] (as jsd_normalize_bwd), Jacobian of F.normalize, LayerNorm backward:
 *     acc       fp32 [n_slices][rows, D], slices `slice_stride` elements apart, summed in order.  The gradient with
 *               respect to the unit rows (index mode: dF / dG of jsd_index_fwd_bwd on the unit rows), the scaled
 *               accumulators of jsd_dense_bwd_du / _dv (acc_scale = 0: taken as they are), or the unscaled slices
 *               of jsd_dense_fused_fwd_bwd (acc_scale = 1 / (B (B - 1)): multiplied by gamma * tau * acc_scale)
 *     gdiag     NULL = no positive-pair term (partner, t_dev, M_rows unused unless acc_scale > 0 needs t_dev)
 *     dX        [rows, D] in `dtype`: gradient of the head output;  dw, db fp32 [D] (may be NULL): LayerNorm's
 *               weight / bias gradients, summed over the rows per block in registers and over the blocks in a
 *               fixed order by a second small launch (deterministic, no atomics)
 *     rowdot    optional [rows]: <u_row, dU_row> of row set 0; dt_out (optional, needs rowdot) = their sum
 *               = gamma * dL/dt of the dense estimator
 *     workspace >= jsd_ln_workspace_bytes(rows, D) bytes (per-block column partials; no initialisation needed)
 * D <= 4096 when rows are 16-byte aligned and D % 4 == 0, else D <= 1024 (the heads' width is 2048). */
size_t jsd_ln_workspace_bytes(int64_t rows, int64_t D);
int jsd_ln_normalize_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                          const float* b0, float eps0, const float* w1, const float* b1, float eps1, int out_bf16,
                          void* out0, void* out1, float* stats0, float* stats1, jsd_stream_t stream);
int jsd_ln_normalize_bwd_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                              const float* b0, const float* w1, const float* b1, const float* stats0,
                              const float* stats1, const float* acc0, const float* acc1, int64_t n_slices,
                              int64_t slice_stride, float acc_scale, const void* partner0_bf16,
                              int64_t partner_offset0, const void* partner1_bf16, int64_t partner_offset1,
                              const float* gdiag, const float* t_dev, const float* gamma_dev, int64_t M_rows,
                              void* workspace, void* dX0, void* dX1, float* dw0, float* db0, float* dw1, float* db1,
                              float* rowdot, float* dt_out, jsd_stream_t stream);

/* ------------------------------------------------------------------ peer-memory exchange (multi-GPU dense path)
 * One process per GPU; the two exchange steps of the sharded path (all-gather of the text rows, sum of
 * the dV partials on the owner rank) are fused into the kernels that produce / consume the data, over NVLink peer
 * memory, instead of separate NCCL collectives:
 *   jsd_peer_normalize_push      the local half: normalises F -> U, G -> this rank's row block of its OWN gathered V
 *                                buffer, and advances the buffer's step counter
 *   jsd_peer_dense_fwd           jsd_dense_fwd on (U, this rank's gathered V) with the all-gather fused into it:
 *                                a spare warp of every CTA copies the rank's own rows into the other ranks'
 *                                buffers (rank - 1, rank - 2, ... in turn, one "rows of rank r are in" flag per
 *                                destination) while the other warps already score the own column block; every
 *                                tile waits only for the flags of the ranks whose rows it loads, and the walk goes
 *                                rank, rank + 1, ... -- the order in which the blocks arrive.  Must directly follow
 *                                jsd_peer_normalize_push of the same parity on the same stream.
 *   jsd_peer_dense_bwd_dv        dV partial over all text rows.  partials_bf16 = 1: the contraction's epilogue
 *                                pushes every output tile as bf16 by TMA store into its OWNER's slot buffer (owner
 *                                blocks walked from rank + 1 on, one flag per owner) -- the reduce-scatter's traffic
 *                                is hidden behind the contraction; partials_bf16 = 0: the fp32 partial stays in this
 *                                rank's peer-mapped buffer and the last CTA publishes "complete" to every rank
 *   jsd_peer_normalize_bwd_text  waits for every rank's flag, sums the world partials of its rows in rank order
 *                                (deterministic; bf16: from its local slots, fp32: read over NVLink), then
 *                                positive-pair term and Jacobian of F.normalize
 * (the image side uses jsd_dense_backward_image_side with V_all = v_all[parity][rank]).
 * Buffers come from jsd_peer_alloc (cudaMalloc, zero-filled) and are mapped into the other processes with
 * jsd_peer_export / jsd_peer_open (CUDA IPC).  The gathered V buffer is double-buffered by the step's parity
 * (the caller alternates 0, 1, 0, ...; every rank must use the same sequence), so a rank may start pushing step
 * k+1 while a slower rank still reads step k.  Flags only grow (per-buffer push counters): nothing is reset.
 * Every rank must make the same sequence of calls (collective semantics), one step in flight at a time, and the
 * ranks must stay within the wait limit of each other (jsd_peer_set_timeout, default 300 s): a wait that expires
 * records which flag it was waiting for (jsd_peer_wait_error) and traps. */
#define JSD_MAX_PEERS 8
#define JSD_PEER_HANDLE_BYTES 64
/* layout of a rank's flag block (int32 words) */
#define JSD_PEER_READY_V 0        /* [parity][source rank]: pushes of that rank into this rank's V buffer */
#define JSD_PEER_READY_DV 16      /* [source rank]: dV partial launches of that rank (for this rank's rows) */
#define JSD_PEER_COUNTER_V 24     /* [parity] this rank's own push counter */
#define JSD_PEER_COUNTER_DV 26
#define JSD_PEER_TICKET_DV 28     /* CTAs of the dV launch that have finished */
#define JSD_PEER_TICKET_PUSH 32   /* [destination slot]: CTAs of the forward launch done pushing to that destination */
#define JSD_PEER_FLAG_INTS 64

typedef struct jsd_peer_ctx {
  int32_t rank, world;
  int64_t rows;                          /* rows per rank (the same on every rank) */
  int64_t dim;                           /* D */
  void* v_all[2][JSD_MAX_PEERS];         /* [parity][q]: rank q's gathered V [world * rows, D] bf16, as mapped here */
  void* stage[JSD_MAX_PEERS];            /* rank q's gradient staging, world * rows * D * 4 bytes, as mapped here:
                                            fp32 route: rank q's own partial [world * rows, D] fp32;
                                            bf16 route: [world][rows, D] bf16, slot s = the partial pushed by rank s */
  int32_t* flags[JSD_MAX_PEERS];         /* rank q's flag block, zero before the first step */
} jsd_peer_ctx;

size_t jsd_peer_flag_bytes(void);
int jsd_peer_alloc(size_t bytes, void** out_ptr);
int jsd_peer_free(void* ptr);
int jsd_peer_export(void* ptr, void* handle64);
int jsd_peer_open(const void* handle64, void** out_ptr);
int jsd_peer_close(void* ptr);
/* Time limit of every wait on a peer's flag (seconds; default 300, or the environment variable JSD_PEER_TIMEOUT_S).
 * jsd_peer_wait_error: non-zero iff a wait of this process has expired; kind 1 = a forward waiting for the text rows
 * of rank *index, 2 = a text-side Jacobian waiting for the gradient partial of rank *index; *target = the step
 * count it was waiting for.  The word lives in host memory: it can be read after the trap poisoned the context. */
int jsd_peer_set_timeout(double seconds);
int jsd_peer_wait_error(int* kind, int* index, int* target);
int jsd_peer_normalize_push(const void* F, const void* G, int dtype, const jsd_peer_ctx* ctx, int parity, void* U_bf16,
                            float* inv_f, float* inv_g, jsd_stream_t stream);
int jsd_peer_dense_fwd(const void* U_bf16, const jsd_peer_ctx* ctx, int parity, const float* t_dev, void* Gmat_bf16,
                       int64_t ldg, float* gdiag, void* workspace, float* out4, float* loss_out, jsd_stream_t stream);
int jsd_peer_dense_bwd_dv(const void* Gmat_bf16, int64_t ldg, const void* U_bf16, const jsd_peer_ctx* ctx,
                          const float* t_dev, const float* gamma_dev, int partials_bf16, jsd_stream_t stream);
int jsd_peer_normalize_bwd_text(const void* G, int dtype, const jsd_peer_ctx* ctx, const float* inv_g,
                                const void* U_bf16, const float* gdiag, const float* t_dev, const float* gamma_dev,
                                int partials_bf16, void* dG, jsd_stream_t stream);

/* Whole backward of a peer-exchange step in one call, as two chains forked behind the forward: on the caller's
 * stream the dV partial (its flags go out as early as possible) and then the text-side Jacobian that waits for the
 * peers and sums their partials; on a library-owned helper stream the dU contraction (split-K when the rank's rows
 * underfill the GPU) and the image-side Jacobian (+ dt_out = gamma * dL_r/dt).  Joined before returning.  acc_u fp32 [rows, D] and rowdot fp32 [rows] are scratch; workspace =
 * the forward's; sk_workspace = jsd_streamk_workspace_bytes() bytes (split-K slices). */
int jsd_peer_dense_backward(const void* F, const void* G, int dtype, const jsd_peer_ctx* ctx, int parity,
                            const void* U_bf16, const float* inv_f, const float* inv_g, const void* Gmat_bf16,
                            int64_t ldg, const float* gdiag, const float* t_dev, const float* gamma_dev,
                            int partials_bf16, float* acc_u, float* rowdot, void* workspace, void* sk_workspace,
                            void* dF, void* dG, float* dt_out, jsd_stream_t stream);

/* ------------------------------------------------------------------ retrieval / zero-shot scoring
 * replaces: retrieval.py:143 (`sims_matrix = image_embeds @ text_embeds.t()`, copied to the host) + the NumPy
 * argsort loops of itm_eval (retrieval.py:163-187), and zero_shot.py:155 (`torch.max(features @ prompts.t(), 1)`).
 * S = A B^T runs on the tensor-core kernel and is never stored.
 *
 * jsd_split_bf16x3: fp32-faithful scores from bf16 tensor cores.  out [rows, 3 D] bf16 = (hi, hi, lo) for side 0 and
 *   (hi, lo, hi) for side 1, hi = bf16(x), lo = bf16(x - hi); then <A'_i, B'_j> = <a,b> up to ~2^-17.  normalize != 0
 *   first scales each row to unit L2 norm (F.normalize).  Use the outputs as A / B below with K = 3 D.
 * jsd_score_ranks: rank_row[i] = #{j : S_ij > max_{t in targets(i)} S_it}  (image -> text rank of the best ground
 *   truth: CSR row_tgt_ptr [M+1] / row_tgt_idx), rank_col[j] = #{i : S_ij > S_{col_tgt[j], j}} (text -> image;
 *   col_tgt [N], -1 = none).  Either side may be omitted (NULL).  Two passes over the same tiles: the first extracts
 *   the target scores, the second counts, so a target is compared with bit-identical arithmetic (no tie margin).
 *   thr_row_scratch [M] / thr_col_scratch [N]: 4-byte scratch per row / column.
 * jsd_score_argmax: best[i] = max_j ((ordered bits of S_ij) << 32 | (0xFFFFFFFF - j)): the row maximum and, on
 *   ties, the smallest column; column = 0xFFFFFFFF - (best & 0xFFFFFFFF). */
int jsd_split_bf16x3(const void* X, int dtype, int64_t rows, int64_t D, int side, int normalize, void* out_bf16,
                     jsd_stream_t stream);
int jsd_score_ranks(const void* A_bf16, const void* B_bf16, int64_t M, int64_t N, int64_t K, const int32_t* row_tgt_ptr,
                    const int32_t* row_tgt_idx, const int32_t* col_tgt, void* thr_row_scratch, float* thr_col_scratch,
                    int32_t* rank_row, int32_t* rank_col, jsd_stream_t stream);
int jsd_score_argmax(const void* A_bf16, const void* B_bf16, int64_t M, int64_t N, int64_t K, unsigned long long* best,
                     jsd_stream_t stream);

/* Development aid: device-side event trace.  With a buffer installed (events != NULL; *count zeroed by the caller),
 * one thread of every kernel appends (kernel id << 60 | event << 56 | %globaltimer ns) at entry (event 0), once the
 * flags of its peers are in (1, peer-exchange consumers only) and at exit (2).  Kernel ids: 1 normalise, 2 forward
 * GEMM, 3 backward GEMM, 4 Jacobian, 5 index, 6 score, 7 normalise + push.  events == NULL switches it off (the
 * default; the cost is one load in one thread per launch).  Synchronous (cudaMemcpyToSymbol); affects the current
 * device only; not for use during a CUDA-graph capture. */
int jsd_trace_enable(unsigned long long* events, int capacity, int* count);

/* Plain C [M, N] fp32 = A . B^T on the same tcgen05 kernel, every operand-layout combination:
 * A [M, K] bf16 (a_mn_major = 0) or A^T [K, lda] (a_mn_major = 1); B [N, K] (b_mn_major = 0) or
 * B^T [K, ldb] (b_mn_major = 1).  sk_workspace as above (NULL => whole tiles only).
 * Exposed for the unit tests of the tensor-core path. */
int jsd_gemm_bf16(const void* A_bf16, int64_t lda, int a_mn_major, const void* B_bf16, int64_t ldb, int b_mn_major,
                  int64_t M, int64_t N, int64_t K, void* sk_workspace, float* C, jsd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JSD_B200_H_ */
