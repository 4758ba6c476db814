// TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the row-wise kernels of clip_lite_b200/csrc compiled for the CPU and
// exported with a C interface for tests/test_emu_kernels.py.  Pointers are HOST pointers.  Kernel, variant, block
// size and columns-per-thread are chosen by the same functions the C ABI uses (jsd_heads.cuh: ln_fwd_variant /
// ln_bwd_plan / ln_*_select); only the number of blocks of the backward is a parameter here (the product derives
// it from the occupancy), so that the grid-stride walk over the rows is exercised with few and with many blocks.
#include "jsd_heads.cuh"
#include "../../include/jsd_b200.h"   // the product's prototypes: the same-named exports below must match them

using namespace jsd;

namespace {

template <typename T>
int run_ln_fwd(const LnNormJob& job, int count, int rows, long long D, bool out_bf16, int force_variant) {
  uintptr_t bits = 0;
  for (int i = 0; i < count; ++i)
    bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.out[i]) |
            reinterpret_cast<uintptr_t>(job.w[i]) | reinterpret_cast<uintptr_t>(job.b[i]);
  int variant = ln_fwd_variant(D, (bits & 15) == 0);
  if (force_variant >= 0) {
    if (force_variant < variant) return -1;          // a faster variant than the shape allows
    variant = force_variant;
  }
  const emu::Dim3 grid{(unsigned)((rows + 7) / 8), (unsigned)count, 1}, block{256, 1, 1};
  auto go = [&](auto kernel, int arg) { emu::launch(grid, block, [&]() { kernel(job, rows, arg); }); };
  if (out_bf16) ln_fwd_select<T, __nv_bfloat16>(variant, D, go);
  else ln_fwd_select<T, float>(variant, D, go);
  return variant;
}

template <typename T>
int run_ln_bwd(const LnNormBwdJob& job, int count, int rows, long long D, const float* gdiag, const float* t,
               const float* gamma, float inv_rows, int blocks, int force_scalar) {
  uintptr_t bits = (uintptr_t)job.slice_stride * 4u | reinterpret_cast<uintptr_t>(job.col_partials);
  for (int i = 0; i < count; ++i)
    bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.dX[i]) |
            reinterpret_cast<uintptr_t>(job.acc[i]) | reinterpret_cast<uintptr_t>(job.partner[i]);
  const LnBwdPlan plan = ln_bwd_plan(D, !force_scalar && (bits & 15) == 0);
  if (plan.kch <= 0) return -1;
  const emu::Dim3 grid{(unsigned)blocks, (unsigned)count, 1}, block{(unsigned)plan.threads, 1, 1};
  ln_bwd_select<T>(plan, [&](auto kernel) {
    emu::launch(grid, block, [&]() { kernel(job, rows, (int)D, gdiag, t, gamma, inv_rows); });
  });
  return plan.vec * 1000 + plan.threads * 10 + plan.kch;       // reported to the test: which variant ran
}

}  // namespace

#define EMU_DISPATCH(dtype, CALL)                         \
  switch (dtype) {                                        \
    case 0: { using T = float; return CALL; }             \
    case 1: { using T = __nv_bfloat16; return CALL; }     \
    case 2: { using T = __half; return CALL; }            \
    default: return -2;                                   \
  }

extern "C" {

// mirrors jsd_ln_normalize_pair (include/jsd_b200.h); returns the forward variant that ran (0 reg, 1 vec4, 2 scalar)
int emu_ln_normalize_pair(const void* X0, const void* X1, int dtype, long long rows, long long D, const float* w0,
                          const float* b0, float eps0, const float* w1, const float* b1, float eps1, int out_bf16,
                          void* out0, void* out1, float* stats0, float* stats1, int force_variant) {
  LnNormJob job{};
  job.X[0] = X0; job.w[0] = w0; job.b[0] = b0; job.out[0] = out0; job.eps[0] = eps0;
  job.mean[0] = stats0; job.rstd[0] = stats0 + rows; job.inv_norm[0] = stats0 + 2 * rows;
  const int count = X1 ? 2 : 1;
  if (X1) {
    job.X[1] = X1; job.w[1] = w1; job.b[1] = b1; job.out[1] = out1; job.eps[1] = eps1;
    job.mean[1] = stats1; job.rstd[1] = stats1 + rows; job.inv_norm[1] = stats1 + 2 * rows;
  }
  EMU_DISPATCH(dtype, (run_ln_fwd<T>(job, count, (int)rows, D, out_bf16 != 0, force_variant)));
}

// mirrors jsd_ln_normalize_bwd_pair; workspace = 2 * blocks * 2 * D floats
int emu_ln_normalize_bwd_pair(const void* X0, const void* X1, int dtype, long long rows, long long D, const float* w0,
                              const float* b0, const float* w1, const float* b1, const float* stats0,
                              const float* stats1, const float* acc0, const float* acc1, long long n_slices,
                              long long slice_stride, float acc_scale, const void* partner0, long long partner_offset0,
                              const void* partner1, long long partner_offset1, const float* gdiag, const float* t,
                              const float* gamma, long long M_rows, float* workspace, void* dX0, void* dX1, float* dw0,
                              float* db0, float* dw1, float* db1, float* rowdot, float* dt_out, int blocks,
                              int force_scalar) {
  const bool two = X1 != nullptr;
  LnNormBwdJob job{};
  job.X[0] = X0; job.w[0] = w0; job.b[0] = b0;
  job.mean[0] = stats0; job.rstd[0] = stats0 + rows; job.inv_norm[0] = stats0 + 2 * rows;
  job.acc[0] = acc0; job.partner[0] = (const __nv_bfloat16*)partner0; job.partner_offset[0] = partner_offset0;
  job.dX[0] = dX0;
  if (two) {
    job.X[1] = X1; job.w[1] = w1; job.b[1] = b1;
    job.mean[1] = stats1; job.rstd[1] = stats1 + rows; job.inv_norm[1] = stats1 + 2 * rows;
    job.acc[1] = acc1; job.partner[1] = (const __nv_bfloat16*)partner1; job.partner_offset[1] = partner_offset1;
    job.dX[1] = dX1;
  }
  job.slice_stride = n_slices > 1 ? slice_stride : 0;
  job.n_slices = (int)n_slices;
  job.acc_scale = acc_scale;
  job.col_partials = workspace;
  job.rowdot = rowdot;
  const int count = two ? 2 : 1;
  const float inv_rows = (float)(1.0 / (double)M_rows);
  int rc = -2;
  switch (dtype) {
    case 0: rc = run_ln_bwd<float>(job, count, (int)rows, D, gdiag, t, gamma, inv_rows, blocks, force_scalar); break;
    case 1: rc = run_ln_bwd<__nv_bfloat16>(job, count, (int)rows, D, gdiag, t, gamma, inv_rows, blocks, force_scalar); break;
    case 2: rc = run_ln_bwd<__half>(job, count, (int)rows, D, gdiag, t, gamma, inv_rows, blocks, force_scalar); break;
  }
  if (rc < 0) return rc;
  LnFinalizeJob fin{};
  fin.col_partials = workspace;
  fin.nblocks = blocks;
  fin.dw[0] = dw0; fin.db[0] = db0;
  fin.dw[1] = two ? dw1 : nullptr; fin.db[1] = two ? db1 : nullptr;
  fin.rowdot = rowdot;
  fin.rows = (int)rows;
  fin.dt_out = dt_out;
  const emu::Dim3 fgrid{(unsigned)((D + 255) / 256), (unsigned)(2 * count), 1}, fblock{256, 1, 1};
  emu::launch(fgrid, fblock, [&]() { ln_bwd_finalize_kernel(fin, (int)D); });
  return rc;
}

#ifndef EMU_HEADS_ONLY   // (the sanitizer builds instantiate the head-tail kernels only: a third of the compile time)
// ---- kernels that ARE covered by the GPU tier, run here to validate the shim itself against the oracle
int emu_normalize_cast(const void* X, int dtype, long long rows, long long D, void* Xn_bf16, float* inv_norm,
                       int variant /* 0 reg, 1 vec4, 2 scalar */) {
  NormalizeJob job{};
  job.X[0] = X; job.Xn[0] = (__nv_bfloat16*)Xn_bf16; job.inv_norm[0] = inv_norm;
  const emu::Dim3 grid{(unsigned)((rows + 7) / 8), 1, 1}, block{256, 1, 1};
  auto run = [&](auto tag) {
    using T = decltype(tag);
    if (variant == 0) emu::launch(grid, block, [&]() { normalize_cast_reg_kernel<T>(job, (int)rows, (int)(D / 128)); });
    else if (variant == 1) emu::launch(grid, block, [&]() { normalize_cast_kernel<T, 4>(job, (int)rows, (int)D); });
    else emu::launch(grid, block, [&]() { normalize_cast_kernel<T, 1>(job, (int)rows, (int)D); });
    return 0;
  };
  EMU_DISPATCH(dtype, run(T{}));
}

int emu_normalize_bwd(const void* X, int dtype, long long rows, long long D, const float* inv_norm, const float* acc,
                      const void* partner_bf16, long long partner_offset, const float* gdiag, const float* t,
                      const float* gamma, long long M_rows, void* dX, float* rowdot, int* ticket, float* dt_out,
                      int variant) {
  NormBwdJob job{};
  job.X[0] = X; job.inv_norm[0] = inv_norm; job.acc[0] = acc; job.partner[0] = (const __nv_bfloat16*)partner_bf16;
  job.partner_offset[0] = partner_offset; job.dX[0] = dX; job.rowdot = rowdot; job.ticket = ticket; job.dt_out = dt_out;
  const float inv_rows = (float)(1.0 / (double)M_rows);
  const emu::Dim3 grid{(unsigned)((rows + 7) / 8), 1, 1}, block{256, 1, 1};
  auto run = [&](auto tag) {
    using T = decltype(tag);
    if (variant == 0)
      emu::launch(grid, block, [&]() { normalize_bwd_reg_kernel<T>(job, (int)rows, (int)(D / 128), gdiag, t, gamma, inv_rows); });
    else if (variant == 1)
      emu::launch(grid, block, [&]() { normalize_bwd_kernel<T, 4>(job, (int)rows, (int)D, gdiag, t, gamma, inv_rows); });
    else
      emu::launch(grid, block, [&]() { normalize_bwd_kernel<T, 1>(job, (int)rows, (int)D, gdiag, t, gamma, inv_rows); });
    return 0;
  };
  EMU_DISPATCH(dtype, run(T{}));
}

// index mode: jsd_index_kernel + finalize_kernel as launch_index() in jsd_capi.cu runs them
int emu_index_fwd_bwd(const void* F, const void* G, int dtype, long long B, long long D, const int* neg,
                      const int* inv_ptr, const int* inv_idx, const float* t, float* coefp, float* partials,
                      float* out4, void* dF, void* dG, float grad_scale, const float* gamma, int vec) {
  const int nblk = (int)((B + INDEX_ROWS_PER_CTA - 1) / INDEX_ROWS_PER_CTA);
  const emu::Dim3 grid{(unsigned)nblk, 1, 1}, block{32 * INDEX_ROWS_PER_CTA, 1, 1};
  auto run = [&](auto tag) {
    using T = decltype(tag);
    if (vec)
      emu::launch(grid, block, [&]() {
        jsd_index_kernel<T, 4>((const T*)F, (const T*)G, (int)B, (int)D, neg, inv_ptr, inv_idx, t, coefp, partials, (T*)dF,
                               (T*)dG, grad_scale, gamma);
      });
    else
      emu::launch(grid, block, [&]() {
        jsd_index_kernel<T, 1>((const T*)F, (const T*)G, (int)B, (int)D, neg, inv_ptr, inv_idx, t, coefp, partials, (T*)dF,
                               (T*)dG, grad_scale, gamma);
      });
    const double invB = 1.0 / (double)B;
    emu::launch(emu::Dim3{1, 1, 1}, emu::Dim3{FINALIZE_THREADS, 1, 1},
                [&]() { finalize_kernel(partials, nblk, 3, invB, invB, 1.0, 0.0, out4, nullptr); });
    return 0;
  };
  EMU_DISPATCH(dtype, run(T{}));
}

#endif  // !EMU_HEADS_ONLY

// ---- the head-tail entry points under their PRODUCT names and signatures (include/jsd_b200.h), so that
// clip_lite_b200/kernels.py can be run unmodified against this library on CPU tensors (tests/test_heads_cpu.py):
// a wrong argument order in the 34-argument ctypes call shows up here, not on the GPU.  `stream` is ignored.
static int g_emu_bwd_blocks = 3;     // few blocks: every block walks several rows
void emu_set_bwd_blocks(int n) { g_emu_bwd_blocks = n > 0 ? n : 1; }

size_t jsd_ln_workspace_bytes(int64_t rows, int64_t D) {
  if (rows <= 0 || D <= 0) return 0;
  const int64_t blocks = rows < g_emu_bwd_blocks ? rows : g_emu_bwd_blocks;
  return (size_t)2 * (size_t)blocks * 2 * (size_t)D * sizeof(float);
}

int jsd_ln_normalize_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                          const float* b0, float eps0, const float* w1, const float* b1, float eps1, int out_bf16,
                          void* out0, void* out1, float* stats0, float* stats1, jsd_stream_t) {
  return emu_ln_normalize_pair(X0, X1, dtype, rows, D, w0, b0, eps0, w1, b1, eps1, out_bf16, out0, out1, stats0,
                               stats1, -1) >= 0 ? 0 : 1;
}

int jsd_ln_normalize_bwd_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                              const float* b0, const float* w1, const float* b1, const float* stats0,
                              const float* stats1, const float* acc0, const float* acc1, int64_t n_slices,
                              int64_t slice_stride, float acc_scale, const void* partner0_bf16,
                              int64_t partner_offset0, const void* partner1_bf16, int64_t partner_offset1,
                              const float* gdiag, const float* t_dev, const float* gamma_dev, int64_t M_rows,
                              void* workspace, void* dX0, void* dX1, float* dw0, float* db0, float* dw1, float* db1,
                              float* rowdot, float* dt_out, jsd_stream_t) {
  const int blocks = (int)(rows < g_emu_bwd_blocks ? rows : g_emu_bwd_blocks);
  return emu_ln_normalize_bwd_pair(X0, X1, dtype, rows, D, w0, b0, w1, b1, stats0, stats1, acc0, acc1, n_slices,
                                   slice_stride, acc_scale, partner0_bf16, partner_offset0, partner1_bf16,
                                   partner_offset1, gdiag, t_dev, gamma_dev, M_rows, (float*)workspace, dX0, dX1, dw0,
                                   db0, dw1, db1, rowdot, dt_out, blocks, 0) >= 0 ? 0 : 1;
}

#ifndef EMU_HEADS_ONLY
// index mode under its product name (the L1-based kernel; the ring-staged kernel needs the bulk-copy engine), so
// that a whole index-mode step of the drop-in module can run through kernels.py on CPU tensors
size_t jsd_index_workspace_bytes(int64_t B) { return (size_t)(B > 0 ? B : 0) * 4 * sizeof(float); }

int jsd_index_fwd_bwd(const void* F, const void* G, int dtype, int64_t B, int64_t D, const int32_t* neg_index,
                      const int32_t* inv_ptr, const int32_t* inv_idx, const float* t_dev, void* workspace,
                      float* out4, float* loss_out, void* dF, void* dG, float grad_scale, const float* gamma_dev,
                      jsd_stream_t) {
  float* coefp = (float*)workspace;
  float* partials = coefp + B;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(G) |
                                     reinterpret_cast<uintptr_t>(dF) | reinterpret_cast<uintptr_t>(dG)) & 15) == 0;
  const int rc = emu_index_fwd_bwd(F, G, dtype, B, D, neg_index, inv_ptr, inv_idx, t_dev, coefp, partials, out4, dF, dG,
                                   grad_scale, gamma_dev, vec ? 1 : 0);
  if (rc == 0 && loss_out) *loss_out = out4[2];
  return rc;
}

#endif  // !EMU_HEADS_ONLY

}  // extern "C"
