// Tail of the projection heads fused into the estimator's row passes:
//
//   reference: MILinearBlock.forward ends in nn.LayerNorm (loss.py:36-38), GlobalDiscriminatorDot.forward then
//              applies F.normalize (loss.py:94-95); the dense path additionally casts to bf16 and keeps 1/||.||.
//              In PyTorch that is LayerNorm (read x, write y) + normalise (read y, write u) forward, and in the
//              backward the normalise Jacobian, LayerNorm's dx pass and two more passes for its weight / bias sums.
//
//   ln_normalize_*_kernel      x -> u = LN(x) / max(||LN(x)||, eps) in ONE pass over the row (fp32 unit rows for the
//                              index-mode kernel, bf16 unit rows + 1/||.|| for the tensor-core kernels), keeping
//                              mean / rstd / 1/||.|| per row (12 bytes) instead of the [B, D] LayerNorm output
//   ln_normalize_bwd_kernel    positive-pair term + Jacobian of F.normalize + LayerNorm backward in one pass:
//                              reads x and the gradient accumulator, writes dx, and sums LayerNorm's weight / bias
//                              gradients per column in registers (every thread owns fixed columns of all rows its
//                              block visits; partial sums per block, no atomics)
//   ln_bwd_finalize_kernel     fixed-order sum of the per-block column partials -> dw, db (+ the row dots -> dL/dt)
//
// Plain CUDA C++ only (warp shuffles, block barriers): the same source runs under tests/emu on the CPU.
#pragma once
#include "jsd_rowwise.cuh"

namespace jsd {

constexpr int LN_REG_CHUNKS = 16;     // register-resident forward: D = nch * 128 <= 2048 (the heads' `units`)
constexpr int LN_BWD_THREADS = 256;
constexpr int LN_BWD_MAX_KCH = 4;     // columns per thread = VEC * KCH: D <= 4096 (VEC 4) / 1024 (VEC 1)

// One launch serves up to two row sets of the same shape (blockIdx.y selects): the image and the text head.
struct LnNormJob {
  const void* X[2];          // [rows, D] head outputs BEFORE LayerNorm
  const float* w[2];         // LayerNorm weight [D] (null = 1)
  const float* b[2];         // LayerNorm bias [D] (null = 0)
  void* out[2];              // [rows, D] unit rows, OUT = float or bf16
  float* mean[2];            // [rows]
  float* rstd[2];            // [rows]
  float* inv_norm[2];        // [rows] 1 / max(||LN(x)||, 1e-12)
  float eps[2];              // LayerNorm eps
};

__device__ __forceinline__ float4 ld_param4(const float* p, int d, float dflt) {
  return p ? *reinterpret_cast<const float4*>(p + d) : make_float4(dflt, dflt, dflt, dflt);
}
__device__ __forceinline__ float ld_param1(const float* p, int d, float dflt) { return p ? p[d] : dflt; }

// any D: the row is visited four times (mean, variance, norm, write) -- after the first visit out of L1
template <typename T, typename OUT, int VEC>
__global__ void __launch_bounds__(256)
ln_normalize_kernel(const LnNormJob job, int rows, int D) {
  const bool second = blockIdx.y != 0;
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  const float* __restrict__ W = second ? job.w[1] : job.w[0];
  const float* __restrict__ Bs = second ? job.b[1] : job.b[0];
  OUT* __restrict__ out = static_cast<OUT*>(second ? job.out[1] : job.out[0]);
  const float eps = second ? job.eps[1] : job.eps[0];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* x = X + (size_t)row * D;
  const float invD = 1.f / (float)D;
  float s = 0.f;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d);
      s += (f.x + f.y) + (f.z + f.w);
    } else {
      s += to_f32(x[d]);
    }
  });
  const float mean = warp_sum(s) * invD;
  float q = 0.f;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d);
      const float a = f.x - mean, b = f.y - mean, c = f.z - mean, e = f.w - mean;
      q += a * a + b * b + c * c + e * e;
    } else {
      const float a = to_f32(x[d]) - mean;
      q += a * a;
    }
  });
  const float rstd = 1.f / sqrtf(warp_sum(q) * invD + eps);
  float ss = 0.f;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d), w = ld_param4(W, d, 1.f), b = ld_param4(Bs, d, 0.f);
      const float y0 = fmaf((f.x - mean) * rstd, w.x, b.x), y1 = fmaf((f.y - mean) * rstd, w.y, b.y),
                  y2 = fmaf((f.z - mean) * rstd, w.z, b.z), y3 = fmaf((f.w - mean) * rstd, w.w, b.w);
      ss += y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3;
    } else {
      const float y = fmaf((to_f32(x[d]) - mean) * rstd, ld_param1(W, d, 1.f), ld_param1(Bs, d, 0.f));
      ss += y * y;
    }
  });
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), kNormEps);
  OUT* o = out + (size_t)row * D;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d), w = ld_param4(W, d, 1.f), b = ld_param4(Bs, d, 0.f);
      Vec4<OUT>::store(o + d, make_float4(fmaf((f.x - mean) * rstd, w.x, b.x) * inv, fmaf((f.y - mean) * rstd, w.y, b.y) * inv,
                                          fmaf((f.z - mean) * rstd, w.z, b.z) * inv, fmaf((f.w - mean) * rstd, w.w, b.w) * inv));
    } else {
      o[d] = from_f32<OUT>(fmaf((to_f32(x[d]) - mean) * rstd, ld_param1(W, d, 1.f), ld_param1(Bs, d, 0.f)) * inv);
    }
  });
  if (lane == 0) {
    (second ? job.mean[1] : job.mean[0])[row] = mean;
    (second ? job.rstd[1] : job.rstd[0])[row] = rstd;
    (second ? job.inv_norm[1] : job.inv_norm[0])[row] = inv;
  }
}

// D = nch * 128 <= 2048, 16-byte aligned rows: the row is read from memory exactly once (a lane keeps its 4-element
// pieces in registers between the three reductions and the write), all loads of the row in flight together
template <typename T, typename OUT>
__global__ void __launch_bounds__(256)
ln_normalize_reg_kernel(const LnNormJob job, int rows, int nch) {
  const bool second = blockIdx.y != 0;
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  const float* __restrict__ W = second ? job.w[1] : job.w[0];
  const float* __restrict__ Bs = second ? job.b[1] : job.b[0];
  OUT* __restrict__ out = static_cast<OUT*>(second ? job.out[1] : job.out[0]);
  const float eps = second ? job.eps[1] : job.eps[0];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int D = nch * 128;
  const float invD = 1.f / (float)D;
  const T* x = X + (size_t)row * D + lane * 4;
  float4 xv[LN_REG_CHUNKS];
#pragma unroll
  for (int i = 0; i < LN_REG_CHUNKS; ++i)
    if (i < nch) xv[i] = Vec4<T>::load(x + i * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_REG_CHUNKS; ++i)
    if (i < nch) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
  const float mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_REG_CHUNKS; ++i)
    if (i < nch) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
    }
  const float rstd = 1.f / sqrtf(warp_sum(q) * invD + eps);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_REG_CHUNKS; ++i)
    if (i < nch) {
      const float4 w = ld_param4(W, lane * 4 + i * 128, 1.f), b = ld_param4(Bs, lane * 4 + i * 128, 0.f);
      xv[i].x = fmaf(xv[i].x * rstd, w.x, b.x);
      xv[i].y = fmaf(xv[i].y * rstd, w.y, b.y);
      xv[i].z = fmaf(xv[i].z * rstd, w.z, b.z);
      xv[i].w = fmaf(xv[i].w * rstd, w.w, b.w);
      ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
    }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), kNormEps);
  OUT* o = out + (size_t)row * D + lane * 4;
#pragma unroll
  for (int i = 0; i < LN_REG_CHUNKS; ++i)
    if (i < nch)
      Vec4<OUT>::store(o + i * 128, make_float4(xv[i].x * inv, xv[i].y * inv, xv[i].z * inv, xv[i].w * inv));
  if (lane == 0) {
    (second ? job.mean[1] : job.mean[0])[row] = mean;
    (second ? job.rstd[1] : job.rstd[0])[row] = rstd;
    (second ? job.inv_norm[1] : job.inv_norm[0])[row] = inv;
  }
}

// ------------------------------------------------------------------ backward
//   xhat = (x - mean) rstd,  y = xhat w + b,  u = y inv_norm
//   dU   = as * sum_s acc[s] + c * partner          (c = gamma tau / M_rows * gdiag[row]; 0 without gdiag)
//   dy   = (dU - u <u, dU>) inv_norm                (Jacobian of F.normalize)
//   dw  += dy xhat,  db += dy                       (summed over the rows, per column)
//   dx   = rstd (dy w - mean_d(dy w) - xhat mean_d(dy w xhat))
// The whole block works on ONE row at a time (row = blockIdx.x, + gridDim.x, ...); thread t owns the columns
// VEC * (t + blockDim.x * k), k < KCH, of every row, so the column sums live in its registers.
struct LnNormBwdJob {
  const void* X[2];
  const float* w[2];
  const float* b[2];
  const float* mean[2];
  const float* rstd[2];
  const float* inv_norm[2];
  const float* acc[2];                 // [n_slices][rows, D] fp32
  long long slice_stride;              // elements between slices
  int n_slices;                        // >= 1, summed in order
  float acc_scale;                     // > 0: the slices hold unscaled sums, multiplied by gamma * tau * acc_scale here
  const __nv_bfloat16* partner[2];     // positive-pair partner rows (bf16 unit rows); unused without gdiag
  long long partner_offset[2];
  void* dX[2];
  float* col_partials;                 // [jobs][gridDim.x][2][D]: per-block sums of dy xhat and dy
  float* rowdot;                       // job 0 only; may be null: <u, dU> per row (their sum = gamma dL/dt)
};

// Register cap: up to two pieces per thread (D <= 2048, the heads' width) the kernel fits three resident blocks of
// 256 threads per SM (<= 80 registers; ptxas: at most 24 bytes of spill) -- half as many rows again in flight as the
// 126 registers the compiler takes when left alone; the four-piece variant would spill 232 bytes there and keeps two
// resident blocks (128 registers, 16 bytes of spill).
template <typename T, int VEC, int KCH>
__global__ void __launch_bounds__(LN_BWD_THREADS, KCH <= 2 ? 3 : 2)
ln_normalize_bwd_kernel(const LnNormBwdJob job, int rows, int D, const float* __restrict__ gdiag,
                        const float* __restrict__ t_dev, const float* __restrict__ gamma_dev, float inv_rows) {
  __shared__ float scratch[2 * 32];
  const bool second = blockIdx.y != 0;
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  const float* __restrict__ W = second ? job.w[1] : job.w[0];
  const float* __restrict__ Bs = second ? job.b[1] : job.b[0];
  const float* __restrict__ mean_p = second ? job.mean[1] : job.mean[0];
  const float* __restrict__ rstd_p = second ? job.rstd[1] : job.rstd[0];
  const float* __restrict__ inv_p = second ? job.inv_norm[1] : job.inv_norm[0];
  const float* __restrict__ acc = second ? job.acc[1] : job.acc[0];
  const __nv_bfloat16* __restrict__ partner = second ? job.partner[1] : job.partner[0];
  const long long partner_offset = second ? job.partner_offset[1] : job.partner_offset[0];
  T* __restrict__ dX = static_cast<T*>(second ? job.dX[1] : job.dX[0]);
  float* __restrict__ rowdot = second ? nullptr : job.rowdot;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float invD = 1.f / (float)D;
  const float gamma = gamma_dev ? *gamma_dev : 1.f;
  const float tau = t_dev ? expf(*t_dev) : 1.f;
  const float as = job.acc_scale > 0.f ? gamma * tau * job.acc_scale : 1.f;

  // this thread's columns, LayerNorm parameters and column sums
  float wv[KCH][VEC], bv[KCH][VEC], aw[KCH][VEC], ab[KCH][VEC];
#pragma unroll
  for (int k = 0; k < KCH; ++k) {
    const int d0 = VEC * (tid + nthr * k);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      aw[k][e] = 0.f;
      ab[k][e] = 0.f;
      wv[k][e] = (d0 + e < D) ? ld_param1(W, d0 + e, 1.f) : 0.f;
      bv[k][e] = (d0 + e < D) ? ld_param1(Bs, d0 + e, 0.f) : 0.f;
    }
  }

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = mean_p[row], rstd = rstd_p[row], inv = inv_p[row];
    const float c = gdiag ? gamma * tau * inv_rows * gdiag[row] : 0.f;
    const T* x = X + (size_t)row * D;
    const float* a = acc + (size_t)row * D;
    const __nv_bfloat16* pr = gdiag ? partner + (size_t)(row + partner_offset) * D : nullptr;
    float xh[KCH][VEC], du[KCH][VEC];      // xhat, dU (later dy w)
    float red[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
      const int d0 = VEC * (tid + nthr * k);
      if (d0 < D) {                         // D % VEC == 0 for VEC 4: the whole piece is inside the row
        float xs[VEC], gs[VEC], ps[VEC];
        if constexpr (VEC == 4) {
          const float4 f = Vec4<T>::load(x + d0);
          xs[0] = f.x; xs[1] = f.y; xs[2] = f.z; xs[3] = f.w;
          float4 g = __ldcs(reinterpret_cast<const float4*>(a + d0));
          for (int s = 1; s < job.n_slices; ++s) {
            const float4 h = __ldcs(reinterpret_cast<const float4*>(a + (size_t)s * job.slice_stride + d0));
            g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
          }
          gs[0] = g.x; gs[1] = g.y; gs[2] = g.z; gs[3] = g.w;
          if (pr != nullptr) {
            const float4 p = Vec4<__nv_bfloat16>::load(pr + d0);
            ps[0] = p.x; ps[1] = p.y; ps[2] = p.z; ps[3] = p.w;
          } else {
            ps[0] = ps[1] = ps[2] = ps[3] = 0.f;
          }
        } else {
          xs[0] = to_f32(x[d0]);
          float g = a[d0];
          for (int s = 1; s < job.n_slices; ++s) g += a[(size_t)s * job.slice_stride + d0];
          gs[0] = g;
          ps[0] = pr != nullptr ? __bfloat162float(pr[d0]) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          xh[k][e] = (xs[e] - mean) * rstd;
          du[k][e] = fmaf(c, ps[e], gs[e] * as);
          red[0] += fmaf(xh[k][e], wv[k][e], bv[k][e]) * du[k][e];          // y dU
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          xh[k][e] = 0.f;
          du[k][e] = 0.f;
        }
      }
    }
    float one[1] = {red[0]};
    block_sum<1>(one, scratch);
    const float dot = one[0] * inv;          // <u, dU>
    if (rowdot != nullptr && tid == 0) rowdot[row] = dot;
    red[0] = 0.f;
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float u = fmaf(xh[k][e], wv[k][e], bv[k][e]) * inv;
        const float dy = (du[k][e] - u * dot) * inv;           // columns outside the row: xh = du = 0, w = b = 0 => 0
        aw[k][e] = fmaf(dy, xh[k][e], aw[k][e]);
        ab[k][e] += dy;
        const float g = dy * wv[k][e];
        du[k][e] = g;                                          // dy w
        red[0] += g;
        red[1] = fmaf(g, xh[k][e], red[1]);
      }
    }
    block_sum<2>(red, scratch);
    const float m1 = red[0] * invD, m2 = red[1] * invD;
    T* o = dX + (size_t)row * D;
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
      const int d0 = VEC * (tid + nthr * k);
      if (d0 < D) {
        if constexpr (VEC == 4) {
          Vec4<T>::store(o + d0, make_float4(rstd * (du[k][0] - m1 - xh[k][0] * m2), rstd * (du[k][1] - m1 - xh[k][1] * m2),
                                             rstd * (du[k][2] - m1 - xh[k][2] * m2), rstd * (du[k][3] - m1 - xh[k][3] * m2)));
        } else {
          o[d0] = from_f32<T>(rstd * (du[k][0] - m1 - xh[k][0] * m2));
        }
      }
    }
  }

  // per-block column partials: [job][block][0 = dw, 1 = db][D]
  float* part = job.col_partials + ((size_t)(second ? 1 : 0) * gridDim.x + blockIdx.x) * 2 * (size_t)D;
#pragma unroll
  for (int k = 0; k < KCH; ++k) {
    const int d0 = VEC * (tid + nthr * k);
    if (d0 < D) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(part + d0) = make_float4(aw[k][0], aw[k][1], aw[k][2], aw[k][3]);
        *reinterpret_cast<float4*>(part + D + d0) = make_float4(ab[k][0], ab[k][1], ab[k][2], ab[k][3]);
      } else {
        part[d0] = aw[k][0];
        part[D + d0] = ab[k][0];
      }
    }
  }
}

// dw[j][d] = sum over the blocks (in block order, fp64) of the column partials; likewise db.  grid = (ceil(D / 256),
// 2 * jobs): blockIdx.y = 2 * job + (0 = dw, 1 = db).  Block (0, 0) also sums the row dots into *dt_out (optional).
struct LnFinalizeJob {
  const float* col_partials;
  int nblocks;                 // gridDim.x of the backward launch
  float* dw[2];                // may be null (LayerNorm without affine parameters)
  float* db[2];
  const float* rowdot;         // may be null
  int rows;
  float* dt_out;               // may be null
};

__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const LnFinalizeJob job, int D) {
  const int j = blockIdx.y >> 1, which = blockIdx.y & 1;
  float* dst = which ? (j ? job.db[1] : job.db[0]) : (j ? job.dw[1] : job.dw[0]);
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (dst != nullptr && d < D) {
    const float* src = job.col_partials + ((size_t)j * job.nblocks * 2 + which) * (size_t)D + d;
    double s = 0.0;
    for (int blk = 0; blk < job.nblocks; ++blk) s += (double)src[(size_t)blk * 2 * D];
    dst[d] = (float)s;
  }
  if (job.dt_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0) block_reduce_to(job.rowdot, job.rows, job.dt_out);
}

// ------------------------------------------------------------------ host-side kernel selection
// Shared by the C ABI (jsd_capi.cu) and by the CPU emulation of these kernels (tests/emu), so that the tests
// exercise the same variant / block-size / columns-per-thread choice the product makes.
struct LnBwdPlan {
  int vec;       // 4: 16-byte pieces (D % 4 == 0, every pointer 16-byte aligned); 1: element-wise
  int threads;   // block size: one thread per column piece, rounded up to whole warps, at most LN_BWD_THREADS
  int kch;       // column pieces per thread (instantiated: 1, 2, 4); 0 = D too large
};
inline LnBwdPlan ln_bwd_plan(long long D, bool aligned16) {
  LnBwdPlan p;
  p.vec = (D % 4 == 0 && aligned16) ? 4 : 1;
  const long long pieces = (D + p.vec - 1) / p.vec;
  long long threads = (pieces + 31) / 32 * 32;
  if (threads > LN_BWD_THREADS) threads = LN_BWD_THREADS;
  p.threads = (int)threads;
  const long long kch = (pieces + threads - 1) / threads;
  p.kch = kch <= 1 ? 1 : (kch == 2 ? 2 : (kch <= LN_BWD_MAX_KCH ? 4 : 0));
  return p;
}
template <typename T, typename Fn>
inline void ln_bwd_select(const LnBwdPlan& p, Fn&& fn) {
  if (p.vec == 4) {
    if (p.kch == 1) fn(ln_normalize_bwd_kernel<T, 4, 1>);
    else if (p.kch == 2) fn(ln_normalize_bwd_kernel<T, 4, 2>);
    else fn(ln_normalize_bwd_kernel<T, 4, 4>);
  } else {
    if (p.kch == 1) fn(ln_normalize_bwd_kernel<T, 1, 1>);
    else if (p.kch == 2) fn(ln_normalize_bwd_kernel<T, 1, 2>);
    else fn(ln_normalize_bwd_kernel<T, 1, 4>);
  }
}
// forward: 0 = register-resident (argument: D / 128), 1 = 16-byte pieces, 2 = element-wise (argument: D)
inline int ln_fwd_variant(long long D, bool aligned16) {
  if (D % 4 != 0 || !aligned16) return 2;
  return (D % 128 == 0 && D / 128 <= LN_REG_CHUNKS) ? 0 : 1;
}
template <typename T, typename OUT, typename Fn>
inline void ln_fwd_select(int variant, long long D, Fn&& fn) {
  if (variant == 0) fn(ln_normalize_reg_kernel<T, OUT>, (int)(D / 128));
  else if (variant == 1) fn(ln_normalize_kernel<T, OUT, 4>, (int)D);
  else fn(ln_normalize_kernel<T, OUT, 1>, (int)D);
}

}  // namespace jsd
