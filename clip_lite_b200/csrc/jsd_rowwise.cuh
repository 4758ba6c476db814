// HBM-bound row-wise kernels of the JSD hot path:
//
//   jsd_index_kernel      the reference's estimator (one indexed negative per row,
//                         loss.py:94-105,204-254) fused forward + backward
//   normalize_cast_kernel F.normalize (loss.py:94-95) + cast to bf16 + 1/||x||
//   normalize_bwd_kernel  diagonal (positive-pair) term + Jacobian of F.normalize
//   finalize_*_kernel     deterministic fp64 reduction of the per-CTA loss partials
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifdef JSD_HOST_EMU
#include "ptx_emu.cuh"   // tests/emu: the kernels of this header run on the CPU under the test shim (tests only)
#else
#include "ptx.cuh"
#endif

namespace jsd {

constexpr float kNormEps = 1e-12f;   // F.normalize default eps

// ------------------------------------------------------------------ typed 4-wide access
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 w;
    w.x = pack_bf16x2(v.x, v.y);
    w.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = w;
  }
};
template <>
struct Vec4<__half> {
  static __device__ __forceinline__ float4 load(const __half* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__half* p, float4 v) {
    uint2 w;
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    w.x = *reinterpret_cast<const uint32_t*>(&a);
    w.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = w;
  }
};
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// Visit a row 4 elements at a time (VEC=4: 16 B-aligned vector access) or one at a time.
// loads in flight per lane: 16-bit rows move half the bytes per instruction, so they unroll twice as far
// (index kernel, bf16, B = 8192 x D = 2048: 52.5 -> 47.2 us; fp32 prefers 2: 75.6 vs 78.9 us -- r02n)
template <typename T, int VEC, typename Fn>
__device__ __forceinline__ void for_row(int D, int tid, int nthreads, Fn fn) {
  if constexpr (VEC == 4 && sizeof(T) == 2) {
#pragma unroll 4
    for (int d = tid * 4; d < D; d += nthreads * 4) fn(d);
  } else if constexpr (VEC == 4) {
#pragma unroll 2
    for (int d = tid * 4; d < D; d += nthreads * 4) fn(d);   // two iterations of loads in flight per lane
  } else {
    for (int d = tid; d < D; d += nthreads) fn(d);
  }
}

// Sum NV per-thread values over the block; every thread gets the totals.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch /* [NV*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();   // scratch may still be read from a previous call
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += scratch[i * 32 + w];
    v[i] = s;
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// ------------------------------------------------------------------ index mode
// One WARP per row j, eight consecutive rows per CTA (a row's negative and pre-image are usually its
// neighbours, so their data is found in L1).  The warp computes everything row j of dF and dG needs:
//   the positive pair (j, j), row j's negative (j, n = neg[j]) and the pairs (p, j) of every row p
//   whose negative is j (CSR inverse; NULL => p = j-1, the roll-by-one of loss.py:214-216).
// No block barriers: all reductions are warp shuffles.  Pass 1 accumulates the dot products, pass 2
// re-reads the (L1-resident) rows and writes both gradients, with <u,dU> and <v,dV> in closed form.
#ifndef JSD_INDEX_ROWS
#define JSD_INDEX_ROWS 8
#endif
#ifndef JSD_INDEX_MINBLOCKS
#define JSD_INDEX_MINBLOCKS (2048 / (32 * JSD_INDEX_ROWS) / 2)   // at least half of the SM's warp slots: <= 64 registers
#endif
constexpr int INDEX_ROWS_PER_CTA = JSD_INDEX_ROWS;

template <typename T, int VEC>
__global__ void __launch_bounds__(32 * INDEX_ROWS_PER_CTA, JSD_INDEX_MINBLOCKS)
jsd_index_kernel(const T* __restrict__ F, const T* __restrict__ G, int B, int D, const int* __restrict__ neg_index,
                 const int* __restrict__ inv_ptr, const int* __restrict__ inv_idx, const float* __restrict__ t_dev,
                 float* __restrict__ coefp, float* __restrict__ partials, T* __restrict__ dF, T* __restrict__ dG,
                 float grad_scale, const float* __restrict__ gamma_dev) {
  __shared__ float cta_part[INDEX_ROWS_PER_CTA][3];
  const int lane = threadIdx.x & 31;
  const int wrow = threadIdx.x >> 5;
  const int j = blockIdx.x * INDEX_ROWS_PER_CTA + wrow;
  if (lane < 3) cta_part[wrow][lane] = 0.f;
  if (j < B) {
  const int n = neg_index ? neg_index[j] : (j + 1 == B ? 0 : j + 1);
  const float tau = expf(*t_dev);
  const float invB = 1.f / (float)B;
  const T* fj = F + (size_t)j * D;
  const T* gj = G + (size_t)j * D;
  const T* gn = G + (size_t)n * D;
  // rows p that use text row j as their negative
  const int pbeg = inv_ptr ? inv_ptr[j] : 0;
  const int pend = inv_ptr ? inv_ptr[j + 1] : 1;
  const int p0 = inv_idx ? (pbeg < pend ? inv_idx[pbeg] : 0) : (j == 0 ? B - 1 : j - 1);
  const bool one_pre = (pend - pbeg == 1);          // permutation (normal / cluster mode): fully fused path
  const T* fp0 = F + (size_t)p0 * D;

  float s7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // ff, gg, fg, gngn, f.gn, fpfp, fp.g
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(fj + d), g = Vec4<T>::load(gj + d), h = Vec4<T>::load(gn + d);
      s7[0] += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
      s7[1] += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
      s7[2] += f.x * g.x + f.y * g.y + f.z * g.z + f.w * g.w;
      s7[3] += h.x * h.x + h.y * h.y + h.z * h.z + h.w * h.w;
      s7[4] += f.x * h.x + f.y * h.y + f.z * h.z + f.w * h.w;
      if (one_pre) {
        const float4 q = Vec4<T>::load(fp0 + d);
        s7[5] += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        s7[6] += q.x * g.x + q.y * g.y + q.z * g.z + q.w * g.w;
      }
    } else {
      const float f = to_f32(fj[d]), g = to_f32(gj[d]), h = to_f32(gn[d]);
      s7[0] += f * f;
      s7[1] += g * g;
      s7[2] += f * g;
      s7[3] += h * h;
      s7[4] += f * h;
      if (one_pre) {
        const float q = to_f32(fp0[d]);
        s7[5] += q * q;
        s7[6] += q * g;
      }
    }
  });
#pragma unroll
  for (int i = 0; i < 7; ++i) s7[i] = warp_sum(s7[i]);
  const float inv_f = 1.f / fmaxf(sqrtf(s7[0]), kNormEps);
  const float inv_g = 1.f / fmaxf(sqrtf(s7[1]), kNormEps);
  const float inv_gn = 1.f / fmaxf(sqrtf(s7[3]), kNormEps);
  const float s_pos = tau * s7[2] * inv_f * inv_g;
  const float s_neg = tau * s7[4] * inv_f * inv_gn;
  const float a = -sigmoid_f(-s_pos) * invB;   // dL/ds_pos
  const float b = sigmoid_f(s_neg) * invB;     // dL/ds_neg
  const float udot = a * s_pos + b * s_neg;    // <u_j, dU_j>
  float vdot = a * s_pos;                      // <v_j, dV_j>
  float c0 = 0.f;                              // tau * b_p / ||f_p|| of the single pre-image
  if (one_pre) {
    const float inv_fp = 1.f / fmaxf(sqrtf(s7[5]), kNormEps);
    const float sp = tau * s7[6] * inv_fp * inv_g;
    const float bp = sigmoid_f(sp) * invB;
    vdot += bp * sp;
    c0 = tau * bp * inv_fp;
  } else {
    // general index (a text row may be the negative of any number of rows): one extra sweep per pre-image
    for (int k = pbeg; k < pend; ++k) {
      const int pr = inv_idx[k];
      const T* fp = F + (size_t)pr * D;
      float s2[2] = {0.f, 0.f};   // fpfp, fp.gj
      for_row<T, VEC>(D, lane, 32, [&](int d) {
        if constexpr (VEC == 4) {
          const float4 f = Vec4<T>::load(fp + d), g = Vec4<T>::load(gj + d);
          s2[0] += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
          s2[1] += f.x * g.x + f.y * g.y + f.z * g.z + f.w * g.w;
        } else {
          const float f = to_f32(fp[d]), g = to_f32(gj[d]);
          s2[0] += f * f;
          s2[1] += f * g;
        }
      });
      s2[0] = warp_sum(s2[0]);
      s2[1] = warp_sum(s2[1]);
      const float inv_fp = 1.f / fmaxf(sqrtf(s2[0]), kNormEps);
      const float sp = tau * s2[1] * inv_fp * inv_g;
      const float bp = sigmoid_f(sp) * invB;
      vdot += bp * sp;
      if (lane == 0) coefp[pr] = tau * bp * inv_fp;   // each p has exactly one target row => no race
    }
    __syncwarp();
  }

  // The stored gradients are those of upstream gradient 1 times `grad_scale` (the caller divides it out again in
  // fp32): with fp16 features sigma / (B ||f||) would land in the subnormals before GradScaler's factor is applied.
  // dF == nullptr: forward only (eval / no_grad), the write-back pass is skipped.
  if (dF != nullptr) {
  T* dfj = dF + (size_t)j * D;
  T* dgj = dG + (size_t)j * D;
  const float ca = tau * a, cb = tau * b;
  const float gs = gamma_dev ? grad_scale * *gamma_dev : grad_scale;   // upstream gradient applied in fp32, here
  const float inv_fs = inv_f * gs, inv_gs = inv_g * gs;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(fj + d), g = Vec4<T>::load(gj + d), h = Vec4<T>::load(gn + d);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (one_pre) {
        const float4 q = Vec4<T>::load(fp0 + d);
        acc = make_float4(c0 * q.x, c0 * q.y, c0 * q.z, c0 * q.w);
      } else {
        for (int k = pbeg; k < pend; ++k) {
          const int pr = inv_idx[k];
          const float c = coefp[pr];
          const float4 q = Vec4<T>::load(F + (size_t)pr * D + d);
          acc.x = fmaf(c, q.x, acc.x);
          acc.y = fmaf(c, q.y, acc.y);
          acc.z = fmaf(c, q.z, acc.z);
          acc.w = fmaf(c, q.w, acc.w);
        }
      }
      float4 of, og;
#define JSD_IDX_ELT(c)                                                            \
  {                                                                               \
    const float u = f.c * inv_f, v = g.c * inv_g, vn = h.c * inv_gn;              \
    of.c = (ca * v + cb * vn - u * udot) * inv_fs;                                \
    og.c = (ca * u + acc.c - v * vdot) * inv_gs;                                  \
  }
      JSD_IDX_ELT(x) JSD_IDX_ELT(y) JSD_IDX_ELT(z) JSD_IDX_ELT(w)
#undef JSD_IDX_ELT
      Vec4<T>::store(dfj + d, of);
      Vec4<T>::store(dgj + d, og);
    } else {
      const float f = to_f32(fj[d]), g = to_f32(gj[d]), h = to_f32(gn[d]);
      float acc = 0.f;
      if (one_pre) {
        acc = c0 * to_f32(fp0[d]);
      } else {
        for (int k = pbeg; k < pend; ++k) {
          const int pr = inv_idx[k];
          acc = fmaf(coefp[pr], to_f32(F[(size_t)pr * D + d]), acc);
        }
      }
      const float u = f * inv_f, v = g * inv_g, vn = h * inv_gn;
      dfj[d] = from_f32<T>((ca * v + cb * vn - u * udot) * inv_fs);
      dgj[d] = from_f32<T>((ca * u + acc - v * vdot) * inv_gs);
    }
  });
  }  // dF != nullptr
  __syncwarp();
  if (lane == 0) {
    cta_part[wrow][0] = softplus_f(-s_pos);
    cta_part[wrow][1] = softplus_f(s_neg);
    cta_part[wrow][2] = udot;
  }
  }  // j < B
  __syncthreads();
  if (threadIdx.x < 3) {   // fixed-order sum over the CTA's rows: one partial triple per CTA
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < INDEX_ROWS_PER_CTA; ++w) acc += cta_part[w][threadIdx.x];
    partials[3 * (size_t)blockIdx.x + threadIdx.x] = acc;
  }
}

#ifndef JSD_HOST_EMU   // bulk-copy engine + mbarriers: not emulated
// ------------------------------------------------------------------ index mode, normal (roll-by-one) pairing, staged
// Same arithmetic as jsd_index_kernel, but every row of F and G is fetched from memory EXACTLY ONCE: persistent
// CTAs own contiguous row ranges, a producer warp streams the rows through two shared-memory rings with the
// bulk-copy engine (cp.async.bulk + mbarrier: whole rows in flight, no register or L1 involvement), eight consumer
// warps take rows round-robin and run both passes out of shared memory.  (ncu, r02: the L1-based kernel reads
// 191 MB from DRAM for 134 MB of operands and sits at 55 % DRAM utilisation with 43 % of the warp slots active.)
// Row i of a range needs F[i] (the pre-image row j - 1), F[i + 1] (its own), G[i] (its own) and G[i + 1] (its
// negative, row j + 1) in ring coordinates; every slot therefore has two consumers (the range's end rows arrive
// twice on their outer slots).
constexpr int IR_WARPS = 8;
constexpr int IR_THREADS = 32 * (IR_WARPS + 1);
constexpr int IR_MAX_SLOTS = 16;

template <typename T>
__global__ void __launch_bounds__(IR_THREADS, 1)
jsd_index_ring_kernel(const T* __restrict__ F, const T* __restrict__ G, int B, int D, int slots,
                      const float* __restrict__ t_dev, float* __restrict__ partials, T* __restrict__ dF,
                      T* __restrict__ dG, float grad_scale, const float* __restrict__ gamma_dev) {
  extern __shared__ __align__(16) uint8_t ring_smem[];
  __shared__ __align__(8) unsigned long long bars[4 * IR_MAX_SLOTS];
  __shared__ float cta_part[IR_WARPS][3];
  const uint32_t row_bytes = (uint32_t)D * (uint32_t)sizeof(T);
  const uint32_t ring_f = smem_u32(ring_smem), ring_g = ring_f + (uint32_t)slots * row_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto f_full = [&](int s) { return bar0 + 8u * s; };
  auto f_empty = [&](int s) { return bar0 + 8u * (IR_MAX_SLOTS + s); };
  auto g_full = [&](int s) { return bar0 + 8u * (2 * IR_MAX_SLOTS + s); };
  auto g_empty = [&](int s) { return bar0 + 8u * (3 * IR_MAX_SLOTS + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's contiguous rows [r0, r0 + n)
  const int r0 = (int)((long long)B * blockIdx.x / gridDim.x);
  const int n = (int)((long long)B * (blockIdx.x + 1) / gridDim.x) - r0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) {
      mbar_init(f_full(s), 1);
      mbar_init(g_full(s), 1);
      mbar_init(f_empty(s), 2);
      mbar_init(g_empty(s), 2);
    }
    fence_mbar_init();
  }
  if (warp < IR_WARPS && lane < 3) cta_part[warp][lane] = 0.f;
  __syncthreads();

  if (warp == IR_WARPS) {
    // ===================================================== producer: ring index k <-> F row r0 - 1 + k, G row r0 + k
    if (lane == 0) {
      for (int k = 0; k <= n; ++k) {
        const int s = k % slots;
        const uint32_t ph = (uint32_t)(((k / slots) & 1) ^ 1);
        int fr = r0 - 1 + k;
        if (fr < 0) fr += B;
        int gr = r0 + k;
        if (gr >= B) gr -= B;
        mbar_wait(f_empty(s), ph);
        mbar_arrive_expect_tx(f_full(s), row_bytes);
        bulk_load_1d(ring_f + (uint32_t)s * row_bytes, F + (size_t)fr * D, row_bytes, f_full(s));
        mbar_wait(g_empty(s), ph);
        mbar_arrive_expect_tx(g_full(s), row_bytes);
        bulk_load_1d(ring_g + (uint32_t)s * row_bytes, G + (size_t)gr * D, row_bytes, g_full(s));
      }
    }
  } else {
    // ===================================================== consumers: rows i = warp, warp + 8, ...
    const float tau = expf(*t_dev);
    const float invB = 1.f / (float)B;
    const float gs = gamma_dev ? grad_scale * *gamma_dev : grad_scale;
    float part0 = 0.f, part1 = 0.f, part2 = 0.f;
    for (int i = warp; i < n; i += IR_WARPS) {
      const int sp = i % slots, so = (i + 1) % slots;       // slots of ring indices i and i + 1
      const uint32_t php = (uint32_t)((i / slots) & 1), pho = (uint32_t)(((i + 1) / slots) & 1);
      mbar_wait(f_full(sp), php);
      mbar_wait(g_full(sp), php);
      mbar_wait(f_full(so), pho);
      mbar_wait(g_full(so), pho);
      const T* fp0 = reinterpret_cast<const T*>(ring_smem + (size_t)sp * row_bytes);                          // f_{j-1}
      const T* fj = reinterpret_cast<const T*>(ring_smem + (size_t)so * row_bytes);                           // f_j
      const T* gj = reinterpret_cast<const T*>(ring_smem + (size_t)(slots + sp) * row_bytes);                 // g_j
      const T* gn = reinterpret_cast<const T*>(ring_smem + (size_t)(slots + so) * row_bytes);                 // g_{j+1}
      float s7[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // ff, gg, fg, gngn, f.gn, fpfp, fp.g
      for (int d = lane * 4; d < D; d += 128) {
        const float4 f = Vec4<T>::load(fj + d), g = Vec4<T>::load(gj + d), h = Vec4<T>::load(gn + d),
                     q = Vec4<T>::load(fp0 + d);
        s7[0] += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
        s7[1] += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
        s7[2] += f.x * g.x + f.y * g.y + f.z * g.z + f.w * g.w;
        s7[3] += h.x * h.x + h.y * h.y + h.z * h.z + h.w * h.w;
        s7[4] += f.x * h.x + f.y * h.y + f.z * h.z + f.w * h.w;
        s7[5] += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        s7[6] += q.x * g.x + q.y * g.y + q.z * g.z + q.w * g.w;
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) s7[k] = warp_sum(s7[k]);
      const float inv_f = 1.f / fmaxf(sqrtf(s7[0]), kNormEps);
      const float inv_g = 1.f / fmaxf(sqrtf(s7[1]), kNormEps);
      const float inv_gn = 1.f / fmaxf(sqrtf(s7[3]), kNormEps);
      const float inv_fp = 1.f / fmaxf(sqrtf(s7[5]), kNormEps);
      const float s_pos = tau * s7[2] * inv_f * inv_g;
      const float s_neg = tau * s7[4] * inv_f * inv_gn;
      const float spp = tau * s7[6] * inv_fp * inv_g;       // score of (row j - 1, text row j): j is its negative
      const float a = -sigmoid_f(-s_pos) * invB;            // dL/ds_pos
      const float b = sigmoid_f(s_neg) * invB;              // dL/ds_neg
      const float bp = sigmoid_f(spp) * invB;
      const float udot = a * s_pos + b * s_neg;             // <u_j, dU_j>
      const float vdot = a * s_pos + bp * spp;              // <v_j, dV_j>
      const float c0 = tau * bp * inv_fp;
      if (dF != nullptr) {
        const int j = r0 + i;
        T* dfj = dF + (size_t)j * D;
        T* dgj = dG + (size_t)j * D;
        const float ca = tau * a, cb = tau * b;
        const float inv_fs = inv_f * gs, inv_gs = inv_g * gs;
        for (int d = lane * 4; d < D; d += 128) {
          const float4 f = Vec4<T>::load(fj + d), g = Vec4<T>::load(gj + d), h = Vec4<T>::load(gn + d),
                       q = Vec4<T>::load(fp0 + d);
          float4 of, og;
#define JSD_RING_ELT(c)                                                           \
  {                                                                               \
    const float u = f.c * inv_f, v = g.c * inv_g, vn = h.c * inv_gn;              \
    of.c = (ca * v + cb * vn - u * udot) * inv_fs;                                \
    og.c = (ca * u + c0 * q.c - v * vdot) * inv_gs;                               \
  }
          JSD_RING_ELT(x) JSD_RING_ELT(y) JSD_RING_ELT(z) JSD_RING_ELT(w)
#undef JSD_RING_ELT
          Vec4<T>::store(dfj + d, of);
          Vec4<T>::store(dgj + d, og);
        }
      }
      part0 += softplus_f(-s_pos);
      part1 += softplus_f(s_neg);
      part2 += udot;
      __syncwarp();                                          // every lane is done reading the four slots
      if (lane == 0) {
        const int twice_lo = (i == 0) ? 2 : 1, twice_hi = (i == n - 1) ? 2 : 1;   // the range's outer slots have one consumer
        for (int r = 0; r < twice_lo; ++r) {
          mbar_arrive(f_empty(sp));
          mbar_arrive(g_empty(sp));
        }
        for (int r = 0; r < twice_hi; ++r) {
          mbar_arrive(f_empty(so));
          mbar_arrive(g_empty(so));
        }
      }
    }
    if (lane == 0) {
      cta_part[warp][0] = part0;
      cta_part[warp][1] = part1;
      cta_part[warp][2] = part2;
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {   // fixed-order sum over the CTA's warps: one partial triple per CTA
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < IR_WARPS; ++w) acc += cta_part[w][threadIdx.x];
    partials[3 * (size_t)blockIdx.x + threadIdx.x] = acc;
  }
}

#endif  // !JSD_HOST_EMU

// out4 = {pos, neg, pos + neg, dL/dt} (+ an optional separate copy of the loss); deterministic (fixed order, fp64).
constexpr int FINALIZE_THREADS = 1024;

__global__ void __launch_bounds__(FINALIZE_THREADS)
finalize_kernel(const float* __restrict__ partials, int n, int width, double inv0, double inv1, double inv2,
                double inv3, float* __restrict__ out4, float* __restrict__ loss_out) {
  __shared__ double sh[4][FINALIZE_THREADS];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += FINALIZE_THREADS)
    for (int k = 0; k < width; ++k) acc[k] += (double)partials[(size_t)i * width + k];
  for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = FINALIZE_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double pos = sh[0][0] * inv0, neg = sh[1][0] * inv1;
    const double dt = sh[2][0] * inv2 + sh[3][0] * inv3;
    out4[0] = (float)pos;
    out4[1] = (float)neg;
    out4[2] = (float)(pos + neg);
    out4[3] = (float)dt;
    if (loss_out) *loss_out = (float)(pos + neg);
  }
}

// ------------------------------------------------------------------ normalise + cast (dense pre-pass)
// One warp per row: Xn = bf16(x / max(||x||, eps)), inv_norm = 1 / max(||x||, eps).
// One launch normalises up to two tensors of the same shape (blockIdx.y selects): F and G of a step.
struct NormalizeJob {
  const void* X[2];
  __nv_bfloat16* Xn[2];
  float* inv_norm[2];
  int* bump;          // optional: *bump += 1 by one thread (the peer exchange's per-buffer step counter: the forward
                      // launch behind this one publishes / waits for exactly this value)
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_cast_kernel(const NormalizeJob job, int rows, int D) {
  const bool second = blockIdx.y != 0;      // static selects: a dynamically indexed parameter array goes through local memory
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  __nv_bfloat16* __restrict__ Xn = second ? job.Xn[1] : job.Xn[0];
  float* __restrict__ inv_norm = second ? job.inv_norm[1] : job.inv_norm[0];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (job.bump != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *job.bump += 1;
  if (row >= rows) return;
  const T* x = X + (size_t)row * D;
  float ss = 0.f;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d);
      ss += f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w;
    } else {
      const float f = to_f32(x[d]);
      ss += f * f;
    }
  });
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
  __nv_bfloat16* o = Xn + (size_t)row * D;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d);
      Vec4<__nv_bfloat16>::store(o + d, make_float4(f.x * inv, f.y * inv, f.z * inv, f.w * inv));
    } else {
      o[d] = __float2bfloat16_rn(to_f32(x[d]) * inv);
    }
  });
  if (lane == 0) inv_norm[row] = inv;
}

// ------------------------------------------------------------------ normalise + cast + all-gather by peer stores
// Multi-GPU forward pre-pass in ONE launch: blockIdx.y = 0 normalises the image rows into the local U;
// blockIdx.y = 1 normalises the text rows and writes each bf16 unit row into EVERY rank's gathered V buffer
// (16-byte stores to peer memory over NVLink: the all-gather is fused into the producer).  The last block bumps
// this rank's counter and publishes it to every rank's "rows of rank r are in" flag -- one thread per destination,
// so that the remote release-stores (a ~2 us round trip each) go out together (trace r02a: issued one after the
// other they cost 14 us at 8 GPUs).  Selected by the host when the forward is too short to hide the in-forward
// push (8 GPUs at B = 8192); see jsd_peer_normalize_push.
struct PeerPushJob {
  const void* X[2];              // F rows, G rows [rows, D]
  __nv_bfloat16* U;              // local image unit rows
  float* inv_norm[2];
  __nv_bfloat16* v_dst[8];       // slot k: gathered V buffer of rank (rank - k) % world, offset to this rank's rows
  int* flag_dst[8];              // slot k: that rank's flag word for this rank (slot 0 = this rank: unused)
  int* counter;                  // local: pushes so far into this buffer
  int* ticket;                   // local, zero between launches
  int world;
};

// 8 consecutive elements of a row as fp32
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&x)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  const uint32_t r[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(r[i] << 16);
    x[2 * i + 1] = __uint_as_float(r[i] & 0xFFFF0000u);
  }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&x)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  const uint32_t r[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[i]));
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
}

#ifndef JSD_HOST_EMU   // bulk-copy engine + system-scope flags: not emulated
// VEC = 8: D % 8 == 0 and 16-byte aligned rows (the product's shapes): every warp builds its bf16 unit row in shared
// memory and ONE lane hands it to the bulk-copy engine once per destination (cp.async.bulk shared -> global, a
// whole 2 KB row per instruction) -- per-lane 16-byte stores to 8 peers reach ~420 GB/s (trace r02j: 32-37 us for
// 14.7 MB).  Dynamic shared memory: 8 rows x D bf16.  VEC = 1: anything else, plain stores.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_push_kernel(const __grid_constant__ PeerPushJob job, int rows, int D) {
  extern __shared__ __align__(16) uint8_t push_smem[];
  const int jy = blockIdx.y;
  if (blockIdx.x == 0 && jy == 0 && threadIdx.x == 0) trace_event(TK_PUSH, TE_START);
  const T* __restrict__ X = static_cast<const T*>(job.X[jy]);
  const int wrow = threadIdx.x >> 5;
  const int row = blockIdx.x * (blockDim.x >> 5) + wrow;
  const int lane = threadIdx.x & 31;
  if (row < rows) {
    const T* x = X + (size_t)row * D;
    float ss = 0.f;
    if constexpr (VEC == 8) {
#pragma unroll 2
      for (int d = lane * 8; d < D; d += 256) {
        float v[8];
        load8<T>(x + d, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float f = to_f32(x[d]);
        ss = fmaf(f, f, ss);
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
    if (lane == 0) job.inv_norm[jy][row] = inv;
    const int ndst = jy == 0 ? 1 : job.world;
    if constexpr (VEC == 8) {
      uint4* srow = reinterpret_cast<uint4*>(push_smem + (size_t)wrow * D * 2);
      for (int d = lane * 8; d < D; d += 256) {      // the row is re-read from L1
        float v[8];
        load8<T>(x + d, v);
        uint4 w;
        w.x = pack_bf16x2(v[0] * inv, v[1] * inv);
        w.y = pack_bf16x2(v[2] * inv, v[3] * inv);
        w.z = pack_bf16x2(v[4] * inv, v[5] * inv);
        w.w = pack_bf16x2(v[6] * inv, v[7] * inv);
        srow[d >> 3] = w;
      }
      fence_proxy_async();                           // generic-proxy smem writes -> visible to the bulk copies
      __syncwarp();
      if (lane == 0) {
        const uint32_t src = smem_u32(srow);
        for (int k = 0; k < ndst; ++k)
          bulk_store_1d((jy == 0 ? job.U : job.v_dst[k]) + (size_t)row * D, src, (uint32_t)D * 2);
        tma_store_commit();
        tma_store_wait_all();                        // written (not merely read) before this block takes its ticket
        fence_proxy_async_all();
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const __nv_bfloat16 r = __float2bfloat16_rn(to_f32(x[d]) * inv);
        for (int k = 0; k < ndst; ++k) ((jy == 0 ? job.U : job.v_dst[k]) + (size_t)row * D)[d] = r;
      }
    }
  }
  __shared__ int s_flag[2];
  __syncthreads();                                   // the block's stores, before thread 0's (cumulative) fence
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_flag[0] = (atomicAdd(job.ticket, 1) == (int)(gridDim.x * gridDim.y) - 1);
  }
  __syncthreads();
  if (s_flag[0]) {
    if (threadIdx.x == 0) {
      __threadfence_system();
      const int e = *job.counter + 1;
      *job.counter = e;
      s_flag[1] = e;
      *job.ticket = 0;
    }
    __syncthreads();
    if (threadIdx.x >= 1 && (int)threadIdx.x < job.world) {      // slot 0 is this rank itself: stream-ordered
      __threadfence_system();
      st_release_sys(job.flag_dst[threadIdx.x], s_flag[1]);
    }
    if (threadIdx.x == 0) trace_event(TK_PUSH, TE_END);
  }
}

#endif  // !JSD_HOST_EMU

// Fixed-order fp64 sum of n floats by one block (deterministic); every thread of the block must call it.
__device__ __forceinline__ void block_reduce_to(const float* src, int n, float* out) {
  __shared__ double s_red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)__ldcg(src + i);
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)s_red[0];
}

// Last-block reduction of the row dots: dt_out = sum_i rowdot[i] (fp64, fixed order => deterministic).
// Called by every thread of every block at the end of a normalise-backward kernel when dt_out != nullptr.
__device__ __forceinline__ void reduce_rowdot_last_block(const float* rowdot, int rows, int* ticket, float* dt_out) {
  __shared__ int s_last;
  __shared__ double s_part[256];
  __syncthreads();                                   // every warp has stored AND fenced its rowdot entry
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += (double)__ldcg(rowdot + i);
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *dt_out = (float)s_part[0];
    *ticket = 0;                                     // re-armed for the next launch on this stream
  }
}

// ------------------------------------------------------------------ normalise backward (dense post-pass)
// dU_row = acc_row + coef * gdiag[row] * partner[row + partner_offset]      (positive-pair term, fp32)
// dX_row = (dU_row - u_row <u_row, dU_row>) * inv_norm[row],  u_row = x_row * inv_norm[row]
// coef = gamma * tau / M_rows.  One warp per row.  With dt_out != nullptr the last block also reduces the row
// dots <u_i, dU_i> to gamma * dL/dt (no separate reduction launch).
// One launch serves up to two row sets of the same shape (blockIdx.y selects): the image and the text side of
// a single-GPU step.  Only job 0 (the image side) reports row dots / dL/dt.
struct NormBwdJob {
  const void* X[2];
  const float* inv_norm[2];
  const float* acc[2];
  const __nv_bfloat16* partner[2];
  long long partner_offset[2];
  void* dX[2];
  float* rowdot;     // job 0 only; may be null
  int* ticket;       // zero between launches (needed iff dt_out != nullptr)
  float* dt_out;     // may be null
  // peer exchange (text side of a multi-GPU step): the accumulator is the sum of `acc_slots` partials, one in
  // each rank's memory (slot[q], read over NVLink, already offset to this rank's rows), added here in rank order
  // (deterministic) -- the reduce-scatter is fused into its consumer.  Every block first waits until each rank's
  // flag has reached *wait_counter ("my partial is complete").
  // deferred dL/dt: block (0, 0) of THIS launch sums reduce_src[0 .. reduce_n) -- row dots written by an EARLIER
  // launch (stream order makes them visible) -- into *reduce_out: no ticket, no fence in either kernel
  const float* reduce_src;
  int reduce_n;
  float* reduce_out;
  float acc_scale[2];          // > 0: acc[] holds UNSCALED sums (fused single-pass kernel); the kernel multiplies them
                               // by gamma * tau * acc_scale in fp32.  0: acc[] is already scaled (staged path)
  int acc_slots;               // 0: a single local accumulator (acc[])
  int slot_bf16;               // the slots hold bf16 (partials pushed by the peers' dV contractions), not fp32
  const float* slot[8];        // job 0 (or the only job)
  const float* slot1[8];       // job 1 of a two-job launch (same acc_slots)
  const int* wait_flags;       // null: no wait
  const int* wait_counter;
  int wait_count;
};

__device__ __forceinline__ void normbwd_wait_peers(const NormBwdJob& job) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) trace_event(TK_JACOBIAN, TE_START);
  if (job.wait_flags == nullptr) return;
  if (threadIdx.x == 0) {
    wait_flags_sys(job.wait_flags, job.wait_count, *job.wait_counter, WAIT_GRAD_PARTIALS);
    if (blockIdx.x == 0 && blockIdx.y == 0) trace_event(TK_JACOBIAN, TE_PEERS_IN);
  }
  __syncthreads();
}

// accumulator element(s) at float offset `off` of the row block: the single local accumulator, or the sum over
// the ranks' partials in rank order
__device__ __forceinline__ float4 load_slot4(const float* slot, size_t off, bool bf16) {
  if (!bf16) return __ldcs(reinterpret_cast<const float4*>(slot + off));
  const uint2 w = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(slot) + off));
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xFFFF0000u));
}
__device__ __forceinline__ float4 load_acc4(const NormBwdJob& job, const float* acc, size_t off, bool second = false) {
  if (job.acc_slots <= 0) return __ldcs(reinterpret_cast<const float4*>(acc + off));
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int sl = 0; sl < 8; ++sl)
    if (sl < job.acc_slots) {
      const float4 h = load_slot4(second ? job.slot1[sl] : job.slot[sl], off, job.slot_bf16 != 0);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
  return g;
}
__device__ __forceinline__ float load_acc1(const NormBwdJob& job, const float* acc, size_t off, bool second = false) {
  if (job.acc_slots <= 0) return acc[off];
  float g = 0.f;
#pragma unroll
  for (int sl = 0; sl < 8; ++sl)
    if (sl < job.acc_slots) {
      const float* sp = second ? job.slot1[sl] : job.slot[sl];
      g += job.slot_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(sp)[off]) : sp[off];
    }
  return g;
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const NormBwdJob job, int rows, int D, const float* __restrict__ gdiag,
                     const float* __restrict__ t_dev, const float* __restrict__ gamma_dev, float inv_rows) {
  const bool second = blockIdx.y != 0;
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  const float* __restrict__ inv_norm = second ? job.inv_norm[1] : job.inv_norm[0];
  const float* __restrict__ acc = second ? job.acc[1] : job.acc[0];
  const __nv_bfloat16* __restrict__ partner = second ? job.partner[1] : job.partner[0];
  const long long partner_offset = second ? job.partner_offset[1] : job.partner_offset[0];
  T* __restrict__ dX = static_cast<T*>(second ? job.dX[1] : job.dX[0]);
  float* __restrict__ rowdot = second ? nullptr : job.rowdot;
  normbwd_wait_peers(job);
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row < rows) {
  const float gamma = gamma_dev ? *gamma_dev : 1.f;
  const float c = gdiag ? gamma * expf(*t_dev) * inv_rows * gdiag[row] : 0.f;
  const float asel = second ? job.acc_scale[1] : job.acc_scale[0];
  const float as = asel > 0.f ? gamma * expf(*t_dev) * asel : 1.f;      // x * 1 is exact: the staged path is unchanged
  const float inv = inv_norm[row];
  const T* x = X + (size_t)row * D;
  const size_t aoff = (size_t)row * D;
  const __nv_bfloat16* pr = partner + (size_t)(row + partner_offset) * D;
  float dot = 0.f;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d), g = load_acc4(job, acc, aoff + d, second), q = Vec4<__nv_bfloat16>::load(pr + d);
      dot += f.x * fmaf(c, q.x, g.x * as) + f.y * fmaf(c, q.y, g.y * as) + f.z * fmaf(c, q.z, g.z * as) +
             f.w * fmaf(c, q.w, g.w * as);
    } else {
      dot += to_f32(x[d]) * fmaf(c, __bfloat162float(pr[d]), load_acc1(job, acc, aoff + d, second) * as);
    }
  });
  dot = warp_sum(dot) * inv;   // <u, dU>
  if (rowdot != nullptr && lane == 0) {   // sum over rows = gamma * dL/dt
    rowdot[row] = dot;
    if (job.dt_out != nullptr) __threadfence();   // ticketed reduction in this launch: publish now, not after the row's big stores
  }
  T* o = dX + (size_t)row * D;
  for_row<T, VEC>(D, lane, 32, [&](int d) {
    if constexpr (VEC == 4) {
      const float4 f = Vec4<T>::load(x + d), g = load_acc4(job, acc, aoff + d, second), q = Vec4<__nv_bfloat16>::load(pr + d);
      float4 r;
      r.x = (fmaf(c, q.x, g.x * as) - f.x * inv * dot) * inv;
      r.y = (fmaf(c, q.y, g.y * as) - f.y * inv * dot) * inv;
      r.z = (fmaf(c, q.z, g.z * as) - f.z * inv * dot) * inv;
      r.w = (fmaf(c, q.w, g.w * as) - f.w * inv * dot) * inv;
      Vec4<T>::store(o + d, r);
    } else {
      o[d] = from_f32<T>((fmaf(c, __bfloat162float(pr[d]), load_acc1(job, acc, aoff + d, second) * as) - to_f32(x[d]) * inv * dot) * inv);
    }
  });
  }  // row < rows
  if (job.dt_out != nullptr) reduce_rowdot_last_block(job.rowdot, rows, job.ticket, job.dt_out);
  if (job.reduce_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0) block_reduce_to(job.reduce_src, job.reduce_n, job.reduce_out);
  if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && threadIdx.x == 0) trace_event(TK_JACOBIAN, TE_END);
}

// ------------------------------------------------------------------ register-resident variants (D = nch * 128 <= 1024)
// Same arithmetic as the two kernels above, but every row is read from memory exactly once: a lane keeps
// its (up to) 8 x 4 elements of each operand in registers between the reduction and the write-back, and all
// loads of a row are in flight together.
constexpr int ROW_REG_CHUNKS = 8;

template <typename T>
__global__ void __launch_bounds__(256)
normalize_cast_reg_kernel(const NormalizeJob job, int rows, int nch) {
  const bool second = blockIdx.y != 0;      // static selects: a dynamically indexed parameter array goes through local memory
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  __nv_bfloat16* __restrict__ Xn = second ? job.Xn[1] : job.Xn[0];
  float* __restrict__ inv_norm = second ? job.inv_norm[1] : job.inv_norm[0];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (job.bump != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *job.bump += 1;
  if (row >= rows) return;
  const int D = nch * 128;
  const T* x = X + (size_t)row * D + lane * 4;
  float4 xv[ROW_REG_CHUNKS];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch) xv[i] = Vec4<T>::load(x + i * 128);
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch) ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
  __nv_bfloat16* o = Xn + (size_t)row * D + lane * 4;
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch)
      Vec4<__nv_bfloat16>::store(o + i * 128, make_float4(xv[i].x * inv, xv[i].y * inv, xv[i].z * inv, xv[i].w * inv));
  if (lane == 0) inv_norm[row] = inv;
}

template <typename T>
__global__ void __launch_bounds__(256)
normalize_bwd_reg_kernel(const NormBwdJob job, int rows, int nch, const float* __restrict__ gdiag,
                         const float* __restrict__ t_dev, const float* __restrict__ gamma_dev, float inv_rows) {
  const bool second = blockIdx.y != 0;
  const T* __restrict__ X = static_cast<const T*>(second ? job.X[1] : job.X[0]);
  const float* __restrict__ inv_norm = second ? job.inv_norm[1] : job.inv_norm[0];
  const float* __restrict__ acc = second ? job.acc[1] : job.acc[0];
  const __nv_bfloat16* __restrict__ partner = second ? job.partner[1] : job.partner[0];
  const long long partner_offset = second ? job.partner_offset[1] : job.partner_offset[0];
  T* __restrict__ dX = static_cast<T*>(second ? job.dX[1] : job.dX[0]);
  float* __restrict__ rowdot = second ? nullptr : job.rowdot;
  normbwd_wait_peers(job);
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row < rows) {
  const int D = nch * 128;
  const float gamma = gamma_dev ? *gamma_dev : 1.f;
  const float c = gdiag ? gamma * expf(*t_dev) * inv_rows * gdiag[row] : 0.f;
  const float asel = second ? job.acc_scale[1] : job.acc_scale[0];
  const float as = asel > 0.f ? gamma * expf(*t_dev) * asel : 1.f;      // x * 1 is exact: the staged path is unchanged
  const float inv = inv_norm[row];
  const T* x = X + (size_t)row * D + lane * 4;
  const size_t aoff = (size_t)row * D + lane * 4;
  const __nv_bfloat16* pr = partner + (size_t)(row + partner_offset) * D + lane * 4;
  float4 xv[ROW_REG_CHUNKS], dv[ROW_REG_CHUNKS];
  // dv starts as the accumulator row (the sum of the ranks' partials in rank order when there are several); every
  // chunk load of a pass is issued before the first use so that the whole row is in flight together
  if (job.acc_slots <= 0) {
#pragma unroll
    for (int i = 0; i < ROW_REG_CHUNKS; ++i)
      if (i < nch) dv[i] = __ldcs(reinterpret_cast<const float4*>(acc + aoff + i * 128));
  } else if (job.slot_bf16) {
    // bf16 partials pushed by the peers: FOUR slots' loads are in flight together (raw 8-byte pieces, 16 registers
    // per slot); one slot at a time, 8 GPUs cost eight dependent round trips per row -- 18 us for 1024 rows (trace
    // r02h).  Added in rank order either way (deterministic).
#pragma unroll
    for (int i = 0; i < ROW_REG_CHUNKS; ++i) dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s0 = 0; s0 < 8; s0 += 4) {
      if (s0 < job.acc_slots) {
        uint2 raw[4][ROW_REG_CHUNKS];
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < job.acc_slots) {
            const __nv_bfloat16* a =
                reinterpret_cast<const __nv_bfloat16*>(second ? job.slot1[s0 + s] : job.slot[s0 + s]) + aoff;
#pragma unroll
            for (int i = 0; i < ROW_REG_CHUNKS; ++i)
              if (i < nch) raw[s][i] = __ldcs(reinterpret_cast<const uint2*>(a + i * 128));
          }
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < job.acc_slots) {
#pragma unroll
            for (int i = 0; i < ROW_REG_CHUNKS; ++i)
              if (i < nch) {
                dv[i].x += __uint_as_float(raw[s][i].x << 16);
                dv[i].y += __uint_as_float(raw[s][i].x & 0xFFFF0000u);
                dv[i].z += __uint_as_float(raw[s][i].y << 16);
                dv[i].w += __uint_as_float(raw[s][i].y & 0xFFFF0000u);
              }
          }
      }
    }
  } else {
    // fp32 slots (split-K slices of the image-side contraction, or partials pulled from the peers): TWO slots'
    // loads in flight together, added in slot order (deterministic)
#pragma unroll
    for (int i = 0; i < ROW_REG_CHUNKS; ++i) dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s0 = 0; s0 < 8; s0 += 2) {
      if (s0 < job.acc_slots) {
        float4 h[2][ROW_REG_CHUNKS];
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s0 + s < job.acc_slots) {
#pragma unroll
            for (int i = 0; i < ROW_REG_CHUNKS; ++i)
              if (i < nch)
                h[s][i] = __ldcs(reinterpret_cast<const float4*>((second ? job.slot1[s0 + s] : job.slot[s0 + s]) + aoff +
                                                                 i * 128));
          }
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (s0 + s < job.acc_slots) {
#pragma unroll
            for (int i = 0; i < ROW_REG_CHUNKS; ++i)
              if (i < nch) {
                dv[i].x += h[s][i].x; dv[i].y += h[s][i].y; dv[i].z += h[s][i].z; dv[i].w += h[s][i].w;
              }
          }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch) {
      xv[i] = Vec4<T>::load(x + i * 128);
      const float4 q = Vec4<__nv_bfloat16>::load(pr + i * 128);
      dv[i] = make_float4(fmaf(c, q.x, dv[i].x * as), fmaf(c, q.y, dv[i].y * as), fmaf(c, q.z, dv[i].z * as),
                          fmaf(c, q.w, dv[i].w * as));   // dU
    }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch) dot += xv[i].x * dv[i].x + xv[i].y * dv[i].y + xv[i].z * dv[i].z + xv[i].w * dv[i].w;
  dot = warp_sum(dot) * inv;   // <u, dU>
  if (rowdot != nullptr && lane == 0) {
    rowdot[row] = dot;
    if (job.dt_out != nullptr) __threadfence();   // ticketed reduction in this launch: publish now, not after the row's big stores
  }
  T* o = dX + (size_t)row * D + lane * 4;
  const float k = inv * dot;
#pragma unroll
  for (int i = 0; i < ROW_REG_CHUNKS; ++i)
    if (i < nch) {
      float4 r;
      r.x = (dv[i].x - xv[i].x * k) * inv;
      r.y = (dv[i].y - xv[i].y * k) * inv;
      r.z = (dv[i].z - xv[i].z * k) * inv;
      r.w = (dv[i].w - xv[i].w * k) * inv;
      Vec4<T>::store(o + i * 128, r);
    }
  }  // row < rows
  if (job.dt_out != nullptr) reduce_rowdot_last_block(job.rowdot, rows, job.ticket, job.dt_out);
  if (job.reduce_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0) block_reduce_to(job.reduce_src, job.reduce_n, job.reduce_out);
  if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && threadIdx.x == 0) trace_event(TK_JACOBIAN, TE_END);
}

}  // namespace jsd
