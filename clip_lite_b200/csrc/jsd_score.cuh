// Operand preparation for the retrieval / zero-shot scoring path (reference retrieval.py:143-187,
// zero_shot.py:155): fp32-faithful scores from bf16 tensor cores by operand splitting.
//
//   x = hi + lo,  hi = bf16(x),  lo = bf16(x - hi)       (|x - hi - lo| <= 2^-17 |x|)
//   <a, b> ~= <a_hi, b_hi> + <a_hi, b_lo> + <a_lo, b_hi>  (the dropped lo.lo term is ~2^-18 relative)
//
// written as ONE contraction of length 3D:  A' = [a_hi | a_hi | a_lo],  B' = [b_hi | b_lo | b_hi].
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "jsd_rowwise.cuh"

namespace jsd {

// out [rows, 3 D] bf16.  side = 0: (hi, hi, lo)   side = 1: (hi, lo, hi).  normalize != 0: rows are first scaled to
// unit L2 norm (F.normalize, eps 1e-12) in fp32.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const T* __restrict__ X, int rows, int D, int side, int normalize, __nv_bfloat16* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* x = X + (size_t)row * D;
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float f = to_f32(x[d]);
      ss += f * f;
    }
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
  }
  __nv_bfloat16* o = out + (size_t)row * 3 * D;
  for (int d = lane; d < D; d += 32) {
    const float f = to_f32(x[d]) * inv;
    const __nv_bfloat16 hi = __float2bfloat16_rn(f);
    const __nv_bfloat16 lo = __float2bfloat16_rn(f - __bfloat162float(hi));
    o[d] = hi;
    o[D + d] = side == 0 ? hi : lo;
    o[2 * D + d] = side == 0 ? lo : hi;
  }
}

}  // namespace jsd
