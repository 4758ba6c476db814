// TEST INFRASTRUCTURE ONLY: host stand-ins for the few ptx.cuh helpers the row-wise kernels use, so that
// jsd_rowwise.cuh / jsd_heads.cuh compile with g++ -DJSD_HOST_EMU (see cuda_emu.h).  Same results as the PTX they
// replace: round-to-nearest-even bf16 packing, butterfly warp sum in the same order.
#pragma once
#include "cuda_emu.h"

namespace jsd {

enum TraceKernel { TK_NORMALIZE = 1, TK_FWD = 2, TK_GRAD = 3, TK_JACOBIAN = 4, TK_INDEX = 5, TK_SCORE = 6, TK_PUSH = 7 };
enum TraceEvent { TE_START = 0, TE_PEERS_IN = 1, TE_END = 2 };
enum WaitKind { WAIT_GATHERED_ROWS = 1, WAIT_GRAD_PARTIALS = 2 };
inline void trace_event(int, int) {}
inline void wait_flags_sys(const int* flags, int count, int target, int) {
  for (int i = 0; i < count; ++i)
    if (__atomic_load_n(flags + i, __ATOMIC_ACQUIRE) - target < 0) {
      std::fprintf(stderr, "emu: wait_flags_sys would spin (flag %d)\n", i);
      std::abort();
    }
}
inline uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat16 l = __float2bfloat16_rn(lo), h = __float2bfloat16_rn(hi);
  uint16_t lw, hw;
  std::memcpy(&lw, &l, 2);
  std::memcpy(&hw, &h, 2);
  return (uint32_t)lw | ((uint32_t)hw << 16);
}
inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace jsd
