// TEST INFRASTRUCTURE ONLY -- never part of the product, never loaded by clip_lite_b200.
//
// A small "CUDA block on the CPU" shim: the row-wise kernel headers under clip_lite_b200/csrc are compiled by g++
// with -DJSD_HOST_EMU and their __global__ functions are executed here with ONE OS THREAD PER CUDA THREAD, block after
// block.  __syncthreads / __syncwarp / __shfl_*_sync are real rendezvous between those threads, so a missing barrier,
// a divergent barrier, an out-of-range index or a wrong reduction shows up as a wrong result, a deadlock (caught by
// the test's timeout) or an AddressSanitizer report -- without a GPU.  What it cannot show: anything that depends
// on the hardware (alignment faults of vector accesses are checked explicitly below, occupancy / registers / speed
// are not).  The GPU tier (`-m gpu`) remains the parity gate; this tier keeps kernels that were written without a GPU
// at hand from being wrong in their logic.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#undef __global__
#undef __shared__
#undef __launch_bounds__
#undef __grid_constant__
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__

namespace emu {

struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
};

struct BlockState {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint32_t> xchg;   // [nthreads] shuffle exchange words
};

inline thread_local Dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockState* t_block = nullptr;

inline void check_aligned(const void* p, size_t a, const char* what) {
  if (reinterpret_cast<uintptr_t>(p) % a != 0) {
    std::fprintf(stderr, "emu: misaligned %zu-byte access (%s) at %p\n", a, what, p);
    std::abort();
  }
}

// run `body()` once per thread of every block, blocks in sequence (static `__shared__` storage is one block's at a
// time).  The OS threads are created once per launch and walk the blocks together; every block gets fresh barriers,
// so threads that returned early from one block (CUDA: exited threads are not waited for) take part again in the next.
template <typename Body>
void launch(Dim3 grid, Dim3 block, Body body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  std::vector<BlockState> states(nblocks);
  for (auto& st : states) {
    st.block_bar = std::make_unique<std::barrier<>>(nthreads);
    for (unsigned w = 0; w < nwarps; ++w)
      st.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, nthreads - 32 * w)));
    st.xchg.assign(nthreads, 0);
  }
  std::barrier<> next_block(nthreads);
  std::vector<std::thread> threads;
  threads.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) {
    threads.emplace_back([&, t]() {
      t_threadIdx = Dim3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      t_blockDim = block;
      t_gridDim = grid;
      size_t b = 0;
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx, ++b) {
            BlockState& st = states[b];
            t_blockIdx = Dim3{bx, by, bz};
            t_block = &st;
            body();
            // a thread that has returned no longer takes part in this block's barriers
            st.warp_bar[t / 32]->arrive_and_drop();
            st.block_bar->arrive_and_drop();
            next_block.arrive_and_wait();
          }
    });
  }
  for (auto& th : threads) th.join();
}

inline unsigned linear_tid() {
  return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
}

template <typename T>
inline T shfl_from(T v, unsigned src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  BlockState& st = *t_block;
  const unsigned tid = linear_tid(), warp = tid / 32;
  uint32_t w;
  std::memcpy(&w, &v, 4);
  st.xchg[tid] = w;
  st.warp_bar[warp]->arrive_and_wait();
  const uint32_t r = st.xchg[warp * 32 + (src_lane & 31)];
  st.warp_bar[warp]->arrive_and_wait();
  T out;
  std::memcpy(&out, &r, 4);
  return out;
}

}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

inline void __syncthreads() { emu::t_block->block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_block->warp_bar[emu::linear_tid() / 32]->arrive_and_wait(); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::shfl_from(v, (emu::linear_tid() & 31) ^ (unsigned)lane_mask);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  return emu::shfl_from(v, (unsigned)src_lane);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const unsigned lane = emu::linear_tid() & 31;
  const T r = emu::shfl_from(v, lane + delta < 32 ? lane + delta : lane);
  return r;
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return old;
}
template <typename T>
inline T __ldcs(const T* p) { emu::check_aligned(p, sizeof(T), "__ldcs"); return *p; }
template <typename T>
inline T __ldcg(const T* p) { emu::check_aligned(p, sizeof(T), "__ldcg"); return *p; }
template <typename T>
inline T __ldg(const T* p) { emu::check_aligned(p, sizeof(T), "__ldg"); return *p; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
