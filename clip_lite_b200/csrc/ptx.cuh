// Thin inline-PTX wrappers for the sm_100a features the dense kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the UMMA shared-memory / instruction descriptors.  No CUTLASS dependency.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace jsd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin on the barrier.  A pipeline bug would otherwise hang the GPU forever: after
// ~20 s without progress the kernel traps, which surfaces as a launch failure.
constexpr uint64_t kSpinLimitNs = 20000000000ull;
// (`limit_ns`: kernels whose producer may legitimately wait for a PEER GPU pass the peer-wait limit + this one, so
//  that their inner pipeline waits do not fire before the peer wait has reported what it was waiting for)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint64_t limit_ns = kSpinLimitNs) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > limit_ns) __trap();
    }
  }
}

// Same, for long waits (an epilogue warp waiting ~100 us for a 128-chunk accumulator): sleep between polls.
// The chip runs at its power cap under these kernels, so 16 warps spinning at full rate cost clock.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns,
                                                  uint64_t limit_ns = kSpinLimitNs) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if ((++spins & 0xFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > limit_ns) __trap();
    }
  }
}

// ---------------------------------------------------------------- device-scope flags (stream-K hand-off)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- device-side event trace (development aid)
// nsys is not available on the GPU boxes and ncu serialises launches, so the kernels can stamp %globaltimer into a
// caller-provided buffer: one thread per kernel at entry / once its peers' flags are in / at exit.  Disabled
// (events == nullptr, the default) it costs one 8-byte load in that one thread.  See clip_lite_b200/trace.py.
// Compiled in only with -DJSD_TRACE=1 (clip_lite_b200.build.build_library(trace=True) -> libjsd_b200_trace.so); the
// default build contains no trace code at all.
#ifndef JSD_TRACE
#define JSD_TRACE 0
#endif
struct TraceBuf {
  unsigned long long* events;   // [capacity]: (kernel id << 60) | (event << 56) | (globaltimer ns & (2^56 - 1))
  int* count;                   // next free slot
  int capacity;
};
#if JSD_TRACE
__device__ TraceBuf g_trace = {nullptr, nullptr, 0};
#endif
enum TraceKernel { TK_NORMALIZE = 1, TK_FWD = 2, TK_GRAD = 3, TK_JACOBIAN = 4, TK_INDEX = 5, TK_SCORE = 6, TK_PUSH = 7 };
enum TraceEvent { TE_START = 0, TE_PEERS_IN = 1, TE_END = 2 };
__device__ __forceinline__ void trace_event(int kernel, int event) {
#if JSD_TRACE
  unsigned long long* ev = g_trace.events;
  if (ev == nullptr) return;
  const int i = atomicAdd(g_trace.count, 1);
  if (i < g_trace.capacity)
    ev[i] = ((unsigned long long)kernel << 60) | ((unsigned long long)event << 56) |
            (global_timer_ns() & ((1ull << 56) - 1));
#else
  (void)kernel;
  (void)event;
#endif
}

// ---------------------------------------------------------------- system-scope flags (peer GPUs over NVLink)
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy observations (the flag) ordered before later async-proxy (TMA) accesses of global memory
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
// Peer waits are bounded: a rank that never publishes (crashed, or more than the timeout behind -- a dataloader
// respawn, a checkpoint write outside any barrier) must not hang the GPU forever.  The limit is a library parameter
// (jsd_peer_set_timeout / JSD_PEER_TIMEOUT_S, default 300 s, NCCL-like), and on expiry the waiting thread first
// records WHICH flag it was waiting for in a host-mapped error word -- readable by the host even after the trap
// has poisoned the context (jsd_peer_wait_error) -- and only then traps.
struct WaitCfg {
  unsigned long long timeout_ns;
  int* err_host;                // mapped pinned host memory (may be null): [0] = code, [1] = flag index, [2] = target
};
__device__ WaitCfg g_wait_cfg = {300000000000ull, nullptr};
enum WaitKind { WAIT_GATHERED_ROWS = 1, WAIT_GRAD_PARTIALS = 2 };

__device__ __noinline__ void wait_timed_out(int kind, int index, int target) {
  int* e = g_wait_cfg.err_host;
  if (e != nullptr) {
    e[1] = index;
    e[2] = target;
    __threadfence_system();
    e[0] = kind;
    __threadfence_system();
  }
  __trap();
}
// Spin until flags[i] has reached `target` (flags only grow).
__device__ __forceinline__ void wait_flag_sys(const int* flags, int i, int target, int kind) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (ld_acquire_sys(flags + i) - target < 0) {
    if ((++spins & 0xFFu) == 0) {
      __nanosleep(200);
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > g_wait_cfg.timeout_ns) wait_timed_out(kind, i, target);
    }
  }
}
// ... until every flag[0..count) has
__device__ __forceinline__ void wait_flags_sys(const int* flags, int count, int target, int kind) {
  for (int i = 0; i < count; ++i) wait_flag_sys(flags, i, target, kind);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-range rows/columns of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// 1-D bulk copy global -> shared, completion (bytes) signalled on an mbarrier: `bytes` a multiple of 16, both
// addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D bulk copy shared -> global (bulk async group): `bytes` a multiple of 16, both addresses 16-byte aligned.
// The destination may be peer memory: the copy leaves the SM as full-size packets, no LSU involvement.
__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed entirely
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// ---------------------------------------------------------------- clusters / CTA pairs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// CTA-pair TMA load: data lands in THIS CTA's smem, completion bytes are counted on `cluster_bar`,
// which may live in the peer (leader) CTA.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA-pair (cta_group::2) variants: issued by the same warp index in BOTH CTAs of the pair
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows per CTA), B split along N.
// Issued by one thread of the leader CTA only.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = lane t of the warp's quarter).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same for 16 consecutive columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int CW>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[CW]) {
  static_assert(CW == 16 || CW == 32, "chunk width");
  if constexpr (CW == 32) tmem_ld_32x32(taddr, r);
  else tmem_ld_32x16(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B, 16-bit elements.
//   K-major  tile: rows of 128 B (64 elements of K); 8-row groups 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major tile: rows of 128 B (64 elements of M/N), one row per k; 8-k groups SBO apart,
//                  64-element MN atoms LBO apart.
__host__ __device__ constexpr uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) /* version */ |
         (2ull << 61) /* SWIZZLE_128B */;
}
// Instruction descriptor for kind::f16, bf16 inputs, fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ | (a_mn_major << 15) |
         (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------- misc math
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace jsd
