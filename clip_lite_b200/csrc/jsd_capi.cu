// C ABI of libjsd_b200.so (see include/jsd_b200.h): argument checks, TMA tensor-map
// encoding and kernel launches.  Everything is asynchronous on the caller's stream.
#include "../../include/jsd_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "jsd_dense.cuh"
#include "jsd_fused.cuh"
#include "jsd_heads.cuh"
#include "jsd_rowwise.cuh"
#include "jsd_score.cuh"

#ifndef JSD_L2_PROMO
#define JSD_L2_PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define JSD_CUDA_OK(expr)                                                                    \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e_));      \
  } while (0)

#define JSD_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) return fail(__VA_ARGS__);  \
  } while (0)

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeFn>(p);
  return fn;
}

// bf16 matrix [outer, inner] with row pitch `pitch_elems`; box = box_inner x box_outer, 128B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, int64_t inner, int64_t outer, int64_t pitch_elems, int box_inner,
              int box_outer) {
  EncodeFn enc = get_encode_fn();
  JSD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  JSD_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  JSD_REQUIRE((pitch_elems * 2) % 16 == 0, "TMA operand row pitch must be a multiple of 16 bytes (got %lld elems)",
              (long long)pitch_elems);
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, JSD_L2_PROMO,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JSD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

size_t streamk_flag_bytes() { return (size_t)jsd::SK_MAX_CTAS * jsd::NUM_EPI_WARPS * sizeof(int); }

int sm_count_cached() {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
  }
  return cached;
}

// CTA-group size of the tensor-core kernels: 2 (CTA pairs, 256x256 tiles) whenever there is more than
// one 128-row block; JSD_CTA_GROUP=1|2 overrides (development / A-B timing).
// development knob: JSD_TILE_ORDER=fwd:grad with 0 = m-fastest, 1 = n-fastest (e.g. "1:0")
int tile_order(bool grad, int dflt) {
  static int v[2] = {-1, -1};
  if (v[0] < 0) {
    v[0] = v[1] = -2;
    const char* e = getenv("JSD_TILE_ORDER");
    if (e && strlen(e) >= 3) {
      v[0] = e[0] - '0';
      v[1] = e[2] - '0';
    }
  }
  const int x = v[grad ? 1 : 0];
  return (x == 0 || x == 1) ? x : dflt;
}

int pick_cta_group(int64_t m_rows) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("JSD_CTA_GROUP");
    forced = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  if (forced) return forced;
  return m_rows > jsd::BLOCK_M ? 2 : 1;
}

// Stream-K policies of the GRAD launches (a workspace must be given for any of them):
//   0               never (the product's fused backward entry points: they use plan_split() instead, measured
//                   3x faster on underfilled launches -- stream-K's hand-off costs ~30 us there);
//   SK_UNDERFILLED  cut tiles when they fill at most half of the workers;
//   SK_RAGGED       additionally balance a ragged last wave (1.73 waves at B = 8192, D = 1024).  Implemented and
//                   tested, but measured 2-5 % slower than whole tiles on the power-capped B200,
//                   so only the explicit GEMM entry points ask for it.
enum SkPolicy { SK_UNDERFILLED = 1, SK_RAGGED = 2 };

template <int MODE, bool A_MN, bool B_MN, int CG>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, jsd::GemmParams p, void* sk_workspace,
                cudaStream_t st, int sk_policy = SK_RAGGED, int worker_cap = 0) {
  auto kern = jsd::jsd_gemm_kernel<MODE, A_MN, B_MN, CG>;
  constexpr int smem = jsd::gemm_smem_bytes(CG, MODE);
  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set it once per device
  // (a process normally drives one GPU, but nothing here may assume so)
  static unsigned long long configured_mask = 0;
  int dev = 0;
  JSD_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !((configured_mask >> dev) & 1ull)) {
    JSD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) configured_mask |= 1ull << dev;
  }
  const int sms = sm_count_cached();
  JSD_REQUIRE(sms >= CG, "no CUDA device");
  int max_workers = sms / CG;                               // a worker = one CTA or one CTA pair
  if (worker_cap > 0 && worker_cap / CG < max_workers) max_workers = worker_cap / CG;   // worker_cap: SMs this launch
                                                                                        // may take (a sibling launch
                                                                                        // gets the others)
  JSD_REQUIRE(max_workers >= 1, "worker cap too small");
  const int n_blocks = (p.N + jsd::BLOCK_N - 1) / jsd::BLOCK_N;
  const long long tiles = (long long)((p.M + jsd::BLOCK_M * CG - 1) / (jsd::BLOCK_M * CG)) * n_blocks;
  const long long work_items = tiles * (MODE == jsd::MODE_GRAD && p.ksplit > 1 ? p.ksplit : 1);   // split-K slices
  if (MODE != jsd::MODE_GRAD) p.ksplit = 1;
  int workers = (int)(work_items < max_workers ? work_items : max_workers);
  // stream-K (opt-in: the caller passes a workspace), see SkPolicy
  p.stream_k = 0;
  const int sk_workers = max_workers / n_blocks * n_blocks;   // whole groups of n_blocks workers
  static int sk_enabled = -1;                                  // development knob: JSD_STREAMK=0 disables it
  if (sk_enabled < 0) {
    const char* e = getenv("JSD_STREAMK");
    sk_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (MODE == jsd::MODE_GRAD && sk_enabled && sk_workspace != nullptr && p.n_fastest && sk_workers >= n_blocks) {
    const long long m_blocks = tiles / n_blocks;
    const int groups = sk_workers / n_blocks;
    const long long sk_mb = m_blocks % groups;                 // m-blocks that are cut along the contraction
    const long long nk = (p.K + jsd::BLOCK_K * jsd::k_atoms(MODE) - 1) / (jsd::BLOCK_K * jsd::k_atoms(MODE));
    const bool underfilled = sk_policy >= SK_UNDERFILLED && tiles * 2 <= sk_workers;
    const bool ragged = sk_policy >= SK_RAGGED && tiles > sk_workers && sk_mb != 0;
    if ((underfilled || ragged) && sk_workers * 16 >= max_workers * 15 && sk_workers * CG <= jsd::SK_MAX_CTAS &&
        sk_mb * nk >= groups) {                                // every group receives at least one k-chunk
      p.stream_k = 1;
      p.sk_flags = reinterpret_cast<int*>(sk_workspace);
      p.sk_slots = reinterpret_cast<float*>(static_cast<char*>(sk_workspace) + streamk_flag_bytes());
      workers = sk_workers;
    }
  }
  const int grid = workers * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(jsd::GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  JSD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  return 0;
}

template <int MODE>
int launch_gemm_any(bool a_mn, bool b_mn, int cg, const CUtensorMap& tmA, const CUtensorMap& tmB,
                    const jsd::GemmParams& p, void* sk_workspace, cudaStream_t st, int sk_policy = SK_RAGGED,
                    int worker_cap = 0) {
#define JSD_GEMM_CASE(A, B, C)           \
  if (a_mn == A && b_mn == B && cg == C) \
    return launch_gemm<MODE, A, B, C>(tmA, tmB, p, sk_workspace, st, sk_policy, worker_cap);
  JSD_GEMM_CASE(false, false, 1) JSD_GEMM_CASE(false, true, 1) JSD_GEMM_CASE(true, false, 1) JSD_GEMM_CASE(true, true, 1)
  JSD_GEMM_CASE(false, false, 2) JSD_GEMM_CASE(false, true, 2) JSD_GEMM_CASE(true, false, 2) JSD_GEMM_CASE(true, true, 2)
#undef JSD_GEMM_CASE
  return fail("unsupported GEMM variant");
}

bool fits_int(int64_t v) { return v > 0 && v < (int64_t)1 << 30; }

// JSD_INDEX_RING=0 sends the normal mode through the L1-based kernel as well (A/B timing)
bool index_ring_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JSD_INDEX_RING");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <typename T>
int launch_index(const void* F, const void* G, int64_t B, int64_t D, const int32_t* neg, const int32_t* iptr,
                 const int32_t* iidx, const float* t_dev, float* coefp, float* partials, void* dF, void* dG,
                 float grad_scale, const float* gamma_dev, cudaStream_t st, int* n_partials) {
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(G) |
                                     reinterpret_cast<uintptr_t>(dF) | reinterpret_cast<uintptr_t>(dG)) & 15) == 0;
  *n_partials = (int)((B + jsd::INDEX_ROWS_PER_CTA - 1) / jsd::INDEX_ROWS_PER_CTA);
  // normal (roll-by-one) pairing on 16-byte-granular rows: persistent CTAs, rows staged once through shared-memory
  // rings by the bulk-copy engine
  const size_t row_bytes = (size_t)D * sizeof(T);
  // (long rows only: with 1 KB rows the per-row barrier round trips of eight consumer warps dominate -- 204 vs 95 us
  //  at 65536 x 512 bf16; at 8 KB rows the ring wins 75.6 -> 58.4 us per call, r02s)
  if (neg == nullptr && vec && row_bytes % 16 == 0 && row_bytes >= 8192 && index_ring_enabled()) {
    int slots = (int)((size_t)(196 * 1024) / (2 * row_bytes));
    if (slots > jsd::IR_MAX_SLOTS) slots = jsd::IR_MAX_SLOTS;
    const int sms = sm_count_cached();
    if (slots >= 4 && sms > 0) {
      const size_t smem = 2 * (size_t)slots * row_bytes;
      auto kern = jsd::jsd_index_ring_kernel<T>;
      static unsigned long long configured_mask = 0;
      int dev = 0;
      JSD_CUDA_OK(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64 || !((configured_mask >> dev) & 1ull)) {
        JSD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) configured_mask |= 1ull << dev;
      }
      const int grid = (int)(B < sms ? B : sms);
      kern<<<grid, jsd::IR_THREADS, smem, st>>>((const T*)F, (const T*)G, (int)B, (int)D, slots, t_dev, partials,
                                                (T*)dF, (T*)dG, grad_scale, gamma_dev);
      JSD_CUDA_OK(cudaGetLastError());
      *n_partials = grid;
      return 0;
    }
  }
  const unsigned grid = (unsigned)((B + jsd::INDEX_ROWS_PER_CTA - 1) / jsd::INDEX_ROWS_PER_CTA);
  const int threads = 32 * jsd::INDEX_ROWS_PER_CTA;
  if (vec)
    jsd::jsd_index_kernel<T, 4><<<grid, threads, 0, st>>>((const T*)F, (const T*)G, (int)B, (int)D, neg, iptr, iidx,
                                                          t_dev, coefp, partials, (T*)dF, (T*)dG, grad_scale,
                                                          gamma_dev);
  else
    jsd::jsd_index_kernel<T, 1><<<grid, threads, 0, st>>>((const T*)F, (const T*)G, (int)B, (int)D, neg, iptr, iidx,
                                                          t_dev, coefp, partials, (T*)dF, (T*)dG, grad_scale,
                                                          gamma_dev);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_normalize(const jsd::NormalizeJob& job, int count, int64_t rows, int64_t D, cudaStream_t st) {
  uintptr_t bits = 0;
  for (int i = 0; i < count; ++i) bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.Xn[i]);
  const bool vec = (D % 4 == 0) && (bits & 15) == 0;
  const dim3 grid((unsigned)((rows + 7) / 8), (unsigned)count);
  if (vec && D % 128 == 0 && D / 128 <= jsd::ROW_REG_CHUNKS)
    jsd::normalize_cast_reg_kernel<T><<<grid, 256, 0, st>>>(job, (int)rows, (int)(D / 128));
  else if (vec)
    jsd::normalize_cast_kernel<T, 4><<<grid, 256, 0, st>>>(job, (int)rows, (int)D);
  else
    jsd::normalize_cast_kernel<T, 1><<<grid, 256, 0, st>>>(job, (int)rows, (int)D);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_normalize_bwd(const jsd::NormBwdJob& job, int count, int64_t rows, int64_t D, const float* gdiag,
                         const float* t_dev, const float* gamma_dev, float inv_rows, cudaStream_t st) {
  uintptr_t bits = 0;
  for (int i = 0; i < count; ++i)
    bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.dX[i]) |
            reinterpret_cast<uintptr_t>(job.acc[i]) | reinterpret_cast<uintptr_t>(job.partner[i]);
  for (int q = 0; q < job.acc_slots; ++q)
    bits |= reinterpret_cast<uintptr_t>(job.slot[q]) | (count > 1 ? reinterpret_cast<uintptr_t>(job.slot1[q]) : 0);
  const bool vec = (D % 4 == 0) && (bits & 15) == 0;
  const dim3 grid((unsigned)((rows + 7) / 8), (unsigned)count);
  if (vec && D % 128 == 0 && D / 128 <= jsd::ROW_REG_CHUNKS)
    jsd::normalize_bwd_reg_kernel<T><<<grid, 256, 0, st>>>(job, (int)rows, (int)(D / 128), gdiag, t_dev, gamma_dev,
                                                           inv_rows);
  else if (vec)
    jsd::normalize_bwd_kernel<T, 4><<<grid, 256, 0, st>>>(job, (int)rows, (int)D, gdiag, t_dev, gamma_dev, inv_rows);
  else
    jsd::normalize_bwd_kernel<T, 1><<<grid, 256, 0, st>>>(job, (int)rows, (int)D, gdiag, t_dev, gamma_dev, inv_rows);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

// One non-blocking helper stream and two events per device: the image-side Jacobian kernel (HBM-bound) runs on it
// next to the dV contraction, whose ragged last wave leaves 40 SMs idle for half of its run time.  Fork / join by
// events, so the pattern is legal inside a CUDA-graph capture of the caller's stream.  JSD_OVERLAP=0 disables it.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, mid = nullptr, mid2 = nullptr;
};

SideStream* side_stream() {
  static SideStream per_dev[64];
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("JSD_OVERLAP");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (s.stream == nullptr) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.mid, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.mid2, cudaEventDisableTiming) != cudaSuccess) {
      s = SideStream{};
      return nullptr;
    }
  }
  return &s;
}

// All-gather of the text rows by the COPY ENGINES: a few non-blocking streams per device on which
// jsd_peer_normalize_push enqueues, behind the normalise launch, one peer-to-peer copy of this rank's row block per
// destination, each followed by a 4-byte copy of the step counter into the destination's "rows of rank r are in"
// flag (stream order = the flag lands after the rows).  No SM is involved and nothing is fenced: pushing from the
// forward kernel's idle warps works too (JSD_PEER_GATHER=sm) but a system-scope fence per destination takes 30-40 us
// on an SM that is busy feeding the tensor cores (traces r02d / r02f).  jsd_peer_dense_fwd joins the copy streams
// behind its launch, so the exchange runs underneath the forward and everything later is ordered after it.
constexpr int kCopyStreams = 3;
struct CopyStreams {
  cudaStream_t stream[kCopyStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[kCopyStreams] = {nullptr, nullptr, nullptr};
  bool pending = false;
};

CopyStreams* copy_streams() {
  static CopyStreams per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  CopyStreams& c = per_dev[dev];
  if (c.fork == nullptr) {
    bool ok = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kCopyStreams && ok; ++i)
      ok = cudaStreamCreateWithFlags(&c.stream[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      c = CopyStreams{};
      (void)cudaGetLastError();
      return nullptr;
    }
  }
  return &c;
}

// How the text rows reach the other ranks (measured at B = 8192, D = 1024; traces under profiles/):
//   GATHER_FWD     stores from an idle warp of the FORWARD kernel, one flag per destination: costs no SM time and
//                  no launch, but ~10 us per destination -- hidden completely behind the forward of 2 / 4 GPUs
//                  (E(2) = 0.985), far too slow for the 22 us forward of 8 GPUs (86-110 us);
//   GATHER_KERNEL  the normalise launch stores every row into every rank's buffer and publishes all flags at its
//                  end: ~25-30 us in front of the forward, independent of what the forward does;
//   GATHER_CE      peer-to-peer copies on the copy engines underneath the forward: no SM involved, but ~10 us of
//                  latency per copy node (70-85 us for 7 x 2 MB at 8 GPUs).
// Default: FWD up to 4 ranks, KERNEL above; JSD_PEER_GATHER=fwd|kernel|ce overrides (A/B timing).
enum GatherMode { GATHER_FWD = 0, GATHER_KERNEL = 1, GATHER_CE = 2 };
int peer_gather_mode(int world) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("JSD_PEER_GATHER");
    forced = !e ? -1 : (e[0] == 'f' ? GATHER_FWD : (e[0] == 'k' ? GATHER_KERNEL : (e[0] == 'c' ? GATHER_CE : -1)));
  }
  if (forced >= 0) return forced;
  return world <= 4 ? GATHER_FWD : GATHER_KERNEL;
}

// Peer waits (ptx.cuh: WaitCfg): time limit + host-mapped error word, installed once per device.
double g_wait_timeout_s = -1.0;
unsigned long long g_wait_cfg_mask = 0;
volatile int* g_wait_err_host = nullptr;

int ensure_wait_cfg() {
  int dev = 0;
  JSD_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && ((g_wait_cfg_mask >> dev) & 1ull)) return 0;
  if (g_wait_timeout_s <= 0.0) {
    const char* e = getenv("JSD_PEER_TIMEOUT_S");
    const double v = e ? atof(e) : 0.0;
    g_wait_timeout_s = v > 0.0 ? v : 300.0;
  }
  if (g_wait_err_host == nullptr) {
    int* h = nullptr;
    if (cudaHostAlloc((void**)&h, 4 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      memset(h, 0, 4 * sizeof(int));
      g_wait_err_host = h;
    } else {
      (void)cudaGetLastError();       // no mapped memory: the trap still fires, only the report is lost
    }
  }
  jsd::WaitCfg cfg;
  cfg.timeout_ns = (unsigned long long)(g_wait_timeout_s * 1e9);
  cfg.err_host = const_cast<int*>(g_wait_err_host);
  // (synchronous, not legal during a stream capture: the eager warm-up step every capture needs installs it)
  JSD_CUDA_OK(cudaMemcpyToSymbol(jsd::g_wait_cfg, &cfg, sizeof(cfg)));
  if (dev >= 0 && dev < 64) g_wait_cfg_mask |= 1ull << dev;
  return 0;
}

// development knob: JSD_PEER_WAIT_ALL=1 makes the peer forward wait for every rank before its first load (round 1)
bool peer_wait_per_source() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JSD_PEER_WAIT_ALL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

// the second int of the forward workspace's 16-byte ticket area serialises the dL/dt reduction
int* dt_ticket(void* workspace) { return reinterpret_cast<int*>(workspace) + 1; }

#define JSD_DISPATCH_DTYPE(dtype, CALL)                                    \
  switch (dtype) {                                                         \
    case JSD_F32: { using T = float; return CALL; }                        \
    case JSD_BF16: { using T = __nv_bfloat16; return CALL; }               \
    case JSD_F16: { using T = __half; return CALL; }                       \
    default: return fail("unsupported dtype code %d", (int)(dtype));       \
  }

}  // namespace

extern "C" {

int jsd_abi_version(void) { return 12; }

const char* jsd_last_error(void) { return g_err; }

int jsd_sm_count(void) { return sm_count_cached(); }

size_t jsd_index_workspace_bytes(int64_t B) { return (size_t)(B > 0 ? B : 0) * 4 * sizeof(float); }

int jsd_index_fwd_bwd(const void* F, const void* G, int dtype, int64_t B, int64_t D, const int32_t* neg_index,
                      const int32_t* inv_ptr, const int32_t* inv_idx, const float* t_dev, void* workspace,
                      float* out4, float* loss_out, void* dF, void* dG, float grad_scale, const float* gamma_dev,
                      jsd_stream_t stream) {
  JSD_REQUIRE(F && G && t_dev && workspace && out4, "jsd_index_fwd_bwd: null pointer argument");
  JSD_REQUIRE((dF == nullptr) == (dG == nullptr), "jsd_index_fwd_bwd: dF and dG must be given together");
  JSD_REQUIRE(grad_scale > 0.f, "jsd_index_fwd_bwd: grad_scale must be positive");
  JSD_REQUIRE(fits_int(B) && fits_int(D), "jsd_index_fwd_bwd: B=%lld, D=%lld out of range", (long long)B, (long long)D);
  JSD_REQUIRE((neg_index == nullptr) == (inv_ptr == nullptr) && (inv_ptr == nullptr) == (inv_idx == nullptr),
              "jsd_index_fwd_bwd: neg_index, inv_ptr and inv_idx must be given together");
  cudaStream_t st = (cudaStream_t)stream;
  float* coefp = (float*)workspace;
  float* partials = coefp + B;
  int n_partials = 0;
  int rc = [&]() -> int {
    JSD_DISPATCH_DTYPE(dtype, (launch_index<T>(F, G, B, D, neg_index, inv_ptr, inv_idx, t_dev, coefp, partials, dF,
                                               dG, grad_scale, gamma_dev, st, &n_partials)));
  }();
  if (rc) return rc;
  jsd::finalize_kernel<<<1, jsd::FINALIZE_THREADS, 0, st>>>(partials, n_partials, 3, 1.0 / (double)B, 1.0 / (double)B,
                                                             1.0, 0.0, out4, loss_out);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

int jsd_normalize_cast(const void* X, int dtype, int64_t rows, int64_t D, void* Xn, float* inv_norm,
                       jsd_stream_t stream) {
  JSD_REQUIRE(X && Xn && inv_norm, "jsd_normalize_cast: null pointer argument");
  JSD_REQUIRE(fits_int(rows) && fits_int(D), "jsd_normalize_cast: rows=%lld, D=%lld out of range", (long long)rows,
              (long long)D);
  jsd::NormalizeJob job{};
  job.X[0] = X;
  job.Xn[0] = (__nv_bfloat16*)Xn;
  job.inv_norm[0] = inv_norm;
  cudaStream_t st = (cudaStream_t)stream;
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize<T>(job, 1, rows, D, st)));
}

int jsd_normalize_cast_pair(const void* F, const void* G, int dtype, int64_t rows, int64_t D, void* U, void* V,
                            float* inv_f, float* inv_g, jsd_stream_t stream) {
  JSD_REQUIRE(F && G && U && V && inv_f && inv_g, "jsd_normalize_cast_pair: null pointer argument");
  JSD_REQUIRE(fits_int(rows) && fits_int(D), "jsd_normalize_cast_pair: rows=%lld, D=%lld out of range",
              (long long)rows, (long long)D);
  jsd::NormalizeJob job{};
  job.X[0] = F;
  job.X[1] = G;
  job.Xn[0] = (__nv_bfloat16*)U;
  job.Xn[1] = (__nv_bfloat16*)V;
  job.inv_norm[0] = inv_f;
  job.inv_norm[1] = inv_g;
  cudaStream_t st = (cudaStream_t)stream;
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize<T>(job, 2, rows, D, st)));
}

size_t jsd_dense_workspace_bytes(void) {
  return 16 + (size_t)1024 * jsd::NUM_EPI_WARPS * jsd::PARTIALS_PER_WARP * sizeof(float);   // ticket + up to 1024 CTAs
}

}  // extern "C"

namespace {
struct PeerWait {
  const int* flags = nullptr;
  const int* counter = nullptr;
  int count = 0;
  int rows = 0;        // rows per source rank (0: wait for every rank before the first load)
  const jsd_peer_ctx* ctx = nullptr;   // set: own rows need no flag; with push_in_kernel the forward also copies
  int parity = 0;                      // them to the other ranks
  bool push_in_kernel = false;
};
int dense_fwd_impl(const void* U, const void* V, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                   const float* t_dev, void* Gmat, int64_t ldg, float* gdiag, void* workspace, float* out4,
                   float* loss_out, const PeerWait& wait, jsd_stream_t stream);
}  // namespace

extern "C" {

int jsd_dense_fwd(const void* U, const void* V, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                  const float* t_dev, void* Gmat, int64_t ldg, float* gdiag, void* workspace, float* out4,
                  float* loss_out, jsd_stream_t stream) {
  return dense_fwd_impl(U, V, M, N, D, row_offset, t_dev, Gmat, ldg, gdiag, workspace, out4, loss_out, PeerWait{},
                        stream);
}

}  // extern "C"

namespace {
int dense_fwd_impl(const void* U, const void* V, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                   const float* t_dev, void* Gmat, int64_t ldg, float* gdiag, void* workspace, float* out4,
                   float* loss_out, const PeerWait& wait, jsd_stream_t stream) {
  JSD_REQUIRE(U && V && t_dev && gdiag && workspace && out4, "jsd_dense_fwd: null pointer argument");
  JSD_REQUIRE(fits_int(M) && fits_int(N) && fits_int(D), "jsd_dense_fwd: M=%lld N=%lld D=%lld out of range",
              (long long)M, (long long)N, (long long)D);
  JSD_REQUIRE(D % 8 == 0, "jsd_dense_fwd: D must be a multiple of 8 (got %lld)", (long long)D);
  JSD_REQUIRE(row_offset >= 0 && row_offset + M <= N, "jsd_dense_fwd: positives [%lld, %lld) outside the %lld columns",
              (long long)row_offset, (long long)(row_offset + M), (long long)N);
  if (Gmat) {
    JSD_REQUIRE(ldg >= N && ldg % 64 == 0, "jsd_dense_fwd: ldg must be a multiple of 64 and >= N");
    JSD_REQUIRE((reinterpret_cast<uintptr_t>(Gmat) & 15) == 0, "jsd_dense_fwd: Gmat must be 16-byte aligned");
  }
  const int cg = pick_cta_group(M);
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap(&tmA, U, D, M, D, jsd::BLOCK_K, jsd::BLOCK_M)) return rc;
  if (int rc = make_tmap(&tmB, V, D, N, D, jsd::BLOCK_K, jsd::b_rows_per_cta(cg))) return rc;
  jsd::GemmParams p{};
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)D;
  p.n_fastest = tile_order(false, 0);
  p.row_offset = (int)row_offset;
  p.t_dev = t_dev;
  p.gmat = (__nv_bfloat16*)Gmat;
  p.ldg = ldg;
  p.gdiag = gdiag;
  // workspace = [ticket (16 bytes, zero between launches) | per-warp partial sums]
  p.ticket = (int*)workspace;
  p.partials = (float*)((char*)workspace + 16);
  p.out4 = out4;
  p.loss_out = loss_out;
  p.inv_pos = 1.0 / (double)M;
  p.inv_neg = N > 1 ? 1.0 / ((double)M * (double)(N - 1)) : 0.0;
  p.wait_flags = wait.flags;
  p.wait_counter = wait.counter;
  p.wait_count = wait.count;
  p.wait_rows = wait.rows;
  p.wait_skip = -1;
  if (wait.flags != nullptr && wait.rows > 0) {
    // start on this rank's own column block, then walk the ranks upwards (rank + 1, rank + 2, ...): the order in
    // which their rows arrive (every sender serves rank - 1, rank - 2, ... in turn)
    p.n_rot = (int)(row_offset / jsd::BLOCK_N);
  }
  if (wait.ctx != nullptr) {
    const jsd_peer_ctx* c = wait.ctx;
    p.wait_skip = c->rank;
    p.gath_world = wait.push_in_kernel ? c->world : 0;
    p.gath_chunks = (int)(c->rows * c->dim / 8);
    const size_t own = (size_t)c->rank * c->rows * c->dim;
    p.gath_src = reinterpret_cast<const uint4*>((const __nv_bfloat16*)c->v_all[wait.parity][c->rank] + own);
    for (int k = 1; k < c->world; ++k) {
      const int q = (c->rank - k + c->world) % c->world;
      p.gath_dst[k] = reinterpret_cast<uint4*>((__nv_bfloat16*)c->v_all[wait.parity][q] + own);
      p.gath_flag_dst[k] = c->flags[q] + JSD_PEER_READY_V + wait.parity * JSD_MAX_PEERS + c->rank;
    }
    p.gath_ticket = c->flags[c->rank] + JSD_PEER_TICKET_PUSH;
  }
  if (Gmat) {
    // epilogue TMA stores: box = 64 columns x 32 rows per warp; rows >= M / columns >= N are clipped
    if (int rc = make_tmap(&p.tmG, Gmat, N, M, ldg, jsd::COLS_PER_WARP, 32)) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return cg == 2 ? launch_gemm<jsd::MODE_FWD, false, false, 2>(tmA, tmB, p, nullptr, st)
                 : launch_gemm<jsd::MODE_FWD, false, false, 1>(tmA, tmB, p, nullptr, st);
}
}  // namespace

extern "C" {

size_t jsd_streamk_flag_bytes(void) { return streamk_flag_bytes(); }

// split-K partial slices of an underfilled GRAD launch: (ksplit - 1) * rows * D floats with tiles * ksplit <= 74 CTA
// pairs, i.e. at most 74 * 256 * 256 floats; two regions (dU and dV of one step may be live together)
constexpr size_t kSplitRegionBytes = (size_t)7 * 2048 * 1024 * sizeof(float);   // 7 extra slices of a 2048 x 1024 block

size_t jsd_streamk_workspace_bytes(void) {
  const size_t sk = (size_t)jsd::SK_MAX_CTAS * jsd::SK_SLOT_FLOATS * sizeof(float);
  return streamk_flag_bytes() + (sk > 2 * kSplitRegionBytes ? sk : 2 * kSplitRegionBytes);
}

// Split-K plan of an underfilled GRAD launch whose consumer is the library's own Jacobian kernel: when the output
// tiles fill at most half of the CTA pairs (dU of a 1024-row rank: 16 tiles for 74 pairs), every tile is computed
// in `ksplit` K-slices by different workers; the slices land in `region` of the workspace and the Jacobian kernel
// adds them in order (NormBwdJob::slot).  ksplit = 1: nothing to do.
struct SplitPlan {
  int ksplit = 1;
  float* slice_base = nullptr;
  long long slice_stride = 0;
};

static SplitPlan plan_split(int64_t rows, int64_t D, int64_t kdim, void* workspace, int region) {
  SplitPlan sp;
  static int enabled = -1;                                     // development knob: JSD_SPLITK=0 disables it
  if (enabled < 0) {
    const char* e = getenv("JSD_SPLITK");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || workspace == nullptr) return sp;
  const int cg = pick_cta_group(rows);
  const int workers = sm_count_cached() / cg;
  const long long tiles = ((rows + jsd::BLOCK_M * cg - 1) / (jsd::BLOCK_M * cg)) * ((D + jsd::BLOCK_N - 1) / jsd::BLOCK_N);
  const long long nk = (kdim + jsd::BLOCK_K * jsd::k_atoms(jsd::MODE_GRAD) - 1) / (jsd::BLOCK_K * jsd::k_atoms(jsd::MODE_GRAD));
  if (tiles <= 0 || tiles * 2 > workers) return sp;
  long long kmax = workers / tiles;
  if (kmax > 8) kmax = 8;                                      // NormBwdJob sums at most 8 slices
  if (kmax > nk / 2) kmax = nk / 2;                            // at least two k-chunks per slice
  // Cost model (measured on B200): a k-chunk of a 256 x 256 pair tile takes ~0.7 us, so s slices save
  // (1 - 1/s) nk 0.7 us of serial MMA time; every extra slice is rows x D fp32 written by the GEMM and read back by
  // the Jacobian kernel (~3 TB/s effective).  B = 1024, D = 1024 single GPU: splitting loses (measured 61 -> 68 us
  // per step); dU of a 1024-row rank against 8192 text rows: 53 -> 30 us.
  long long ks = 1;
  double best = 2.0;                                           // demand at least 2 us of gain
  for (long long c = 2; c <= kmax; ++c) {
    const double gain = (1.0 - 1.0 / (double)c) * (double)nk * 0.7 -
                        (double)(c - 1) * (double)rows * (double)D * 8.0 / 3.0e6;
    if (gain > best) {
      best = gain;
      ks = c;
    }
  }
  if (ks < 2) return sp;
  if ((size_t)(ks - 1) * rows * D * sizeof(float) > kSplitRegionBytes) return sp;
  sp.ksplit = (int)ks;
  sp.slice_base = reinterpret_cast<float*>(static_cast<char*>(workspace) + streamk_flag_bytes() + region * kSplitRegionBytes);
  sp.slice_stride = rows * D;
  return sp;
}

static int dense_bwd_common(bool dv, const void* Gmat, int64_t ldg, const void* X, int64_t M, int64_t N, int64_t D,
                            const float* t_dev, const float* gamma_dev, void* sk_workspace, float* out,
                            jsd_stream_t stream, const jsd_peer_ctx* peer = nullptr, int sk_policy = SK_RAGGED,
                            const SplitPlan* split = nullptr, int worker_cap = 0) {
  JSD_REQUIRE(Gmat && X && t_dev && (out || peer), "jsd_dense_bwd: null pointer argument");
  JSD_REQUIRE(fits_int(M) && fits_int(N) && fits_int(D), "jsd_dense_bwd: M=%lld N=%lld D=%lld out of range",
              (long long)M, (long long)N, (long long)D);
  JSD_REQUIRE(D % 8 == 0, "jsd_dense_bwd: D must be a multiple of 8");
  JSD_REQUIRE(ldg >= N && ldg % 8 == 0, "jsd_dense_bwd: bad ldg");
  const int64_t kdim = dv ? M : N;     // contraction length = rows of the bf16 operand X
  const int64_t rows = dv ? N : M;     // rows of the gradient
  CUtensorMap tmA, tmB;
  if (!dv) {
    if (int rc = make_tmap(&tmA, Gmat, N, M, ldg, jsd::BLOCK_K, jsd::BLOCK_M)) return rc;   // [M, N] K-major, K = N
  } else {
    if (int rc = make_tmap(&tmA, Gmat, N, M, ldg, 64, jsd::BLOCK_K)) return rc;             // MN-major: 64 j x 64 i boxes
  }
  if (int rc = make_tmap(&tmB, X, D, kdim, D, 64, jsd::BLOCK_K)) return rc;                  // [kdim, D] MN-major boxes
  jsd::GemmParams p{};
  p.M = (int)rows;
  p.N = (int)D;
  p.K = (int)kdim;
  p.n_fastest = tile_order(true, 1);
  p.t_dev = t_dev;
  p.gamma_dev = gamma_dev;
  p.scale = N > 1 ? (float)(1.0 / ((double)M * (double)(N - 1))) : 0.f;
  p.out = out;
  p.ldo = D;
  p.ksplit = 1;
  if (split != nullptr && split->ksplit > 1) {
    p.ksplit = split->ksplit;
    p.slice_base = split->slice_base;
    p.slice_stride = split->slice_stride;
    sk_workspace = nullptr;                                    // split-K and stream-K are exclusive
  }
  if (peer != nullptr) {
    // the partial over ALL text rows stays in this rank's (peer-mapped) buffer; the owner of each row block
    // reads it from there once this launch has published its flag
    p.out = (float*)peer->stage[peer->rank];
    p.peer_world = peer->world;
    int32_t* mine = peer->flags[peer->rank];
    for (int q = 0; q < peer->world; ++q) p.peer_flag_dst[q] = peer->flags[q] + JSD_PEER_READY_DV + peer->rank;
    p.peer_counter = mine + JSD_PEER_COUNTER_DV;
    p.peer_ticket = mine + JSD_PEER_TICKET_DV;
  }
  return launch_gemm_any<jsd::MODE_GRAD>(dv, true, pick_cta_group(rows), tmA, tmB, p, sk_workspace,
                                          (cudaStream_t)stream, sk_policy, worker_cap);
}

int jsd_dense_bwd_du(const void* Gmat, int64_t ldg, const void* V, int64_t M, int64_t N, int64_t D,
                     const float* t_dev, const float* gamma_dev, void* sk_workspace, float* dUacc,
                     jsd_stream_t stream) {
  return dense_bwd_common(false, Gmat, ldg, V, M, N, D, t_dev, gamma_dev, sk_workspace, dUacc, stream);
}

int jsd_dense_bwd_dv(const void* Gmat, int64_t ldg, const void* U, int64_t M, int64_t N, int64_t D,
                     const float* t_dev, const float* gamma_dev, void* sk_workspace, float* dVacc,
                     jsd_stream_t stream) {
  return dense_bwd_common(true, Gmat, ldg, U, M, N, D, t_dev, gamma_dev, sk_workspace, dVacc, stream);
}

static int normalize_bwd_impl(const void* X, int dtype, int64_t rows, int64_t D, const float* inv_norm,
                              const float* acc, const SplitPlan* split, const void* partner, int64_t partner_offset,
                              const float* gdiag, const float* t_dev, const float* gamma_dev, int64_t M_rows, void* dX,
                              float* rowdot, void* workspace, float* dt_out, jsd_stream_t stream,
                              const float* reduce_src = nullptr, int64_t reduce_n = 0, float* reduce_out = nullptr);

int jsd_normalize_bwd(const void* X, int dtype, int64_t rows, int64_t D, const float* inv_norm, const float* acc,
                      const void* partner, int64_t partner_offset, const float* gdiag, const float* t_dev,
                      const float* gamma_dev, int64_t M_rows, void* dX, float* rowdot, void* workspace, float* dt_out,
                      jsd_stream_t stream) {
  return normalize_bwd_impl(X, dtype, rows, D, inv_norm, acc, nullptr, partner, partner_offset, gdiag, t_dev, gamma_dev,
                            M_rows, dX, rowdot, workspace, dt_out, stream);
}

static int normalize_bwd_impl(const void* X, int dtype, int64_t rows, int64_t D, const float* inv_norm,
                              const float* acc, const SplitPlan* split, const void* partner, int64_t partner_offset,
                              const float* gdiag, const float* t_dev, const float* gamma_dev, int64_t M_rows, void* dX,
                              float* rowdot, void* workspace, float* dt_out, jsd_stream_t stream,
                              const float* reduce_src, int64_t reduce_n, float* reduce_out) {
  JSD_REQUIRE(X && inv_norm && acc && partner && t_dev && dX, "jsd_normalize_bwd: null pointer argument");
  JSD_REQUIRE(fits_int(rows) && fits_int(D) && M_rows > 0, "jsd_normalize_bwd: bad shape");
  JSD_REQUIRE(dt_out == nullptr || (rowdot && workspace), "jsd_normalize_bwd: dt_out needs rowdot and the workspace");
  jsd::NormBwdJob job{};
  job.X[0] = X;
  job.inv_norm[0] = inv_norm;
  job.acc[0] = acc;
  job.partner[0] = (const __nv_bfloat16*)partner;
  job.partner_offset[0] = partner_offset;
  job.dX[0] = dX;
  job.rowdot = rowdot;
  job.ticket = dt_out ? dt_ticket(workspace) : nullptr;
  job.dt_out = dt_out;
  job.reduce_src = reduce_src;
  job.reduce_n = (int)reduce_n;
  job.reduce_out = reduce_src ? reduce_out : nullptr;
  if (split != nullptr && split->ksplit > 1) {          // the accumulator is the sum of the split-K slices, in order
    job.acc_slots = split->ksplit;
    job.slot[0] = acc;
    for (int q = 1; q < split->ksplit; ++q) job.slot[q] = split->slice_base + (size_t)(q - 1) * split->slice_stride;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_rows = (float)(1.0 / (double)M_rows);
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize_bwd<T>(job, 1, rows, D, gdiag, t_dev, gamma_dev, inv_rows, st)));
}

int jsd_dense_forward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const float* t_dev, void* U,
                      void* V, float* inv_f, float* inv_g, void* Gmat, int64_t ldg, float* gdiag, void* workspace,
                      float* out4, float* loss_out, jsd_stream_t stream) {
  if (int rc = jsd_normalize_cast_pair(F, G, dtype, B, D, U, V, inv_f, inv_g, stream)) return rc;
  return jsd_dense_fwd(U, V, B, B, D, 0, t_dev, Gmat, ldg, gdiag, workspace, out4, loss_out, stream);
}

int jsd_dense_backward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const void* U, const void* V,
                       const float* inv_f, const float* inv_g, const void* Gmat, int64_t ldg, const float* gdiag,
                       const float* t_dev, const float* gamma_dev, float* acc_u, float* acc_v, float* rowdot,
                       void* workspace, void* sk_workspace, void* dF, void* dG, float* dt_out, jsd_stream_t stream) {
  JSD_REQUIRE(F && G && U && V && inv_f && inv_g && gdiag && dF && dG, "jsd_dense_backward: null pointer argument");
  JSD_REQUIRE(dt_out && acc_u && acc_v && rowdot && workspace, "jsd_dense_backward: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const SplitPlan su = plan_split(B, D, B, sk_workspace, 0), sv = plan_split(B, D, B, sk_workspace, 1);
  SideStream* side = side_stream();
  // Small batches (B = 1024, D = 1024: 16 pair tiles per contraction, ~20 us per launch almost all of it fixed cost
  // -- prologue, pipeline fill, un-overlapped epilogue): both contractions fit on the GPU together, so they run
  // SIDE BY SIDE (dV on the helper stream), then both Jacobians (+ dL/dt) in one launch.
  {
    const int cg = pick_cta_group(B);
    const int64_t tiles = ((B + jsd::BLOCK_M * cg - 1) / (jsd::BLOCK_M * cg)) * ((D + jsd::BLOCK_N - 1) / jsd::BLOCK_N);
    static int pair_small = -1;                                  // development knob: JSD_PAIRED=0 disables it
    if (pair_small < 0) {
      const char* e = getenv("JSD_PAIRED");
      pair_small = (e && e[0] == '0') ? 0 : 1;
    }
    if (pair_small && side != nullptr && su.ksplit == 1 && sv.ksplit == 1 && 2 * tiles * cg <= sm_count_cached()) {
      JSD_CUDA_OK(cudaEventRecord(side->fork, st));
      JSD_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
      if (int rc = dense_bwd_common(false, Gmat, ldg, V, B, B, D, t_dev, gamma_dev, nullptr, acc_u, stream, nullptr, 0))
        return rc;
      if (int rc = dense_bwd_common(true, Gmat, ldg, U, B, B, D, t_dev, gamma_dev, nullptr, acc_v,
                                    (jsd_stream_t)side->stream, nullptr, 0))
        return rc;
      JSD_CUDA_OK(cudaEventRecord(side->join, side->stream));
      JSD_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
      jsd::NormBwdJob job{};
      job.X[0] = F;
      job.X[1] = G;
      job.inv_norm[0] = inv_f;
      job.inv_norm[1] = inv_g;
      job.acc[0] = acc_u;
      job.acc[1] = acc_v;
      job.partner[0] = (const __nv_bfloat16*)V;
      job.partner[1] = (const __nv_bfloat16*)U;
      job.dX[0] = dF;
      job.dX[1] = dG;
      job.rowdot = rowdot;
      job.ticket = dt_ticket(workspace);
      job.dt_out = dt_out;
      const float inv_rows = (float)(1.0 / (double)B);
      JSD_DISPATCH_DTYPE(dtype, (launch_normalize_bwd<T>(job, 2, B, D, gdiag, t_dev, gamma_dev, inv_rows, st)));
    }
  }
  if (int rc = dense_bwd_common(false, Gmat, ldg, V, B, B, D, t_dev, gamma_dev, nullptr, acc_u, stream, nullptr, 0, &su))
    return rc;
  if (side != nullptr || su.ksplit > 1 || sv.ksplit > 1) {
    // image-side Jacobian (+ gamma * dL/dt) on the helper stream, next to the dV contraction.  The contraction
    // is enqueued FIRST so that its persistent CTAs take the SMs and the Jacobian's blocks fill in as they retire.
    cudaStream_t js = side ? side->stream : st;
    if (side) JSD_CUDA_OK(cudaEventRecord(side->fork, st));
    if (int rc = dense_bwd_common(true, Gmat, ldg, U, B, B, D, t_dev, gamma_dev, nullptr, acc_v, stream, nullptr, 0, &sv))
      return rc;
    if (side) JSD_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
    // the image side only stores its row dots; gamma * dL/dt = their sum is formed by block 0 of the text-side
    // launch, which is stream-ordered behind it: no ticket, no device-wide fence in either kernel
    if (int rc = normalize_bwd_impl(F, dtype, B, D, inv_f, acc_u, &su, V, 0, gdiag, t_dev, gamma_dev, B, dF, rowdot,
                                    nullptr, nullptr, js))
      return rc;
    if (side) {
      JSD_CUDA_OK(cudaEventRecord(side->join, side->stream));
      JSD_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
    }
    return normalize_bwd_impl(G, dtype, B, D, inv_g, acc_v, &sv, U, 0, gdiag, t_dev, gamma_dev, B, dG, nullptr, nullptr,
                              nullptr, stream, rowdot, B, dt_out);
  }
  if (int rc = dense_bwd_common(true, Gmat, ldg, U, B, B, D, t_dev, gamma_dev, nullptr, acc_v, stream, nullptr, 0, &sv))
    return rc;
  // both Jacobians (image side = job 0, text side = job 1) and gamma * dL/dt = sum_i <u_i, dU_i> in one launch
  jsd::NormBwdJob job{};
  job.X[0] = F;
  job.X[1] = G;
  job.inv_norm[0] = inv_f;
  job.inv_norm[1] = inv_g;
  job.acc[0] = acc_u;
  job.acc[1] = acc_v;
  job.partner[0] = (const __nv_bfloat16*)V;
  job.partner[1] = (const __nv_bfloat16*)U;
  job.dX[0] = dF;
  job.dX[1] = dG;
  job.rowdot = rowdot;
  job.ticket = dt_ticket(workspace);
  job.dt_out = dt_out;
  const float inv_rows = (float)(1.0 / (double)B);
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize_bwd<T>(job, 2, B, D, gdiag, t_dev, gamma_dev, inv_rows, st)));
}

/* ------------------------------------------------------------------ fused single-pass path (D <= 256) */
int jsd_dense_fused_supported(int64_t B, int64_t D) {
  return (D > 0 && D <= 256 && D % 64 == 0 && B >= 2 && (B + jsd::FB - 1) / jsd::FB * 2 <= 1024) ? 1 : 0;
}

// CTAs per 128-row block: a small batch is cut along the other modality so that the launch fills the GPU (every CTA
// writes its own accumulator slice, summed in order by the Jacobian kernel: at most 8)
int jsd_dense_fused_splits(int64_t B, int64_t D) {
  (void)D;
  if (B < 2) return 1;
  const int64_t blocks = (B + jsd::FB - 1) / jsd::FB;
  const int sms = sm_count_cached() > 0 ? sm_count_cached() : 148;
  int64_t ns = sms / (2 * blocks);
  if (ns > 8) ns = 8;
  if (ns > blocks) ns = blocks;            // at least one 128-row tile of the other modality per CTA
  return (int)(ns < 1 ? 1 : ns);
}

}  // extern "C"

template <int NKA>
static int launch_fused(const CUtensorMap& tmU, const CUtensorMap& tmV, const jsd::FusedParams& p, int grid,
                        cudaStream_t st) {
  auto kern = jsd::jsd_fused_kernel<NKA>;
  constexpr int smem = jsd::fused_smem_bytes(NKA);
  static unsigned long long configured_mask = 0;
  int dev = 0;
  JSD_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !((configured_mask >> dev) & 1ull)) {
    JSD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) configured_mask |= 1ull << dev;
  }
  kern<<<grid, jsd::F_THREADS, smem, st>>>(tmU, tmV, p);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" {

int jsd_dense_fused_fwd_bwd(const void* U, const void* V, int64_t B, int64_t D, const float* t_dev, float* acc_u,
                            float* acc_v, float* gdiag, void* workspace, float* out4, float* loss_out,
                            jsd_stream_t stream) {
  JSD_REQUIRE(U && V && t_dev && acc_u && acc_v && gdiag && workspace && out4,
              "jsd_dense_fused_fwd_bwd: null pointer argument");
  JSD_REQUIRE(jsd_dense_fused_supported(B, D), "jsd_dense_fused_fwd_bwd: needs D in {64, 128, 192, 256} and "
              "2 <= B <= 65536 (got B=%lld, D=%lld)", (long long)B, (long long)D);
  CUtensorMap tmU, tmV;
  if (int rc = make_tmap(&tmU, U, D, B, D, jsd::BLOCK_K, jsd::FB)) return rc;
  if (int rc = make_tmap(&tmV, V, D, B, D, jsd::BLOCK_K, jsd::FB)) return rc;
  jsd::FusedParams p{};
  p.D = (int)D;
  p.t_dev = t_dev;
  const int blocks = (int)((B + jsd::FB - 1) / jsd::FB);
  const int ns = jsd_dense_fused_splits(B, D);
  p.prob[0] = jsd::FusedProblem{(int)B, (int)B, 0, acc_u, gdiag, 1};     // image rows x all text rows
  p.prob[1] = jsd::FusedProblem{(int)B, (int)B, 0, acc_v, nullptr, 0};   // text rows x all image rows (recomputed)
  p.blocks0 = blocks * ns;
  p.nsplit = ns;
  p.acc_stride = B * D;
  p.ticket = (int*)workspace;
  p.partials = (float*)((char*)workspace + 16);
  p.out4 = out4;
  p.loss_out = loss_out;
  p.inv_pos = 1.0 / (double)B;
  p.inv_neg = 1.0 / ((double)B * (double)(B - 1));
  cudaStream_t st = (cudaStream_t)stream;
  switch (D / 64) {
    case 1: return launch_fused<1>(tmU, tmV, p, 2 * blocks * ns, st);
    case 2: return launch_fused<2>(tmU, tmV, p, 2 * blocks * ns, st);
    case 3: return launch_fused<3>(tmU, tmV, p, 2 * blocks * ns, st);
    default: return launch_fused<4>(tmU, tmV, p, 2 * blocks * ns, st);
  }
}

int jsd_dense_fused_forward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const float* t_dev, void* U,
                            void* V, float* inv_f, float* inv_g, float* acc_u, float* acc_v, float* gdiag,
                            void* workspace, float* out4, float* loss_out, jsd_stream_t stream) {
  if (int rc = jsd_normalize_cast_pair(F, G, dtype, B, D, U, V, inv_f, inv_g, stream)) return rc;
  return jsd_dense_fused_fwd_bwd(U, V, B, D, t_dev, acc_u, acc_v, gdiag, workspace, out4, loss_out, stream);
}

int jsd_dense_fused_backward(const void* F, const void* G, int dtype, int64_t B, int64_t D, const void* U,
                             const void* V, const float* inv_f, const float* inv_g, const float* gdiag,
                             const float* t_dev, const float* gamma_dev, const float* acc_u, const float* acc_v,
                             float* rowdot, void* workspace, void* dF, void* dG, float* dt_out, jsd_stream_t stream) {
  JSD_REQUIRE(F && G && U && V && inv_f && inv_g && gdiag && t_dev && acc_u && acc_v && rowdot && workspace && dF &&
              dG && dt_out, "jsd_dense_fused_backward: null pointer argument");
  JSD_REQUIRE(fits_int(B) && fits_int(D) && B >= 2, "jsd_dense_fused_backward: bad shape");
  // both Jacobians (image side = job 0, text side = job 1) and gamma * dL/dt = sum_i <u_i, dU_i> in ONE launch; the
  // accumulators are the fused kernel's unscaled sums: gamma * tau / (B (B - 1)) is applied here, in fp32
  jsd::NormBwdJob job{};
  job.X[0] = F;
  job.X[1] = G;
  job.inv_norm[0] = inv_f;
  job.inv_norm[1] = inv_g;
  job.acc[0] = acc_u;
  job.acc[1] = acc_v;
  job.partner[0] = (const __nv_bfloat16*)V;
  job.partner[1] = (const __nv_bfloat16*)U;
  job.dX[0] = dF;
  job.dX[1] = dG;
  job.rowdot = rowdot;
  job.ticket = dt_ticket(workspace);
  job.dt_out = dt_out;
  job.acc_scale[0] = job.acc_scale[1] = (float)(1.0 / ((double)B * (double)(B - 1)));
  const int ns = jsd_dense_fused_splits(B, D);
  if (ns > 1) {                           // the accumulators are sums of the column splits' slices, in order
    job.acc_slots = ns;
    for (int q = 0; q < ns; ++q) {
      job.slot[q] = acc_u + (size_t)q * B * D;
      job.slot1[q] = acc_v + (size_t)q * B * D;
    }
  }
  const float inv_rows = (float)(1.0 / (double)B);
  cudaStream_t st = (cudaStream_t)stream;
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize_bwd<T>(job, 2, B, D, gdiag, t_dev, gamma_dev, inv_rows, st)));
}

int jsd_dense_backward_image_side(const void* F, int dtype, int64_t M, int64_t N, int64_t D, int64_t row_offset,
                                  const void* V_all, const float* inv_f, const void* Gmat, int64_t ldg,
                                  const float* gdiag, const float* t_dev, const float* gamma_dev, float* acc_u,
                                  float* rowdot, void* workspace, void* sk_workspace, void* dF, float* dt_out,
                                  jsd_stream_t stream) {
  JSD_REQUIRE(acc_u && rowdot && dt_out && workspace, "jsd_dense_backward_image_side: null pointer argument");
  const SplitPlan su = plan_split(M, D, N, sk_workspace, 0);
  if (int rc = dense_bwd_common(false, Gmat, ldg, V_all, M, N, D, t_dev, gamma_dev, nullptr, acc_u, stream, nullptr, 0,
                                &su))
    return rc;
  return normalize_bwd_impl(F, dtype, M, D, inv_f, acc_u, &su, V_all, row_offset, gdiag, t_dev, gamma_dev, M, dF,
                            rowdot, workspace, dt_out, stream);
}

/* ------------------------------------------------------------------ peer-memory exchange */
static int check_peer_ctx(const jsd_peer_ctx* c, const char* who) {
  JSD_REQUIRE(c != nullptr, "%s: null context", who);
  JSD_REQUIRE(c->world >= 1 && c->world <= JSD_MAX_PEERS && c->rank >= 0 && c->rank < c->world,
              "%s: bad rank %d / world %d", who, c->rank, c->world);
  JSD_REQUIRE(fits_int(c->rows) && fits_int(c->dim) && fits_int(c->rows * c->world), "%s: bad shape", who);
  for (int q = 0; q < c->world; ++q)
    JSD_REQUIRE(c->v_all[0][q] && c->v_all[1][q] && c->stage[q] && c->flags[q], "%s: null buffer of rank %d", who, q);
  return 0;
}

size_t jsd_peer_flag_bytes(void) { return JSD_PEER_FLAG_INTS * sizeof(int32_t); }

int jsd_peer_alloc(size_t bytes, void** out) {
  JSD_REQUIRE(out && bytes > 0, "jsd_peer_alloc: bad argument");
  JSD_CUDA_OK(cudaMalloc(out, bytes));
  JSD_CUDA_OK(cudaMemset(*out, 0, bytes));
  JSD_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

int jsd_peer_free(void* ptr) {
  JSD_CUDA_OK(cudaFree(ptr));
  return 0;
}

int jsd_peer_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == JSD_PEER_HANDLE_BYTES, "IPC handle size");
  JSD_REQUIRE(ptr && handle64, "jsd_peer_export: null pointer argument");
  cudaIpcMemHandle_t h;
  JSD_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

int jsd_peer_open(const void* handle64, void** out) {
  JSD_REQUIRE(handle64 && out, "jsd_peer_open: null pointer argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  JSD_CUDA_OK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int jsd_peer_close(void* ptr) {
  JSD_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int jsd_peer_normalize_push(const void* F, const void* G, int dtype, const jsd_peer_ctx* ctx, int parity, void* U,
                            float* inv_f, float* inv_g, jsd_stream_t stream) {
  if (int rc = check_peer_ctx(ctx, "jsd_peer_normalize_push")) return rc;
  JSD_REQUIRE(F && G && U && inv_f && inv_g && (parity == 0 || parity == 1), "jsd_peer_normalize_push: bad argument");
  if (int rc = ensure_wait_cfg()) return rc;
  const int64_t rows = ctx->rows, D = ctx->dim;
  const int mode = peer_gather_mode(ctx->world);
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* mine = ctx->flags[ctx->rank];
  if (mode == GATHER_KERNEL) {   // (also at world 1 when forced: the single-GPU tests cover the bulk-store path)
    jsd::PeerPushJob job{};
    job.X[0] = F;
    job.X[1] = G;
    job.U = (__nv_bfloat16*)U;
    job.inv_norm[0] = inv_f;
    job.inv_norm[1] = inv_g;
    uintptr_t bits = reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(U);
    for (int k = 0; k < ctx->world; ++k) {
      const int q = (ctx->rank - k + ctx->world) % ctx->world;
      job.v_dst[k] = (__nv_bfloat16*)ctx->v_all[parity][q] + (size_t)ctx->rank * rows * D;
      job.flag_dst[k] = ctx->flags[q] + JSD_PEER_READY_V + parity * JSD_MAX_PEERS + ctx->rank;
      bits |= reinterpret_cast<uintptr_t>(job.v_dst[k]);
    }
    job.counter = mine + JSD_PEER_COUNTER_V + parity;
    job.ticket = mine + JSD_PEER_TICKET_PUSH;
    job.world = ctx->world;
    const size_t smem = (size_t)8 * D * sizeof(__nv_bfloat16);      // one bf16 row per warp
    const bool vec = (D % 8 == 0) && (bits & 15) == 0 && smem <= 160 * 1024;
    const dim3 grid((unsigned)((rows + 7) / 8), 2);
    switch (dtype) {
#define JSD_PUSH_CASE(code, T)                                                                        \
      case code:                                                                                      \
        if (vec) {                                                                                    \
          if (smem > 48 * 1024)                                                                       \
            JSD_CUDA_OK(cudaFuncSetAttribute(jsd::normalize_push_kernel<T, 8>,                        \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
          jsd::normalize_push_kernel<T, 8><<<grid, 256, smem, st>>>(job, (int)rows, (int)D);          \
        } else {                                                                                      \
          jsd::normalize_push_kernel<T, 1><<<grid, 256, 0, st>>>(job, (int)rows, (int)D);             \
        }                                                                                             \
        break;
      JSD_PUSH_CASE(JSD_F32, float)
      JSD_PUSH_CASE(JSD_BF16, __nv_bfloat16)
      JSD_PUSH_CASE(JSD_F16, __half)
#undef JSD_PUSH_CASE
      default: return fail("unsupported dtype code %d", dtype);
    }
    JSD_CUDA_OK(cudaGetLastError());
    return 0;
  }
  // local part only: F -> U, G -> this rank's row block of its OWN gathered V buffer, step counter of the buffer
  // + 1.  The copies into the other ranks' buffers are made underneath the forward launch that follows
  // (jsd_peer_dense_fwd): by one of its idle warps, or by the copy engines.
  jsd::NormalizeJob job{};
  job.X[0] = F;
  job.X[1] = G;
  job.Xn[0] = (__nv_bfloat16*)U;
  job.Xn[1] = (__nv_bfloat16*)ctx->v_all[parity][ctx->rank] + (size_t)ctx->rank * rows * D;
  job.inv_norm[0] = inv_f;
  job.inv_norm[1] = inv_g;
  job.bump = mine + JSD_PEER_COUNTER_V + parity;
  int rc = [&]() -> int { JSD_DISPATCH_DTYPE(dtype, (launch_normalize<T>(job, 2, rows, D, st))); }();
  if (rc) return rc;
  if (ctx->world > 1 && mode == GATHER_CE) {
    CopyStreams* cs = copy_streams();
    JSD_REQUIRE(cs != nullptr, "jsd_peer_normalize_push: could not create the copy streams");
    JSD_CUDA_OK(cudaEventRecord(cs->fork, st));
    for (int i = 0; i < kCopyStreams; ++i) JSD_CUDA_OK(cudaStreamWaitEvent(cs->stream[i], cs->fork, 0));
    const size_t own = (size_t)ctx->rank * rows * D, bytes = (size_t)rows * D * sizeof(__nv_bfloat16);
    const __nv_bfloat16* src = (const __nv_bfloat16*)ctx->v_all[parity][ctx->rank] + own;
    for (int k = 1; k < ctx->world; ++k) {
      // destination rank - k (mod world): a receiver q is served by q + 1, q + 2, ... in turn, the order in which
      // its forward walks the column blocks
      const int q = (ctx->rank - k + ctx->world) % ctx->world;
      cudaStream_t c = cs->stream[(k - 1) % kCopyStreams];
      JSD_CUDA_OK(cudaMemcpyAsync((__nv_bfloat16*)ctx->v_all[parity][q] + own, src, bytes, cudaMemcpyDeviceToDevice, c));
      JSD_CUDA_OK(cudaMemcpyAsync(ctx->flags[q] + JSD_PEER_READY_V + parity * JSD_MAX_PEERS + ctx->rank, job.bump,
                                  sizeof(int32_t), cudaMemcpyDeviceToDevice, c));
    }
    for (int i = 0; i < kCopyStreams; ++i) JSD_CUDA_OK(cudaEventRecord(cs->join[i], cs->stream[i]));
    cs->pending = true;
  }
  return 0;
}

int jsd_peer_dense_fwd(const void* U, const jsd_peer_ctx* ctx, int parity, const float* t_dev, void* Gmat,
                       int64_t ldg, float* gdiag, void* workspace, float* out4, float* loss_out,
                       jsd_stream_t stream) {
  if (int rc = check_peer_ctx(ctx, "jsd_peer_dense_fwd")) return rc;
  JSD_REQUIRE(parity == 0 || parity == 1, "jsd_peer_dense_fwd: parity must be 0 or 1");
  JSD_REQUIRE(ctx->dim % 8 == 0, "jsd_peer_dense_fwd: D must be a multiple of 8");
  if (int rc = ensure_wait_cfg()) return rc;
  int32_t* mine = ctx->flags[ctx->rank];
  PeerWait w;
  w.flags = mine + JSD_PEER_READY_V + parity * JSD_MAX_PEERS;
  w.counter = mine + JSD_PEER_COUNTER_V + parity;
  w.count = ctx->world;
  w.rows = peer_wait_per_source() ? (int)ctx->rows : 0;
  w.ctx = ctx;
  w.parity = parity;
  w.push_in_kernel = peer_gather_mode(ctx->world) == GATHER_FWD;
  if (int rc = dense_fwd_impl(U, ctx->v_all[parity][ctx->rank], ctx->rows, ctx->rows * ctx->world, ctx->dim,
                              ctx->rows * ctx->rank, t_dev, Gmat, ldg, gdiag, workspace, out4, loss_out, w, stream))
    return rc;
  if (CopyStreams* cs = copy_streams()) {
    if (cs->pending) {          // the copies of jsd_peer_normalize_push ran underneath the forward: join them here
      for (int i = 0; i < kCopyStreams; ++i) JSD_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, cs->join[i], 0));
      cs->pending = false;
    }
  }
  return 0;
}

// dV partial as bf16 tiles pushed by TMA stores from the contraction's epilogue into the OWNER's slots
// (stage[q] viewed as [world][rows][D] bf16: slot = source rank), owner blocks walked from rank + 1 on
static int peer_dv_push(const void* Gmat, int64_t ldg, const void* U, const jsd_peer_ctx* ctx, const float* t_dev,
                        const float* gamma_dev, jsd_stream_t stream, int worker_cap = 0) {
  const int64_t M = ctx->rows, N = ctx->rows * ctx->world, D = ctx->dim;
  JSD_REQUIRE(Gmat && U && t_dev, "jsd_peer_dense_bwd_dv: null pointer argument");
  JSD_REQUIRE(D % 8 == 0 && ldg >= N && ldg % 8 == 0, "jsd_peer_dense_bwd_dv: bad D / ldg");
  JSD_REQUIRE(M % 32 == 0, "jsd_peer_dense_bwd_dv: bf16 partials need rows per rank %% 32 == 0 (got %lld)", (long long)M);
  const int cg = pick_cta_group(N);
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap(&tmA, Gmat, N, M, ldg, 64, jsd::BLOCK_K)) return rc;       // Gmat^T read MN-major
  if (int rc = make_tmap(&tmB, U, D, M, D, 64, jsd::BLOCK_K)) return rc;            // U [M, D] read MN-major
  jsd::GemmParams p{};
  p.M = (int)N;                     // rows of the output: every text row
  p.N = (int)D;
  p.K = (int)M;
  p.n_fastest = tile_order(true, 1);
  p.t_dev = t_dev;
  p.gamma_dev = gamma_dev;
  p.scale = N > 1 ? (float)(1.0 / ((double)M * (double)(N - 1))) : 0.f;
  p.ksplit = 1;
  p.peer_world = ctx->world;
  int32_t* mine = ctx->flags[ctx->rank];
  for (int q = 0; q < ctx->world; ++q) {
    p.peer_flag_dst[q] = ctx->flags[q] + JSD_PEER_READY_DV + ctx->rank;
    void* slot = (__nv_bfloat16*)ctx->stage[q] + (size_t)ctx->rank * M * D;         // this rank's slot at owner q
    if (int rc = make_tmap(&p.tmPush[q], slot, D, M, D, jsd::COLS_PER_WARP, 32)) return rc;
  }
  p.peer_counter = mine + JSD_PEER_COUNTER_DV;
  p.peer_ticket = mine + JSD_PEER_TICKET_DV;
  p.push_rows = (int)M;
  // first m-block of rank + 1's rows (own rows last: they need no link)
  const int tile_m = jsd::BLOCK_M * cg;
  p.m_rot = (int)((((int64_t)(ctx->rank + 1) % ctx->world) * M) / tile_m);
  cudaStream_t st = (cudaStream_t)stream;
  return cg == 2 ? launch_gemm<jsd::MODE_GRADPUSH, true, true, 2>(tmA, tmB, p, nullptr, st, 0, worker_cap)
                 : launch_gemm<jsd::MODE_GRADPUSH, true, true, 1>(tmA, tmB, p, nullptr, st, 0, worker_cap);
}

static int peer_dv_impl(const void* Gmat, int64_t ldg, const void* U, const jsd_peer_ctx* ctx, const float* t_dev,
                        const float* gamma_dev, int partials_bf16, jsd_stream_t stream, int cap_sms) {
  if (partials_bf16) return peer_dv_push(Gmat, ldg, U, ctx, t_dev, gamma_dev, stream, cap_sms);
  // fp32: the partial is read by the peers as one buffer: whole tiles only
  return dense_bwd_common(true, Gmat, ldg, U, ctx->rows, ctx->rows * ctx->world, ctx->dim, t_dev, gamma_dev, nullptr,
                          nullptr, stream, ctx, 0, nullptr, cap_sms);
}

int jsd_peer_dense_bwd_dv(const void* Gmat, int64_t ldg, const void* U, const jsd_peer_ctx* ctx, const float* t_dev,
                          const float* gamma_dev, int partials_bf16, jsd_stream_t stream) {
  if (int rc = check_peer_ctx(ctx, "jsd_peer_dense_bwd_dv")) return rc;
  return peer_dv_impl(Gmat, ldg, U, ctx, t_dev, gamma_dev, partials_bf16, stream, 0);
}

// Small slabs (4 / 8 GPUs at B = 8192): each backward contraction is one or two rounds of tiles, so a launch is
// mostly fixed cost -- prologue, pipeline fill, the un-overlapped last epilogue, the drain of the pushed tiles
// (trace r02j: 28 + 24 us for 2 x 17 GFLOP).  Run side by side on HALF of the SMs each, the two launches pay those
// costs at the same time and each half stays busy for 3-4 rounds.  The image-side contraction is cut along K into
// as many slices as make its units as long as a dV tile (K_dU / K_dV = world), summed in order by the Jacobian.
struct PairedPlan {
  int cap_sms = 0;          // 0: not paired (full-width launches, dU split by plan_split)
  SplitPlan su;
};

static PairedPlan plan_paired(int64_t M, int64_t N, int64_t D, void* sk_workspace) {
  PairedPlan pp;
  static int enabled = -1;                                     // development knob: JSD_PAIRED=0 disables it
  if (enabled < 0) {
    const char* e = getenv("JSD_PAIRED");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int sms = sm_count_cached();
  if (!enabled || sk_workspace == nullptr || side_stream() == nullptr || sms < 8 || M <= jsd::BLOCK_M) return pp;
  const int64_t tiles_dv = ((N + 255) / 256) * ((D + jsd::BLOCK_N - 1) / jsd::BLOCK_N);
  if (tiles_dv > sms || M > 2048) return pp;                   // many rounds, or long tiles (K = M): fixed costs are
                                                               // small next to the tiles themselves, leave it
  const int64_t nk = (N + 127) / 128;                          // k-chunks of the dU contraction (2 k-atoms each)
  int64_t ks = N / M;                                          // = world
  if (ks > 8) ks = 8;
  while (ks > 1 && (nk / ks < 2 || (size_t)(ks - 1) * M * D * sizeof(float) > kSplitRegionBytes)) --ks;
  pp.cap_sms = (sms / 4) * 2;                                  // an even number of SMs: whole CTA pairs
  if (ks > 1) {
    pp.su.ksplit = (int)ks;
    pp.su.slice_base = reinterpret_cast<float*>(static_cast<char*>(sk_workspace) + streamk_flag_bytes());
    pp.su.slice_stride = M * D;
  }
  return pp;
}

int jsd_peer_normalize_bwd_text(const void* G, int dtype, const jsd_peer_ctx* ctx, const float* inv_g, const void* U,
                                const float* gdiag, const float* t_dev, const float* gamma_dev, int partials_bf16,
                                void* dG, jsd_stream_t stream) {
  if (int rc = check_peer_ctx(ctx, "jsd_peer_normalize_bwd_text")) return rc;
  JSD_REQUIRE(G && inv_g && U && gdiag && t_dev && dG, "jsd_peer_normalize_bwd_text: null pointer argument");
  if (int rc = ensure_wait_cfg()) return rc;
  const int32_t* mine = ctx->flags[ctx->rank];
  jsd::NormBwdJob job{};
  job.X[0] = G;
  job.inv_norm[0] = inv_g;
  job.partner[0] = (const __nv_bfloat16*)U;
  job.partner_offset[0] = 0;
  job.dX[0] = dG;
  job.acc_slots = ctx->world;
  job.slot_bf16 = partials_bf16 ? 1 : 0;
  for (int q = 0; q < ctx->world; ++q) {
    if (partials_bf16)      // rank q's partial for this rank's rows was pushed into slot q of the LOCAL buffer
      job.slot[q] = (const float*)((const __nv_bfloat16*)ctx->stage[ctx->rank] + (size_t)q * ctx->rows * ctx->dim);
    else                    // rank q's partial over all rows stays in rank q's memory: read this rank's rows of it
      job.slot[q] = (const float*)ctx->stage[q] + (size_t)ctx->rank * ctx->rows * ctx->dim;
  }
  job.wait_flags = mine + JSD_PEER_READY_DV;
  job.wait_counter = mine + JSD_PEER_COUNTER_DV;
  job.wait_count = ctx->world;
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_rows = (float)(1.0 / (double)ctx->rows);
  JSD_DISPATCH_DTYPE(dtype, (launch_normalize_bwd<T>(job, 1, ctx->rows, ctx->dim, gdiag, t_dev, gamma_dev, inv_rows,
                                                     st)));
}

int jsd_peer_dense_backward(const void* F, const void* G, int dtype, const jsd_peer_ctx* ctx, int parity,
                            const void* U, const float* inv_f, const float* inv_g, const void* Gmat, int64_t ldg,
                            const float* gdiag, const float* t_dev, const float* gamma_dev, int partials_bf16,
                            float* acc_u, float* rowdot, void* workspace, void* sk_workspace, void* dF, void* dG,
                            float* dt_out, jsd_stream_t stream) {
  if (int rc = check_peer_ctx(ctx, "jsd_peer_dense_backward")) return rc;
  JSD_REQUIRE(parity == 0 || parity == 1, "jsd_peer_dense_backward: parity must be 0 or 1");
  const void* V_all = ctx->v_all[parity][ctx->rank];
  const int64_t M = ctx->rows, N = ctx->rows * ctx->world, D = ctx->dim, off = ctx->rows * ctx->rank;
  cudaStream_t st = (cudaStream_t)stream;
  // Two independent contractions side by side, forked right behind the forward, then -- once BOTH have finished --
  // the two Jacobians side by side:
  //   caller's stream : dV contraction (+ push / publish of the partial) | text-side Jacobian
  //   helper stream   : dU contraction (split-K when underfilled)        | image-side Jacobian
  // Both contractions are persistent kernels with static work lists, so any number of their CTAs may be resident:
  // whichever gets the SMs first, the other's CTA pairs move in as they retire (no ragged last wave is wasted),
  // and the partials travel while the rest computes.  The Jacobians are held back until both contractions are
  // done: a row kernel started next to a persistent contraction has all of its blocks made resident on the few
  // SMs that contraction leaves free and crawls there (traces r02d / r02g: 35-48 us instead of 6).
  SideStream* side = side_stream();
  cudaStream_t is = side ? side->stream : st;
  if (side) {
    JSD_CUDA_OK(cudaEventRecord(side->fork, st));
    JSD_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  const PairedPlan pp = plan_paired(M, N, D, sk_workspace);
  if (int rc = peer_dv_impl(Gmat, ldg, U, ctx, t_dev, gamma_dev, partials_bf16, stream, pp.cap_sms)) return rc;
  const SplitPlan su = pp.cap_sms ? pp.su : plan_split(M, D, N, sk_workspace, 0);
  if (int rc = dense_bwd_common(false, Gmat, ldg, V_all, M, N, D, t_dev, gamma_dev, nullptr, acc_u, (jsd_stream_t)is,
                                nullptr, 0, &su, pp.cap_sms))
    return rc;
  if (side) {
    JSD_CUDA_OK(cudaEventRecord(side->mid, side->stream));      // dU done
    JSD_CUDA_OK(cudaEventRecord(side->mid2, st));               // dV done
    JSD_CUDA_OK(cudaStreamWaitEvent(st, side->mid, 0));
    JSD_CUDA_OK(cudaStreamWaitEvent(side->stream, side->mid2, 0));
  }
  if (int rc = normalize_bwd_impl(F, dtype, M, D, inv_f, acc_u, &su, V_all, off, gdiag, t_dev, gamma_dev, M, dF,
                                  rowdot, workspace, dt_out, (jsd_stream_t)is))
    return rc;
  if (side) JSD_CUDA_OK(cudaEventRecord(side->join, side->stream));
  if (int rc = jsd_peer_normalize_bwd_text(G, dtype, ctx, inv_g, U, gdiag, t_dev, gamma_dev, partials_bf16, dG, stream))
    return rc;
  if (side) JSD_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
  return 0;
}

/* ------------------------------------------------------------------ peer waits: time limit and error report */
int jsd_peer_set_timeout(double seconds) {
  JSD_REQUIRE(seconds > 0.0 && seconds < 1e7, "jsd_peer_set_timeout: seconds out of range");
  g_wait_timeout_s = seconds;
  g_wait_cfg_mask = 0;                // re-install on every device at its next peer call
  return 0;
}

int jsd_peer_wait_error(int* kind, int* index, int* target) {
  if (g_wait_err_host == nullptr || g_wait_err_host[0] == 0) return 0;
  if (kind) *kind = g_wait_err_host[0];
  if (index) *index = g_wait_err_host[1];
  if (target) *target = g_wait_err_host[2];
  return 1;
}

/* ------------------------------------------------------------------ device-side event trace */
int jsd_trace_enable(unsigned long long* events, int capacity, int* count) {
#if !JSD_TRACE
  (void)events;
  (void)capacity;
  (void)count;
  return fail("this libjsd_b200.so was built without -DJSD_TRACE=1 (build.build_library(trace=True))");
#else
  JSD_REQUIRE((events == nullptr) || (count != nullptr && capacity > 0), "jsd_trace_enable: bad argument");
  jsd::TraceBuf tb;
  tb.events = events;
  tb.count = events ? count : nullptr;
  tb.capacity = events ? capacity : 0;
  JSD_CUDA_OK(cudaMemcpyToSymbol(jsd::g_trace, &tb, sizeof(tb)));
  return 0;
#endif
}

/* ------------------------------------------------------------------ retrieval / zero-shot scoring */
int jsd_split_bf16x3(const void* X, int dtype, int64_t rows, int64_t D, int side, int normalize, void* out,
                     jsd_stream_t stream) {
  JSD_REQUIRE(X && out && (side == 0 || side == 1), "jsd_split_bf16x3: bad argument");
  JSD_REQUIRE(fits_int(rows) && fits_int(3 * D), "jsd_split_bf16x3: bad shape");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case JSD_F32:
      jsd::split_bf16x3_kernel<float><<<grid, 256, 0, st>>>((const float*)X, (int)rows, (int)D, side, normalize,
                                                           (__nv_bfloat16*)out);
      break;
    case JSD_BF16:
      jsd::split_bf16x3_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)X, (int)rows, (int)D, side,
                                                                   normalize, (__nv_bfloat16*)out);
      break;
    case JSD_F16:
      jsd::split_bf16x3_kernel<__half><<<grid, 256, 0, st>>>((const __half*)X, (int)rows, (int)D, side, normalize,
                                                            (__nv_bfloat16*)out);
      break;
    default: return fail("unsupported dtype code %d", dtype);
  }
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

static int score_launch(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const jsd::GemmParams& proto,
                        cudaStream_t st) {
  const int cg = pick_cta_group(M);
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap(&tmA, A, K, M, K, jsd::BLOCK_K, jsd::BLOCK_M)) return rc;
  if (int rc = make_tmap(&tmB, B, K, N, K, jsd::BLOCK_K, jsd::b_rows_per_cta(cg))) return rc;
  jsd::GemmParams p = proto;
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.n_fastest = 0;
  return cg == 2 ? launch_gemm<jsd::MODE_SCORE, false, false, 2>(tmA, tmB, p, nullptr, st)
                 : launch_gemm<jsd::MODE_SCORE, false, false, 1>(tmA, tmB, p, nullptr, st);
}

int jsd_score_ranks(const void* A, const void* B, int64_t M, int64_t N, int64_t K, const int32_t* row_tgt_ptr,
                    const int32_t* row_tgt_idx, const int32_t* col_tgt, void* thr_row_scratch, float* thr_col_scratch,
                    int32_t* rank_row, int32_t* rank_col, jsd_stream_t stream) {
  JSD_REQUIRE(A && B, "jsd_score_ranks: null pointer argument");
  JSD_REQUIRE(fits_int(M) && fits_int(N) && fits_int(K) && K % 8 == 0, "jsd_score_ranks: bad shape (K %% 8 == 0)");
  const bool rows = row_tgt_ptr != nullptr, cols = col_tgt != nullptr;
  JSD_REQUIRE(rows || cols, "jsd_score_ranks: neither row nor column targets given");
  JSD_REQUIRE(!rows || (row_tgt_idx && thr_row_scratch && rank_row), "jsd_score_ranks: row side incomplete");
  JSD_REQUIRE(!cols || (thr_col_scratch && rank_col), "jsd_score_ranks: column side incomplete");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows) {
    JSD_CUDA_OK(cudaMemsetAsync(thr_row_scratch, 0, (size_t)M * 4, st));        // encoding 0 = no target
    JSD_CUDA_OK(cudaMemsetAsync(rank_row, 0, (size_t)M * 4, st));
  }
  if (cols) {
    JSD_CUDA_OK(cudaMemsetAsync(thr_col_scratch, 0x7f, (size_t)N * 4, st));     // 3.4e38: nothing beats "no target"
    JSD_CUDA_OK(cudaMemsetAsync(rank_col, 0, (size_t)N * 4, st));
  }
  jsd::GemmParams p{};
  p.row_tgt_ptr = rows ? row_tgt_ptr : nullptr;
  p.row_tgt_idx = row_tgt_idx;
  p.col_tgt = col_tgt;
  p.thr_row_enc = rows ? (unsigned*)thr_row_scratch : nullptr;
  p.thr_col = cols ? thr_col_scratch : nullptr;
  p.cnt_row = rank_row;
  p.cnt_col = rank_col;
  p.score_pass = 0;                 // pass 0: the target scores, taken from the very tiles pass 1 will recompute
  if (int rc = score_launch(A, B, M, N, K, p, st)) return rc;
  p.score_pass = 1;                 // pass 1: count the entries that beat them
  return score_launch(A, B, M, N, K, p, st);
}

int jsd_score_argmax(const void* A, const void* B, int64_t M, int64_t N, int64_t K, unsigned long long* best,
                     jsd_stream_t stream) {
  JSD_REQUIRE(A && B && best, "jsd_score_argmax: null pointer argument");
  JSD_REQUIRE(fits_int(M) && fits_int(N) && fits_int(K) && K % 8 == 0, "jsd_score_argmax: bad shape (K %% 8 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  JSD_CUDA_OK(cudaMemsetAsync(best, 0, (size_t)M * 8, st));
  jsd::GemmParams p{};
  p.best = best;
  p.score_pass = 1;
  return score_launch(A, B, M, N, K, p, st);
}

int jsd_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, int64_t M,
                  int64_t N, int64_t K, void* sk_workspace, float* C, jsd_stream_t stream) {
  JSD_REQUIRE(A && B && C, "jsd_gemm_bf16: null pointer argument");
  JSD_REQUIRE(fits_int(M) && fits_int(N) && fits_int(K), "jsd_gemm_bf16: bad shape");
  JSD_REQUIRE(N % 4 == 0, "jsd_gemm_bf16: N must be a multiple of 4");
  const int cg = pick_cta_group(M);
  CUtensorMap tmA, tmB;
  if (a_mn_major) {
    if (int rc = make_tmap(&tmA, A, M, K, lda, 64, jsd::BLOCK_K)) return rc;     // A^T stored [K, lda]
  } else {
    if (int rc = make_tmap(&tmA, A, K, M, lda, jsd::BLOCK_K, jsd::BLOCK_M)) return rc;
  }
  if (b_mn_major) {
    if (int rc = make_tmap(&tmB, B, N, K, ldb, 64, jsd::BLOCK_K)) return rc;     // B^T stored [K, ldb]
  } else {
    if (int rc = make_tmap(&tmB, B, K, N, ldb, jsd::BLOCK_K, jsd::b_rows_per_cta(cg))) return rc;
  }
  jsd::GemmParams p{};
  p.M = (int)M;
  p.N = (int)N;
  p.K = (int)K;
  p.n_fastest = 1;
  p.t_dev = nullptr;
  p.gamma_dev = nullptr;
  p.scale = 1.f;
  p.out = C;
  p.ldo = N;
  return launch_gemm_any<jsd::MODE_GRAD>(a_mn_major != 0, b_mn_major != 0, cg, tmA, tmB, p, sk_workspace,
                                          (cudaStream_t)stream);
}

}  // extern "C"

/* ------------------------------------------------------------------ projection-head tail (LayerNorm + normalise) */
namespace {

// blocks of the fused LayerNorm / normalise backward: every block walks rows blockIdx.x, + gridDim.x, ... and keeps
// the LayerNorm weight / bias sums of its rows in registers.  One wave of resident blocks (occupancy x SMs, at most
// 8 per SM -- the bound the workspace is sized for), never more blocks than rows.
constexpr int kLnBwdMaxBlocksPerSm = 8;
int ln_bwd_block_cap(int64_t rows) {
  const int64_t cap = kLnBwdMaxBlocksPerSm * (int64_t)(sm_count_cached() > 0 ? sm_count_cached() : 148);
  return (int)(rows < cap ? rows : cap);
}
template <typename Kernel>
int ln_bwd_blocks(Kernel kernel, int threads, int64_t rows) {
  // resident blocks per SM, asked once per (kernel, block size): the eager warm-up step every CUDA-graph capture
  // needs fills the table, so a capture makes no runtime query
  struct Entry { const void* fn; int threads, per_sm; };
  static Entry table[32];
  static int used = 0;
  static std::mutex mu;
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < used; ++i)
      if (table[i].fn == (const void*)kernel && table[i].threads == threads) per_sm = table[i].per_sm;
    if (per_sm == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
      }
      if (used < 32) table[used++] = Entry{(const void*)kernel, threads, per_sm};
    }
  }
  if (per_sm > kLnBwdMaxBlocksPerSm) per_sm = kLnBwdMaxBlocksPerSm;
  const int64_t wave = (int64_t)per_sm * (sm_count_cached() > 0 ? sm_count_cached() : 148);
  return (int)(rows < wave ? rows : wave);
}

template <typename T>
int launch_ln_normalize(const jsd::LnNormJob& job, int count, int64_t rows, int64_t D, bool out_bf16, cudaStream_t st) {
  uintptr_t bits = 0;
  for (int i = 0; i < count; ++i)
    bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.out[i]) |
            reinterpret_cast<uintptr_t>(job.w[i]) | reinterpret_cast<uintptr_t>(job.b[i]);
  const int variant = jsd::ln_fwd_variant(D, (bits & 15) == 0);
  const dim3 grid((unsigned)((rows + 7) / 8), (unsigned)count);
  auto go = [&](auto kernel, int arg) { kernel<<<grid, 256, 0, st>>>(job, (int)rows, arg); };
  if (out_bf16) jsd::ln_fwd_select<T, __nv_bfloat16>(variant, D, go);
  else jsd::ln_fwd_select<T, float>(variant, D, go);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_ln_normalize_bwd(const jsd::LnNormBwdJob& job, int count, int64_t rows, int64_t D, const float* gdiag,
                            const float* t_dev, const float* gamma_dev, float inv_rows, int* blocks_out,
                            cudaStream_t st) {
  uintptr_t bits = (uintptr_t)job.slice_stride * 4u | reinterpret_cast<uintptr_t>(job.col_partials);
  for (int i = 0; i < count; ++i)
    bits |= reinterpret_cast<uintptr_t>(job.X[i]) | reinterpret_cast<uintptr_t>(job.dX[i]) |
            reinterpret_cast<uintptr_t>(job.acc[i]) | reinterpret_cast<uintptr_t>(job.partner[i]);
  const jsd::LnBwdPlan plan = jsd::ln_bwd_plan(D, (bits & 15) == 0);
  JSD_REQUIRE(plan.kch > 0, "LayerNorm/normalise backward: D=%lld too large (max %d%s)", (long long)D,
              jsd::LN_BWD_MAX_KCH * jsd::LN_BWD_THREADS * plan.vec, plan.vec == 4 ? "" : " for unaligned rows");
  jsd::ln_bwd_select<T>(plan, [&](auto kernel) {
    const int blocks = ln_bwd_blocks(kernel, plan.threads, rows);
    *blocks_out = blocks;
    kernel<<<dim3((unsigned)blocks, (unsigned)count), plan.threads, 0, st>>>(job, (int)rows, (int)D, gdiag, t_dev,
                                                                              gamma_dev, inv_rows);
  });
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

size_t jsd_ln_workspace_bytes(int64_t rows, int64_t D) {
  if (rows <= 0 || D <= 0) return 0;
  return (size_t)2 * (size_t)ln_bwd_block_cap(rows) * 2 * (size_t)D * sizeof(float);
}

int jsd_ln_normalize_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                          const float* b0, float eps0, const float* w1, const float* b1, float eps1, int out_bf16,
                          void* out0, void* out1, float* stats0, float* stats1, jsd_stream_t stream) {
  JSD_REQUIRE(X0 && out0 && stats0, "jsd_ln_normalize_pair: null pointer argument");
  JSD_REQUIRE((X1 == nullptr) == (out1 == nullptr) && (X1 == nullptr) == (stats1 == nullptr),
              "jsd_ln_normalize_pair: X1, out1 and stats1 must be given together");
  JSD_REQUIRE(fits_int(rows) && fits_int(D), "jsd_ln_normalize_pair: rows=%lld, D=%lld out of range", (long long)rows,
              (long long)D);
  JSD_REQUIRE(eps0 >= 0.f && eps1 >= 0.f, "jsd_ln_normalize_pair: negative eps");
  jsd::LnNormJob job{};
  job.X[0] = X0; job.w[0] = w0; job.b[0] = b0; job.out[0] = out0; job.eps[0] = eps0;
  job.mean[0] = stats0; job.rstd[0] = stats0 + rows; job.inv_norm[0] = stats0 + 2 * rows;
  const int count = X1 ? 2 : 1;
  if (X1) {
    job.X[1] = X1; job.w[1] = w1; job.b[1] = b1; job.out[1] = out1; job.eps[1] = eps1;
    job.mean[1] = stats1; job.rstd[1] = stats1 + rows; job.inv_norm[1] = stats1 + 2 * rows;
  }
  cudaStream_t st = (cudaStream_t)stream;
  JSD_DISPATCH_DTYPE(dtype, (launch_ln_normalize<T>(job, count, rows, D, out_bf16 != 0, st)));
}

int jsd_ln_normalize_bwd_pair(const void* X0, const void* X1, int dtype, int64_t rows, int64_t D, const float* w0,
                              const float* b0, const float* w1, const float* b1, const float* stats0,
                              const float* stats1, const float* acc0, const float* acc1, int64_t n_slices,
                              int64_t slice_stride, float acc_scale, const void* partner0_bf16,
                              int64_t partner_offset0, const void* partner1_bf16, int64_t partner_offset1,
                              const float* gdiag, const float* t_dev, const float* gamma_dev, int64_t M_rows,
                              void* workspace, void* dX0, void* dX1, float* dw0, float* db0, float* dw1, float* db1,
                              float* rowdot, float* dt_out, jsd_stream_t stream) {
  JSD_REQUIRE(X0 && stats0 && acc0 && dX0 && workspace, "jsd_ln_normalize_bwd_pair: null pointer argument");
  const bool two = X1 != nullptr;
  JSD_REQUIRE(!two || (stats1 && acc1 && dX1), "jsd_ln_normalize_bwd_pair: second row set is incomplete");
  JSD_REQUIRE(fits_int(rows) && fits_int(D) && M_rows > 0, "jsd_ln_normalize_bwd_pair: bad shape");
  JSD_REQUIRE(n_slices >= 1 && n_slices <= 64 && (n_slices == 1 || slice_stride >= rows * D),
              "jsd_ln_normalize_bwd_pair: bad accumulator slices (%lld, stride %lld)", (long long)n_slices,
              (long long)slice_stride);
  JSD_REQUIRE(acc_scale >= 0.f, "jsd_ln_normalize_bwd_pair: negative acc_scale");
  JSD_REQUIRE(gdiag == nullptr || (partner0_bf16 && (!two || partner1_bf16) && t_dev),
              "jsd_ln_normalize_bwd_pair: the positive-pair term needs partner rows and the temperature");
  JSD_REQUIRE(acc_scale == 0.f || t_dev, "jsd_ln_normalize_bwd_pair: acc_scale needs the temperature");
  JSD_REQUIRE(dt_out == nullptr || rowdot != nullptr, "jsd_ln_normalize_bwd_pair: dt_out needs rowdot");
  jsd::LnNormBwdJob job{};
  job.X[0] = X0; job.w[0] = w0; job.b[0] = b0;
  job.mean[0] = stats0; job.rstd[0] = stats0 + rows; job.inv_norm[0] = stats0 + 2 * rows;
  job.acc[0] = acc0; job.partner[0] = (const __nv_bfloat16*)partner0_bf16; job.partner_offset[0] = partner_offset0;
  job.dX[0] = dX0;
  if (two) {
    job.X[1] = X1; job.w[1] = w1; job.b[1] = b1;
    job.mean[1] = stats1; job.rstd[1] = stats1 + rows; job.inv_norm[1] = stats1 + 2 * rows;
    job.acc[1] = acc1; job.partner[1] = (const __nv_bfloat16*)partner1_bf16; job.partner_offset[1] = partner_offset1;
    job.dX[1] = dX1;
  }
  job.slice_stride = n_slices > 1 ? slice_stride : 0;
  job.n_slices = (int)n_slices;
  job.acc_scale = acc_scale;
  job.col_partials = (float*)workspace;
  job.rowdot = rowdot;
  const int count = two ? 2 : 1;
  const float inv_rows = (float)(1.0 / (double)M_rows);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 1, blocks = 0;
  switch (dtype) {
    case JSD_F32: rc = launch_ln_normalize_bwd<float>(job, count, rows, D, gdiag, t_dev, gamma_dev, inv_rows, &blocks, st); break;
    case JSD_BF16: rc = launch_ln_normalize_bwd<__nv_bfloat16>(job, count, rows, D, gdiag, t_dev, gamma_dev, inv_rows, &blocks, st); break;
    case JSD_F16: rc = launch_ln_normalize_bwd<__half>(job, count, rows, D, gdiag, t_dev, gamma_dev, inv_rows, &blocks, st); break;
    default: return fail("unsupported dtype code %d", dtype);
  }
  if (rc) return rc;
  jsd::LnFinalizeJob fin{};
  fin.col_partials = (const float*)workspace;
  fin.nblocks = blocks;
  fin.dw[0] = dw0; fin.db[0] = db0;
  fin.dw[1] = two ? dw1 : nullptr; fin.db[1] = two ? db1 : nullptr;
  fin.rowdot = rowdot;
  fin.rows = (int)rows;
  fin.dt_out = dt_out;
  const dim3 fgrid((unsigned)((D + 255) / 256), (unsigned)(2 * count));
  jsd::ln_bwd_finalize_kernel<<<fgrid, 256, 0, st>>>(fin, (int)D);
  JSD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
