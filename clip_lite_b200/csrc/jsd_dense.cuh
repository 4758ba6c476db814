// Dense (all-pairs) JSD estimator on sm_100a tensor cores.
//
// One warp-specialised, persistent tcgen05 GEMM kernel, instantiated three ways:
//
//   FWD      S = U . V^T                (A = U [M,D] K-major,      B = V [N,D] K-major)
//            epilogue: x = tau*S -> softplus / sigmoid -> per-warp loss partials,
//            Gmat[i,j] = sigma(x_ij) (bf16, 0 on the positives), gdiag[i] = -sigma(-x_ii')
//   GRAD_DU  dUacc = scale * Gmat   . V (A = Gmat [M,N] K-major,   B = V [N,D] read MN-major)
//   GRAD_DV  dVacc = scale * Gmat^T . U (A = Gmat read MN-major,   B = U [M,D] read MN-major)
//
// The B x B score matrix itself is never written; what crosses the fwd/bwd boundary is
// the bf16 sigmoid-coefficient matrix Gmat (DESIGN.md explains why the dU/dV accumulators
// of a D=1024 problem cannot live in TMEM next to the S tiles).  No operand is ever
// transposed in memory: the MN-major UMMA descriptors read U, V and Gmat as they lie.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4.. = epilogue
// (four TMEM lane quarters x 2 or 4 column groups; 16 epilogue warps by default: the softplus/sigmoid
// epilogue is latency-bound with only two warps per scheduler).
// Pipelines: smem ring (full/empty mbarriers), 2-stage TMEM accumulator ring
// (tfull/tempty) so the epilogue of one tile overlaps the MMAs of the next.
//
// CTA pairs (template CG = 2): the kernel is fed from L2, and 128x256 tiles pull 48 KB per
// 128x256x64 MACs through it.  With tcgen05 cta_group::2 two CTAs of a cluster share one
// 256x256 tile: each loads its own 128 rows of A and HALF of B (32 KB per CTA and k-chunk,
// 6 stages), the leader CTA issues M=256 MMAs that read both CTAs' shared memory and write
// both CTAs' TMEM, and tcgen05.commit multicasts the barrier arrivals to both.
//
// Scheduling: FWD walks whole tiles round-robin.  The GRAD GEMMs (1.73 waves at B=8192,
// D=1024) use stream-K per group of workers (see SegmentIter): a worker that ends up with
// the tail of a tile publishes its fp32 partial accumulator to a workspace slot, the worker
// holding the head of that tile adds the partials in its epilogue.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace jsd {

constexpr int BLOCK_M = 128;    // rows per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;    // accumulator columns per tile
constexpr int BLOCK_K = 64;     // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int MN_ATOM_BYTES = 64 * BLOCK_K * 2;       // one 64-wide MN-major atom: 64 k-rows x 128 B
constexpr int NUM_CTRL_WARPS = 4;
constexpr int NUM_EPI_WARPS = 16;                             // 4 TMEM lane quarters x 4 column groups
constexpr int EPI_COL_GROUPS = NUM_EPI_WARPS / 4;
constexpr int COLS_PER_WARP = 256 / EPI_COL_GROUPS;           // 64 accumulator columns per epilogue warp and tile
constexpr int CW = COLS_PER_WARP / 4;                         // chunk width 16: 4 register-resident chunks per warp and tile
constexpr int G_STAGE_BYTES = 32 * COLS_PER_WARP * 2;         // per-warp Gmat staging box: 32 rows x 64 bf16 = 4 KB (128 B rows)
static_assert(COLS_PER_WARP * 2 == 128, "the Gmat staging box is one 128-byte swizzle row wide");
constexpr int GEMM_THREADS = 32 * (NUM_CTRL_WARPS + NUM_EPI_WARPS);
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;       // 512
constexpr int PARTIALS_PER_WARP = 4;                  // sum sp(-x_pos), sum max(s,0), sum log2(1+e), spare
constexpr int SK_SLOT_FLOATS = BLOCK_M * BLOCK_N;     // one fp32 partial accumulator tile (per CTA)
constexpr int SK_MAX_CTAS = 256;                      // flags / slots reserved in the workspace
constexpr int MAX_PEERS = 8;                          // GPUs of one NVSwitch domain

// per-CTA shared-memory budget as a function of the CTA-group size (1 = single CTA, 2 = CTA pair)
__host__ __device__ constexpr int b_rows_per_cta(int cg) { return BLOCK_N / cg; }
__host__ __device__ constexpr int stage_bytes(int cg) { return A_TILE_BYTES + b_rows_per_cta(cg) * BLOCK_K * 2; }
// the forward kernel gives 64 KB to the Gmat staging boxes of its 16 epilogue warps (TMA stores); so does the
// GRADPUSH variant of the backward contraction, whose bf16 output tiles leave by TMA stores into PEER memory
__host__ __device__ constexpr int g_staging_bytes(int mode) {
  return (mode == 0 || mode == 3) ? NUM_EPI_WARPS * G_STAGE_BYTES : 0;
}
// k-atoms (64-wide, one 128B swizzle row each) staged per pipeline slot, and slots per ring
#ifndef JSD_GRAD_KATOMS
#define JSD_GRAD_KATOMS 2   // two k-atoms per slot: half as many barrier round trips per MMA (measured +3.6 %)
#endif
#ifndef JSD_GRAD_STAGES
#define JSD_GRAD_STAGES 3
#endif
__host__ __device__ constexpr int k_atoms(int mode) { return (mode == 0 || mode == 3) ? 1 : JSD_GRAD_KATOMS; }
__host__ __device__ constexpr int num_stages(int cg, int mode) {
  return (mode == 0 || mode == 3) ? (cg == 2 ? 5 : 3) : (cg == 2 ? JSD_GRAD_STAGES : 4 / JSD_GRAD_KATOMS);
}
__host__ __device__ constexpr int gemm_smem_bytes(int cg, int mode) {
  return num_stages(cg, mode) * k_atoms(mode) * stage_bytes(cg) + g_staging_bytes(mode) + 1024 /* align slack */ +
         256 /* barriers */;
}

// SCORE shares GRAD's pipeline shape; GRADPUSH is GRAD with the forward's shared-memory budget (one k-atom per
// slot + 64 KB of store staging): out = bf16(gscale * acc), pushed tile by tile into the owner rank's memory
enum GemmMode { MODE_FWD = 0, MODE_GRAD = 1, MODE_SCORE = 2, MODE_GRADPUSH = 3 };
__host__ __device__ constexpr bool is_grad_mode(int mode) { return mode == MODE_GRAD || mode == MODE_GRADPUSH; }

struct GemmParams {
  alignas(64) CUtensorMap tmG;   // FWD: Gmat [M, N] bf16 as the target of the epilogue's TMA stores (box 64 x 32)
  int M;            // rows of the output tile space (FWD: image rows; GRAD: rows of the gradient)
  int N;            // cols of the output tile space (FWD: text rows;  GRAD: D)
  int K;            // contraction length
  int n_fastest;    // tile order: 1 = consecutive tiles walk N first (A tile shared through L2)
  int m_rot, n_rot; // the walk starts at this m-block / n-block (rotation): a peer forward starts on its OWN column
                    // block (no exchange needed), a pushing dV contraction on the rows of the NEXT rank
  // FWD
  int row_offset;   // column of the positive of local row 0
  const float* t_dev;
  __nv_bfloat16* gmat;   // may be null (loss only)
  long long ldg;
  float* gdiag;          // [M]
  float* partials;       // [grid * NUM_EPI_WARPS * PARTIALS_PER_WARP]
  int* ticket;           // zero before the launch; the last CTA to finish reduces the partials and re-zeroes it
  float* out4;           // {pos, neg, pos + neg, 0}
  float* loss_out;       // optional separate copy of the loss
  double inv_pos, inv_neg;   // 1 / M,  1 / (M (N - 1))
  // GRAD
  const float* gamma_dev;  // may be null (gamma = 1)
  float scale;             // 1 / (M_rows * (N_cols - 1))
  float* out;              // [M, N] fp32, pitch ldo
  long long ldo;
  // stream-K (GRAD): sk_flags [SK_MAX_CTAS * NUM_EPI_WARPS] ints, zero between launches;
  // sk_slots [grid * SK_SLOT_FLOATS] floats.  stream_k = 0 -> whole tiles, round-robin.
  int stream_k;
  int* sk_flags;
  float* sk_slots;
  // split-K (GRAD, whole tiles): each output tile is computed `ksplit` times over consecutive K ranges by
  // different workers; slice 0 goes to `out`, slice s > 0 to slice_base + (s - 1) * slice_stride (same pitch).
  // The consumer (the Jacobian kernel) adds the slices in order: deterministic, no atomics, no hand-off flags.
  int ksplit;
  float* slice_base;
  long long slice_stride;
  // SCORE (retrieval / zero-shot scoring, S = A B^T never stored).  Pass 0 extracts the target scores, pass 1
  // counts the entries that beat them -- the same tiles, the same accumulation order, so a target never beats
  // itself and no tie margin is needed.
  int score_pass;
  const int* row_tgt_ptr;        // CSR of each row's target columns (may be null)
  const int* row_tgt_idx;
  const int* col_tgt;            // [N] target row of each column, -1 = none (may be null)
  unsigned* thr_row_enc;         // [M] max target score per row, order-preserving encoding (0 = no target)
  float* thr_col;                // [N] target score per column
  int* cnt_row;                  // [M] += #{j : S_ij > thr_row[i]}
  int* cnt_col;                  // [N] += #{i : S_ij > thr_col[j]}
  unsigned long long* best;      // [M] max over j of (enc(S_ij) << 32 | ~j): row maximum, smallest column on ties
  // peer-memory exchange (multi-GPU).  wait_*: the TMA producer holds its first load until every rank's flag has
  // reached *wait_counter (the gathered B operand was written by peer GPUs).  peer_*: the output of a GRAD launch
  // is read by the other ranks straight out of this GPU's memory; the last CTA of the launch bumps *peer_counter
  // and publishes it to every rank's flag (system scope) once all CTAs' stores are fenced.
  const int* wait_flags;         // FWD: [wait_count] one flag per source rank
  const int* wait_counter;
  int wait_count;
  int wait_rows;                 // FWD: B rows (columns of S) owned by each source rank; a tile waits only for the
                                 // ranks whose rows it loads (0: wait for every rank before the first load)
  int peer_world;                // 0 = nobody to notify
  int* peer_flag_dst[MAX_PEERS];
  int* peer_counter;
  int* peer_ticket;
  // GRADPUSH: the [M, N] output is cut into `peer_world` row blocks of push_rows rows; block q belongs to rank q
  // and is written there (tmPush[q]: [push_rows, N] bf16 in rank q's memory, box 64 x 32) by one TMA store per
  // epilogue warp and tile.  "Complete" is published to every rank by the last CTA, as for the fp32 partial: the
  // consumer (the owner's text-side Jacobian) runs a whole contraction later, early per-owner flags buy nothing
  // and a system-scope fence per box costs the epilogue 15 us per launch (trace r02b).
  int push_rows;
  alignas(64) CUtensorMap tmPush[MAX_PEERS];
  // FWD, all-gather fused into the CONSUMER: the spare control warp of every CTA copies its share of this rank's
  // own bf16 text rows (gath_src, written by the normalise launch just before) into the gathered-V buffer of the
  // other ranks -- destination slot k = rank - k (mod world), one destination after the other, each with its own
  // "rows of rank r are in" flag once every CTA is done with it -- while the other warps already score the own
  // column block.  The exchange runs underneath the forward instead of in front of it, and no rank ever waits
  // before it has pushed (no cross-rank cycle).
  int gath_world;                        // 0 / 1: nothing to push
  int gath_chunks;                       // 16-byte pieces of the own block (rows * D / 8)
  const uint4* gath_src;
  uint4* gath_dst[MAX_PEERS];            // slot k (k >= 1)
  int* gath_flag_dst[MAX_PEERS];
  int* gath_ticket;                      // [MAX_PEERS], zero between launches
  int wait_skip;                         // source rank whose rows need no flag (this rank itself), -1: none
};

// Score of a masked (out-of-range) pair: tau * kMaskedScore is finite and so negative that
// softplus, sigmoid and sigmoid * x are exactly 0 in fp32.
constexpr float kMaskedScore = -30000.f;

// Negative-pair terms of one raw score s = <u_i, v_j> (x = tau s):
//   e = exp(-|x|) (MUFU.EX2), d = 1 + e, sigmoid(|x|) = 1/d (MUFU.RCP), sg = sigmoid(x).
// softplus(x) = max(x, 0) + log(d) is NOT formed per element: the epilogue accumulates sum max(s, 0)
// (scaled by tau once, in the finalize kernel) and the PRODUCT of the d's of a chunk (each in [1, 2], so
// 16 of them cannot overflow), and takes one MUFU.LG2 per chunk: log-sum = log-product.
//
// The epilogue is bound by the XU (MUFU) pipe, not by instruction issue (ncu: XU is the busiest non-tensor pipe,
// warps stall in mio_throttle), so 1/d does NOT go through MUFU.RCP: d lies in [1, 2], a minimax cubic gives
// 1/d to 1.7e-3 and one Newton step squares that to 3e-6 -- five FMA-pipe instructions, far below the bf16
// rounding (2e-3) of the only consumer, Gmat.  sigma(x < 0) = e / d is formed as e * (1/d): no cancellation.
#ifndef JSD_FWD_PACKED_F32X2
#define JSD_FWD_PACKED_F32X2 1
#endif
#ifndef JSD_FWD_RCP_POLY
#define JSD_FWD_RCP_POLY 0
#endif
__device__ __forceinline__ float rcp_1to2(float d) {
#if JSD_FWD_RCP_POLY
  float r = fmaf(d, -0.22183725f, 1.33102348f);
  r = fmaf(d, r, -2.93934318f);
  r = fmaf(d, r, 2.82842386f);
  return r * fmaf(-d, r, 2.f);
#else
  return rcp_approx(d);
#endif
}
__device__ __forceinline__ void neg_terms(float s, float tau_l2, float& d, float& sg) {
  const float e = ex2_approx(-fabsf(s * tau_l2));
  d = 1.f + e;
  const float rr = rcp_1to2(d);
#if JSD_FWD_RCP_POLY
  sg = s >= 0.f ? rr : e * rr;
#else
  sg = s >= 0.f ? rr : 1.f - rr;
#endif
}

// Order-preserving map float -> unsigned (so that atomicMax on the encoding is a float max); 0 is below every float.
__device__ __forceinline__ unsigned enc_ordered(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e ^ 0x80000000u) : ~e);
}

// Work iterator shared by the three roles: yields (tile, k_begin, k_end) segments.
//
// Stream-K (opt-in, GRAD) follows the hybrid schedule of the Stream-K paper, applied per GROUP of
// num_n_blocks workers (worker j of a group owns column block j, so a group walks the same A tiles in
// lock step and the A operand is shared through L2 exactly as in the whole-tile schedule):
//   1. the m-blocks that do not fill a whole wave of groups ("sk_mb" of them) are cut, as one
//      (m-block, k-chunk) unit space, into one contiguous range per group -- every group gets the same
//      fraction of a tile, a tile split between groups is finished by the group holding its head;
//   2. the remaining m-blocks are whole tiles, one wave after the other.
struct SegmentIter {
  int num_k, num_tiles, stride, tile;   // round-robin whole tiles (x ksplit K-slices)
  int ksplit, slice;                    // split-K: slices per tile, slice of the segment last returned
  int n_blocks, n_blk;                  // stream-K: column blocks per m-block, this worker's column block
  int groups, group, sk_mb, dp_i, dp_waves;
  long long u, u_end;                   // stream-K: unit range of this worker's group in the sk_mb space
  bool stream_k;
  __device__ SegmentIter(bool sk, int worker, int n_workers, int tiles, int nk, int nb, int ks = 1)
      : num_k(nk), num_tiles(tiles), stride(n_workers), tile(worker), ksplit(ks < 1 ? 1 : ks), slice(0), n_blocks(nb),
        n_blk(0), groups(1), group(0), sk_mb(0), dp_i(0), dp_waves(0), u(0), u_end(0), stream_k(sk) {
    if (sk) {
      groups = n_workers / nb;
      group = worker / nb;
      n_blk = worker - group * nb;
      const int num_mb = tiles / nb;
      dp_waves = num_mb / groups;
      sk_mb = num_mb - dp_waves * groups;
      u = group_begin(worker, n_workers, tiles, nk, nb);
      u_end = group_begin(worker + nb, n_workers, tiles, nk, nb);
    }
  }
  // first stream-K unit of the group `worker` belongs to (== total for the group past the last)
  __device__ static long long group_begin(int worker, int n_workers, int tiles, int nk, int nb) {
    const int groups = n_workers / nb, num_mb = tiles / nb;
    const int sk = num_mb - (num_mb / groups) * groups;
    return (long long)sk * nk * (worker / nb) / groups;
  }
  __device__ bool next(int& t, int& k0, int& k1) {
    if (stream_k) {
      if (u < u_end) {
        const int m_blk = (int)(u / num_k);
        t = m_blk * n_blocks + n_blk;       // n-fastest tile numbering
        k0 = (int)(u - (long long)m_blk * num_k);
        const long long rem = u_end - u;
        k1 = (k0 + rem < num_k) ? (int)(k0 + rem) : num_k;
        u += k1 - k0;
        return true;
      }
      if (dp_i < dp_waves) {
        t = (sk_mb + group + dp_i * groups) * n_blocks + n_blk;
        k0 = 0;
        k1 = num_k;
        ++dp_i;
        return true;
      }
      return false;
    }
    if (tile >= num_tiles * ksplit) return false;
    t = tile / ksplit;                    // neighbouring workers share an output tile, each takes one K range
    slice = tile - t * ksplit;
    k0 = (int)((long long)slice * num_k / ksplit);
    k1 = (int)((long long)(slice + 1) * num_k / ksplit);
    tile += stride;
    return true;
  }
};

template <int MODE, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
jsd_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmParams p) {
  constexpr int STAGES = num_stages(CG, MODE);
  constexpr int KA = k_atoms(MODE);                // 64-wide k-atoms per pipeline slot
  constexpr int SUB_BYTES = stage_bytes(CG);       // one k-atom of A and B
  constexpr int STAGE_BYTES = KA * SUB_BYTES;
  constexpr int CHUNK_K = KA * BLOCK_K;            // contraction elements consumed per pipeline slot
  constexpr int BN_CTA = b_rows_per_cta(CG);     // rows of the B operand this CTA stages
  constexpr int TILE_M = BLOCK_M * CG;           // rows of one worker tile

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t g_stage_base = smem_base + STAGES * STAGE_BYTES;          // 1024-aligned (FWD only)
  const uint32_t bar_base = g_stage_base + g_staging_bytes(MODE);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 160u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // a "worker" is one CTA (CG = 1) or one CTA pair (CG = 2); rank = this CTA's position in the pair
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int worker = blockIdx.x / CG;
  const int n_workers = gridDim.x / CG;

  const int num_m_blocks = (p.M + TILE_M - 1) / TILE_M;
  const int num_n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k = (p.K + CHUNK_K - 1) / CHUNK_K;
  constexpr bool IS_GRAD = is_grad_mode(MODE);
  // a forward fed by peer GPUs may legitimately stall for as long as the peer-wait limit allows
  const uint64_t spin_limit =
      (MODE == MODE_FWD && p.wait_flags != nullptr) ? g_wait_cfg.timeout_ns + kSpinLimitNs : kSpinLimitNs;
  constexpr bool PUSH = MODE == MODE_GRADPUSH;
  const bool stream_k = (MODE == MODE_GRAD) && p.stream_k != 0;

  constexpr int TRACE_ID = MODE == MODE_FWD ? TK_FWD : (IS_GRAD ? TK_GRAD : TK_SCORE);
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(TRACE_ID, TE_START);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (MODE == MODE_FWD) {
      if (p.gmat != nullptr) tma_prefetch_desc(&p.tmG);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * CG);   // the leader collects both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();   // the peer CTA's barriers and TMEM are set up as well
  __syncthreads();                             // (also after the cluster barrier: compute-sanitizer's racecheck does not
                                               //  model barrier.cluster as ordering the allocator's shared-memory write)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    if (p.n_fastest) {
      m_blk = tile / num_n_blocks;
      n_blk = tile - m_blk * num_n_blocks;
    } else {
      n_blk = tile / num_m_blocks;
      m_blk = tile - n_blk * num_m_blocks;
    }
    m_blk += p.m_rot;                       // rotations (< the block counts; 0 outside the peer paths)
    if (m_blk >= num_m_blocks) m_blk -= num_m_blocks;
    n_blk += p.n_rot;
    if (n_blk >= num_n_blocks) n_blk -= num_n_blocks;
  };

  if (warp == 0) {
    // ===================================================== TMA producer (every CTA stages its own A rows / B rows)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // completion bytes of both CTAs of a pair are counted on the LEADER's full barrier
      const uint32_t full0 = (CG == 2) ? mapa_shared(full_bar(0), 0) : full_bar(0);
      auto load = [&](uint32_t dst, const CUtensorMap* m, int s, int c0, int c1) {
        if constexpr (CG == 2) tma_load_2d_pair(dst, m, full0 + 8u * s, c0, c1);
        else tma_load_2d(dst, m, full0 + 8u * s, c0, c1);
      };
      // The B operand of a peer forward is gathered by the other GPUs writing into this GPU's memory, one flag per
      // source rank ("my rows are in").  A tile waits only for the ranks whose rows it is about to load, so the walk
      // (rotated to start on this rank's own column block) consumes the peers' rows in the order they land.
      uint32_t peers_in = (p.wait_skip >= 0) ? (1u << p.wait_skip) : 0u;   // own rows: stream-ordered, no flag
      int wait_target = 0;
      if (p.wait_flags != nullptr) {
        wait_target = *p.wait_counter;
        if (p.wait_rows <= 0) {
          for (int q = 0; q < p.wait_count; ++q)
            if (q != p.wait_skip) wait_flag_sys(p.wait_flags, q, wait_target, WAIT_GATHERED_ROWS);
          fence_proxy_async_all();
          peers_in = 0xFFFFFFFFu;
          if (blockIdx.x == 0) trace_event(TRACE_ID, TE_PEERS_IN);
        }
      }
      SegmentIter it(stream_k, worker, n_workers, num_tiles, num_k, num_n_blocks, MODE == MODE_GRAD ? p.ksplit : 1);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + rank * BLOCK_M;     // this CTA's A rows
        const int n0 = n_blk * BLOCK_N + rank * BN_CTA;     // this CTA's share of the B rows
        if constexpr (MODE == MODE_FWD) {
          if (p.wait_flags != nullptr && peers_in != 0xFFFFFFFFu && n0 < p.N) {
            const int q_lo = n0 / p.wait_rows;
            const int q_hi = (min(n0 + BN_CTA, p.N) - 1) / p.wait_rows;
            bool waited = false;
            for (int q = q_lo; q <= q_hi && q < p.wait_count; ++q)
              if (!((peers_in >> q) & 1u)) {
                wait_flag_sys(p.wait_flags, q, wait_target, WAIT_GATHERED_ROWS);
                peers_in |= 1u << q;
                waited = true;
              }
            if (waited) {
              fence_proxy_async_all();      // flag observed (generic proxy) before the TMA (async proxy) reads
              if (blockIdx.x == 0 && peers_in == (1u << p.wait_count) - 1u) trace_event(TRACE_ID, TE_PEERS_IN);
            }
          }
        }
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u, spin_limit);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES * CG);
#pragma unroll
          for (int ka = 0; ka < KA; ++ka) {
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES + ka * SUB_BYTES;
            const uint32_t b_dst = a_dst + A_TILE_BYTES;
            const int kk = (kc * KA + ka) * BLOCK_K;       // out-of-range k reads as zeros (TMA fill)
            if constexpr (!A_MN) {
              load(a_dst, &tmA, stage, kk, m0);
            } else {
              // A stored [K, M] with M contiguous: 64-wide MN atoms of 64 k-rows each
#pragma unroll
              for (int a = 0; a < BLOCK_M / 64; ++a) load(a_dst + a * MN_ATOM_BYTES, &tmA, stage, m0 + 64 * a, kk);
            }
            if constexpr (!B_MN) {
              load(b_dst, &tmB, stage, kk, n0);
            } else {
#pragma unroll
              for (int a = 0; a < BN_CTA / 64; ++a) load(b_dst + a * MN_ATOM_BYTES, &tmB, stage, n0 + 64 * a, kk);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread of the leader CTA)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      SegmentIter it(stream_k, worker, n_workers, num_tiles, num_k, num_n_blocks, MODE == MODE_GRAD ? p.ksplit : 1);
      int tile, k0, k1;
      while (it.next(tile, k0, k1)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, spin_limit);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(full_bar(stage), phase, spin_limit);
          tc_fence_after();
#pragma unroll
          for (int ka = 0; ka < KA; ++ka) {
            const uint32_t a_base = smem_base + stage * STAGE_BYTES + ka * SUB_BYTES;
            const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // K-major: step 16 elements (32 B) inside the 128 B swizzle row; MN-major: step 16 k-rows (2 KB)
              const uint64_t a_desc = A_MN ? umma_smem_desc(a_base + k * (UMMA_K * 128), MN_ATOM_BYTES, 1024)
                                           : umma_smem_desc(a_base + k * (UMMA_K * 2), 0, 1024);
              const uint64_t b_desc = B_MN ? umma_smem_desc(b_base + k * (UMMA_K * 128), MN_ATOM_BYTES, 1024)
                                           : umma_smem_desc(b_base + k * (UMMA_K * 2), 0, 1024);
              const uint32_t accumulate = (kc > k0 || ka > 0 || k > 0) ? 1u : 0u;
              if constexpr (CG == 2) umma_bf16_pair(d_tmem, a_desc, b_desc, idesc, accumulate);
              else umma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
            }
          }
          // smem slot free (in both CTAs) once these MMAs retire; accumulator ready after the last chunk
          if constexpr (CG == 2) {
            umma_commit_pair(empty_bar(stage), 0x3);
            if (kc == k1 - 1) umma_commit_pair(tfull_bar(acc), 0x3);
          } else {
            umma_commit(empty_bar(stage));
            if (kc == k1 - 1) umma_commit(tfull_bar(acc));
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================== all-gather by peer stores (FWD of a multi-GPU step)
    // The spare control warp of every CTA copies its share of this rank's own text rows to the other ranks,
    // destination after destination, each with its own flag.  Slow per destination (a system-scope fence on an SM
    // that is busy feeding the tensor cores takes 10-30 us: traces r02d / r02f) but free -- no SM time, no launch --
    // so the host side selects it when the forward is long enough to hide it (2 / 4 GPUs at B = 8192).
    if constexpr (MODE == MODE_FWD) {
      if (p.gath_world > 1) {
        const int e = *p.wait_counter;               // this step's count (bumped by the normalise launch before)
        const int first = (int)blockIdx.x * 32 + lane, stride = (int)gridDim.x * 32;
        for (int k = 1; k < p.gath_world; ++k) {
          uint4* dst = p.gath_dst[k];
          int c = first;
          for (; c + 3 * stride < p.gath_chunks; c += 4 * stride) {      // four 16-byte pieces in flight per lane
            const uint4 a0 = __ldg(p.gath_src + c), a1 = __ldg(p.gath_src + c + stride);
            const uint4 a2 = __ldg(p.gath_src + c + 2 * stride), a3 = __ldg(p.gath_src + c + 3 * stride);
            dst[c] = a0;
            dst[c + stride] = a1;
            dst[c + 2 * stride] = a2;
            dst[c + 3 * stride] = a3;
          }
          for (; c < p.gath_chunks; c += stride) dst[c] = __ldg(p.gath_src + c);
          __threadfence_system();                      // this lane's stores have reached the destination
          __syncwarp();
          if (lane == 0) {
            if (atomicAdd(p.gath_ticket + k, 1) == (int)gridDim.x - 1) {   // every CTA is done with destination k
              __threadfence_system();
              p.gath_ticket[k] = 0;
              st_release_sys(p.gath_flag_dst[k], e);
              trace_event(TK_PUSH, k == p.gath_world - 1 ? TE_END : TE_PEERS_IN);
            }
          }
        }
      }
    }
  } else if (warp >= NUM_CTRL_WARPS) {
    // ===================================================== epilogue (each CTA drains its own 128 TMEM lanes)
    const int ew = warp - NUM_CTRL_WARPS;   // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int cgrp = ew >> 2;               // column group of the tile this warp drains
    const int row_in_tile = 32 * q + lane;
    const uint32_t tempty0 = (CG == 2) ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    const uint32_t g_stage = g_stage_base + ew * G_STAGE_BYTES;   // this warp's Gmat staging box (FWD)
    const int cta = blockIdx.x;             // stream-K slots / flags are per CTA
    const int sk_step = num_n_blocks * CG;  // CTA holding the same rows / column block in the next group

    float tau = 1.f, tau_l2 = 1.f, gscale = 1.f;
    if constexpr (MODE == MODE_FWD) {
      tau = expf(*p.t_dev);
      tau_l2 = tau * 1.4426950408889634f;
    } else {
      const float gamma = p.gamma_dev ? *p.gamma_dev : 1.f;
      gscale = gamma * (p.t_dev ? expf(*p.t_dev) : 1.f) * p.scale;
    }
    float pos_sum = 0.f, relu_sum = 0.f, lg_sum = 0.f;   // sum softplus(-x_pos), sum max(s,0), sum log2(1+e)
    int acc = 0;
    uint32_t acc_phase = 0;
    SegmentIter it(stream_k, worker, n_workers, num_tiles, num_k, num_n_blocks, MODE == MODE_GRAD ? p.ksplit : 1);
    int tile, k0, k1;
    while (it.next(tile, k0, k1)) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int m0 = m_blk * TILE_M + rank * BLOCK_M, n0 = n_blk * BLOCK_N;   // this CTA's rows, the tile's columns
      const int grow = m0 + row_in_tile;
      const bool row_ok = grow < p.M;

      // stream-K bookkeeping for this segment (GRAD only)
      const bool sk_partial = stream_k && k0 > 0;                    // tail/middle of a tile: publish, no output
      const bool sk_finish = stream_k && k0 == 0 && k1 < num_k;      // head of a split tile: add the others' partials
      int sk_last = cta;                                             // last CTA contributing to this CTA's rows
      if (sk_finish) {
        const long long m_end = (long long)(tile / num_n_blocks + 1) * num_k;   // end of this m-block's units
        while (sk_last + sk_step < (int)gridDim.x &&
               SegmentIter::group_begin((sk_last + sk_step) / CG, n_workers, num_tiles, num_k, num_n_blocks) < m_end)
          sk_last += sk_step;               // group_begin counts units of the stream-K m-blocks only
        if (lane == 0) {
          for (int c = cta + sk_step; c <= sk_last; c += sk_step) {
            const int* flag = p.sk_flags + c * NUM_EPI_WARPS + ew;
            uint32_t spins = 0;
            uint64_t t_start = 0;
            while (ld_acquire_gpu(flag) == 0) {
              if ((++spins & 0x3FFu) == 0) {
                const uint64_t now = global_timer_ns();
                if (t_start == 0) t_start = now;
                else if (now - t_start > 20000000000ull) __trap();
              }
            }
          }
        }
        __syncwarp();
      }

#ifndef JSD_GRAD_EPI_SLEEP_NS
#define JSD_GRAD_EPI_SLEEP_NS 1000
#endif
      if constexpr (IS_GRAD && JSD_GRAD_EPI_SLEEP_NS > 0) {
        mbar_wait_relaxed(tfull_bar(acc), acc_phase, JSD_GRAD_EPI_SLEEP_NS, spin_limit);   // long wait: poll gently
      } else {
        mbar_wait(tfull_bar(acc), acc_phase, spin_limit);   // short waits: all lanes poll (4 % faster than one lane + __syncwarp)
      }
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BLOCK_N + cgrp * COLS_PER_WARP + ((uint32_t)(32 * q) << 16);

      // SCORE: per-row state of this tile (thread = row)
      [[maybe_unused]] int sc_cnt = 0, sc_tb = 0, sc_te = 0, sc_bj = 0;
      [[maybe_unused]] float sc_thr = 0.f, sc_bv = 0.f;
      [[maybe_unused]] bool sc_has_best = false;
      if constexpr (MODE == MODE_SCORE) {
        if (row_ok) {
          if (p.score_pass == 0 && p.row_tgt_ptr != nullptr) {
            sc_tb = p.row_tgt_ptr[grow];
            sc_te = p.row_tgt_ptr[grow + 1];
          }
          if (p.score_pass == 1 && p.thr_row_enc != nullptr) sc_thr = dec_ordered(p.thr_row_enc[grow]);
        }
      }

      // One 32-row x CW-column chunk held in registers (thread = row, v[j] = column col0 + j).
      auto process_chunk = [&](uint32_t(&v)[CW], const int c) {
        const int col_in_tile = cgrp * COLS_PER_WARP + CW * c;
        const int col0 = n0 + col_in_tile;
        if constexpr (MODE == MODE_SCORE) {
          const int ncols = min(CW, p.N - col0);            // existing columns of this chunk (may be <= 0)
          const int nvalid = row_ok ? ncols : 0;            // ... that this thread's row may look at
          if (p.score_pass == 0) {
            for (int k = sc_tb; k < sc_te; ++k) {
              const int tc = p.row_tgt_idx[k] - col0;
              if ((unsigned)tc < (unsigned)max(nvalid, 0)) {
                float val = 0.f;
#pragma unroll
                for (int j = 0; j < CW; ++j) val = (j == tc) ? __uint_as_float(v[j]) : val;
                atomicMax(p.thr_row_enc + grow, enc_ordered(val));
              }
            }
            if (p.col_tgt != nullptr) {
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (j < nvalid && p.col_tgt[col0 + j] == grow) p.thr_col[col0 + j] = __uint_as_float(v[j]);
            }
          } else {
            if (p.thr_row_enc != nullptr) {
#pragma unroll
              for (int j = 0; j < CW; ++j) sc_cnt += (j < nvalid && __uint_as_float(v[j]) > sc_thr) ? 1 : 0;
            }
            if (p.thr_col != nullptr) {
              int mine = 0;
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const bool g = j < nvalid && __uint_as_float(v[j]) > p.thr_col[col0 + j];
                const unsigned b = __ballot_sync(0xffffffffu, g);
                if (lane == j) mine = __popc(b);
              }
              if (lane < ncols && mine > 0) atomicAdd(p.cnt_col + col0 + lane, mine);
            }
            if (p.best != nullptr) {
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const float sv = __uint_as_float(v[j]);
                if (j < nvalid && (!sc_has_best || sv > sc_bv)) {
                  sc_bv = sv;
                  sc_bj = col0 + j;
                  sc_has_best = true;
                }
              }
            }
          }
        } else if constexpr (MODE == MODE_FWD) {
          // ---- scores -> softplus / sigmoid.  Every element first takes the negative-pair path;
          //      the (rare) chunk holding this warp's positives is corrected afterwards.
          if ((col0 + CW > p.N) || (m0 + BLOCK_M > p.M)) {
            // edge tile: rows/columns beyond the problem read as s = 0 (TMA zero fill); push them to a
            // large negative score so that softplus, sigmoid and sigma*x all vanish exactly
            const int nvalid = row_ok ? p.N - col0 : 0;
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (j >= nvalid) v[j] = __float_as_uint(kMaskedScore);
          }
          uint32_t packed[CW / 2];
#if JSD_FWD_PACKED_F32X2
          // two elements per FMA-pipe instruction (sm_100 f32x2 arithmetic): the epilogue's ALU work runs just as
          // long as the MMAs of a tile, so every instruction saved here comes off the kernel time
          float2 dprod2 = make_float2(1.f, 1.f), relu2 = make_float2(0.f, 0.f);
          const float2 tl2 = make_float2(tau_l2, tau_l2), one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f);
#pragma unroll
          for (int j = 0; j < CW; j += 2) {
            const float2 sv = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            const float2 y = __fmul2_rn(sv, tl2);
            const float2 d = __fadd2_rn(make_float2(ex2_approx(-fabsf(y.x)), ex2_approx(-fabsf(y.y))), one2);
            const float2 rr = make_float2(rcp_approx(d.x), rcp_approx(d.y));
            const float2 om = __ffma2_rn(rr, mone2, one2);                      // 1 - rr
            dprod2 = __fmul2_rn(dprod2, d);
            relu2 = __fadd2_rn(relu2, make_float2(fmaxf(sv.x, 0.f), fmaxf(sv.y, 0.f)));
            packed[j >> 1] = pack_bf16x2(sv.x >= 0.f ? rr.x : om.x, sv.y >= 0.f ? rr.y : om.y);
          }
          relu_sum += relu2.x + relu2.y;
          lg_sum += lg2_approx(dprod2.x * dprod2.y);      // 16 factors in [1, 2]: no overflow
#else
          float dprod = 1.f;
#pragma unroll
          for (int j = 0; j < CW; j += 2) {
            float sg[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float sv = __uint_as_float(v[j + h]);
              float d;
              neg_terms(sv, tau_l2, d, sg[h]);
              dprod *= d;
              relu_sum += fmaxf(sv, 0.f);
            }
            packed[j >> 1] = pack_bf16x2(sg[0], sg[1]);
          }
          lg_sum += lg2_approx(dprod);
#endif
          const int dj = p.row_offset + grow - col0;          // this row's positive sits at v[dj] if 0 <= dj < CW
          if (__any_sync(0xffffffffu, row_ok && (unsigned)dj < (unsigned)CW)) {
            float s_d = 0.f;
#pragma unroll
            for (int j = 0; j < CW; ++j) s_d = (j == dj) ? __uint_as_float(v[j]) : s_d;
            if (row_ok && (unsigned)dj < (unsigned)CW) {
              // take the positive pair back out of the negative sums, then add it in full precision
              const float x = s_d * tau;
              const float e = expf(-fabsf(x));
              relu_sum -= fmaxf(s_d, 0.f);
              lg_sum -= log2f(1.f + ex2_approx(-fabsf(s_d * tau_l2)));
              const float rr = 1.f / (1.f + e);
              const float gneg = x >= 0.f ? -(e * rr) : -rr;    // -sigma(-x)
              pos_sum += fmaxf(-x, 0.f) + log1pf(e);            // softplus(-x)
              p.gdiag[grow] = gneg;
            }
          }
          if (p.gmat != nullptr) {
            // stage the bf16 chunk in this warp's 32 x 64 box (128-byte rows, 128B swizzle: 16-byte piece
            // index XOR row%8 -> conflict-free v4 stores); one TMA store per warp and tile writes it out
            const uint32_t row_base = g_stage + lane * 128;
#pragma unroll
            for (int k4 = 0; k4 < CW / 8; ++k4) {
              const uint32_t piece = (uint32_t)((CW / 8) * c + k4) ^ (uint32_t)(lane & 7);
              st_shared_v4(row_base + piece * 16, packed[4 * k4], packed[4 * k4 + 1], packed[4 * k4 + 2],
                           packed[4 * k4 + 3]);
            }
            if (row_ok && (unsigned)dj < (unsigned)CW) {     // Gmat carries 0 on the positives
              const uint32_t piece = (uint32_t)((CW / 8) * c + (dj >> 3)) ^ (uint32_t)(lane & 7);
              st_shared_u16(row_base + piece * 16 + (dj & 7) * 2, 0);
            }
          }
        } else if constexpr (PUSH) {
          // bf16(gscale * acc) into this warp's staging box (same swizzled layout as the forward's Gmat box)
          uint32_t packed[CW / 2];
#pragma unroll
          for (int j = 0; j < CW; j += 2)
            packed[j >> 1] = pack_bf16x2(__uint_as_float(v[j]) * gscale, __uint_as_float(v[j + 1]) * gscale);
          const uint32_t row_base = g_stage + lane * 128;
#pragma unroll
          for (int k4 = 0; k4 < CW / 8; ++k4) {
            const uint32_t piece = (uint32_t)((CW / 8) * c + k4) ^ (uint32_t)(lane & 7);
            st_shared_v4(row_base + piece * 16, packed[4 * k4], packed[4 * k4 + 1], packed[4 * k4 + 2],
                         packed[4 * k4 + 3]);
          }
        } else {
          const long long slot_off = (long long)row_in_tile * BLOCK_N + col_in_tile;
          if (sk_partial) {
            // raw fp32 partial accumulator -> this CTA's workspace slot (consumed by the tile's head CTA)
            float4* dst = reinterpret_cast<float4*>(p.sk_slots + (long long)cta * SK_SLOT_FLOATS + slot_off);
#pragma unroll
            for (int k4 = 0; k4 < CW / 4; ++k4)
              dst[k4] = make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]),
                                    __uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3]));
            return;
          }
          if (sk_finish) {
            for (int cc = cta + sk_step; cc <= sk_last; cc += sk_step) {
              const float4* src = reinterpret_cast<const float4*>(p.sk_slots + (long long)cc * SK_SLOT_FLOATS + slot_off);
#pragma unroll
              for (int k4 = 0; k4 < CW / 4; ++k4) {
                const float4 a = __ldcg(src + k4);
                v[4 * k4] = __float_as_uint(__uint_as_float(v[4 * k4]) + a.x);
                v[4 * k4 + 1] = __float_as_uint(__uint_as_float(v[4 * k4 + 1]) + a.y);
                v[4 * k4 + 2] = __float_as_uint(__uint_as_float(v[4 * k4 + 2]) + a.z);
                v[4 * k4 + 3] = __float_as_uint(__uint_as_float(v[4 * k4 + 3]) + a.w);
              }
            }
          }
          if (row_ok) {
            float* obase = it.slice == 0 ? p.out : p.slice_base + (long long)(it.slice - 1) * p.slice_stride;
            float* dst = obase + (long long)grow * p.ldo + col0;
            if (col0 + CW <= p.N) {
#pragma unroll
              for (int k4 = 0; k4 < CW / 4; ++k4) {
                float4 o;
                o.x = __uint_as_float(v[4 * k4]) * gscale;
                o.y = __uint_as_float(v[4 * k4 + 1]) * gscale;
                o.z = __uint_as_float(v[4 * k4 + 2]) * gscale;
                o.w = __uint_as_float(v[4 * k4 + 3]) * gscale;
                reinterpret_cast<float4*>(dst)[k4] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (col0 + j < p.N) dst[j] = __uint_as_float(v[j]) * gscale;
            }
          }
        }
      };

      // 4 chunks per warp and tile, software-pipelined over two register buffers; the loop is kept
      // rolled (2 chunk bodies in the binary) so that the epilogue stays inside the instruction cache
      if constexpr (MODE == MODE_FWD) {
        if (p.gmat != nullptr) {       // the previous tile's TMA store must have read the staging box
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
      }
      if constexpr (PUSH) {            // the previous tile's TMA store must have read the staging box
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      uint32_t ra[CW], rb[CW];
      tmem_ld_chunk<CW>(t_base, ra);
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {
        tmem_ld_wait();
        tmem_ld_chunk<CW>(t_base + CW * (2 * cp + 1), rb);
        process_chunk(ra, 2 * cp);
        tmem_ld_wait();
        if (cp == 0) {
          tmem_ld_chunk<CW>(t_base + 2 * CW, ra);
        } else {
          // every TMEM read of this accumulator stage has landed: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);   // the leader's barrier
            else mbar_arrive(tempty0 + 8u * acc);
          }
        }
        process_chunk(rb, 2 * cp + 1);
      }

      if constexpr (MODE == MODE_FWD) {
        if (p.gmat != nullptr) {
          fence_proxy_async();         // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmG, g_stage, n0 + cgrp * COLS_PER_WARP, m0 + 32 * q);
            tma_store_commit();
          }
        }
      }
      if constexpr (PUSH) {
        fence_proxy_async();           // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        const int box_row = m0 + 32 * q, box_col = n0 + cgrp * COLS_PER_WARP;
        if (box_row < p.M && box_col < p.N) {
          const int owner = box_row / p.push_rows;
          if (lane == 0) {
            tma_store_2d(&p.tmPush[owner], g_stage, box_col, box_row - owner * p.push_rows);
            tma_store_commit();
          }
        }
      }
      if constexpr (MODE == MODE_SCORE) {
        if (row_ok && p.score_pass == 1) {
          if (sc_cnt > 0) atomicAdd(p.cnt_row + grow, sc_cnt);
          if (sc_has_best)
            atomicMax(p.best + grow, ((unsigned long long)enc_ordered(sc_bv) << 32) | (0xFFFFFFFFu - (unsigned)sc_bj));
        }
      }
      if constexpr (MODE == MODE_GRAD) {
        if (sk_partial) {
          __threadfence();          // partial tile visible device-wide before the flag
          __syncwarp();
          if (lane == 0) st_release_gpu(p.sk_flags + cta * NUM_EPI_WARPS + ew, 1);
        } else if (sk_finish) {
          __syncwarp();             // all lanes have consumed the partials: re-arm the flags for the next launch
          if (lane == 0)
            for (int c = cta + sk_step; c <= sk_last; c += sk_step) p.sk_flags[c * NUM_EPI_WARPS + ew] = 0;
        }
      }
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if constexpr (PUSH) {
      // every box of this warp has been WRITTEN to its owner's memory (not merely read from the staging box)
      // before the CTA takes its ticket below
      if (lane == 0) {
        tma_store_wait_all();
        fence_proxy_async_all();
      }
    }
    if constexpr (MODE == MODE_FWD) {
      if (p.gmat != nullptr && lane == 0) tma_store_wait_all();   // smem must outlive the last store
      pos_sum = warp_sum(pos_sum);
      relu_sum = warp_sum(relu_sum);
      lg_sum = warp_sum(lg_sum);
      if (lane == 0) {
        float* dst = p.partials + ((long long)blockIdx.x * NUM_EPI_WARPS + ew) * PARTIALS_PER_WARP;
        dst[0] = pos_sum;
        dst[1] = relu_sum;
        dst[2] = lg_sum;
        dst[3] = 0.f;
        __threadfence();   // visible device-wide before this CTA takes its finalisation ticket
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer may still signal / read this CTA
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }

  if constexpr (MODE != MODE_FWD) {
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_event(TRACE_ID, TE_END);   // block 0's end (no last-CTA ticket here)
  }
  if constexpr (IS_GRAD) {
    // Peer exchange.  fp32 route: once every CTA's stores are fenced, the last CTA publishes "this rank's partial
    // is complete" -- one thread per destination, so that the remote release-stores (each a ~2 us round trip) go
    // out together instead of one after the other.  GRADPUSH: the flags went out per owner block already; the
    // last CTA only advances the launch counter they were derived from.
    if (p.peer_world > 0) {
      // (the barrier above orders every thread's stores before thread 0's fence, which is cumulative: one
      //  system-scope fence per CTA instead of one per epilogue thread)
      volatile int* s_last = reinterpret_cast<volatile int*>(smem_raw);
      if (threadIdx.x == 0) {
        __threadfence_system();
        s_last[0] = (atomicAdd(p.peer_ticket, 1) == (int)gridDim.x - 1);
      }
      __syncthreads();
      if (s_last[0]) {
        if (threadIdx.x == 0) {
          __threadfence_system();
          const int e = *p.peer_counter + 1;
          *p.peer_counter = e;
          s_last[1] = e;
          *p.peer_ticket = 0;
        }
        __syncthreads();
        if ((int)threadIdx.x < p.peer_world) {
          __threadfence_system();
          st_release_sys(p.peer_flag_dst[threadIdx.x], s_last[1]);
        }
      }
    }
  }

  if constexpr (MODE == MODE_FWD) {
    // Loss finalisation without a second launch: the last CTA to get here reduces every warp's partial sums
    // in a fixed order (fp64, deterministic).  The pipeline is drained, so the operand ring is free scratch.
    //   pos = P0 / M,   neg = (tau * P1 + ln2 * P2) / (M (N - 1))
    int* s_last = reinterpret_cast<int*>(smem_raw);
    double* sh = reinterpret_cast<double*>(smem_raw + 16);
    if (threadIdx.x == 0) {
      __threadfence();                                  // this CTA's partials (written before the barrier above)
      *s_last = (atomicAdd(p.ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      const int n = (int)gridDim.x * NUM_EPI_WARPS;
      double acc[3] = {0.0, 0.0, 0.0};
      for (int i = threadIdx.x; i < n; i += GEMM_THREADS)
        for (int k = 0; k < 3; ++k) acc[k] += (double)__ldcg(p.partials + (size_t)i * PARTIALS_PER_WARP + k);
      // fixed reduction tree (warp butterflies, then the 20 warp totals in order): short dependent fp64 chains
#pragma unroll
      for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        if (lane == 0) sh[k * 32 + warp] = acc[k];
      }
      __syncthreads();
      if (threadIdx.x < 3) {
        double tot = 0.0;
        for (int w = 0; w < GEMM_THREADS / 32; ++w) tot += sh[threadIdx.x * 32 + w];
        sh[96 + threadIdx.x] = tot;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const double tau = exp((double)*p.t_dev);
        const double pos = sh[96] * p.inv_pos;
        const double neg = (tau * sh[97] + 0.6931471805599453 * sh[98]) * p.inv_neg;
        p.out4[0] = (float)pos;
        p.out4[1] = (float)neg;
        p.out4[2] = (float)(pos + neg);
        p.out4[3] = 0.f;
        if (p.loss_out) *p.loss_out = (float)(pos + neg);
        *p.ticket = 0;                                  // re-armed for the next launch on this stream
        trace_event(TRACE_ID, TE_END);
      }
    }
  }
}

}  // namespace jsd
