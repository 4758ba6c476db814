// Dense (all-pairs) JSD estimator on sm_100a tensor cores.
//
// One warp-specialised, persistent tcgen05 GEMM kernel, instantiated three ways:
//
//   FWD      S = U . V^T                       (A = U  [M,D]  K-major, B = V   [N,D] K-major)
//            epilogue: x = tau*S -> softplus / sigmoid -> per-CTA loss partials,
//            Gmat[i,j] = sigma(x_ij) (bf16, 0 on the positive diagonal), gdiag[i] = -sigma(-x_ii)
//   GRAD_DU  dUacc = scale * Gmat . V          (A = Gmat [M,N] K-major, B = V^T [D,N] K-major)
//   GRAD_DV  dVacc = scale * Gmat^T . U        (A = Gmat read MN-major,  B = U^T [D,M] K-major)
//
// The B x B score matrix itself is never written; what crosses the fwd/bwd
// boundary is the bf16 sigmoid-coefficient matrix Gmat (see DESIGN.md for why the
// dU/dV accumulators of a D=1024 problem cannot live in TMEM next to S tiles).
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue (two column halves x four TMEM lane quarters).
// Pipelines: 4-stage smem ring (full/empty mbarriers), 2-stage TMEM accumulator
// ring (tfull/tempty) so the epilogue of tile t overlaps the MMAs of tile t+1.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace jsd {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int NUM_CTRL_WARPS = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 32 * (NUM_CTRL_WARPS + NUM_EPI_WARPS);
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;       // 512
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* align slack */ + 256 /* barriers */;
constexpr int PARTIALS_PER_WARP = 4;                  // pos, neg, dt_pos, dt_neg

enum GemmMode { MODE_FWD = 0, MODE_GRAD = 1 };

struct GemmParams {
  int M;            // rows of the output tile space (FWD: image rows; GRAD: rows of the gradient)
  int N;            // cols of the output tile space (FWD: text rows;  GRAD: D)
  int K;            // contraction length
  int n_fastest;    // tile order: 1 = consecutive CTAs walk N first (A tile shared through L2)
  // FWD
  int row_offset;   // column of the positive of local row 0
  const float* t_dev;
  __nv_bfloat16* gmat;   // may be null (loss only)
  long long ldg;
  float* gdiag;          // [M]
  float* partials;       // [grid * NUM_EPI_WARPS * PARTIALS_PER_WARP]
  // GRAD
  const float* gamma_dev;  // may be null (gamma = 1)
  float scale;             // 1 / (M_rows * (N_cols - 1))
  float* out;              // [M, N] fp32, pitch ldo
  long long ldo;
};

template <int MODE, bool A_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
jsd_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 32u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 64u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 80u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 96u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
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
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        for (int kc = 0; kc < num_k; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_TILE_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, m0);
          } else {
            // A is stored [K, M] with M contiguous: two 64-wide MN atoms of 64 k-rows each
            tma_load_2d(a_dst, &tmA, full_bar(stage), m0, kc * BLOCK_K);
            tma_load_2d(a_dst + A_TILE_BYTES / 2, &tmA, full_bar(stage), m0 + 64, kc * BLOCK_K);
          }
          tma_load_2d(b_dst, &tmB, full_bar(stage), kc * BLOCK_K, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, 0u);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kc = 0; kc < num_k; ++kc) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_base = smem_base + stage * STAGE_BYTES;
          const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = A_MN ? umma_smem_desc(a_base + k * (UMMA_K * 128), A_TILE_BYTES / 2, 1024)
                                         : umma_smem_desc(a_base + k * (UMMA_K * 2), 0, 1024);
            const uint64_t b_desc = umma_smem_desc(b_base + k * (UMMA_K * 2), 0, 1024);
            umma_bf16(d_tmem, a_desc, b_desc, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
          if (kc == num_k - 1) umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
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
  } else if (warp >= NUM_CTRL_WARPS) {
    // ===================================================== epilogue
    const int ew = warp - NUM_CTRL_WARPS;   // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = ew >> 2;               // column half of the tile
    const int row_in_tile = 32 * q + lane;

    float tau = 1.f, tau_l2 = 1.f, gscale = 1.f;
    if constexpr (MODE == MODE_FWD) {
      tau = expf(*p.t_dev);
      tau_l2 = tau * 1.4426950408889634f;
    } else {
      const float gamma = p.gamma_dev ? *p.gamma_dev : 1.f;
      gscale = gamma * (p.t_dev ? expf(*p.t_dev) : 1.f) * p.scale;
    }
    float pos_sum = 0.f, neg_sum = 0.f, dtp_sum = 0.f, dtn_sum = 0.f;

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
      const int grow = m0 + row_in_tile;
      const bool row_ok = grow < p.M;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BLOCK_N + half * (BLOCK_N / 2) + ((uint32_t)(32 * q) << 16);

      uint32_t r[2][32];
      tmem_ld_32x32(t_base, r[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c + 1 < 4) tmem_ld_32x32(t_base + 32 * (c + 1), r[(c + 1) & 1]);
        const uint32_t(&v)[32] = r[c & 1];
        const int col0 = n0 + half * (BLOCK_N / 2) + 32 * c;

        if constexpr (MODE == MODE_FWD) {
          const int wdiag0 = p.row_offset + m0 + 32 * q;   // positive column of this warp's first row
          const bool has_diag = (wdiag0 < col0 + 32) && (col0 < wdiag0 + 32);
          const bool edge = (col0 + 32 > p.N) || (m0 + BLOCK_M > p.M);
          uint32_t packed[16];
          if (!has_diag && !edge) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float sg[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float s = __uint_as_float(v[j + h]);
                const float x = s * tau;
                const float e = ex2_approx(-fabsf(s * tau_l2));
                const float d = 1.f + e;
                const float rr = rcp_approx(d);
                const float lp = lg2_approx(d);
                neg_sum += fmaf(lp, 0.6931471805599453f, fmaxf(x, 0.f));
                const float sig = x >= 0.f ? rr : 1.f - rr;
                dtn_sum = fmaf(sig, x, dtn_sum);
                sg[h] = sig;
              }
              packed[j >> 1] = pack_bf16x2(sg[0], sg[1]);
            }
          } else {
            const int dcol = p.row_offset + grow;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float sg[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int col = col0 + j + h;
                const float s = __uint_as_float(v[j + h]);
                const float x = s * tau;
                const float e = ex2_approx(-fabsf(s * tau_l2));
                const float d = 1.f + e;
                const float rr = rcp_approx(d);
                const float sp = fmaf(lg2_approx(d), 0.6931471805599453f, fmaxf(x, 0.f));
                const float sig = x >= 0.f ? rr : 1.f - rr;
                const bool ok = row_ok && col < p.N;
                const bool is_diag = ok && col == dcol;
                if (is_diag) {
                  // positive pair: softplus(-x) = softplus(x) - x, dL/dx ~ -sigma(-x) = sigma(x) - 1
                  const float gneg = x >= 0.f ? -(e * rr) : -rr;   // -(1 - sigma(x)) without cancellation
                  pos_sum += fmaxf(-x, 0.f) + log1pf(e);           // rare path: full-precision softplus(-x)
                  dtp_sum = fmaf(gneg, x, dtp_sum);
                  p.gdiag[grow] = gneg;
                  sg[h] = 0.f;
                } else if (ok) {
                  neg_sum += sp;
                  dtn_sum = fmaf(sig, x, dtn_sum);
                  sg[h] = sig;
                } else {
                  sg[h] = 0.f;
                }
              }
              packed[j >> 1] = pack_bf16x2(sg[0], sg[1]);
            }
          }
          if (p.gmat != nullptr && row_ok && col0 < p.ldg) {
            uint4* dst = reinterpret_cast<uint4*>(p.gmat + (long long)grow * p.ldg + col0);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              dst[k4] = make_uint4(packed[4 * k4], packed[4 * k4 + 1], packed[4 * k4 + 2], packed[4 * k4 + 3]);
          }
        } else {
          if (row_ok) {
            float* dst = p.out + (long long)grow * p.ldo + col0;
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int k4 = 0; k4 < 8; ++k4) {
                float4 o;
                o.x = __uint_as_float(v[4 * k4]) * gscale;
                o.y = __uint_as_float(v[4 * k4 + 1]) * gscale;
                o.z = __uint_as_float(v[4 * k4 + 2]) * gscale;
                o.w = __uint_as_float(v[4 * k4 + 3]) * gscale;
                reinterpret_cast<float4*>(dst)[k4] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) dst[j] = __uint_as_float(v[j]) * gscale;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if constexpr (MODE == MODE_FWD) {
      pos_sum = warp_sum(pos_sum);
      neg_sum = warp_sum(neg_sum);
      dtp_sum = warp_sum(dtp_sum);
      dtn_sum = warp_sum(dtn_sum);
      if (lane == 0) {
        float* dst = p.partials + ((long long)blockIdx.x * NUM_EPI_WARPS + ew) * PARTIALS_PER_WARP;
        dst[0] = pos_sum;
        dst[1] = neg_sum;
        dst[2] = dtp_sum;
        dst[3] = dtn_sum;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace jsd
