// Single-pass fused forward + backward of the dense JSD estimator for D <= 256 (BASELINE north star (1) + (2) as
// written): the score tile lives in TMEM, sigma(S) goes through shared memory straight back into the tensor cores,
// the gradient accumulator stays in TMEM -- neither S nor sigma(S) ever reaches HBM.
//
// One CTA owns 128 rows of one modality (A block, resident in shared memory) and walks the other modality in blocks
// of 128 rows (B tiles, TMA-staged).  Per B tile j:
//
//   MMA1   S   = A . B_j^T            M = 128, N = 128, K = D     -> TMEM, double-buffered (2 x 128 columns)
//   epi    x = tau S -> softplus terms (loss), P = bf16 sigma(x), 0 on the positives -> shared memory, laid out as
//          a K-major SWIZZLE_128B A operand (two 64-wide k-atoms)
//   MMA2   acc += P . B_j              M = 128, N = D,   K = 128   -> TMEM (D <= 256 columns), B_j read MN-major
//                                                                    from the very tile MMA1 read K-major
//
// TMEM: 2 x 128 (S) + D (acc) <= 512 columns -- this is why the fused form exists for D <= 256 only (DESIGN.md
// section 3); D = 1024 keeps the staged path through the bf16 Gmat.  The text-side gradient needs the transposed
// coefficients, which live in other CTAs' tiles: the same kernel runs a second set of CTAs with the modalities
// swapped (A = text rows, B = image rows) and recomputes those score tiles -- 8 B^2 D executed FLOPs for 6 B^2 D
// algorithmic ones, no atomics, no B x B traffic.  Both directions are ONE launch (blockIdx selects).
//
// Outputs (the upstream gradient is not known yet when autograd runs the forward, so the accumulators are left
// unscaled; the Jacobian kernel applies gamma tau / (M (N - 1)) in fp32):
//   acc[dir]  [rows, D] fp32   sum_j sigma(x_ij) b_j  over the negatives
//   gdiag     [M] fp32         -sigma(-x_ii')                     (direction 0)
//   loss      as jsd_gemm_kernel<MODE_FWD>: per-warp partials, last CTA reduces in fp64 (direction 0 only)
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..19 epilogue: two groups of 8
// (4 TMEM lane quarters x 2 column halves) working on alternate tiles (ping-pong).
// Measured and rejected: sigma from one half-precision tanh per pair of elements instead of ex2 + rcp per element
// (half the MUFU operations): no change (172.6 vs 162.8 us at B = 8192, D = 128, r02o) -- the epilogue was bound by
// latency in lock step, not by the MUFU pipe.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "jsd_dense.cuh"


namespace jsd {

constexpr int FB = 128;                       // rows per CTA and rows of the other modality per iteration
constexpr int F_ATOM_BYTES = FB * 128;        // one 64-wide k-atom of a 128-row tile: 16 KB
constexpr int F_P_BYTES = 2 * F_ATOM_BYTES;   // P tile: 128 x 128 bf16
constexpr int F_EPI_WARPS = 16;
constexpr int F_THREADS = 32 * (4 + F_EPI_WARPS);
constexpr int F_ACC_COL = 2 * FB;             // TMEM column of the gradient accumulator

__host__ __device__ constexpr int fused_v_stages(int nka) { return nka <= 2 ? 3 : 2; }
__host__ __device__ constexpr int fused_p_bufs(int nka) { return nka >= 4 ? 1 : 2; }
__host__ __device__ constexpr int fused_smem_bytes(int nka) {
  return nka * F_ATOM_BYTES * (1 + fused_v_stages(nka)) + fused_p_bufs(nka) * F_P_BYTES + 1024 + 256;
}

struct FusedProblem {
  int M;              // rows owned by this direction's CTAs (A)
  int N;              // rows of the other modality (B), walked in blocks of 128
  int row_offset;     // the positive of local row i is B row row_offset + i
  float* acc;         // [M, D] fp32
  float* gdiag;       // [M], direction 0 only (null otherwise)
  int want_loss;
};

struct FusedParams {
  int D;
  const float* t_dev;
  FusedProblem prob[2];
  int blocks0;        // CTAs of direction 0; blockIdx.x >= blocks0 belongs to direction 1
  int nsplit;         // small batches: every row block is handled by nsplit CTAs, each walking 1/nsplit of the other
                      // modality and writing its own accumulator slice (acc + split * acc_stride); the Jacobian
                      // kernel adds the slices in order.  B = 1024: 16 CTAs x 8 serial tiles -> 128 CTAs x 1 tile
  long long acc_stride;
  float* partials;    // [grid * F_EPI_WARPS * PARTIALS_PER_WARP]
  int* ticket;
  float* out4;
  float* loss_out;
  double inv_pos, inv_neg;
};

template <int NKA>
__global__ void __launch_bounds__(F_THREADS, 1)
jsd_fused_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ FusedParams p) {
  constexpr int VS = fused_v_stages(NKA);
  constexpr int PB = fused_p_bufs(NKA);
  constexpr int TILE_BYTES = NKA * F_ATOM_BYTES;
  constexpr int D = NKA * 64;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  auto v_base = [&](int s) { return smem_base + TILE_BYTES * (1 + s); };
  auto p_base = [&](int b) { return smem_base + TILE_BYTES * (1 + VS) + F_P_BYTES * b; };
  const uint32_t bar_base = smem_base + TILE_BYTES * (1 + VS) + F_P_BYTES * PB;
  const uint32_t a_full = bar_base;
  auto v_full = [&](int s) { return bar_base + 8u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 40u + 8u * s; };
  auto s_full = [&](int b) { return bar_base + 72u + 8u * b; };
  auto s_empty = [&](int b) { return bar_base + 88u + 8u * b; };
  auto p_full = [&](int b) { return bar_base + 104u + 8u * b; };
  auto p_empty = [&](int b) { return bar_base + 120u + 8u * b; };
  const uint32_t acc_full = bar_base + 136u;
  const uint32_t tmem_slot = bar_base + 144u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int dir = (int)blockIdx.x >= p.blocks0 ? 1 : 0;
  const FusedProblem& pr = p.prob[dir];
  const CUtensorMap* tmA = dir ? &tmV : &tmU;
  const CUtensorMap* tmB = dir ? &tmU : &tmV;
  const int local = (int)blockIdx.x - (dir ? p.blocks0 : 0);
  const int split = local % p.nsplit;
  const int m0 = (local / p.nsplit) * FB;
  const int nj_all = (pr.N + FB - 1) / FB;
  const int j_begin = (int)((long long)split * nj_all / p.nsplit);      // this CTA's share of the other modality
  const int nj = (int)((long long)(split + 1) * nj_all / p.nsplit) - j_begin;
  float* const acc_out = pr.acc + (long long)split * p.acc_stride;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < VS; ++s) {
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), F_EPI_WARPS / 2);            // the group of 8 warps that owns this score buffer
    }
    for (int b = 0; b < PB; ++b) {
      mbar_init(p_full(b), F_EPI_WARPS / 2);
      mbar_init(p_empty(b), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(a_full, TILE_BYTES);
#pragma unroll
      for (int ka = 0; ka < NKA; ++ka) tma_load_2d(a_base + ka * F_ATOM_BYTES, tmA, a_full, ka * 64, m0);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nj; ++j) {
        mbar_wait(v_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(v_full(s), TILE_BYTES);
#pragma unroll
        for (int ka = 0; ka < NKA; ++ka)
          tma_load_2d(v_base(s) + ka * F_ATOM_BYTES, tmB, v_full(s), ka * 64, (j_begin + j) * FB);
        if (++s == VS) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(FB, FB, 0u, 0u);      // S = A . B^T, both K-major
      constexpr uint32_t idesc2 = umma_idesc_bf16(FB, D, 0u, 1u);       // acc += P . B, B read MN-major
      auto mma1 = [&](int j) {                 // score tile j -> S buffer j & 1
        const int s = j % VS, b = j & 1;
        mbar_wait(v_full(s), (uint32_t)((j / VS) & 1));
        mbar_wait(s_empty(b), (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * FB;
#pragma unroll
        for (int ka = 0; ka < NKA; ++ka)
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k) {
            const uint64_t ad = umma_smem_desc(a_base + ka * F_ATOM_BYTES + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bd = umma_smem_desc(v_base(s) + ka * F_ATOM_BYTES + k * (UMMA_K * 2), 0, 1024);
            umma_bf16(d_tmem, ad, bd, idesc1, (ka > 0 || k > 0) ? 1u : 0u);
          }
        umma_commit(s_full(b));
      };
      mbar_wait(a_full, 0);
      tc_fence_after();
      mma1(0);
      for (int j = 0; j < nj; ++j) {
        if (j + 1 < nj) mma1(j + 1);           // keeps the tensor pipe busy while the epilogue works on tile j
        const int s = j % VS, pb = j % PB;
        mbar_wait(p_full(pb), (uint32_t)((j / PB) & 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + F_ACC_COL;
#pragma unroll
        for (int pa = 0; pa < 2; ++pa)
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k) {
            // A = P, K-major: 32 bytes per K step inside the swizzled 128-byte row
            const uint64_t ad = umma_smem_desc(p_base(pb) + pa * F_ATOM_BYTES + k * (UMMA_K * 2), 0, 1024);
            // B = B_j read MN-major: K = the tile's rows (16 rows = 2 KB per step), N = d, 64-wide atoms 16 KB apart
            const uint64_t bd = umma_smem_desc(v_base(s) + (pa * 64 + k * UMMA_K) * 128, F_ATOM_BYTES, 1024);
            umma_bf16(d_tmem, ad, bd, idesc2, (j > 0 || pa > 0 || k > 0) ? 1u : 0u);
          }
        umma_commit(p_empty(pb));              // P buffer and B tile are free once these MMAs retire
        umma_commit(v_empty(s));
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int ew = warp - 4;
    const int q = warp & 3;                   // TMEM lane quarter
    const int cg = ew >> 2;                   // column group of the final accumulator read-out (D / 4 columns each)
    const int grow = m0 + 32 * q + lane;      // this thread's row
    const bool row_ok = grow < pr.M;
    const float tau = expf(*p.t_dev);
    const float tau_l2 = tau * 1.4426950408889634f;
    float pos_sum = 0.f, relu_sum = 0.f, lg_sum = 0.f;
    const bool want_loss = pr.want_loss != 0;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;

    // Ping-pong: the 16 epilogue warps form two groups of 8 (4 lane quarters x 2 column halves of 64); group g takes
    // the tiles with (j & 1) == g, i.e. always score buffer g, so one group's TMEM loads, barrier round trips and
    // hand-over to MMA2 overlap the other group's arithmetic.  (Measured neutral against 16 warps in lock step,
    // r02p: the epilogue is issue-bound -- kept because it halves the warps arriving on each barrier.)
    const int grp = ew >> 3;
    const int ch = (ew >> 2) & 1;             // 64-column half of the 128-wide score tile = k-atom of P
    for (int j = grp; j < nj; j += 2) {
      const int b = j & 1, pb = j % PB;
      const int n0 = (j_begin + j) * FB;
      mbar_wait(s_full(b), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t t_addr = tmem_base + b * FB + ch * 64 + lane_off;
      // P tile: K-major SWIZZLE_128B; k-atom = ch, row = 32 q + lane, 16-byte pieces 2 c + h of chunk c.  The buffer
      // must be free (MMA2 of tile j - PB has retired) before the first chunk is WRITTEN (its arithmetic need not wait).
      const uint32_t row_base = p_base(pb) + ch * F_ATOM_BYTES + (32 * q + lane) * 128;
      uint32_t va[16], vb[16];
      // The epilogue is bound by instruction issue (halving the MUFU work, or letting two warp groups ping-pong,
      // changes nothing), so every tile takes the cheapest of four code paths:
      //   DIAG  only the tiles crossed by the positives (one or two per CTA) pay the per-element position test;
      //   LOSS  only the direction that reports the loss forms the softplus sums; the other one needs nothing but
      //         sigma(x) = 1 / (1 + 2^(-x log2 e)): four instructions per element.
      auto chunk = [&](uint32_t(&v)[16], const int c, auto diag_tag, auto loss_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value, LOSS = decltype(loss_tag)::value;
        const int col0 = n0 + ch * 64 + 16 * c;
        if (col0 + 16 > pr.N || m0 + FB > pr.M) {        // edge tile: rows / columns beyond the problem vanish
          const int nvalid = row_ok ? pr.N - col0 : 0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i >= nvalid) v[i] = __float_as_uint(kMaskedScore);
        }
        // the positive of this row sits at v[dj] if 0 <= dj < 16: it takes the negative-pair path like every other
        // element, P gets 0 there (compile-time indices only: a dynamically indexed register array would live
        // in local memory), and its contribution to the negative sums is taken back out below
        const int dj = (DIAG && row_ok) ? pr.row_offset + grow - col0 : -1;
        float s_d = 0.f;
        float dprod = 1.f;
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float sg[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float sv = __uint_as_float(v[i + h]);
            if constexpr (LOSS) {
              float d;
              neg_terms(sv, tau_l2, d, sg[h]);
              dprod *= d;
              relu_sum += fmaxf(sv, 0.f);
            } else {
              sg[h] = rcp_approx(1.f + ex2_approx(-sv * tau_l2));   // |x| <= tau: no overflow; masked -> 1 / inf = 0
            }
            if constexpr (DIAG) {
              if (i + h == dj) {
                s_d = sv;
                sg[h] = 0.f;
              }
            }
          }
          packed[i >> 1] = pack_bf16x2(sg[0], sg[1]);
        }
        if constexpr (LOSS) lg_sum += lg2_approx(dprod);
        if constexpr (DIAG) {
          if ((unsigned)dj < 16u) {
            const float x = s_d * tau;
            const float e = expf(-fabsf(x));
            const float rr = 1.f / (1.f + e);
            if constexpr (LOSS) {
              relu_sum -= fmaxf(s_d, 0.f);
              lg_sum -= log2f(1.f + ex2_approx(-fabsf(s_d * tau_l2)));
              pos_sum += fmaxf(-x, 0.f) + log1pf(e);           // softplus(-x)
            }
            if (pr.gdiag != nullptr) pr.gdiag[grow] = x >= 0.f ? -(e * rr) : -rr;   // -sigma(-x)
          }
        }
        if (c == 0) mbar_wait(p_empty(pb), (uint32_t)(((j / PB) & 1) ^ 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t piece = (uint32_t)(2 * c + h) ^ (uint32_t)(lane & 7);
          st_shared_v4(row_base + piece * 16, packed[4 * h], packed[4 * h + 1], packed[4 * h + 2], packed[4 * h + 3]);
        }
      };
      // four chunks of 16 columns in a rolled loop of two (two chunk bodies per code path in the binary: the
      // epilogue must stay inside the instruction cache), the TMEM load of the next chunk in flight meanwhile
      auto tile = [&](auto diag_tag, auto loss_tag) {
        tmem_ld_32x16(t_addr, va);
#pragma unroll 1
        for (int cp = 0; cp < 2; ++cp) {
          tmem_ld_wait();
          tmem_ld_32x16(t_addr + 16 * (2 * cp + 1), vb);
          chunk(va, 2 * cp, diag_tag, loss_tag);
          tmem_ld_wait();
          if (cp == 0) {
            tmem_ld_32x16(t_addr + 32, va);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(b));        // the score buffer may be overwritten by tile j + 2
          }
          chunk(vb, 2 * cp + 1, diag_tag, loss_tag);
        }
      };
      // does this tile hold positives of this CTA's rows?  (rows m0 .. m0 + 127 <-> columns row_offset + row)
      const bool has_diag = n0 < pr.row_offset + m0 + FB && n0 + FB > pr.row_offset + m0;
      if (want_loss) {
        if (has_diag) tile(std::true_type{}, std::true_type{});
        else tile(std::false_type{}, std::true_type{});
      } else {
        if (has_diag) tile(std::true_type{}, std::false_type{});
        else tile(std::false_type{}, std::false_type{});
      }
      fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(pb));
    }

    // ---- the gradient accumulator: D columns, D / 4 per column group
    mbar_wait(acc_full, 0);
    tc_fence_after();
    constexpr int CPW = D / 4;                              // 16, 32, 48 or 64 columns per warp
#pragma unroll
    for (int c = 0; c < CPW / 16; ++c) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + F_ACC_COL + cg * CPW + 16 * c + lane_off, r);
      tmem_ld_wait();
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(acc_out + (size_t)grow * D + cg * CPW + 16 * c);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          dst[k4] = make_float4(__uint_as_float(r[4 * k4]), __uint_as_float(r[4 * k4 + 1]),
                                __uint_as_float(r[4 * k4 + 2]), __uint_as_float(r[4 * k4 + 3]));
      }
    }
    pos_sum = warp_sum(pos_sum);
    relu_sum = warp_sum(relu_sum);
    lg_sum = warp_sum(lg_sum);
    if (lane == 0) {
      float* dst = p.partials + ((long long)blockIdx.x * F_EPI_WARPS + ew) * PARTIALS_PER_WARP;
      const bool on = pr.want_loss != 0;
      dst[0] = on ? pos_sum : 0.f;
      dst[1] = on ? relu_sum : 0.f;
      dst[2] = on ? lg_sum : 0.f;
      dst[3] = 0.f;
      __threadfence();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }

  // loss finalisation by the last CTA (same fixed-order fp64 reduction as jsd_gemm_kernel<MODE_FWD>)
  int* s_last = reinterpret_cast<int*>(smem_raw);
  double* sh = reinterpret_cast<double*>(smem_raw + 16);
  if (threadIdx.x == 0) {
    __threadfence();
    *s_last = (atomicAdd(p.ticket, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (*s_last) {
    __threadfence();
    const int n = (int)gridDim.x * F_EPI_WARPS;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += F_THREADS)
      for (int k = 0; k < 3; ++k) acc[k] += (double)__ldcg(p.partials + (size_t)i * PARTIALS_PER_WARP + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
      if (lane == 0) sh[k * 32 + warp] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double tot = 0.0;
      for (int w = 0; w < F_THREADS / 32; ++w) tot += sh[threadIdx.x * 32 + w];
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
      *p.ticket = 0;
    }
  }
}

}  // namespace jsd
