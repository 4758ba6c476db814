// TEST INFRASTRUCTURE ONLY: the projection-head tail kernels under ThreadSanitizer.  Every CUDA thread of a block is an
// OS thread (cuda_emu.h), __syncthreads / shuffles are std::barrier rendezvous (acquire / release), so a shared-memory
// or global-memory access that is not ordered by a barrier -- what compute-sanitizer --tool racecheck reports on the
// GPU -- is a data race TSan reports here.  Built with -fsanitize=thread by tests/test_emu_kernels.py; exits non-zero
// when TSan saw a race (TSAN_OPTIONS=halt_on_error=1 exitcode=66) or a result is not finite.
#define EMU_HEADS_ONLY 1
#include "emu_kernels.cpp"

#include <random>
#include <string>

static std::vector<float> randn(size_t n, unsigned seed, float scale = 1.f, float shift = 0.f) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd;
  std::vector<float> v(n);
  for (auto& x : v) x = nd(rng) * scale + shift;
  return v;
}

static int run_case(long long rows, long long D, int n_slices, int blocks, bool dense, int self_check) {
  auto x0 = randn(rows * D, 1, 2.f, 0.5f), x1 = randn(rows * D, 2, 1.f, -1.f);
  auto w0 = randn(D, 3, 0.3f, 1.f), b0 = randn(D, 4, 0.2f), w1 = randn(D, 5, 0.3f, 1.f), b1 = randn(D, 6, 0.2f);
  std::vector<__nv_bfloat16> u(rows * D), v(rows * D);
  std::vector<float> st0(3 * rows), st1(3 * rows);
  if (emu_ln_normalize_pair(x0.data(), x1.data(), 0, rows, D, w0.data(), b0.data(), 1e-5f, w1.data(), b1.data(), 1e-5f, 1,
                            u.data(), v.data(), st0.data(), st1.data(), -1) < 0)
    return 1;
  const long long stride = rows * D;
  auto acc0 = randn(n_slices * stride, 7), acc1 = randn(n_slices * stride, 8);
  auto gdiag = randn(rows, 9);
  float t = 1.1f, gamma = 0.6f, dt = 0.f;
  std::vector<float> ws(2 * blocks * 2 * D), dx0(rows * D), dx1(rows * D), dw0(D), db0(D), dw1(D), db1(D), rowdot(rows);
  const int rc = emu_ln_normalize_bwd_pair(
      x0.data(), x1.data(), 0, rows, D, w0.data(), b0.data(), w1.data(), b1.data(), st0.data(), st1.data(), acc0.data(),
      acc1.data(), n_slices, stride, dense ? 1.f / (float)(rows * (rows - 1)) : 0.f, dense ? v.data() : nullptr, 0,
      dense ? u.data() : nullptr, 0, dense ? gdiag.data() : nullptr, &t, &gamma, rows, ws.data(), dx0.data(), dx1.data(),
      dw0.data(), db0.data(), dw1.data(), db1.data(), rowdot.data(), &dt, blocks, 0);
  if (rc < 0) return 2;
  double chk = dt;
  for (float f : dx0) chk += f;
  for (float f : dw1) chk += f;
  if (!std::isfinite(chk)) return 3;
  (void)self_check;
  return 0;
}

// negative control: a block reduction whose scratch is re-used WITHOUT the barrier in front of the second round --
// the kind of bug this check exists for.  `--racy` must make TSan fire (the test asserts that it does).
static void racy_kernel(float* out) {
  __shared__ float scratch[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int round = 0; round < 2; ++round) {
    if (lane == 0) scratch[warp] = (float)(threadIdx.x + round);
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
    if (threadIdx.x == 0) out[round] = s;
    // missing __syncthreads(): the next round's writes race with these reads
  }
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--racy") {
    float out[2] = {0.f, 0.f};
    emu::launch(emu::Dim3{1, 1, 1}, emu::Dim3{256, 1, 1}, [&]() { racy_kernel(out); });
    std::printf("racy kernel done (%g %g)\n", out[0], out[1]);
    return 0;
  }
  int rc = 0;
  rc |= run_case(9, 2048, 1, 3, false, 0);    // the heads' width: 256 threads, two 16-byte pieces per thread
  rc |= run_case(7, 256, 3, 2, true, 0);      // 64 threads, accumulator slices, positive-pair term
  rc |= run_case(5, 102, 1, 5, true, 0);      // element-wise variant
  rc |= run_case(4, 2176, 2, 1, true, 0);     // 4-piece instantiation, one block walks every row
  std::printf("racecheck cases done rc=%d\n", rc);
  return rc;
}
