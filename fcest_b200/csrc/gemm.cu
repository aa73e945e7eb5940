// fp64 tensor-core GEMM for sm_100a: C = alpha * op(A) op(B) + beta * C (+ bias), DMMA.8x8x4.
//
// This is the dominant kernel of the SVGP/VGP projection (see DESIGN.md "packed-quadratic-form
// projection"): the three large products  V = Pk * Sk^T,  dSk = g^T * Pk,  T = g * Sk  and all the
// mean / chain-rule products run through it.  fp64 has no tcgen05/TMEM form on Blackwell; the only
// fp64 tensor path is mma.sync.m8n8k4 (SASS DMMA.8x8x4, measured 37.0 TFLOP/s issue peak on B200),
// so the design is a classic multistage cp.async pipeline feeding register-tiled warp MMAs:
//   CTA tile BM x BN x 16, 4-stage cp.async ring in shared memory (zero-filled at the edges),
//   warp tile WM x WN of 8x8 DMMA tiles held in registers, conflict-free padded smem layouts
//   (pitch = 4 mod 16 doubles) for all four operand layouts.
// Operand layouts ("K-major" = contraction index contiguous):
//   A_KMAJOR : A[m*lda + k]   else A[k*lda + m]
//   B_KMAJOR : B[n*ldb + k]   else B[k*ldb + n]
// Split-K (gridDim.z) writes partial tiles to a workspace which splitk_reduce_kernel folds
// deterministically (fixed summation order) together with alpha/beta/bias.
#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

struct GemmParams {
  int M, N, K;
  const double* A; long long lda, strideA;
  const double* B; long long ldb, strideB;
  double* C; long long ldc, strideC;
  double alpha, beta;
  const double* bias_row;  // optional: C[m][n] += bias_row[m]
  const double* bias_col;  // optional: C[m][n] += bias_col[n]
  int splitk; int k_per_split;  // k_per_split is a multiple of BK
  double* ws;  // split-K partials [splitk][M][ldc]
};

template <int BMN, int BK, bool KMAJOR>
struct TileLayout {
  static constexpr int PITCH = KMAJOR ? (BK + 4) : (BMN + 4);
  static constexpr int ROWS = KMAJOR ? BMN : BK;
  static constexpr int DOUBLES = PITCH * ROWS;
};

// Load one BMN x BK operand tile with 16-byte cp.async; out-of-range elements are zero-filled.
template <int BMN, int BK, bool KMAJOR, int NTHREADS>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, long long ld,
                                          int mn0, int MN, int k0, int kend, int tid) {
  constexpr int CHUNKS = BMN * BK / 2;
  constexpr int PER = (CHUNKS + NTHREADS - 1) / NTHREADS;
  using L = TileLayout<BMN, BK, KMAJOR>;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = tid + i * NTHREADS;
    if (CHUNKS % NTHREADS != 0 && c >= CHUNKS) break;
    if (KMAJOR) {
      const int row = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      const int gm = mn0 + row, gk = k0 + kc;
      int bytes = (gm < MN) ? (kend - gk) * 8 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const double* src = bytes > 0 ? g + (long long)gm * ld + gk : g;
      cp_async16(s + row * L::PITCH + kc, src, bytes);
    } else {
      const int krow = c / (BMN / 2), mc = (c % (BMN / 2)) * 2;
      const int gk = k0 + krow, gm = mn0 + mc;
      int bytes = (gk < kend) ? (MN - gm) * 8 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const double* src = bytes > 0 ? g + (long long)gk * ld + gm : g;
      cp_async16(s + krow * L::PITCH + mc, src, bytes);
    }
  }
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, 1)
dgemm_dmma_kernel(const GemmParams p) {
  constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
  constexpr int TM = WM / 8, TN = WN / 8;
  constexpr int KS = BK / 4;  // k4-steps per tile (even)
  using LA = TileLayout<BM, BK, A_KMAJOR>;
  using LB = TileLayout<BN, BK, B_KMAJOR>;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * LA::DOUBLES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);
  // 1-D tile index, fastest along the dimension with fewer tiles: the ~148 CTAs resident at any time
  // then cover all of the short dimension and only a few tiles of the long one, so both operand
  // panels of a wave stay L2 resident (DRAM traffic ~ algorithmic instead of one re-read per tile row).
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  int tm, tn;
  if (tiles_m <= tiles_n) { tm = blockIdx.x % tiles_m; tn = blockIdx.x / tiles_m; }
  else { tn = blockIdx.x % tiles_n; tm = blockIdx.x / tiles_n; }
  const int m0 = tm * BM, n0 = tn * BN;

  int z = blockIdx.z;
  int kbeg = 0, kend = p.K;
  const double* A = p.A;
  const double* B = p.B;
  double* C = p.C;
  if (p.splitk > 1) {
    kbeg = z * p.k_per_split;
    kend = min(p.K, kbeg + p.k_per_split);
  } else {
    A += (long long)z * p.strideA;
    B += (long long)z * p.strideB;
    C += (long long)z * p.strideC;
  }
  const int ktiles = (kend - kbeg + BK - 1) / BK;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Fragment loads for k4-step kk of the tile held in ring slot `slot`.
  auto load_frags = [&](double (&a)[TM], double (&b)[TN], int slot, int kk) {
    const double* a_s = sA + slot * LA::DOUBLES;
    const double* b_s = sB + slot * LB::DOUBLES;
    const int k = kk * 4 + tig;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = wm * WM + 8 * i + gid;
      a[i] = A_KMAJOR ? a_s[r * LA::PITCH + k] : a_s[k * LA::PITCH + r];
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = wn * WN + 8 * j + gid;
      b[j] = B_KMAJOR ? b_s[c * LB::PITCH + k] : b_s[k * LB::PITCH + c];
    }
  };
  auto mma_step = [&](const double (&a)[TM], const double (&b)[TN]) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  };
  auto issue_tile = [&](int t) {
    if (t < ktiles) {
      const int s = t % STAGES;
      load_tile<BM, BK, A_KMAJOR, NTHREADS>(sA + s * LA::DOUBLES, A, p.lda, m0, p.M, kbeg + t * BK, kend, tid);
      load_tile<BN, BK, B_KMAJOR, NTHREADS>(sB + s * LB::DOUBLES, B, p.ldb, n0, p.N, kbeg + t * BK, kend, tid);
    }
    cp_async_commit();
  };

  // Software pipeline: the ring holds tiles kt+1 .. kt+STAGES in flight while tile kt is consumed.
  // The block barrier sits *before the last k4-step* of a tile, after that step's fragments are
  // already in registers, so the first fragments of the next tile are fetched (LDS latency) and the
  // barrier skew is absorbed underneath the last TM*TN DMMAs instead of stalling the tensor pipe.
#pragma unroll
  for (int s = 0; s < STAGES; ++s) issue_tile(s);
  cp_async_wait<STAGES - 1>();
  __syncthreads();
  double fa[2][TM], fb[2][TN];
  load_frags(fa[0], fb[0], 0, 0);

  for (int kt = 0; kt < ktiles; ++kt) {
    const int slot = kt % STAGES;
#pragma unroll
    for (int kk = 0; kk < KS - 1; ++kk) {
      load_frags(fa[(kk + 1) & 1], fb[(kk + 1) & 1], slot, kk + 1);
      mma_step(fa[kk & 1], fb[kk & 1]);
    }
    if (kt + 1 < ktiles) {
      cp_async_wait<STAGES - 2>();  // tile kt+1 has landed (kt+2, kt+3 may still be in flight)
      __syncthreads();              // ... for every thread; nobody reads ring slot `slot` any more
      load_frags(fa[0], fb[0], (kt + 1) % STAGES, 0);
      mma_step(fa[1], fb[1]);
      issue_tile(kt + STAGES);      // refill the slot that was just drained
    } else {
      mma_step(fa[1], fb[1]);
    }
  }
  cp_async_wait<0>();

  // epilogue
  const bool split = p.splitk > 1;
  double* out = split ? p.ws + (long long)z * p.M * p.ldc : C;
  const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + wm * WM + 8 * i + gid;
    if (row >= p.M) continue;
    const double br = (!split && p.bias_row) ? p.bias_row[row] : 0.0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + wn * WN + 8 * j + 2 * tig;
      if (col >= p.N) continue;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      double* dst = out + (long long)row * p.ldc + col;
      const bool has1 = col + 1 < p.N;
      if (!split) {
        v0 = p.alpha * v0 + br;
        v1 = p.alpha * v1 + br;
        if (p.bias_col) {
          v0 += p.bias_col[col];
          if (has1) v1 += p.bias_col[col + 1];
        }
        if (p.beta != 0.0) {
          v0 += p.beta * dst[0];
          if (has1) v1 += p.beta * dst[1];
        }
      }
      if (has1 && vec_ok) {
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      } else {
        dst[0] = v0;
        if (has1) dst[1] = v1;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const GemmParams p) {
  const long long total = (long long)p.M * p.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / p.N), col = (int)(idx % p.N);
    const long long off = (long long)row * p.ldc + col;
    double s = 0.0;
    for (int z = 0; z < p.splitk; ++z) s += p.ws[(long long)z * p.M * p.ldc + off];
    double v = p.alpha * s;
    if (p.bias_row) v += p.bias_row[row];
    if (p.bias_col) v += p.bias_col[col];
    if (p.beta != 0.0) v += p.beta * p.C[off];
    p.C[off] = v;
  }
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, bool AK, bool BKM>
static int launch_cfg(const GemmParams& p, int batch, cudaStream_t stream) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr size_t smem =
      sizeof(double) * STAGES * (TileLayout<BM, BK, AK>::DOUBLES + TileLayout<BN, BK, BKM>::DOUBLES);
  static_assert(smem <= 227 * 1024, "tile ring exceeds shared memory");
  auto kern = dgemm_dmma_kernel<BM, BN, WM, WN, BK, STAGES, AK, BKM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM), 1, p.splitk > 1 ? p.splitk : batch);
  kern<<<grid, NT, smem, stream>>>(p);
  FCEST_CHECK_LAUNCH();
  if (p.splitk > 1) {
    const long long total = (long long)p.M * p.N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p);
    FCEST_CHECK_LAUNCH();
  }
  return 0;
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES>
static int launch_layout(const GemmParams& p, int a_kmajor, int b_kmajor, int batch, cudaStream_t s) {
  if (a_kmajor && b_kmajor) return launch_cfg<BM, BN, WM, WN, BK, STAGES, true, true>(p, batch, s);
  if (a_kmajor && !b_kmajor) return launch_cfg<BM, BN, WM, WN, BK, STAGES, true, false>(p, batch, s);
  if (!a_kmajor && b_kmajor) return launch_cfg<BM, BN, WM, WN, BK, STAGES, false, true>(p, batch, s);
  return launch_cfg<BM, BN, WM, WN, BK, STAGES, false, false>(p, batch, s);
}

}  // namespace fcest

namespace fcest {
// Tile / split heuristic.
//  * "big" problems (>= one wave of 128x128 tiles) use a 12-warp CTA with a 120x128 or 128x128 tile,
//    whichever pads the M dimension less (N = 1200 time points is 10 x 120 exactly);
//  * small problems use 64x64 tiles;
//  * K is split only when the tile count leaves SMs idle (< 4 waves), because every split costs
//    2 * M * N * 8 bytes of partial-tile traffic plus a reduce launch.
struct GemmPlan { int cfg; int splitk; int bk; };  // cfg 0: 64x64x16, 1: 128x128x32, 2: 120x128x32
constexpr int BK_SMALL = 16, BK_BIG = 32;

static GemmPlan plan_gemm(int M, int N, int K, int batch) {
  const int sms = 148;
  GemmPlan pl;
  const bool big = ((long long)((M + 127) / 128) * ((N + 127) / 128) * batch >= sms);
  int bm = 64;
  pl.cfg = 0;
  if (big) {
    const long long pad128 = (long long)((M + 127) / 128) * 128, pad120 = (long long)((M + 119) / 120) * 120;
    pl.cfg = pad120 < pad128 ? 2 : 1;
    bm = pl.cfg == 2 ? 120 : 128;
  }
  const int bn = big ? 128 : 64;
  const int BK = big ? BK_BIG : BK_SMALL;
  pl.bk = BK;
  const long long tiles = (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn);
  int splitk = 1;
  if (batch == 1) {
    const int ktiles = (K + BK - 1) / BK;
    if (tiles < sms) {
      splitk = (int)((2 * sms + tiles - 1) / tiles);
    } else if (tiles < 4 * sms) {
      double best = 0;
      for (int s = 1; s <= 16; ++s) {
        const double waves = (double)tiles * s / sms;
        const double eff = waves / ceil(waves);
        if (eff > best + 0.02) { best = eff; splitk = s; }
      }
    }
    if (splitk > ktiles / 4) splitk = ktiles / 4;
    if (splitk < 1) splitk = 1;
  }
  pl.splitk = splitk;
  return pl;
}
}  // namespace fcest

#ifdef FCEST_GEMM_EXPERIMENTS
static int g_force_cfg = 0;
extern "C" void fcest_dgemm_force_cfg(int cfg) { g_force_cfg = cfg; }
#endif

extern "C" long long fcest_dgemm_workspace_bytes(int M, int N, int K, long long ldc, int batch) {
  const int splitk = fcest::plan_gemm(M, N, K, batch).splitk;
  return splitk > 1 ? (long long)splitk * M * ldc * 8 : 0;
}

extern "C" int fcest_dgemm(int a_kmajor, int b_kmajor, int M, int N, int K, double alpha,
                           const double* A, long long lda, long long strideA, const double* B,
                           long long ldb, long long strideB, double beta, double* C, long long ldc,
                           long long strideC, const double* bias_row, const double* bias_col,
                           int batch, double* workspace, long long workspace_bytes,
                           cudaStream_t stream) {
  using namespace fcest;
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15) || (strideA & 1) || (strideB & 1))
    return FCEST_ERR_ALIGNMENT;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.strideA = strideA;
  p.B = B; p.ldb = ldb; p.strideB = strideB;
  p.C = C; p.ldc = ldc; p.strideC = strideC;
  p.alpha = alpha; p.beta = beta; p.bias_row = bias_row; p.bias_col = bias_col;
  const GemmPlan pl = plan_gemm(M, N, K, batch);
  int splitk = pl.splitk;
  if (splitk > 1 && (workspace == nullptr || workspace_bytes < (long long)splitk * M * ldc * 8))
    splitk = 1;  // no workspace supplied: run unsplit (still correct, just fewer CTAs)
  p.splitk = splitk;
  const int ktiles = (K + pl.bk - 1) / pl.bk;
  p.k_per_split = ((ktiles + splitk - 1) / splitk) * pl.bk;
  p.ws = workspace;
#ifdef FCEST_GEMM_EXPERIMENTS
  if (g_force_cfg == 3) return launch_layout<120, 128, 40, 32, 16, 4>(p, a_kmajor, b_kmajor, batch, stream);
  if (g_force_cfg == 4) return launch_layout<128, 128, 32, 32, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
  if (g_force_cfg == 5) return launch_layout<120, 128, 40, 16, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
  if (g_force_cfg == 6) return launch_layout<128, 128, 64, 32, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
  if (g_force_cfg == 7) return launch_layout<120, 128, 40, 64, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
#endif
  if (pl.cfg == 2) return launch_layout<120, 128, 40, 32, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
  if (pl.cfg == 1) return launch_layout<128, 128, 32, 32, BK_BIG, 3>(p, a_kmajor, b_kmajor, batch, stream);
  return launch_layout<64, 64, 32, 32, BK_SMALL, 4>(p, a_kmajor, b_kmajor, batch, stream);
}
