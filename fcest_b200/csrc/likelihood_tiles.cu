// Shared-memory-tiled Wishart likelihood (forward + analytic backward) for 56 <= D <= 207: one CTA per
// D x D problem, the Cholesky factor held as packed lower 8x8 tiles in shared memory (180 KB at D = 200).
//
// Reference semantics: fcest/models/likelihoods.py:126-206, :253-264 (math in likelihood.cu /
// likelihood_warp.cu; same bordered-system and Q'' = Sigma^-1 G - alpha (alpha^T G) identities).
//
// Why: the L2-scratch variant in likelihood.cu fetches every DMMA operand fragment from global memory and is
// L2-bandwidth bound (3.9 TFLOP/s at D = 200).  Here every operand of the O(D^3) phases comes from shared
// memory or registers:
//   P0  (separate streaming kernel, per chunk of time steps) G = A o (fmean + sqrt(fvar) o eps) for all samples of the
//       chunk, written once at HBM speed (`wishart_g_fill_kernel`); filling G inside the persistent kernel cost 30 % of
//       its time (148 CTAs x 1.6 MB of per-CTA state thrash the L2 and the loads are latency-exposed).
//   P1  Sigma = G G^T: output-stationary 16 x 16 macro tiles in registers (6 per warp and pass); G arrives as
//       4-column panels (one DMMA k-step; dense 32-byte rows are conflict-free for the fragment reads) through a ring
//       of six TMA tensor-map copies (`cp.async.bulk.tensor.3d`, SASS UTMALDG; rows >= D zero-filled by the copy
//       engine) with full / empty mbarriers: no registers and no block barrier in the k loop.
//   P2  left-looking blocked Cholesky on the shared-memory tiles (bordered y row, unit pivot); diagonal tiles
//       factored + inverted in registers (lik_tiles.cuh); three block barriers per block column.
//   P4  per 8-column block of G (one warp each, no block barrier): right-looking block forward substitution
//       H = L^-1 G, row D negated, right-looking back substitution Q'' = L^-T H, all 26 row tiles of the column
//       block in registers (accumulator layout), layout changes by shuffles; column nu carries the unit vector
//       e_D whose solution is -alpha.  Per-sample cotangents accumulate in the CTA's global slot (U1, U2).
//   P5  lam_bar, ve; after the last sample one coalesced pass turns U1, U2 into mu_bar, v_bar, A_bar partials.
// Deterministic: every accumulation has a fixed owner and order.
#include <cuda.h>
#include <math.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"
#include "lik_tiles.cuh"

namespace fcest {

// Warps per CTA, by row-block capacity: as many as give the column-block phase (P4, latency bound: one warp per 8-column
// block of G) a balanced hand-out -- NB column blocks on NB / 2 or NB warps -- at 128 - 168 registers per thread.
// D = 200: 8 warps 36.1 ms, 10 warps 34.5 ms, 13 warps 31.6 ms.
#ifndef FCEST_LT_W26
#define FCEST_LT_W26 13
#endif
__host__ __device__ constexpr int lt_warps(int NB) { return NB == 26 ? FCEST_LT_W26 : NB == 22 ? 11 : NB == 18 ? 9 : NB == 14 ? 14 : 10; }
// macro tiles per warp and SYRK pass (16 accumulator registers each)
__host__ __device__ constexpr int lt_mq(int NB) {
  const int nbm = (NB + 1) / 2, nmac = nbm * (nbm + 1) / 2, w = lt_warps(NB);
  const int per = (nmac + w - 1) / w;
  return per > 4 ? (per + 1) / 2 > 4 ? 4 : (per + 1) / 2 : per;
}
constexpr int LT_PC = 4;     // columns of a G panel = one DMMA k-step (dense 32-byte rows: fragment reads are contiguous)
constexpr int LT_ST = 6;     // panel stages in the TMA ring

struct TilesLayout {
  int NB, Dp, Dm, NT;
  int off_pan, off_vec, off_bar, total;  // doubles
};

// NB = row-block capacity of the instantiation (>= D / 8 + 1; rows > D are identity padding)
__host__ __device__ constexpr TilesLayout lt_layout(int NB) {
  TilesLayout L = {};
  L.NB = NB;
  L.Dp = 8 * L.NB;
  L.Dm = (L.Dp + 15) / 16 * 16;
  L.NT = L.NB * (L.NB + 1) / 2;
  L.off_pan = L.NT * 64;
  int pan = LT_ST * L.Dm * LT_PC;                   // panel ring (P1); reused as 8 x Dp row-sum partials (P4)
  if (pan < lt_warps(NB) * L.Dp) pan = lt_warps(NB) * L.Dp;
  L.off_vec = L.off_pan + pan;
  L.off_bar = L.off_vec + 4 * L.Dp + 16;            // dinv, yv, spl, alpha, misc
  L.total = L.off_bar + 2 * LT_ST;                  // full / empty mbarriers
  return L;
}

__device__ __forceinline__ void lt_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void lt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lt_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void lt_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA tensor-map copy of one panel (4 columns x Dm rows of this CTA's G slot) global -> shared (SASS UTMALDG)
__device__ __forceinline__ void lt_tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// accumulator-layout tile (c0 = [gid][2 tig], c1 = [gid][2 tig + 1]) -> B-operand fragments
// b0 = tile[tig][gid], b1 = tile[4 + tig][gid]
__device__ __forceinline__ void cvt_c_to_b(double c0, double c1, int gid, int tig, double& b0, double& b1) {
  const int src0 = 4 * tig + (gid >> 1), src1 = 4 * (tig + 4) + (gid >> 1);
  const double x0 = __shfl_sync(0xffffffffu, c0, src0), x1 = __shfl_sync(0xffffffffu, c1, src0);
  const double y0 = __shfl_sync(0xffffffffu, c0, src1), y1 = __shfl_sync(0xffffffffu, c1, src1);
  b0 = (gid & 1) ? x1 : x0;
  b1 = (gid & 1) ? y1 : y0;
}


// Context of one column-block solve (P4); passed by reference through the unrolled step templates.
struct LtCtx {
  const double* tiles; const double* G; const double* ep;   // G slot (row pitch ldgs), noise of this sample
  int D, nu, ldgs, colC, ycol, gid, tig, oA0, oA1, oT0, oT1;
};

// G tile (ib, cb) in accumulator layout; column ycol carries e_D
__device__ __forceinline__ void lt_load_g(const LtCtx& c, int ib, double (&g)[2]) {
  const int d = 8 * ib + c.gid;
  double2 v2 = make_double2(0.0, 0.0);
  if (d < c.D && c.colC < c.nu) v2 = *reinterpret_cast<const double2*>(c.G + d * c.ldgs + c.colC);   // one 16-byte load
  g[0] = v2.x;
  g[1] = (c.colC + 1 < c.nu) ? v2.y : 0.0;   // odd nu: the pad column of the slab is not written
  if (d == c.D) {
    if (c.colC == c.ycol) g[0] = 1.0;
    if (c.colC + 1 == c.ycol) g[1] = 1.0;
  }
}

// The loops over tiles are written as template recursion: nvcc stops honouring `#pragma unroll` on the
// 26 x 26 / 2 nest and would index the accumulator array dynamically (local memory).
template <int NB, int KB, int IB>
__device__ __forceinline__ void lt_trail_fwd(double (&acc)[NB][2], const double* tiles, int off, double b) {
  if constexpr (IB < NB) {
    dmma(acc[IB][0], acc[IB][1], tiles[TI(IB, KB) * 64 + off], b);
    lt_trail_fwd<NB, KB, IB + 1>(acc, tiles, off, b);
  }
}
template <int NB, int KB, int IB>
__device__ __forceinline__ void lt_trail_bwd(double (&acc)[NB][2], const double* tiles, int off, double b) {
  if constexpr (IB < KB) {
    dmma(acc[IB][0], acc[IB][1], tiles[TI(KB, IB) * 64 + off], b);
    lt_trail_bwd<NB, KB, IB + 1>(acc, tiles, off, b);
  }
}
constexpr int LT_PF = 3;  // G tiles in flight ahead of the forward substitution
// forward, right-looking: H[kb] = I_kb (G[kb] + acc[kb]); acc[ib] -= L[ib][kb] H[kb] (ib > kb)
template <int NB, int KB>
__device__ __forceinline__ void lt_fwd(double (&acc)[NB][2], double (&gq)[LT_PF][2], const LtCtx& c) {
  if constexpr (KB < NB) {
    double b0, b1;
    cvt_c_to_b(acc[KB][0] + gq[KB % LT_PF][0], acc[KB][1] + gq[KB % LT_PF][1], c.gid, c.tig, b0, b1);
    if constexpr (KB + LT_PF < NB) lt_load_g(c, KB + LT_PF, gq[KB % LT_PF]);
    const double* Ik = c.tiles + TI(KB, KB) * 64;
    double h0 = 0.0, h1 = 0.0;
    dmma(h0, h1, Ik[c.oA0], b0);
    dmma(h0, h1, Ik[c.oA1], b1);
    acc[KB][0] = h0; acc[KB][1] = h1;
    cvt_c_to_b(-h0, -h1, c.gid, c.tig, b0, b1);
    lt_trail_fwd<NB, KB, KB + 1>(acc, c.tiles, c.oA0, b0);
    lt_trail_fwd<NB, KB, KB + 1>(acc, c.tiles, c.oA1, b1);
    lt_fwd<NB, KB + 1>(acc, gq, c);
  }
}
// backward, right-looking: Q[kb] = I_kb^T acc[kb]; acc[ib] -= L[kb][ib]^T Q[kb] (ib < kb)
template <int NB, int KB>
__device__ __forceinline__ void lt_bwd(double (&acc)[NB][2], const LtCtx& c) {
  if constexpr (KB >= 0) {
    double b0, b1;
    cvt_c_to_b(acc[KB][0], acc[KB][1], c.gid, c.tig, b0, b1);
    const double* Ik = c.tiles + TI(KB, KB) * 64;
    double q0 = 0.0, q1 = 0.0;
    dmma(q0, q1, Ik[c.oT0], b0);
    dmma(q0, q1, Ik[c.oT1], b1);
    acc[KB][0] = q0; acc[KB][1] = q1;
    cvt_c_to_b(-q0, -q1, c.gid, c.tig, b0, b1);
    lt_trail_bwd<NB, KB, 0>(acc, c.tiles, c.oT0, b0);
    lt_trail_bwd<NB, KB, 0>(acc, c.tiles, c.oT1, b1);
    lt_bwd<NB, KB - 1>(acc, c);
  }
}
// epilogue of a column block, two row tiles at a time with every global load issued before the first use:
// alpha from the e_D column; row sums of Q'' o G; U1 += Gbar, U2 += Gbar o eps (this lane owns its elements)
template <int NB, int IB>
__device__ __forceinline__ void lt_epilogue(double (&acc)[NB][2], const LtCtx& c, double* U1, double* U2, double* alpha,
                                            double* rsw, double wS, bool need_grad, bool first_sample) {
  if constexpr (IB < NB) {
    constexpr int T = (IB + 1 < NB) ? 2 : 1;
    double ee[T][2], gg[T][2], u1[T][2], u2[T][2];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int d = 8 * (IB + t) + c.gid;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = c.colC + e;
        const bool ok = need_grad && d < c.D && k < c.nu;
        const int r = ok ? d * c.nu + k : 0;
        ee[t][e] = ok ? __ldg(c.ep + r) : 0.0;
        gg[t][e] = ok ? c.G[d * c.ldgs + k] : 0.0;
        u1[t][e] = (ok && !first_sample) ? U1[r] : 0.0;
        u2[t][e] = (ok && !first_sample) ? U2[r] : 0.0;
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int d = 8 * (IB + t) + c.gid;
      double rs = 0.0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = c.colC + e;
        const double q = acc[IB + t][e];
        if (d < c.D && k == c.ycol) alpha[d] = -q;
        if (need_grad && d < c.D && k < c.nu) {
          const int r = d * c.nu + k;
          rs += q * gg[t][e];
          const double gb = -wS * q;
          U1[r] = u1[t][e] + gb;
          U2[r] = u2[t][e] + gb * ee[t][e];
        }
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);  // all lanes take part (rows >= D contribute zero)
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      if (c.tig == 0 && d < c.D) rsw[d] += rs;
    }
    lt_epilogue<NB, IB + T>(acc, c, U1, U2, alpha, rsw, wS, need_grad, first_sample);
  }
}

template <int NB>
__global__ void __launch_bounds__(lt_warps(NB) * 32, 1)
wishart_loglik_tiles_kernel(const __grid_constant__ CUtensorMap tmG, const LikParams p, const int ldgs, const double* Gc,
                            const int n0, const int nc) {
  extern __shared__ __align__(128) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu, N = p.N;
  constexpr TilesLayout L = lt_layout(NB);
  constexpr int Dp = L.Dp, Dm = L.Dm;
  constexpr int STAGE = Dm * LT_PC;   // doubles per panel stage
  constexpr int LT_WARPS = lt_warps(NB), LT_MQ = lt_mq(NB);
  double* tiles = sm;
  double* pan = sm + L.off_pan;
  double* dinv = sm + L.off_vec;
  double* yv = dinv + Dp;
  double* spl = yv + Dp;
  double* alpha = spl + Dp;
  __shared__ int bad_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int yr = D & 7, yb = D >> 3;  // the bordered y row is row D = 8 yb + yr
  const double wS = 1.0 / p.S;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig);
  const int oT0 = swz(tig, gid), oT1 = swz(4 + tig, gid);
  const int oC = swz(gid, 2 * tig);
  // per-CTA global slot: U1, U2 (sample sums of Gbar, Gbar o eps); Gc: G of the chunk's time steps [n0, n0 + nc),
  // (S, nc, D) rows of pitch ldgs, filled by wishart_g_fill_kernel
  double* U1 = p.ws + (long long)blockIdx.x * p.ws_stride;
  double* U2 = U1 + R;
  const int nbm = (NB + 1) / 2, nmac = nbm * (nbm + 1) / 2;
  const int npan = (nu + LT_PC - 1) / LT_PC;
  const uint32_t pan_u32 = smem_u32(pan);
  const uint32_t full_bar = smem_u32(sm + L.off_bar), empty_bar = full_bar + 8 * LT_ST;
  uint32_t pbase = 0;   // panels consumed so far (kernel lifetime; uniform over the CTA)
  if (tid == 0) {
    for (int i = 0; i < LT_ST; ++i) { lt_mbar_init(full_bar + 8 * i, 1); lt_mbar_init(empty_bar + 8 * i, LT_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
  }
  // panel g (sequence number over the kernel's lifetime) into its ring stage; issued by thread 0 only
  auto issue_panel = [&](uint32_t g, int k0, int z) {
    const uint32_t stage = g % LT_ST, round = g / LT_ST;
    if (round > 0) lt_mbar_wait(empty_bar + 8 * stage, (round & 1u) ^ 1u);
    lt_mbar_expect_tx(full_bar + 8 * stage, (uint32_t)(STAGE * 8));
    lt_tma_load_3d(pan_u32 + stage * STAGE * 8, &tmG, full_bar + 8 * stage, k0, 0, z);
  };
  const int ycol = nu, ncb = nu / 8 + 1;  // column `nu` carries e_D (its solution is -alpha)

  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  double glam_acc = 0.0;  // thread d < D (blockDim = 32 lt_warps(NB) >= 8 NB > D)

  for (int n = n0 + blockIdx.x; n < n0 + nc; n += gridDim.x) {
    const double* fm = p.fmean + (long long)n * p.ldf;
    const double* fv = p.fvar + (long long)n * p.ldf;
    __syncthreads();
    for (int d = tid; d < Dp; d += blockDim.x) yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
    if (tid == 0) bad_s = 0;
    double ve_acc = 0.0;
    __syncthreads();

    for (int s = 0; s < p.S; ++s) {
      const double* ep = p.eps + ((long long)s * N + n) * R;
      const int zs = s * nc + (n - n0);                       // this problem's slab of the chunk's G
      const double* Gs = Gc + (long long)zs * D * ldgs;
      asm volatile("fence.proxy.async;" ::: "memory");        // generic writes to the ring (row-sum scratch) before the TMA writes
      __syncthreads();
      // ---------------- P1. Sigma tiles = G G^T ----------------
      for (int mbase = 0; mbase < nmac; mbase += LT_WARPS * LT_MQ) {
        int Iq[LT_MQ], Jq[LT_MQ];
        bool on[LT_MQ];
        double acc[LT_MQ][2][2][2];
#pragma unroll
        for (int q = 0; q < LT_MQ; ++q) {
          const int m = mbase + warp + q * LT_WARPS;
          on[q] = m < nmac;
          const int mm = on[q] ? m : 0;   // macro index -> (I, J), J <= I (closed form: no loop inside the unrolled q loop)
          int I = (int)((sqrtf(8.0f * (float)mm + 1.0f) - 1.0f) * 0.5f);
          if ((I + 1) * (I + 2) / 2 <= mm) ++I;
          if (I * (I + 1) / 2 > mm) --I;
          Iq[q] = I; Jq[q] = mm - I * (I + 1) / 2;
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int v = 0; v < 2; ++v) acc[q][u][v][0] = acc[q][u][v][1] = 0.0;
        }
        if (tid == 0) {
          for (int j = 0; j < LT_ST - 1 && j < npan; ++j) issue_panel(pbase + j, LT_PC * j, zs);
        }
        for (int pi = 0; pi < npan; ++pi) {
          const uint32_t g = pbase + pi, stage = g % LT_ST;
          lt_mbar_wait(full_bar + 8 * stage, (g / LT_ST) & 1u);
          // refill the stage every warp released one k-step ago
          if (tid == 0 && pi + LT_ST - 1 < npan) issue_panel(g + LT_ST - 1, LT_PC * (pi + LT_ST - 1), zs);
          const double* P = pan + stage * STAGE;
#pragma unroll
          for (int q = 0; q < LT_MQ; ++q) {
            if (!on[q]) continue;
            const double* pa = P + (16 * Iq[q] + gid) * LT_PC + tig;
            const double* pb = P + (16 * Jq[q] + gid) * LT_PC + tig;
            const double a0 = pa[0], a1 = pa[8 * LT_PC], b0 = pb[0], b1 = pb[8 * LT_PC];
            dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
            dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
            dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
            dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
          }
          __syncwarp();
          if (lane == 0) lt_mbar_arrive(empty_bar + 8 * stage);
        }
        pbase += (uint32_t)npan;
        // + diag(softplus(lam)), bordered y row, identity padding; store the lower tiles
#pragma unroll
        for (int q = 0; q < LT_MQ; ++q) {
          if (!on[q]) continue;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int ib = 2 * Iq[q] + u, jb = 2 * Jq[q] + v;
              if (ib >= NB || jb > ib) continue;
              const int i = 8 * ib + gid;
              double c[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = 8 * jb + 2 * tig + e;
                double x = acc[q][u][v][e];
                if (i < D) { if (j == i) x += spl[i]; }
                else if (i == D) x = (j < D) ? yv[j] : (j == D ? 1.0 : 0.0);
                else x = (j == i) ? 1.0 : 0.0;
                c[e] = x;
              }
              *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(c[0], c[1]);
            }
          }
        }
      }
      __syncthreads();

      // ---------------- P2. left-looking blocked Cholesky on the shared-memory tiles ----------------
      int bad = 0;
      for (int jb = 0; jb < NB; ++jb) {
        // (a) column jb: S[ib][jb] -= sum_{kb < jb} L[ib][kb] L[jb][kb]^T ; up to four tiles per warp in flight
        {
          double c[4][2], c2[4][2];   // two partial sums per tile (even / odd kb): half the dependent DMMA chain
          int ibs[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ibs[t] = jb + warp + t * LT_WARPS;
            c2[t][0] = c2[t][1] = 0.0;
            if (ibs[t] < NB) {
              const double2 v = *reinterpret_cast<const double2*>(tiles + TI(ibs[t], jb) * 64 + oC);
              c[t][0] = v.x; c[t][1] = v.y;
            } else { c[t][0] = c[t][1] = 0.0; }
          }
          int kb = 0;
          for (; kb + 1 < jb; kb += 2) {
            const double* Tj = tiles + TI(jb, kb) * 64;   // tiles (jb, kb), (jb, kb + 1) are adjacent
            const double nb0 = -Tj[oA0], nb1 = -Tj[oA1], nd0 = -Tj[64 + oA0], nd1 = -Tj[64 + oA1];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (ibs[t] < NB) {
                const double* Ti = tiles + TI(ibs[t], kb) * 64;
                dmma(c[t][0], c[t][1], Ti[oA0], nb0);
                dmma(c2[t][0], c2[t][1], Ti[64 + oA0], nd0);
              }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (ibs[t] < NB) {
                const double* Ti = tiles + TI(ibs[t], kb) * 64;
                dmma(c[t][0], c[t][1], Ti[oA1], nb1);
                dmma(c2[t][0], c2[t][1], Ti[64 + oA1], nd1);
              }
            }
          }
          if (kb < jb) {
            const double* Tj = tiles + TI(jb, kb) * 64;
            const double nb0 = -Tj[oA0], nb1 = -Tj[oA1];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (ibs[t] < NB) dmma(c[t][0], c[t][1], tiles[TI(ibs[t], kb) * 64 + oA0], nb0);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (ibs[t] < NB) dmma(c[t][0], c[t][1], tiles[TI(ibs[t], kb) * 64 + oA1], nb1);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (ibs[t] < NB)
              *reinterpret_cast<double2*>(tiles + TI(ibs[t], jb) * 64 + oC) = make_double2(c[t][0] + c2[t][0], c[t][1] + c2[t][1]);
        }
        __syncthreads();
        // (b) diagonal tile -> I_jj = L_jj^-1 (in place), dinv
        if (warp == 0) bad = factor_diag_swz(tiles + TI(jb, jb) * 64, dinv + 8 * jb, lane, (jb == yb) ? yr : -1, 8 * jb, D, bad);
        __syncthreads();
        // (c) panel: L[ib][jb] = S[ib][jb] I_jj^T
        {
          const double* Ij = tiles + TI(jb, jb) * 64;
          const double b0 = Ij[oA0], b1 = Ij[oA1];
          double r[4][2];
          int ibs[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ibs[t] = jb + 1 + warp + t * LT_WARPS;
            r[t][0] = r[t][1] = 0.0;
            if (ibs[t] < NB) {
              const double* X = tiles + TI(ibs[t], jb) * 64;
              const double x0 = X[oA0], x1 = X[oA1];
              dmma(r[t][0], r[t][1], x0, b0);
              dmma(r[t][0], r[t][1], x1, b1);
            }
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (ibs[t] < NB) *reinterpret_cast<double2*>(tiles + TI(ibs[t], jb) * 64 + oC) = make_double2(r[t][0], r[t][1]);
        }
        __syncthreads();
      }
      if (warp == 0 && lane == 0 && bad) bad_s = bad_s ? bad_s : bad;

      // ---------------- P4. column blocks: Q'' = L^-T (L^-1 G with row D negated) ----------------
      double* rsw = pan + warp * Dp;  // this warp's row sums of Q'' o G (the panel buffers are free now)
      for (int d = lane; d < Dp; d += 32) rsw[d] = 0.0;
      __syncwarp();
      // forward only: just the e_D column block (alpha -> quadratic form), on warp 0
      const int cb_first = p.need_grad ? warp : (warp == 0 ? ncb - 1 : ncb);
      for (int cb = cb_first; cb < ncb; cb += LT_WARPS) {
        LtCtx c;
        c.tiles = tiles; c.G = Gs; c.ep = ep; c.D = D; c.nu = nu; c.ldgs = ldgs; c.colC = 8 * cb + 2 * tig;
        c.ycol = ycol; c.gid = gid; c.tig = tig; c.oA0 = oA0; c.oA1 = oA1; c.oT0 = oT0; c.oT1 = oT1;
        double acc[NB][2], gq[LT_PF][2];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) acc[ib][0] = acc[ib][1] = 0.0;
#pragma unroll
        for (int t = 0; t < LT_PF; ++t) {
          gq[t][0] = gq[t][1] = 0.0;
          if (t < NB) lt_load_g(c, t, gq[t]);
        }
        lt_fwd<NB, 0>(acc, gq, c);
        // row D of H is -alpha^T G: negate it so that the back substitution folds in - alpha (alpha^T G)
        // (the e_D column keeps its sign: its solution is column D of L^-T = (-alpha; 1))
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
          if (ib == yb && gid == yr) {
            if (c.colC != ycol) acc[ib][0] = -acc[ib][0];
            if (c.colC + 1 != ycol) acc[ib][1] = -acc[ib][1];
          }
        }
        lt_bwd<NB, NB - 1>(acc, c);
        lt_epilogue<NB, 0>(acc, c, U1, U2, alpha, rsw, wS, p.need_grad != 0, s == 0);
      }
      __syncthreads();

      // ---------------- P5. lam_bar, ve of this sample ----------------
      if (p.need_grad && tid < D) {
        double sr = 0.0;
        for (int w2 = 0; w2 < LT_WARPS; ++w2) sr += pan[w2 * Dp + tid];
        glam_acc += sigmoid_d(p.lam[tid]) * wS * (-0.5) * (1.0 - alpha[tid] * yv[tid] - sr) / spl[tid];
      }
      if (warp == 0) {
        double q = 0.0, ld = 0.0;
        for (int j = lane; j < D; j += 32) { q += yv[j] * alpha[j]; ld += log(dinv[j]); }
        q = warp_sum(q);
        ld = warp_sum(ld);
        ve_acc += wS * (-0.5 * D * 1.8378770664093453 + ld - 0.5 * q);
      }
      __syncthreads();
    }  // samples

    // mu_bar = a U1, v_bar = a U2 / (2 sd), A_bar partial (per CTA) += fm U1 + sd U2
    if (p.need_grad) {
      const bool first = (n == (int)blockIdx.x);   // chunk 0, first time step of this CTA
      for (int r = tid; r < R; r += blockDim.x) {
        const double a = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + r % nu);
        const double f = fm[r], v = fv[r], rsq = rsqrt(v), sd = v * rsq;
        const double u1 = U1[r], u2 = U2[r];
        p.g_fmean[(long long)n * p.ldg + r] = a * u1;
        p.g_fvar[(long long)n * p.ldg + r] = a * u2 * (0.5 * rsq);
        const long long oa = (long long)blockIdx.x * R + r;
        p.gA_part[oa] = (first ? 0.0 : p.gA_part[oa]) + f * u1 + sd * u2;
      }
    }
    if (tid == 0) {
      p.ve[n] = bad_s ? nan("") : ve_acc;
      p.info[n] = bad_s;
    }
  }
  if (p.need_grad && tid < D) {   // per-CTA partial, accumulated over the chunks in launch order
    const long long o = (long long)blockIdx.x * D + tid;
    p.glam_part[o] = (n0 == 0 ? 0.0 : p.glam_part[o]) + glam_acc;
  }
}

static constexpr long long kSmemLimitT = 227 * 1024;

// row-block capacities instantiated (the padded system has 8 NB rows; rows > D are identity)
static int tiles_capacity(int D) {
  const int need = D / 8 + 1;
  const int caps[] = {10, 14, 18, 22, 26};
  for (int c : caps) if (need <= c) return c;
  return 0;
}

bool loglik_tiles_supported(int D, int nu) {
  if (D < 8 || nu < 1) return false;
  const int NB = tiles_capacity(D);
  if (!NB) return false;
  return (long long)lt_layout(NB).total * 8 + 64 <= kSmemLimitT;
}

// Scratch layout: [U1, U2 slots: 2 R doubles per CTA][G of one chunk of time steps: (S, nc, D) rows of pitch ldgs = nu
// rounded up to 4 (32-byte aligned rows for the tensor map)].  The chunk budget is 8 samples x one time step per SM; other
// sample counts get proportionally more / fewer time steps per chunk.
static int tiles_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return sms;
}
static long long tiles_gchunk_doubles(int D, int nu, int sms) { return 8LL * sms * D * ((nu + 3) & ~3); }

long long loglik_tiles_scratch_bytes(int D, int nu) {
  const int sms = tiles_sms();
  return (long long)sizeof(double) * (2LL * D * nu * sms + tiles_gchunk_doubles(D, nu, sms));
}

// G = A o (fmean + sqrt(fvar) o eps) for the samples / time steps of one chunk: slab z = s nc + (n - n0), D rows of pitch ldgs.
// One thread per element (d, k) of a time step: A o fmean and A o sqrt(fvar) once, then one fused multiply-add and one
// streaming load / store per sample (the square root per (sample, element) made the first version compute bound).
__global__ void __launch_bounds__(256) wishart_g_fill_kernel(const LikParams p, const int n0, const int nc, const int ldgs,
                                                             double* __restrict__ Gc) {
  const int R = p.D * p.nu;
  const int i = blockIdx.y, n = n0 + i;
  const double* fm = p.fmean + (long long)n * p.ldf;
  const double* fv = p.fvar + (long long)n * p.ldf;
  const long long slab = (long long)p.D * ldgs;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int r = (blockIdx.x * 2 + u) * 256 + threadIdx.x;
    if (r < R) {
      const int d = r / p.nu, k = r - d * p.nu;
      const double a = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + k);
      const double gm = a * fm[r], gs = a * sqrt(fv[r]);
      const double* ep = p.eps + (long long)n * R + r;
      double* G = Gc + (long long)i * slab + d * ldgs + k;
      int s = 0;
      for (; s + 4 <= p.S; s += 4) {   // four samples in flight
        double e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = __ldg(ep + (long long)(s + q) * p.N * R);
#pragma unroll
        for (int q = 0; q < 4; ++q) G[(long long)(s + q) * nc * slab] = gm + gs * e[q];
      }
      for (; s < p.S; ++s) G[(long long)s * nc * slab] = gm + gs * __ldg(ep + (long long)s * p.N * R);
    }
  }
}

int launch_loglik_g_fill(const LikParams& p, int n0, int nc, int ldgs, double* Gc, cudaStream_t stream) {
  const int R = p.D * p.nu;
  wishart_g_fill_kernel<<<dim3((R + 511) / 512, nc), 256, 0, stream>>>(p, n0, nc, ldgs, Gc);
  FCEST_CHECK_LAUNCH();
  return 0;
}

typedef CUresult (*LtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static LtEncodeTiledFn lt_encode_fn() {
  static LtEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (LtEncodeTiledFn)ptr;
  }
  return fn;
}

template <int NB>
static int launch_tiles_nb(const LikParams& p, cudaStream_t stream, int* part_rows) {
  constexpr TilesLayout L = lt_layout(NB);
  const long long smem = (long long)L.total * 8;
  auto kern = wishart_loglik_tiles_kernel<NB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int sms = tiles_sms();
  const int ldgs = (p.nu + 3) & ~3;
  const long long slab = (long long)p.D * ldgs;                               // doubles per (sample, time step)
  long long NC = tiles_gchunk_doubles(p.D, p.nu, sms) / (slab * p.S);          // time steps per chunk
  if (NC < 1) return FCEST_ERR_UNSUPPORTED;                                    // S > 8 sms samples
  if (NC > p.N) NC = p.N;
  const int grid = (int)(NC < sms ? NC : sms);                                 // the same CTAs (partial rows) in every chunk
  NC = NC / grid * grid;                                                       // whole rounds of the persistent CTAs per chunk
  double* Gc = p.ws + 2LL * p.D * p.nu * sms;
  // tensor map over the chunk's G: (nu columns, D rows, S x NC slabs); a box is one panel = 4 columns x Dm rows of one
  // slab, rows >= D and columns >= nu are zero-filled by the copy engine
  LtEncodeTiledFn enc = lt_encode_fn();
  if (!enc) return FCEST_ERR_UNSUPPORTED;
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)p.nu, (cuuint64_t)p.D, (cuuint64_t)(p.S * NC)};
  const cuuint64_t strides[2] = {(cuuint64_t)ldgs * 8, (cuuint64_t)slab * 8};
  const cuuint32_t box[3] = {(cuuint32_t)LT_PC, (cuuint32_t)L.Dm, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)Gc, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return FCEST_ERR_ARGUMENT;
  for (int n0 = 0; n0 < p.N; n0 += (int)NC) {
    const int nc = p.N - n0 < NC ? p.N - n0 : (int)NC;
    const int rc = launch_loglik_g_fill(p, n0, nc, ldgs, Gc, stream);
    if (rc) return rc;
    kern<<<grid, lt_warps(NB) * 32, smem, stream>>>(tm, p, ldgs, Gc, n0, nc);
    if (n0 + nc < p.N) FCEST_CHECK_LAUNCH();   // the caller counts / checks the last launch
  }
  *part_rows = grid;
  return 0;
}

int launch_loglik_tiles(LikParams p, cudaStream_t stream, int* part_rows) {
  if (!loglik_tiles_supported(p.D, p.nu)) return -1;
  if (p.ws == nullptr) return FCEST_ERR_ARGUMENT;
  p.ws_stride = 2LL * p.D * p.nu;
  switch (tiles_capacity(p.D)) {
    case 10: return launch_tiles_nb<10>(p, stream, part_rows);
    case 14: return launch_tiles_nb<14>(p, stream, part_rows);
    case 18: return launch_tiles_nb<18>(p, stream, part_rows);
    case 22: return launch_tiles_nb<22>(p, stream, part_rows);
    case 26: return launch_tiles_nb<26>(p, stream, part_rows);
    default: return -1;
  }
}

}  // namespace fcest
