// Shared-memory-tiled Wishart likelihood (forward + analytic backward) for 56 <= D <= 207: one CTA per
// D x D problem, the Cholesky factor held as packed lower 8x8 tiles in shared memory (180 KB at D = 200).
//
// Reference semantics: fcest/models/likelihoods.py:126-206, :253-264 (math in likelihood.cu /
// likelihood_warp.cu; same bordered-system and Q'' = Sigma^-1 G - alpha (alpha^T G) identities).
//
// Why: the L2-scratch variant in likelihood.cu fetches every DMMA operand fragment from global memory and is
// L2-bandwidth bound (3.9 TFLOP/s at D = 200).  Here every operand of the O(D^3) phases comes from shared
// memory or registers:
//   P0  per time step: am = A o fmean, as = A o sqrt(fvar) into a per-CTA global slot (L2 resident), shared by
//       the S samples the CTA walks through.
//   P1  Sigma = G G^T: output-stationary 16 x 16 macro tiles in registers (6 per warp and pass), G streamed
//       through a double-buffered 8-column shared-memory panel built on the fly from am, as and eps.
//   P2  left-looking blocked Cholesky on the shared-memory tiles (bordered y row, unit pivot); diagonal tiles
//       factored + inverted in registers (lik_tiles.cuh); three block barriers per block column.
//   P4  per 8-column block of G (one warp each, no block barrier): right-looking block forward substitution
//       H = L^-1 G, row D negated, right-looking back substitution Q'' = L^-T H, all 26 row tiles of the column
//       block in registers (accumulator layout), layout changes by shuffles; column nu carries the unit vector
//       e_D whose solution is -alpha.  Per-sample cotangents accumulate in the CTA's global slot (U1, U2).
//   P5  lam_bar, ve; after the last sample one coalesced pass turns U1, U2 into mu_bar, v_bar, A_bar partials.
// Deterministic: every accumulation has a fixed owner and order.
#include <math.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"
#include "lik_tiles.cuh"

namespace fcest {

constexpr int LT_WARPS = 8;
constexpr int LT_MQ = 6;     // macro tiles per warp and SYRK pass
constexpr int LT_PP = 12;    // pitch of the 8-column G panel (4 mod 8: conflict-free fragment reads)

struct TilesLayout {
  int NB, Dp, Dm, NT;
  int off_pan, off_vec, total;  // doubles
};

// NB = row-block capacity of the instantiation (>= D / 8 + 1; rows > D are identity padding)
__host__ __device__ inline TilesLayout lt_layout(int NB) {
  TilesLayout L;
  L.NB = NB;
  L.Dp = 8 * L.NB;
  L.Dm = (L.Dp + 15) / 16 * 16;
  L.NT = L.NB * (L.NB + 1) / 2;
  L.off_pan = L.NT * 64;
  int pan = 2 * L.Dm * LT_PP;                       // two panel buffers (P1); reused as 8 x Dp row-sum partials (P4)
  if (pan < LT_WARPS * L.Dp) pan = LT_WARPS * L.Dp;
  L.off_vec = L.off_pan + pan;
  L.total = L.off_vec + 4 * L.Dp + 16;              // dinv, yv, spl, alpha, misc
  return L;
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
  const double* tiles; const double* am; const double* as_; const double* ep;
  int D, nu, colC, ycol, gid, tig, oA0, oA1, oT0, oT1;
};

// G tile (ib, cb) in accumulator layout; column ycol carries e_D
__device__ __forceinline__ void lt_load_g(const LtCtx& c, int ib, double (&g)[2]) {
  const int d = 8 * ib + c.gid;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = c.colC + e;
    double v = 0.0;
    if (d < c.D && k < c.nu) { const int r = d * c.nu + k; v = c.am[r] + c.as_[r] * __ldg(c.ep + r); }
    else if (d == c.D && k == c.ycol) v = 1.0;
    g[e] = v;
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
    double ee[T][2], gm[T][2], gs[T][2], u1[T][2], u2[T][2];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int d = 8 * (IB + t) + c.gid;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = c.colC + e;
        const bool ok = need_grad && d < c.D && k < c.nu;
        const int r = ok ? d * c.nu + k : 0;
        ee[t][e] = ok ? __ldg(c.ep + r) : 0.0;
        gm[t][e] = ok ? c.am[r] : 0.0;
        gs[t][e] = ok ? c.as_[r] : 0.0;
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
          rs += q * (gm[t][e] + gs[t][e] * ee[t][e]);
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
__global__ void __launch_bounds__(LT_WARPS * 32, 1)
wishart_loglik_tiles_kernel(const LikParams p) {
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu, N = p.N;
  const TilesLayout L = lt_layout(NB);
  const int Dp = L.Dp, Dm = L.Dm;
  double* tiles = sm;
  double* pan = sm + L.off_pan;
  double* dinv = sm + L.off_vec;
  double* yv = dinv + Dp;
  double* spl = yv + Dp;
  double* alpha = spl + Dp;
  double* misc = alpha + Dp;
  __shared__ int bad_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int yr = D & 7, yb = D >> 3;  // the bordered y row is row D = 8 yb + yr
  const double wS = 1.0 / p.S;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig);
  const int oT0 = swz(tig, gid), oT1 = swz(4 + tig, gid);
  const int oC = swz(gid, 2 * tig);
  // per-CTA global slot: am, as (A o fmean, A o sqrt(fvar)), U1, U2 (sample sums of Gbar, Gbar o eps)
  double* am = p.ws + (long long)blockIdx.x * p.ws_stride;
  double* as_ = am + R;
  double* U1 = as_ + R;
  double* U2 = U1 + R;
  const int nbm = (NB + 1) / 2, nmac = nbm * (nbm + 1) / 2;
  const int npan = (nu + 7) / 8;
  const int ycol = nu, ncb = nu / 8 + 1;  // column `nu` carries e_D (its solution is -alpha)
  const bool vec = (nu & 1) == 0;

  for (int e = tid; e < 2 * Dm * LT_PP; e += blockDim.x) pan[e] = 0.0;  // panel rows >= D stay zero
  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  double glam_acc = 0.0;  // thread d < D (blockDim = 256 >= D)

  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const double* fm = p.fmean + (long long)n * p.ldf;
    const double* fv = p.fvar + (long long)n * p.ldf;
    __syncthreads();
    // ---------------- P0. per-time-step staging ----------------
    for (int r = tid; r < R; r += blockDim.x) {
      const double a = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + r % nu);
      am[r] = a * fm[r];
      as_[r] = a * sqrt(fv[r]);
    }
    for (int d = tid; d < Dp; d += blockDim.x) yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
    if (tid == 0) bad_s = 0;
    double ve_acc = 0.0;
    __syncthreads();

    for (int s = 0; s < p.S; ++s) {
      const double* ep = p.eps + ((long long)s * N + n) * R;
      // ---------------- P1. Sigma tiles = G G^T ----------------
      // panel = 8 columns of G; its raw ingredients are loaded into registers before the DMMAs of the current
      // panel and combined / stored to the other buffer after them (D * 8 / 256 <= 7 elements per thread)
      constexpr int PE = 7;
      double pa_[PE], ps_[PE], pe_[PE];
      auto panel_load = [&](int pi) {
        const int k0 = 8 * pi;
#pragma unroll
        for (int i = 0; i < PE; ++i) {
          const int e = tid + i * LT_WARPS * 32;
          const int d = e >> 3, k = k0 + (e & 7);
          const bool ok = d < D && k < nu;
          const int r = ok ? d * nu + k : 0;
          pa_[i] = ok ? am[r] : 0.0;
          ps_[i] = ok ? as_[r] : 0.0;
          pe_[i] = ok ? __ldg(ep + r) : 0.0;
        }
      };
      auto panel_store = [&](int buf) {
        double* dst = pan + buf * Dm * LT_PP;
#pragma unroll
        for (int i = 0; i < PE; ++i) {
          const int e = tid + i * LT_WARPS * 32;
          const int d = e >> 3;
          if (d < D) dst[d * LT_PP + (e & 7)] = pa_[i] + ps_[i] * pe_[i];
        }
      };
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
        __syncthreads();  // previous users of the panel buffers are done
        panel_load(0);
        panel_store(0);
        __syncthreads();
        for (int pi = 0; pi < npan; ++pi) {
          if (pi + 1 < npan) panel_load(pi + 1);
          const double* P = pan + (pi & 1) * Dm * LT_PP;
#pragma unroll
          for (int q = 0; q < LT_MQ; ++q) {
            if (!on[q]) continue;
            const double* pa = P + (16 * Iq[q] + gid) * LT_PP + tig;
            const double* pb = P + (16 * Jq[q] + gid) * LT_PP + tig;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double a0 = pa[4 * h], a1 = pa[8 * LT_PP + 4 * h], b0 = pb[4 * h], b1 = pb[8 * LT_PP + 4 * h];
              dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
              dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
              dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
              dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
            }
          }
          if (pi + 1 < npan) panel_store((pi + 1) & 1);
          __syncthreads();
        }
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
          double c[4][2];
          int ibs[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ibs[t] = jb + warp + t * LT_WARPS;
            if (ibs[t] < NB) {
              const double2 v = *reinterpret_cast<const double2*>(tiles + TI(ibs[t], jb) * 64 + oC);
              c[t][0] = v.x; c[t][1] = v.y;
            } else { c[t][0] = c[t][1] = 0.0; }
          }
          for (int kb = 0; kb < jb; ++kb) {
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
            if (ibs[t] < NB) *reinterpret_cast<double2*>(tiles + TI(ibs[t], jb) * 64 + oC) = make_double2(c[t][0], c[t][1]);
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
        c.tiles = tiles; c.am = am; c.as_ = as_; c.ep = ep; c.D = D; c.nu = nu; c.colC = 8 * cb + 2 * tig;
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
      const bool first = (n == (int)blockIdx.x);
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
  if (p.need_grad && tid < D) p.glam_part[(long long)blockIdx.x * D + tid] = glam_acc;
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

static int tiles_grid(int N) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return N < sms ? N : sms;
}

long long loglik_tiles_scratch_bytes(int D, int nu) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return (long long)sizeof(double) * 4LL * D * nu * sms;
}

template <int NB>
static int launch_tiles_nb(const LikParams& p, cudaStream_t stream, int grid) {
  const long long smem = (long long)lt_layout(NB).total * 8;
  auto kern = wishart_loglik_tiles_kernel<NB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, LT_WARPS * 32, smem, stream>>>(p);
  return 0;
}

int launch_loglik_tiles(LikParams p, cudaStream_t stream, int* part_rows) {
  if (!loglik_tiles_supported(p.D, p.nu)) return -1;
  if (p.ws == nullptr) return FCEST_ERR_ARGUMENT;
  const int grid = tiles_grid(p.N);
  p.ws_stride = 4LL * p.D * p.nu;
  int rc = -1;
  switch (tiles_capacity(p.D)) {
    case 10: rc = launch_tiles_nb<10>(p, stream, grid); break;
    case 14: rc = launch_tiles_nb<14>(p, stream, grid); break;
    case 18: rc = launch_tiles_nb<18>(p, stream, grid); break;
    case 22: rc = launch_tiles_nb<22>(p, stream, grid); break;
    case 26: rc = launch_tiles_nb<26>(p, stream, grid); break;
    default: return -1;
  }
  if (rc == 0) *part_rows = grid;
  return rc;
}

}  // namespace fcest
