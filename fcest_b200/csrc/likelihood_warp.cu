// Warp-per-matrix Wishart likelihood (forward + analytic backward) for D <= 55.
//
// Reference semantics: fcest/models/likelihoods.py:126-206, :253-264 (see likelihood.cu for the math).
//
// Mapping.  A persistent CTA owns one time step n at a time; warp s of the CTA owns MC sample s, i.e. ONE
// D x D problem per warp, with no block barrier inside the factorisation:
//   B1  Sigma = G G^T accumulates in registers (all lower 8x8 DMMA tiles of the warp's matrix); the G
//       fragments are built on the fly from the CTA-shared staging  am = A o fmean, as = A o sqrt(fvar)
//       and eps streamed straight from global memory (prefetched two k-steps ahead) - G is never stored.
//   B2  bordered system: row D of the padded matrix is set to y^T (unit pivot), so the Cholesky of the
//       padded tile grid yields z = L^-1 y as row D of L and -alpha^T = -(Sigma^-1 y)^T as row D of L^-1
//       for free (the padding rows of the last 8-row block would otherwise be wasted work).
//   B3  left-looking blocked Cholesky on register tiles; finished tiles live in swizzled 8x8 shared-memory
//       tiles; diagonal tiles are factored + inverted in registers with shuffles.
//   B4  Li = L^-1 row block by row block (independent DMMA chains across the block row), in place.
//   C   the CTA re-partitions: warp c owns COLUMN block c of G for all samples, so the sums over samples
//       (mu_bar, v_bar, A_bar) stay in registers: H = Li G[:, c], row D of H (= -alpha^T G) is negated so that
//       Q'' = Li^T H = Sigma^-1 G - alpha (alpha^T G) = -Gbar / w comes out of the DMMA chain directly.
//   D   lam_bar from the identity  M_dd Lam_d = -1/2 (1 - alpha_d y_d - sum_k Q''_dk G_dk)  (M = dl/dSigma).
// Everything is deterministic (fixed-order reductions, no atomics).
#include <math.h>
#include <stdlib.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"
#include "lik_tiles.cuh"

namespace fcest {

struct WarpLikLayout {
  int Dp, PA, nbN, K4, NT;
  int off_as, off_y, off_spl, off_warp, wstride;  // in doubles
};

__host__ __device__ inline WarpLikLayout wl_layout(int NB, int nu) {
  WarpLikLayout L;
  L.Dp = 8 * NB;
  L.K4 = (nu + 3) & ~3;
  L.nbN = (nu + 7) >> 3;
  L.PA = pitch4(8 * L.nbN);
  L.NT = NB * (NB + 1) / 2;
  L.off_as = L.Dp * L.PA;
  L.off_y = 2 * L.Dp * L.PA;
  L.off_spl = L.off_y + L.Dp;
  L.off_warp = L.off_spl + L.Dp;
  L.wstride = L.NT * 64 + NB * 64 + 3 * L.Dp + 8;  // tiles, staging, dinv, alpha, rs, misc
  return L;
}


template <int NB>
__global__ void __launch_bounds__(256, 1)
wishart_loglik_warp_kernel(const LikParams p) {
  constexpr int NT = NB * (NB + 1) / 2;
  constexpr int Dp = 8 * NB;
  extern __shared__ __align__(16) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu, N = p.N;
  const WarpLikLayout L = wl_layout(NB, nu);
  const int PA = L.PA, K4 = L.K4, nbN = L.nbN;
  double* am = sm;
  double* as_ = sm + L.off_as;
  double* yv = sm + L.off_y;
  double* spl = sm + L.off_spl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  double* tiles = sm + L.off_warp + warp * L.wstride;
  double* stg = tiles + NT * 64;
  double* dinv = stg + NB * 64;
  double* alpha = dinv + Dp;
  double* rsv = alpha + Dp;
  double* misc = rsv + Dp;
  const int yr = D & 7;  // the bordered y row is row D = 8 (NB - 1) + yr
  const double wS = 1.0 / p.S;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig);  // [gid][tig (+4)]
  const int oT0 = swz(tig, gid), oT1 = swz(4 + tig, gid);  // [tig (+4)][gid]
  const int oC = swz(gid, 2 * tig);                        // [gid][2 tig], [gid][2 tig + 1]

  for (int e = tid; e < 2 * Dp * PA; e += blockDim.x) sm[e] = 0.0;  // padding of am / as stays zero
  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  const int nchunk = (p.S + W - 1) / W;
  double glam_acc = 0.0;  // thread d < D: this CTA's partial of lam_bar[d] over its time steps (blockDim >= D)

  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const double* fm = p.fmean + (long long)n * p.ldf;
    const double* fv = p.fvar + (long long)n * p.ldf;
    double ve_acc = 0.0;
    int bad_acc = 0;
    for (int ch = 0; ch < nchunk; ++ch) {
      const int s0 = ch * W;
      const int Sc = min(W, p.S - s0);
      __syncthreads();
      if (ch == 0) {
        // A. stage am = A o fmean, as = A o sqrt(fvar), y  (shared by all samples of this time step);
        //    loads batched four deep (the rows were L2-prefetched one time step ahead)
        constexpr int PF = 4;
        for (int base = 0; base < R; base += PF * blockDim.x) {
          double f4[PF], v4[PF], a4[PF];
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const int r = base + u * blockDim.x + tid;
            if (r < R) {
              f4[u] = fm[r];
              v4[u] = fv[r];
              a4[u] = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + r % nu);
            }
          }
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const int r = base + u * blockDim.x + tid;
            if (r < R) {
              const int d = r / nu, k = r - d * nu;
              am[d * PA + k] = a4[u] * f4[u];
              as_[d * PA + k] = a4[u] * sqrt(v4[u]);
            }
          }
        }
        for (int d = tid; d < Dp; d += blockDim.x) yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
      }
      __syncthreads();

      if (warp < Sc) {
        // ---------------- B1. Sigma tiles = G G^T in registers ----------------
        double t[NT][2];
#pragma unroll
        for (int i = 0; i < NT; ++i) t[i][0] = t[i][1] = 0.0;
        const double* ep = p.eps + ((long long)(s0 + warp) * N + n) * R;
        double eA[NB], eB[NB], eC[NB], a_cur[NB], a_nxt[NB];
#define LD_E(k0, e)                                                                   \
  {                                                                                   \
    const int col_ = (k0) + tig;                                                      \
    _Pragma("unroll") for (int ib = 0; ib < NB; ++ib) {                               \
      const int row_ = 8 * ib + gid;                                                  \
      e[ib] = (row_ < D && col_ < nu) ? __ldg(ep + row_ * nu + col_) : 0.0;           \
    }                                                                                 \
  }
#define MK_A(k0, e, a)                                                                \
  {                                                                                   \
    _Pragma("unroll") for (int ib = 0; ib < NB; ++ib) {                               \
      const int o_ = (8 * ib + gid) * PA + (k0) + tig;                                \
      a[ib] = am[o_] + as_[o_] * e[ib];                                               \
    }                                                                                 \
  }
        LD_E(0, eC);
        LD_E(4, eA);
        LD_E(8, eB);
        MK_A(0, eC, a_cur);
        for (int k0 = 0; k0 < K4; k0 += 4) {
          LD_E(k0 + 12, eC);
          MK_A(k0 + 4, eA, a_nxt);
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) {
#pragma unroll
            for (int jb = 0; jb <= ib; ++jb) dmma(t[TI(ib, jb)][0], t[TI(ib, jb)][1], a_cur[ib], a_cur[jb]);
          }
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) { a_cur[ib] = a_nxt[ib]; eA[ib] = eB[ib]; eB[ib] = eC[ib]; }
        }
#undef LD_E
#undef MK_A
        // ---------------- B2. + diag(softplus(lam)), bordered y row, identity padding ----------------
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
#pragma unroll
          for (int jb = 0; jb <= ib; ++jb) {
            if (ib == jb || ib == NB - 1) {
              const int i = 8 * ib + gid;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = 8 * jb + 2 * tig + e;
                double v = t[TI(ib, jb)][e];
                if (i < D) { if (j == i) v += spl[i]; }
                else if (i == D) v = (j < D) ? yv[j] : (j == D ? 1.0 : 0.0);
                else v = (j == i) ? 1.0 : 0.0;
                t[TI(ib, jb)][e] = v;
              }
            }
          }
        }
        // ---------------- B3. left-looking blocked Cholesky ----------------
        int bad = 0;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
#pragma unroll
          for (int kb = 0; kb < jb; ++kb) {
            const double* Tj = tiles + TI(jb, kb) * 64;
            const double nb0 = -Tj[oA0], nb1 = -Tj[oA1];
            double a0[NB], a1[NB];
#pragma unroll
            for (int ib = jb; ib < NB; ++ib) a0[ib] = tiles[TI(ib, kb) * 64 + oA0];
#pragma unroll
            for (int ib = jb; ib < NB; ++ib) dmma(t[TI(ib, jb)][0], t[TI(ib, jb)][1], a0[ib], nb0);
#pragma unroll
            for (int ib = jb; ib < NB; ++ib) a1[ib] = tiles[TI(ib, kb) * 64 + oA1];
#pragma unroll
            for (int ib = jb; ib < NB; ++ib) dmma(t[TI(ib, jb)][0], t[TI(ib, jb)][1], a1[ib], nb1);
          }
#pragma unroll
          for (int ib = jb; ib < NB; ++ib)
            *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(t[TI(ib, jb)][0], t[TI(ib, jb)][1]);
          __syncwarp();
          bad = factor_diag_swz(tiles + TI(jb, jb) * 64, dinv + 8 * jb, lane, (jb == NB - 1) ? yr : -1, 8 * jb, D, bad);
          __syncwarp();
          if (jb < NB - 1) {
            const double* Ij = tiles + TI(jb, jb) * 64;
            const double b0 = Ij[oA0], b1 = Ij[oA1];
            double r[NB][2], x0[NB], x1[NB];
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib) { r[ib][0] = r[ib][1] = 0.0; x0[ib] = tiles[TI(ib, jb) * 64 + oA0]; }
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib) dmma(r[ib][0], r[ib][1], x0[ib], b0);
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib) x1[ib] = tiles[TI(ib, jb) * 64 + oA1];
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib) dmma(r[ib][0], r[ib][1], x1[ib], b1);
            __syncwarp();
#pragma unroll
            for (int ib = jb + 1; ib < NB; ++ib)
              *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(r[ib][0], r[ib][1]);
            __syncwarp();
          }
        }
        // ---------------- B4. Li = L^-1, block row by block row, in place ----------------
#pragma unroll
        for (int ib = 1; ib < NB; ++ib) {
          double x[NB][2];
#pragma unroll
          for (int jb = 0; jb < ib; ++jb) x[jb][0] = x[jb][1] = 0.0;
#pragma unroll
          for (int kb = 0; kb < ib; ++kb) {
            const double a0 = tiles[TI(ib, kb) * 64 + oA0], a1 = tiles[TI(ib, kb) * 64 + oA1];
            double b0[NB], b1[NB];
#pragma unroll
            for (int jb = 0; jb <= kb; ++jb) b0[jb] = tiles[TI(kb, jb) * 64 + oT0];
#pragma unroll
            for (int jb = 0; jb <= kb; ++jb) dmma(x[jb][0], x[jb][1], a0, b0[jb]);
#pragma unroll
            for (int jb = 0; jb <= kb; ++jb) b1[jb] = tiles[TI(kb, jb) * 64 + oT1];
#pragma unroll
            for (int jb = 0; jb <= kb; ++jb) dmma(x[jb][0], x[jb][1], a1, b1[jb]);
          }
          __syncwarp();
#pragma unroll
          for (int jb = 0; jb < ib; ++jb)
            *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(x[jb][0], x[jb][1]);
          __syncwarp();
          const double i0 = tiles[TI(ib, ib) * 64 + oA0], i1 = tiles[TI(ib, ib) * 64 + oA1];
          double r[NB][2], c0[NB], c1[NB];
#pragma unroll
          for (int jb = 0; jb < ib; ++jb) { r[jb][0] = r[jb][1] = 0.0; c0[jb] = tiles[TI(ib, jb) * 64 + oT0]; }
#pragma unroll
          for (int jb = 0; jb < ib; ++jb) dmma(r[jb][0], r[jb][1], i0, c0[jb]);
#pragma unroll
          for (int jb = 0; jb < ib; ++jb) c1[jb] = tiles[TI(ib, jb) * 64 + oT1];
#pragma unroll
          for (int jb = 0; jb < ib; ++jb) dmma(r[jb][0], r[jb][1], i1, c1[jb]);
          __syncwarp();
#pragma unroll
          for (int jb = 0; jb < ib; ++jb)
            *reinterpret_cast<double2*>(tiles + TI(ib, jb) * 64 + oC) = make_double2(-r[jb][0], -r[jb][1]);
          __syncwarp();
        }
        // ---------------- B5. alpha = Sigma^-1 y (= - row D of Li), quadratic form, log det ----------------
        for (int j = lane; j < Dp; j += 32)
          alpha[j] = (j < D) ? -tiles[TI(NB - 1, j >> 3) * 64 + swz(yr, j & 7)] : 0.0;
        __syncwarp();
        double q = 0.0, ld = 0.0;
        for (int j = lane; j < D; j += 32) { q += yv[j] * alpha[j]; ld += log(dinv[j]); }
        q = warp_sum(q);
        ld = warp_sum(ld);
        if (lane == 0) {
          misc[0] = -0.5 * D * 1.8378770664093453 + ld - 0.5 * q;
          misc[1] = (double)bad;
        }
      }
      __syncthreads();

      // next time step of this CTA: pull its noise block (one sample per warp) and its fmean / fvar rows into L2
      if (ch == nchunk - 1 && n + (int)gridDim.x < N) {
        const int nn = n + gridDim.x;
        for (int s2 = warp; s2 < min(W, p.S); s2 += W)
          prefetch_l2_warp(p.eps + ((long long)s2 * N + nn) * R, (long long)R * 8, lane);
        if (warp == W - 1) {
          prefetch_l2_warp(p.fmean + (long long)nn * p.ldf, (long long)R * 8, lane);
          prefetch_l2_warp(p.fvar + (long long)nn * p.ldf, (long long)R * 8, lane);
        }
      }

      // ---------------- C. column-block warps: Q'' = Sigma^-1 G - alpha (alpha^T G), sums over samples ------
      if (p.need_grad) {
        double rs[NB];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) rs[ib] = 0.0;
        for (int cb = warp; cb < nbN; cb += W) {
          double U1[NB][2], U2[NB][2];
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) U1[ib][0] = U1[ib][1] = U2[ib][0] = U2[ib][1] = 0.0;
          const int colB = 8 * cb + gid;
          const int colC = 8 * cb + 2 * tig;
          // B-type noise fragments of the first sample; the next sample's are fetched one iteration ahead
          double en0[NB], en1[NB];
#define LD_EB(sl_, e0_, e1_)                                                                       \
  {                                                                                                \
    const double* ep_ = p.eps + ((long long)(s0 + (sl_)) * N + n) * R;                             \
    _Pragma("unroll") for (int kb = 0; kb < NB; ++kb) {                                            \
      const int k_ = 8 * kb + tig;                                                                 \
      e0_[kb] = (k_ < D && colB < nu) ? __ldg(ep_ + k_ * nu + colB) : 0.0;                         \
      e1_[kb] = (k_ + 4 < D && colB < nu) ? __ldg(ep_ + (k_ + 4) * nu + colB) : 0.0;               \
    }                                                                                              \
  }
          LD_EB(0, en0, en1);
          for (int sl = 0; sl < Sc; ++sl) {
            const double* Ts = sm + L.off_warp + sl * L.wstride;
            const double* ep = p.eps + ((long long)(s0 + sl) * N + n) * R;
            double g0[NB], g1[NB], ec[NB][2];
#pragma unroll
            for (int kb = 0; kb < NB; ++kb) {
              const int k = 8 * kb + tig;
              g0[kb] = am[k * PA + colB] + as_[k * PA + colB] * en0[kb];
              g1[kb] = am[(k + 4) * PA + colB] + as_[(k + 4) * PA + colB] * en1[kb];
            }
            if (sl + 1 < Sc) LD_EB(sl + 1, en0, en1);
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) {
              const int d = 8 * ib + gid;
              ec[ib][0] = (d < D && colC < nu) ? __ldg(ep + d * nu + colC) : 0.0;
              ec[ib][1] = (d < D && colC + 1 < nu) ? __ldg(ep + d * nu + colC + 1) : 0.0;
            }
            double h[NB][2];
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) h[ib][0] = h[ib][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < NB; ++kb) {
              double a[NB];
#pragma unroll
              for (int ib = kb; ib < NB; ++ib) a[ib] = Ts[TI(ib, kb) * 64 + oA0];
#pragma unroll
              for (int ib = kb; ib < NB; ++ib) dmma(h[ib][0], h[ib][1], a[ib], g0[kb]);
#pragma unroll
              for (int ib = kb; ib < NB; ++ib) a[ib] = Ts[TI(ib, kb) * 64 + oA1];
#pragma unroll
              for (int ib = kb; ib < NB; ++ib) dmma(h[ib][0], h[ib][1], a[ib], g1[kb]);
            }
            if (gid == yr) { h[NB - 1][0] = -h[NB - 1][0]; h[NB - 1][1] = -h[NB - 1][1]; }  // row D: -c -> +c
            __syncwarp();
#pragma unroll
            for (int ib = 0; ib < NB; ++ib)
              *reinterpret_cast<double2*>(stg + ib * 64 + oC) = make_double2(h[ib][0], h[ib][1]);
            __syncwarp();
            double qq[NB][2];
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) qq[ib][0] = qq[ib][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < NB; ++kb) {
              const double b0 = stg[kb * 64 + oT0], b1 = stg[kb * 64 + oT1];
              double a[NB];
#pragma unroll
              for (int ib = 0; ib <= kb; ++ib) a[ib] = Ts[TI(kb, ib) * 64 + oT0];
#pragma unroll
              for (int ib = 0; ib <= kb; ++ib) dmma(qq[ib][0], qq[ib][1], a[ib], b0);
#pragma unroll
              for (int ib = 0; ib <= kb; ++ib) a[ib] = Ts[TI(kb, ib) * 64 + oT1];
#pragma unroll
              for (int ib = 0; ib <= kb; ++ib) dmma(qq[ib][0], qq[ib][1], a[ib], b1);
            }
#pragma unroll
            for (int ib = 0; ib < NB; ++ib) {
              const int d = 8 * ib + gid;
              if (d < D) {
                const int o = d * PA + colC;
                const double2 m2 = *reinterpret_cast<const double2*>(am + o);
                const double2 s2 = *reinterpret_cast<const double2*>(as_ + o);
                const double ga = m2.x + s2.x * ec[ib][0], gb = m2.y + s2.y * ec[ib][1];
                rs[ib] += qq[ib][0] * ga + qq[ib][1] * gb;
                const double gb0 = -wS * qq[ib][0], gb1 = -wS * qq[ib][1];
                U1[ib][0] += gb0; U2[ib][0] += gb0 * ec[ib][0];
                U1[ib][1] += gb1; U2[ib][1] += gb1 * ec[ib][1];
              }
            }
          }
#undef LD_EB
          // outputs for this column block: mu_bar = a U1, v_bar = a U2 / (2 sd), A_bar part = fm U1 + sd U2
          // (loads batched first; sd = fv rsqrt(fv), 1 / (2 sd) = rsqrt(fv) / 2)
          // A_bar partials are accumulated per CTA (row blockIdx.x of the workspace): read-modify-write by the
          // same lane every time step, so the order of additions is fixed
          const bool first = (n == (int)blockIdx.x) && ch == 0;
          double oa_[NB][2], of_[NB][2], ov_[NB][2], og_[NB][2];
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) {
            const int d = 8 * ib + gid;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int k = colC + e;
              const bool ok = d < D && k < nu;
              const int r = d * nu + k;
              oa_[ib][e] = ok ? (p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + k)) : 0.0;
              of_[ib][e] = ok ? fm[r] : 0.0;
              ov_[ib][e] = ok ? fv[r] : 1.0;
              og_[ib][e] = (ok && !first) ? p.gA_part[(long long)blockIdx.x * R + r] : 0.0;
            }
          }
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) {
            const int d = 8 * ib + gid;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int k = colC + e;
              if (d < D && k < nu) {
                const int r = d * nu + k;
                const double rsq = rsqrt(ov_[ib][e]), sd = ov_[ib][e] * rsq;
                const double vm = oa_[ib][e] * U1[ib][e], vv = oa_[ib][e] * U2[ib][e] * (0.5 * rsq);
                const double va = of_[ib][e] * U1[ib][e] + sd * U2[ib][e];
                const long long o = (long long)n * p.ldg + r;
                p.gA_part[(long long)blockIdx.x * R + r] = og_[ib][e] + va;
                if (ch == 0) { p.g_fmean[o] = vm; p.g_fvar[o] = vv; }
                else { p.g_fmean[o] += vm; p.g_fvar[o] += vv; }
              }
            }
          }
        }
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
          double v = rs[ib];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (tig == 0) rsv[8 * ib + gid] = v;
        }
      }
      __syncthreads();

      // ---------------- D. lam_bar, ve, info ----------------
      if (p.need_grad) {
        for (int d = tid; d < D; d += blockDim.x) {  // one trip: blockDim >= 32 and the launcher keeps D <= blockDim
          double sa = 0.0, sr = 0.0;
          for (int sl = 0; sl < Sc; ++sl) sa += (sm + L.off_warp + sl * L.wstride + NT * 64 + NB * 64 + Dp)[d];
          for (int w2 = 0; w2 < W; ++w2) sr += (sm + L.off_warp + w2 * L.wstride + NT * 64 + NB * 64 + 2 * Dp)[d];
          const double v = sigmoid_d(p.lam[d]) * wS * (-0.5) * ((double)Sc - sa * yv[d] - sr) / spl[d];
          glam_acc += v;
        }
      }
      if (tid == 0) {
        for (int sl = 0; sl < Sc; ++sl) {
          const double* ms = sm + L.off_warp + sl * L.wstride + NT * 64 + NB * 64 + 3 * Dp;
          ve_acc += wS * ms[0];
          const int b = (int)ms[1];
          if (b && !bad_acc) bad_acc = b;
        }
        if (ch == nchunk - 1) {
          p.ve[n] = bad_acc ? nan("") : ve_acc;
          p.info[n] = bad_acc;
        }
      }
    }
  }
  if (p.need_grad && tid < D) p.glam_part[(long long)blockIdx.x * D + tid] = glam_acc;
}


static constexpr int kWarpLikMaxNB = 7;
static constexpr long long kSmemLimit = 227 * 1024;

static long long wl_smem_bytes(int NB, int nu, int W) {
  const WarpLikLayout L = wl_layout(NB, nu);
  return (long long)sizeof(double) * ((long long)L.off_warp + (long long)W * L.wstride);
}

bool loglik_warp_supported(int D, int nu) {
  const int NB = D / 8 + 1;
  if (D < 1 || nu < 1 || NB > kWarpLikMaxNB) return false;
  return wl_smem_bytes(NB, nu, 1) <= kSmemLimit;
}

template <int NB>
static int launch_nb(const LikParams& p, int W, long long smem, cudaStream_t stream, int* part_rows) {
  auto kern = wishart_loglik_warp_kernel<NB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, (size_t)smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) occ = 1;
  long long grid = (long long)sms * occ;
  if (grid > p.N) grid = p.N;
  kern<<<(int)grid, 32 * W, smem, stream>>>(p);
  *part_rows = (int)grid;
  return 0;
}

int launch_loglik_warp(const LikParams& p, cudaStream_t stream, int* part_rows) {
  if (!loglik_warp_supported(p.D, p.nu)) return -1;
  const int NB = p.D / 8 + 1;
  int W = p.S < 8 ? p.S : 8;
  static const int forced_w = getenv("FCEST_LIK_WARPS") ? atoi(getenv("FCEST_LIK_WARPS")) : 0;  // A/B runs
  if (forced_w > 0 && forced_w < W) W = forced_w;
  if (p.D > 32 * W) W = (p.D + 31) / 32;  // phase D keeps one lam_bar element per thread
  while (W > 1 && wl_smem_bytes(NB, p.nu, W) > kSmemLimit) --W;
  const long long smem = wl_smem_bytes(NB, p.nu, W);
  switch (NB) {
    case 1: return launch_nb<1>(p, W, smem, stream, part_rows);
    case 2: return launch_nb<2>(p, W, smem, stream, part_rows);
    case 3: return launch_nb<3>(p, W, smem, stream, part_rows);
    case 4: return launch_nb<4>(p, W, smem, stream, part_rows);
    case 5: return launch_nb<5>(p, W, smem, stream, part_rows);
    case 6: return launch_nb<6>(p, W, smem, stream, part_rows);
    case 7: return launch_nb<7>(p, W, smem, stream, part_rows);
    default: return -1;
  }
}

}  // namespace fcest
