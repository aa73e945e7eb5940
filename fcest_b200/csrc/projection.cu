// Packed-quadratic-form projection kernels for the whitened SVGP / VGP conditional.
//
// GPflow's base_conditional materialises LTA = tril(q_sqrt_r)^T P for every
// latent r, an (R, M, N) tensor.  Here the predictive variance is rewritten as a quadratic form
//     v[n, r] = c[n] + p_n^T S_r p_n ,   S_r = L_r L_r^T ,   p_n = P[:, n]
// and both sides are packed over the lower triangle (K = M(M+1)/2 entries, row-major (i,j), i>=j):
//     Sk[r, (i,j)] = S_r[i,j]                    (tri_syrk_pack_kernel)
//     Pk[n, (i,j)] = w_ij P[i,n] P[j,n], w=1|2   (pack_outer_kernel)
// so that  V = Pk Sk^T,  dELBO/dSk = g^T Pk  and  T = g Sk  are three *plain* fp64 GEMMs (gemm.cu)
// with the algorithmic flop count 3 M^2 N R and no recompute.  The kernels in this file are the
// (cheap) packing / unpacking steps around those GEMMs.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fcest_b200.h"

namespace fcest {

// ---------------------------------------------------------------------------------------------
// Pk[n, idx(i,j)] = w_ij * P[i,n] * P[j,n];  bias[n] = (use_prior ? var - sum_m P[m,n]^2 : 0)
// grid = N, block = 256, dynamic smem = M doubles (+32 for the reduction)
// ---------------------------------------------------------------------------------------------
__global__ void pack_outer_kernel(const double* __restrict__ P, long long ldp, int M, int N,
                                  double* __restrict__ Pk, long long ldk, double* __restrict__ bias,
                                  const double* __restrict__ kvar, int use_prior) {
  extern __shared__ double sm[];
  double* p = sm;
  double* red = sm + M;
  const int n = blockIdx.x;
  double sq = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double v = P[(long long)m * ldp + n];
    p[m] = v;
    sq += v * v;
  }
  sq = block_sum(sq, red);  // contains __syncthreads -> p[] visible
  if (threadIdx.x == 0) bias[n] = use_prior ? (kvar[0] - sq) : 0.0;
  const long long K = (long long)M * (M + 1) / 2;
  double* out = Pk + (long long)n * ldk;
  // thread t walks idx = t, t + blockDim, ...: (i, j) of the first element from one square root, then advanced
  // incrementally (a square root per element made this kernel compute bound: 0.09 ms for a 193 MB store)
  int i = (int)((sqrt(8.0 * (double)threadIdx.x + 1.0) - 1.0) * 0.5);
  while ((long long)(i + 1) * (i + 2) / 2 <= (long long)threadIdx.x) ++i;
  while ((long long)i * (i + 1) / 2 > (long long)threadIdx.x) --i;
  int j = (int)(threadIdx.x - (long long)i * (i + 1) / 2);
  for (long long idx = threadIdx.x; idx < ldk; idx += blockDim.x) {
    double v = 0.0;
    if (idx < K) {
      v = p[i] * p[j];
      if (i != j) v *= 2.0;
    }
    out[idx] = v;
    j += blockDim.x;
    while (j > i) { j -= i + 1; ++i; }   // next row(s) of the packed triangle
  }
}

// ---------------------------------------------------------------------------------------------
// Sk[r, idx(i,j)] = sum_{k<=j} L_r[i,k] L_r[j,k]  (L_r = tril(q_sqrt_r)), DMMA 8x8 tiles read
// straight from global/L1 (L_r is 8*M*M bytes, L1/L2 resident for the CTA that owns r).
// Also emits the two KL reductions per latent: klp[r,0] = ||tril L_r||_F^2 = tr(S_r),
// klp[r,1] = sum_m log(L_r[m,m]^2).                              grid = R, block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tri_syrk_pack_kernel(const double* __restrict__ q_sqrt, int M, double* __restrict__ Sk,
                     long long ldk, double* __restrict__ klp) {
  __shared__ double red[32];
  const int r = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r * M * M;
  double* __restrict__ out = Sk + (long long)r * ldk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nb = (M + 7) / 8;
  const int ntiles = nb * (nb + 1) / 2;
  double tr = 0.0;
  // heavier tiles (large jb) first so the tail is short; tiles are dealt round-robin to warps
  for (int t = ntiles - 1 - warp; t >= 0; t -= nwarps) {
    int ib = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
    while (ib * (ib + 1) / 2 > t) --ib;
    const int jb = t - ib * (ib + 1) / 2;
    const int ra = 8 * ib + gid, rb = 8 * jb + gid;
    const double* La = L + (long long)ra * M;
    const double* Lb = L + (long long)rb * M;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < 8 * (jb + 1); k0 += 4) {
      const int k = k0 + tig;
      const double a = (ra < M && k <= ra) ? La[k] : 0.0;
      const double b = (rb < M && k <= rb) ? Lb[k] : 0.0;
      dmma(c0, c1, a, b);
    }
    const int row = ra, col = 8 * jb + 2 * tig;
    if (row < M) {
      const long long base = (long long)row * (row + 1) / 2;
      if (col <= row) { out[base + col] = c0; if (col == row) tr += c0; }
      if (col + 1 <= row) { out[base + col + 1] = c1; if (col + 1 == row) tr += c1; }
    }
  }
  // zero the row padding [K, ldk)
  const long long K = (long long)M * (M + 1) / 2;
  for (long long idx = K + threadIdx.x; idx < ldk; idx += blockDim.x) out[idx] = 0.0;
  double ld = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double d = L[(long long)m * M + m];
    ld += log(d * d);
  }
  tr = block_sum(tr, red);
  ld = block_sum(ld, red);
  if (threadIdx.x == 0) { klp[2 * r] = tr; klp[2 * r + 1] = ld; }
}

// ---------------------------------------------------------------------------------------------
// dq_sqrt_r = tril( 2 W_r L_r - kl * L_r ) + kl * diag(1 / L_r[m,m]),
//   W_r symmetric, unpacked from Gk[r, idx(i,j)] = w_ij W_r[i,j]   (Gk = g_var^T Pk).
// This is d ELBO / d q_sqrt including the KL term (kl = 1) or the bare projection adjoint (kl = 0).
// grid = R, block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tri_grad_kernel(const double* __restrict__ Gk, long long ldk, const double* __restrict__ q_sqrt,
                int M, double* __restrict__ dq_sqrt, double kl) {
  const int r = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r * M * M;
  const double* __restrict__ G = Gk + (long long)r * ldk;
  double* __restrict__ out = dq_sqrt + (long long)r * M * M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nb = (M + 7) / 8;
  const int ntiles = nb * (nb + 1) / 2;
  // tile (ib, jb): cost ~ (nb - jb); enumerate column-major so that heavy (small jb) tiles come first
  for (int t = warp; t < ntiles; t += nwarps) {
    // column-major enumeration of the lower triangle: column jb holds (nb - jb) tiles
    int jb = 0, rem = t;
    while (rem >= nb - jb) { rem -= nb - jb; ++jb; }
    const int ib = jb + rem;
    const int ra = 8 * ib + gid;       // row of W held by this lane (A fragment)
    const int cb = 8 * jb + gid;       // column of L held by this lane (B fragment)
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 8 * jb; k0 < 8 * nb; k0 += 4) {
      const int k = k0 + tig;
      double a = 0.0, b = 0.0;
      if (ra < M && k < M) {
        a = (k <= ra) ? G[(long long)ra * (ra + 1) / 2 + k] : G[(long long)k * (k + 1) / 2 + ra];
        if (k != ra) a *= 0.5;
      }
      if (k < M && cb < M && cb <= k) b = L[(long long)k * M + cb];
      dmma(c0, c1, a, b);
    }
    const int row = ra, col = 8 * jb + 2 * tig;
    if (row < M) {
      if (col < M) {
        double v = 0.0;
        if (col <= row) {
          const double l = L[(long long)row * M + col];
          v = 2.0 * c0 - kl * l;
          if (col == row) v += kl / l;
        }
        out[(long long)row * M + col] = v;
      }
      if (col + 1 < M) {
        double v = 0.0;
        if (col + 1 <= row) {
          const double l = L[(long long)row * M + col + 1];
          v = 2.0 * c1 - kl * l;
          if (col + 1 == row) v += kl / l;
        }
        out[(long long)row * M + col + 1] = v;
      }
    }
  }
  // strictly-upper tiles are never visited: zero them
  for (long long e = threadIdx.x; e < (long long)M * M; e += blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    if ((j >> 3) > (i >> 3)) out[e] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Output-stationary variants of the two kernels above for M <= 200 (the SVGP default M = 200):
// one CTA (16 warps) per latent keeps ALL lower 8x8 output tiles of the M x M result in registers
// (pairs of horizontally adjacent tiles dealt cyclically to warps: balanced triangular work, the A
// fragment is shared inside a pair) and streams the operands through shared memory in k-panels of 24
// columns with a double-buffered cp.async pipeline, so q_sqrt / Gk are read from HBM exactly once.
//   tri_syrk : S[ib][jb]   = sum_{kb <= jb} L[ib][kb] L[jb][kb]^T
//   tri_grad : out[ib][jb] = sum_{kb >= jb} Gs[ib][kb] L[kb][jb]   (Gs = symmetric unpack of Gk)
//              dq_sqrt = tril(out + diag(Gk_ii) L - kl L) + kl diag(1/L_mm)      (2 W = Gs + diag(Gk_ii))
// ---------------------------------------------------------------------------------------------
constexpr int TRI_WARPS = 16;
#ifndef FCEST_TRI_KP
#define FCEST_TRI_KP 24
#endif
constexpr int TRI_KP = FCEST_TRI_KP;  // panel width (TRI_KB blocks of 8)
constexpr int TRI_KB = TRI_KP / 8;
constexpr int TRI_PITCH = TRI_KP + 4;
constexpr int TRI_MAXPAIRS = 11;    // ceil(169 / 16) pairs per warp at nb = 25

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// pair u (row-major over rows ib with (ib+2)/2 pairs each) -> (ib, jb0)
__device__ __forceinline__ int tri_pair_decode(int u) {
  // rows 2a and 2a+1 both have a+1 pairs; pairs before row ib: (ib/2)(ib/2+1) + (ib odd ? ib/2+1 : 0)
  int ib = (int)sqrtf((float)u) * 2;
  while (true) {
    const int h = ib >> 1;
    const int before = h * (h + 1) + ((ib & 1) ? h + 1 : 0);
    const int cnt = h + 1;
    if (u < before) { ib -= 1; continue; }
    if (u >= before + cnt) { ib += 1; continue; }
    return (ib << 8) | (2 * (u - before));
  }
}

// ADAM (with GRAD): instead of writing d ELBO / d q_sqrt, the epilogue applies the Keras-Adam update of run_adam to
// the lower triangle of q_sqrt in place (first / second moments am, av; step size from the device scalar
// alpha_dev): the 8 M^2 R byte gradient is never written or re-read and the optimiser touches half of q_sqrt, m, v.
struct TriAdam {
  double* m;
  double* v;
  const double* alpha_dev;
  double b1, b2, eps;
};

template <bool GRAD, bool ADAM = false>
__global__ void __launch_bounds__(TRI_WARPS * 32, 1)
tri_stationary_kernel(const double* __restrict__ q_sqrt, int M, const double* __restrict__ Gk,
                      double* __restrict__ out_packed, long long ldk, double* __restrict__ klp,
                      double* __restrict__ dq_sqrt, double kl, TriAdam ad = TriAdam()) {
  extern __shared__ __align__(16) double sm[];
  const int Mp = (M + 7) / 8 * 8, nb = Mp / 8;
  const int PL = Mp + 4;                                   // pitch of the L row panel (GRAD)
  const int stage_doubles = Mp * TRI_PITCH + (GRAD ? TRI_KP * PL : 0);
  const int r = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r * M * M;
  const double* __restrict__ G = GRAD ? Gk + (long long)r * ldk : nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int npairs = (nb / 2) * (nb / 2 + 1) + ((nb & 1) ? nb / 2 + 1 : 0);
  const int npanels = (nb + TRI_KB - 1) / TRI_KB;

  // (ib, jb0) of this warp's pairs, 10 bits each (ib << 5 | jb0; 1023 = none), three per register
  unsigned pairs3[(TRI_MAXPAIRS + 2) / 3];
  double acc[TRI_MAXPAIRS][2][2];
#pragma unroll
  for (int q = 0; q < (TRI_MAXPAIRS + 2) / 3; ++q) pairs3[q] = 0;
#pragma unroll
  for (int q = 0; q < TRI_MAXPAIRS; ++q) {
    const int u = warp + q * TRI_WARPS;
    unsigned code = 1023u;
    if (u < npairs) { const int d = tri_pair_decode(u); code = (unsigned)(((d >> 8) << 5) | (d & 255)); }
    pairs3[q / 3] |= code << (10 * (q % 3));
    acc[q][0][0] = acc[q][0][1] = acc[q][1][0] = acc[q][1][1] = 0.0;
  }

  // syrk: the diagonal of L (KL piece sum log L_mm^2) is fetched now so its latency hides under the panel loop
  const double diag_l = (!GRAD && tid < M) ? L[(long long)tid * M + tid] : 1.0;
  // zero both stages once: rows >= M (padding) are never written by the loaders
  for (int e = tid; e < 2 * stage_doubles; e += blockDim.x) sm[e] = 0.0;
  __syncthreads();

  double diag_pending = 0.0;   // grad: this thread's diagonal element of the panel in flight
  auto load_panel = [&](int pnl, int buf) {
    double* Wc = sm + buf * stage_doubles;                 // Mp x TRI_PITCH: column slab (L for syrk, Gs for grad)
    const int k0 = pnl * TRI_KP;
    if (!GRAD) {
      // rows i >= k0 of L[:, k0:k0+24): 12 x 16-byte chunks per row (M even), zero-filled past column M
      const int i0 = k0;
      for (int e = tid; e < (M - i0) * (TRI_KP / 2); e += blockDim.x) {
        const int i = i0 + e / (TRI_KP / 2), c2 = 2 * (e % (TRI_KP / 2)), k = k0 + c2;
        // columns k <= i (L is lower triangular: the strict upper part of the slab's diagonal square reads as
        // zero) and k < M, through the copy size
        int bytes = min(M - k, i - k + 1) * 8;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        cp_async16(Wc + i * TRI_PITCH + c2, bytes > 0 ? L + (long long)i * M + k : L, bytes);
      }
    } else {
      if (lane < TRI_KP) {  // one warp per slab row, lanes over the panel's 24 columns
        const int kk = lane, k = k0 + kk;
        for (int i = warp; i < M; i += TRI_WARPS) {
          double* dst = Wc + i * TRI_PITCH + kk;
          if (k < M) {
            const int idx = (k <= i) ? i * (i + 1) / 2 + k : k * (k + 1) / 2 + i;
            if (i == k) diag_pending = G[idx];   // lands as 2 Gk_ii - kl at the top of the panel's iteration
            else cp_async8(dst, G + idx);
          } else {
            *dst = 0.0;
          }
        }
      }
      double* Lr = Wc + Mp * TRI_PITCH;                    // TRI_KP x PL: row panel of L (cols <= row)
      const int rows = min(TRI_KP, M - k0);
      for (int kk = warp; kk < rows; kk += TRI_WARPS) {       // one warp per row of the L panel
        const int k = k0 + kk;
        for (int j = 2 * lane; j < Mp; j += 64) {
          int bytes = (min(M, k + 1) - j) * 8;              // only columns j <= k (and < M) are non-zero
          bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
          cp_async16(Lr + kk * PL + j, bytes > 0 ? L + (long long)k * M + j : L, bytes);
        }
      }
      for (int e = tid; e < (TRI_KP - rows) * Mp; e += blockDim.x)
        Lr[(rows + e / Mp) * PL + e % Mp] = 0.0;
    }
    cp_async_commit();
  };
  // One block barrier per panel: it publishes panel pnl (every thread has waited for its own copies) and says
  // that everyone is done with panel pnl - 1, whose buffer the copies of panel pnl + 1 may now overwrite.
  load_panel(0, 0);
  for (int pnl = 0; pnl < npanels; ++pnl) {
    const int buf = pnl & 1;
    cp_async_wait<0>();
    if (GRAD && lane < TRI_KP) {
      // out = (Gs + diag(Gk_ii) - kl I) L: the packed diagonal holds W_ii (not 2 W_ii) and the KL term is -kl L, so
      // the slab's diagonal becomes 2 Gk_ii - kl and the epilogue needs no L / G reloads
      const int k = pnl * TRI_KP + lane;
      if (k < M && (k % TRI_WARPS) == warp) sm[buf * stage_doubles + k * TRI_PITCH + lane] = 2.0 * diag_pending - kl;
    }
    __syncthreads();
    if (pnl + 1 < npanels) load_panel(pnl + 1, buf ^ 1);
    const double* Wc = sm + buf * stage_doubles;
    const double* Lr = Wc + Mp * TRI_PITCH;
    const int kb_lo = TRI_KB * pnl, kb_hi = min(nb, TRI_KB * pnl + TRI_KB) - 1;   // blocks of this panel (inclusive)
#pragma unroll
    for (int q = 0; q < TRI_MAXPAIRS; ++q) {
      const unsigned pr = (pairs3[q / 3] >> (10 * (q % 3))) & 1023u;
      if (pr == 1023u) continue;
      const int ib = pr >> 5, jb0 = pr & 31;
      const bool two = (jb0 + 1 <= ib);
      // syrk: tile (ib,jb) takes kb <= jb ; grad: kb >= jb   (block ranges, inclusive)
      const int lo0 = GRAD ? max(kb_lo, jb0) : kb_lo, hi0 = GRAD ? kb_hi : min(kb_hi, jb0);
      const int lo1 = GRAD ? max(kb_lo, jb0 + 1) : kb_lo, hi1 = GRAD ? kb_hi : min(kb_hi, jb0 + 1);
      const int lo = min(lo0, lo1), hi = two ? max(hi0, hi1) : hi0;
      if (lo > hi) continue;
      // fast path (all but one panel of every pair): both tiles take the whole 24-column panel -- six k4-steps,
      // no predicates (the generic loop below spends ~20 instructions per DMMA on its range logic)
      if (two && kb_hi == kb_lo + TRI_KB - 1 && (GRAD ? kb_lo >= jb0 + 1 : kb_hi <= jb0)) {
        const double* fa = Wc + (8 * ib + gid) * TRI_PITCH + tig;
        const double* fb = GRAD ? Lr + tig * PL + 8 * jb0 + gid : Wc + (8 * jb0 + gid) * TRI_PITCH + tig;
#pragma unroll
        for (int st = 0; st < TRI_KP / 4; ++st) {
          const double a = fa[4 * st];
          const double b0 = GRAD ? fb[4 * st * PL] : fb[4 * st];
          const double b1 = GRAD ? fb[4 * st * PL + 8] : fb[4 * st + 8 * TRI_PITCH];
          dmma(acc[q][0][0], acc[q][0][1], a, b0);
          dmma(acc[q][1][0], acc[q][1][1], a, b1);
        }
        continue;
      }
      const double* pa = Wc + (8 * ib + gid) * TRI_PITCH + tig - 8 * kb_lo;
      const double* pb = GRAD ? Lr + (tig - 8 * kb_lo) * PL + 8 * jb0 + gid
                              : Wc + (8 * jb0 + gid) * TRI_PITCH + tig - 8 * kb_lo;
      for (int kb = lo; kb <= hi; ++kb) {
        const bool u0 = (kb >= lo0 && kb <= hi0), u1 = two && (kb >= lo1 && kb <= hi1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ko = 8 * kb + 4 * h;
          const double a = pa[ko];
          if (u0) dmma(acc[q][0][0], acc[q][0][1], a, GRAD ? pb[ko * PL] : pb[ko]);
          if (u1) dmma(acc[q][1][0], acc[q][1][1], a, GRAD ? pb[ko * PL + 8] : pb[ko + 8 * TRI_PITCH]);
        }
      }
    }
  }

  // epilogue
  double tr = 0.0;
#pragma unroll
  for (int q = 0; q < TRI_MAXPAIRS; ++q) {
    const unsigned pr = (pairs3[q / 3] >> (10 * (q % 3))) & 1023u;
    if (pr == 1023u) continue;
    const int ib = pr >> 5, jb0 = pr & 31;
    const int row = 8 * ib + gid;
    if (row >= M) continue;
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2) {
      if (jb0 + t2 > ib) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * (jb0 + t2) + 2 * tig + e;
        if (col > row) continue;
        const double c = acc[q][t2][e];
        if (!GRAD) {
          out_packed[(long long)r * ldk + (long long)row * (row + 1) / 2 + col] = c;
          if (col == row) tr += c;
        } else {
          double v = c;
          // only the 25 diagonal tiles pay for the load + fp64 division (warp-uniform test first: the compiler
          // otherwise evaluates kl / L_mm for every element, 21 % of the kernel's instructions)
          if (jb0 + t2 == ib) {
            if (col == row) v += kl / L[(long long)row * M + row];
          }
          const long long off = (long long)r * M * M + (long long)row * M + col;
          if (ADAM) {
            double* theta = const_cast<double*>(q_sqrt);   // every element is read (above) before it is written
            double th = theta[off], mi = ad.m[off], vi = ad.v[off];
            adam_update(th, mi, vi, -v, ad.alpha_dev[0], ad.b1, ad.b2, ad.eps);   // v = d ELBO / d theta: minimise -ELBO
            ad.m[off] = mi;
            ad.v[off] = vi;
            theta[off] = th;
          } else {
            dq_sqrt[off] = v;
          }
        }
      }
    }
  }
  if (!GRAD) {
    __shared__ double red[32];
    const long long K = (long long)M * (M + 1) / 2;
    for (long long idx = K + tid; idx < ldk; idx += blockDim.x) out_packed[(long long)r * ldk + idx] = 0.0;
    double ld = (tid < M) ? log(diag_l * diag_l) : 0.0;  // diag_l was fetched at kernel start (M <= 200 < blockDim)
    tr = block_sum(tr, red);
    ld = block_sum(ld, red);
    if (tid == 0) { klp[2 * r] = tr; klp[2 * r + 1] = ld; }
  } else if (!ADAM) {
    // strict upper triangle of dq_sqrt is zero
    double* o = dq_sqrt + (long long)r * M * M;
    for (int e = tid; e < M * M; e += blockDim.x) {
      const int i = e / M, j = e - i * M;
      if (j > i) o[e] = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tri_syrk with the whole factor resident in shared memory (M <= 200, even): the lower 8x8 blocks of L_r
// (nb (nb + 1) / 2 blocks of 512 bytes, 166 KB at M = 200) are copied in once, then 32 warps compute the output
// tiles independently -- no k-panels, no block barrier in the main loop, 8 warps per scheduler to cover the LDS /
// DMMA latencies that bound the output-stationary kernel above (4 warps per scheduler, one barrier per panel).
// Work unit = a pair of horizontally adjacent tiles (ib, jb0), (ib, jb0 + 1) sharing the A fragment; units are
// dealt to the warps in descending weight (2 jb0 + 3 block products), serpentine, so the triangular work balances.
// Block (ib, kb) sits at index ib (ib + 1) / 2 + kb, element (r, c) at r * 8 + (c ^ ((r & 2) << 1)): the fragment
// pattern [gid][tig + 4h] then touches every bank pair once per half-warp.
// ---------------------------------------------------------------------------------------------
constexpr int TSM_WARPS = 32;
// threads per CTA of the resident-factor kernels: a 50 x 50 factor (C5: 40 000 latents) has ~10 work units, so 8 warps and
// four or more CTAs per SM; 32 warps from M = 129 (the factor fills the shared memory of an SM at M = 200)
static inline int tsm_threads(int M) { return M <= 64 ? 256 : (M <= 128 ? 512 : TSM_WARPS * 32); }

__device__ __forceinline__ int tsm_off(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }

__global__ void __launch_bounds__(TSM_WARPS * 32, 1)
tri_syrk_smem_kernel(const double* __restrict__ q_sqrt, int M, double* __restrict__ Sk, long long ldk,
                     double* __restrict__ klp) {
  extern __shared__ __align__(16) double Lb[];
  __shared__ double red[32];
  const int Mp = (M + 7) / 8 * 8, nb = Mp / 8;
  const int r_lat = blockIdx.x;
  const double* __restrict__ L = q_sqrt + (long long)r_lat * M * M;
  double* __restrict__ out = Sk + (long long)r_lat * ldk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  // ---- copy the lower blocks in: one warp per row, lanes over the 16-byte chunks of columns [0, 8 (ib + 1))
  const int nwarps = blockDim.x >> 5;   // 32 warps at M = 200; fewer for small factors (several CTAs per SM)
  for (int i = warp; i < Mp; i += nwarps) {
    const int ib = i >> 3, r = i & 7;
    double* brow = Lb + (size_t)(ib * (ib + 1) / 2) * 64;
    for (int col = 2 * lane; col < 8 * (ib + 1); col += 64) {
      const int kb = col >> 3, c = col & 7;
      // columns <= row (lower triangle) and < M, through the copy size; rows >= M are zero
      int bytes = (i < M) ? min(M - col, i - col + 1) * 8 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      cp_async16(brow + kb * 64 + tsm_off(r, c), bytes > 0 ? L + (long long)i * M + col : L, bytes);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // KL piece: sum_m log L_mm^2 from the resident diagonal blocks
  double ld = 0.0;
  if (tid < M) {
    const int ib = tid >> 3, r = tid & 7;
    const double dgl = Lb[(size_t)(ib * (ib + 1) / 2 + ib) * 64 + tsm_off(r, r)];
    ld = log(dgl * dgl);
  }
  // ---- units in descending weight: column pair jb0 = jtop, jtop - 2, ..., 0; rows ib = jb0 .. nb - 1
  const int fa_off = tsm_off(gid, tig), fb_off = tsm_off(gid, tig + 4);   // [gid][tig], [gid][tig + 4]
  const long long K = (long long)M * (M + 1) / 2;
  double tr = 0.0;
  // unit table (ib << 8 | jb0) in descending weight, built once per CTA; warp w takes entries w, 63 - w, 64 + w, ...
  // (a shared-counter hand-out, which pays in tri_grad, was slower here: 0.53 vs 0.47 ms)
  __shared__ unsigned short units[13 * 26];
  const int jtop = (nb - 1) & ~1;
  int nunits = 0;
  for (int j = jtop; j >= 0; j -= 2) nunits += nb - j;
  for (int u = tid; u < nunits; u += blockDim.x) {
    int j = jtop, rem = u;
    while (rem >= nb - j) { rem -= nb - j; j -= 2; }
    units[u] = (unsigned short)(((j + rem) << 8) | j);
  }
  __syncthreads();
  for (int round = 0; round * nwarps < nunits; ++round) {
    {
      const int idx = round * nwarps + ((round & 1) ? nwarps - 1 - warp : warp);
      if (idx >= nunits) continue;
      const int ib = units[idx] >> 8, jb0 = units[idx] & 255;
      const bool two = jb0 + 1 <= ib;
      const double* pa = Lb + (size_t)(ib * (ib + 1) / 2) * 64;
      const double* pb0 = Lb + (size_t)(jb0 * (jb0 + 1) / 2) * 64;
      const double* pb1 = Lb + (size_t)((jb0 + 1) * (jb0 + 2) / 2) * 64;
      double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
      for (int kb = 0; kb <= jb0; ++kb) {   // S[ib][jb] = sum_{kb <= jb} L[ib][kb] L[jb][kb]^T
        const double a0 = pa[kb * 64 + fa_off], a1 = pa[kb * 64 + fb_off];
        const double b0 = pb0[kb * 64 + fa_off], b1 = pb0[kb * 64 + fb_off];
        dmma(c00, c01, a0, b0);
        dmma(c00, c01, a1, b1);
        if (two) {
          const double d0 = pb1[kb * 64 + fa_off], d1 = pb1[kb * 64 + fb_off];
          dmma(c10, c11, a0, d0);
          dmma(c10, c11, a1, d1);
        }
      }
      if (two) {   // the second tile also takes kb = jb0 + 1 (its own diagonal block)
        const int kb = jb0 + 1;
        dmma(c10, c11, pa[kb * 64 + fa_off], pb1[kb * 64 + fa_off]);
        dmma(c10, c11, pa[kb * 64 + fb_off], pb1[kb * 64 + fb_off]);
      }
      const int row = 8 * ib + gid;
      if (row < M) {
        double* orow = out + (long long)row * (row + 1) / 2;
        const int col0 = 8 * jb0 + 2 * tig;
        if (col0 <= row) { orow[col0] = c00; if (col0 == row) tr += c00; }
        if (col0 + 1 <= row) { orow[col0 + 1] = c01; if (col0 + 1 == row) tr += c01; }
        if (two) {
          const int col1 = col0 + 8;
          if (col1 <= row) { orow[col1] = c10; if (col1 == row) tr += c10; }
          if (col1 + 1 <= row) { orow[col1 + 1] = c11; if (col1 + 1 == row) tr += c11; }
        }
      }
    }
  }
  for (long long idx = K + tid; idx < ldk; idx += blockDim.x) out[idx] = 0.0;
  tr = block_sum(tr, red);
  ld = block_sum(ld, red);
  if (tid == 0) { klp[2 * r_lat] = tr; klp[2 * r_lat + 1] = ld; }
}

// ---------------------------------------------------------------------------------------------
// tri_grad with the factor resident in shared memory (same layout and occupancy argument as tri_syrk_smem_kernel):
//   out[ib][jb] = sum_{kb >= jb} Gs[ib][kb] L[kb][jb],   dq_sqrt = tril(out) + kl diag(1 / L_mm)
// B fragments (L[kb][jb], pattern [tig + 4h][gid]) come from the resident blocks; A fragments (the symmetric unpack of
// the packed cotangent, diagonal replaced by 2 Gk_ii - kl) are read from global memory with full-sector accesses
// (rows of 4 / columns of 8 consecutive doubles).  Units (pairs of adjacent tiles of one block row) are handed out in
// block-row order, so the 32 warps of a CTA work on two or three block rows at a time and the 13 KB of Gs they share
// stay in L1; the packed cotangent is then read from L2 about once.
// ---------------------------------------------------------------------------------------------
constexpr int TG_NC = 4;   // column blocks per work unit of tri_grad_smem_kernel

template <bool ADAM>
__global__ void __launch_bounds__(TSM_WARPS * 32, 1)
tri_grad_smem_kernel(const double* __restrict__ Gk, long long ldk, const double* __restrict__ q_sqrt, int M,
                     double* __restrict__ dq_sqrt, double kl, TriAdam ad) {
  extern __shared__ __align__(16) double Lb[];
  __shared__ unsigned short units[13 * 26];
  __shared__ int next_unit;
  const int Mp = (M + 7) / 8 * 8, nb = Mp / 8;
  const int r_lat = blockIdx.x;
  if (threadIdx.x == 0) next_unit = 0;
  const double* __restrict__ L = q_sqrt + (long long)r_lat * M * M;
  const double* __restrict__ G = Gk + (long long)r_lat * ldk;
  double* __restrict__ dq = dq_sqrt + (long long)r_lat * M * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int nwarps = blockDim.x >> 5;
  for (int i = warp; i < Mp; i += nwarps) {   // resident lower blocks of L (see tri_syrk_smem_kernel)
    const int ib = i >> 3, r = i & 7;
    double* brow = Lb + (size_t)(ib * (ib + 1) / 2) * 64;
    for (int col = 2 * lane; col < 8 * (ib + 1); col += 64) {
      int bytes = (i < M) ? min(M - col, i - col + 1) * 8 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      cp_async16(brow + (col >> 3) * 64 + tsm_off(r, col & 7), bytes > 0 ? L + (long long)i * M + col : L, bytes);
    }
  }
  cp_async_commit();
  // strict upper triangle of dq_sqrt is zero (under the copy); the fused optimiser never touches it
  if (!ADAM) {
    for (int i = warp; i < M; i += nwarps)
      for (int j = i + 1 + lane; j < M; j += 32) dq[(long long)i * M + j] = 0.0;
  }
  // unit table: block rows from the bottom, groups of TG_NC column blocks left to right (heavy first within a row);
  // one A fragment pair feeds up to TG_NC output tiles (half the non-DMMA instructions per DMMA of the pair version:
  // 0.93 -> 0.88 ms; pairing two block rows on top of it, for half the B reads, was slower: 0.95 ms at the 64-register cap)
  int nunits = 0;
  for (int ib = 0; ib < nb; ++ib) nunits += ib / TG_NC + 1;
  for (int u = tid; u < nunits; u += blockDim.x) {
    int ib = nb - 1, rem = u;
    while (rem >= ib / TG_NC + 1) { rem -= ib / TG_NC + 1; --ib; }
    units[u] = (unsigned short)((ib << 8) | (TG_NC * rem));
  }
  cp_async_wait<0>();
  __syncthreads();
  const int ob0 = tsm_off(tig, gid), ob1 = tsm_off(tig + 4, gid);   // [tig][gid], [tig + 4][gid]
  // units are handed out in table order from a shared counter: the triangular weights (1 .. 49 block products)
  // balance dynamically while concurrent warps stay on neighbouring block rows; which warp computes a unit does
  // not change its result
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&next_unit, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= nunits) break;
    const int ib = units[idx] >> 8, jb0 = units[idx] & 255;
    const int ncol = min(TG_NC, ib + 1 - jb0);   // output tiles (ib, jb0 .. jb0 + ncol - 1)
    const int i = 8 * ib + gid;
    const bool iok = i < M;
    const double* Gi = G + (long long)i * (i + 1) / 2;   // packed row i (columns <= i)
    double c[TG_NC][2];
#pragma unroll
    for (int t = 0; t < TG_NC; ++t) c[t][0] = c[t][1] = 0.0;
    for (int kb = jb0; kb < nb; ++kb) {
      // A fragment: Gs[i][k], k = 8 kb + tig (+ 4)
      const int k0 = 8 * kb + tig, k1 = k0 + 4;
      double a0 = 0.0, a1 = 0.0;
      if (kb < ib) {
        if (iok) { a0 = Gi[k0]; a1 = Gi[k1]; }
      } else if (kb > ib) {
        if (iok && k0 < M) a0 = G[(long long)k0 * (k0 + 1) / 2 + i];
        if (iok && k1 < M) a1 = G[(long long)k1 * (k1 + 1) / 2 + i];
      } else {
        if (iok && k0 < M) a0 = (k0 <= i) ? Gi[k0] : G[(long long)k0 * (k0 + 1) / 2 + i];
        if (iok && k1 < M) a1 = (k1 <= i) ? Gi[k1] : G[(long long)k1 * (k1 + 1) / 2 + i];
        // out = (Gs + diag(Gk_ii) - kl I) L: the packed diagonal holds W_ii, the KL term is -kl L
        if (k0 == i) a0 = 2.0 * a0 - kl;
        if (k1 == i) a1 = 2.0 * a1 - kl;
      }
      const double* pb = Lb + (size_t)(kb * (kb + 1) / 2 + jb0) * 64;   // block (kb, jb0); (kb, jb0 + 1) follows
      const int nact = min(ncol, kb - jb0 + 1);   // block (kb, jb0 + t) is zero above the diagonal
#pragma unroll
      for (int t = 0; t < TG_NC; ++t)
        if (t < nact) dmma(c[t][0], c[t][1], a0, pb[t * 64 + ob0]);
#pragma unroll
      for (int t = 0; t < TG_NC; ++t)
        if (t < nact) dmma(c[t][0], c[t][1], a1, pb[t * 64 + ob1]);
    }
    if (iok) {
      const long long rbase = (long long)r_lat * M * M + (long long)i * M;
      const double dinv = kl / Lb[(size_t)(ib * (ib + 1) / 2 + ib) * 64 + tsm_off(gid, gid)];
      // ADAM: the Keras-Adam update of run_adam applied in place (theta = q_sqrt, all of whose reads went through the
      // resident copy), else the gradient itself
      auto emit = [&](int col, double v0, double v1, bool both) {
        if (ADAM) {
          double* theta = const_cast<double*>(q_sqrt);
          const double alpha = ad.alpha_dev[0];
          for (int e = 0; e < (both ? 2 : 1); ++e) {
            const long long off = rbase + col + e;
            double th = theta[off], mi = ad.m[off], vi = ad.v[off];
            adam_update(th, mi, vi, -(e ? v1 : v0), alpha, ad.b1, ad.b2, ad.eps);
            ad.m[off] = mi; ad.v[off] = vi; theta[off] = th;
          }
        } else if (both) {
          *reinterpret_cast<double2*>(dq_sqrt + rbase + col) = make_double2(v0, v1);
        } else {
          dq_sqrt[rbase + col] = v0;
        }
      };
#pragma unroll
      for (int t = 0; t < TG_NC; ++t) {
        if (t < ncol) {
          const int col = 8 * (jb0 + t) + 2 * tig;
          if (col + 1 <= i) emit(col, c[t][0], col + 1 == i ? c[t][1] + dinv : c[t][1], true);
          else if (col == i) emit(col, c[t][0] + dinv, 0.0, false);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dP[:, n] (+)= 2 T_n p_n - 2 gs[n] p_n * use_prior,   T_n symmetric from packed Tk[n, idx(i,j)].
// gs[n] = sum_r g_var[n, r] (row sums, computed here).   grid = N, block = 256, smem = 2M + 32
// ---------------------------------------------------------------------------------------------
template <int MAXC>   // columns per lane: M <= 32 MAXC
__global__ void sym_matvec_kernel(const double* __restrict__ Tk, long long ldk,
                                  const double* __restrict__ P, long long ldp, int M, int N,
                                  const double* __restrict__ g_var, long long ldg, int R,
                                  int use_prior, double* __restrict__ dP, long long lddp,
                                  double* __restrict__ gs_out, int accumulate) {
  extern __shared__ double sm[];
  double* p = sm;
  double* acc = sm + M;
  double* red = sm + 2 * M;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int m = threadIdx.x; m < M; m += blockDim.x) p[m] = P[(long long)m * ldp + n];
  double gs = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) gs += g_var[(long long)n * ldg + r];
  gs = block_sum(gs, red);
  if (threadIdx.x == 0 && gs_out) gs_out[n] = gs;
  const double* T = Tk + (long long)n * ldk;
  // single pass over the packed lower triangle, one warp per row i (coalesced): lane j accumulates
  //   row part    acc_row  += T[i,j] p[j]          (warp-reduced -> out[i])
  //   column part col[j]   += T[i,j] p[i], j < i   (lane-private registers, columns j = lane + 32 c)
  double col[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) col[c] = 0.0;
  double* colpart = red + 32;  // nwarps x M per-warp column partials (no atomics: deterministic)
  // four rows of a warp in flight (loads of all four issued before the first use; same order of additions as row by row)
  constexpr int RU = 4;
  for (int i0 = warp; i0 < M; i0 += RU * nwarps) {
    double t[RU][MAXC];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int i = i0 + u * nwarps;
      const double* Ti = T + (long long)i * (i + 1) / 2;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int j = lane + 32 * c;
        t[u][c] = (i < M && j <= i) ? Ti[j] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int i = i0 + u * nwarps;
      if (i < M) {   // warp-uniform
        const double pi = p[i];
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int j = lane + 32 * c;
          if (j <= i) {
            s += t[u][c] * p[j];
            if (j < i) col[c] += t[u][c] * pi;
          }
        }
        s = warp_sum(s);
        if (lane == 0) acc[i] = s;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = lane + 32 * c;
    if (j < M) colpart[warp * M + j] = col[c];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    double t = acc[j];
    for (int w2 = 0; w2 < nwarps; ++w2) t += colpart[w2 * M + j];
    double v = 2.0 * t;
    if (use_prior) v -= 2.0 * gs * p[j];
    double* dst = dP + (long long)j * lddp + n;
    *dst = accumulate ? (*dst + v) : v;
  }
}

}  // namespace fcest

using namespace fcest;

extern "C" int fcest_pack_outer(const double* P, long long ldp, int M, int N, double* Pk,
                                long long ldk, double* bias, const double* kvar, int use_prior,
                                cudaStream_t stream) {
  if (N <= 0) return 0;
  const size_t smem = sizeof(double) * (M + 32);
  if (smem > 200 * 1024) return FCEST_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pack_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pack_outer_kernel<<<N, 256, smem, stream>>>(P, ldp, M, N, Pk, ldk, bias, kvar, use_prior);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tri_syrk_pack(const double* q_sqrt, int R, int M, double* Sk, long long ldk,
                                   double* kl_parts, cudaStream_t stream) {
  if (R <= 0) return 0;
  // default: the shared-memory-resident kernel (0.70 -> 0.49 ms at C4); FCEST_TRI_SYRK_IMPL=stationary selects the
  // output-stationary panel kernel
  static const int use_smem = getenv("FCEST_TRI_SYRK_IMPL") ? strcmp(getenv("FCEST_TRI_SYRK_IMPL"), "stationary") != 0 : 1;
  if (use_smem && M <= 200 && (M & 1) == 0) {
    const int nb = (M + 7) / 8;
    const size_t smem = sizeof(double) * 64 * (size_t)(nb * (nb + 1) / 2);
    cudaError_t e = cudaFuncSetAttribute(tri_syrk_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tri_syrk_smem_kernel<<<R, tsm_threads(M), smem, stream>>>(q_sqrt, M, Sk, ldk, kl_parts);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  if (M <= 200 && (M & 1) == 0) {
    const int Mp = (M + 7) / 8 * 8;
    const size_t smem = sizeof(double) * 2 * (size_t)Mp * TRI_PITCH;
    cudaError_t e = cudaFuncSetAttribute(tri_stationary_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tri_stationary_kernel<false><<<R, TRI_WARPS * 32, smem, stream>>>(q_sqrt, M, nullptr, Sk, ldk, kl_parts, nullptr, 0.0);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  tri_syrk_pack_kernel<<<R, 256, 0, stream>>>(q_sqrt, M, Sk, ldk, kl_parts);
  FCEST_CHECK_LAUNCH();
  return 0;
}

static int tri_grad_use_smem() {
  static const int v = getenv("FCEST_TRI_GRAD_IMPL") ? strcmp(getenv("FCEST_TRI_GRAD_IMPL"), "stationary") != 0 : 1;
  return v;
}

extern "C" int fcest_tri_grad(const double* Gk, long long ldk, const double* q_sqrt, int R, int M,
                              double* dq_sqrt, double kl, cudaStream_t stream) {
  if (R <= 0) return 0;
  // default: the shared-memory-resident kernel (1.28 -> 0.97 ms at C4); FCEST_TRI_GRAD_IMPL=stationary selects the
  // output-stationary panel kernel
  if (tri_grad_use_smem() && M <= 200 && (M & 1) == 0) {
    const int nb = (M + 7) / 8;
    const size_t smem = sizeof(double) * 64 * (size_t)(nb * (nb + 1) / 2);
    cudaError_t e = cudaFuncSetAttribute(tri_grad_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tri_grad_smem_kernel<false><<<R, tsm_threads(M), smem, stream>>>(Gk, ldk, q_sqrt, M, dq_sqrt, kl, TriAdam());
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  if (M <= 200 && (M & 1) == 0) {
    const int Mp = (M + 7) / 8 * 8;
    const size_t smem = sizeof(double) * 2 * ((size_t)Mp * TRI_PITCH + (size_t)TRI_KP * (Mp + 4));
    cudaError_t e = cudaFuncSetAttribute(tri_stationary_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tri_stationary_kernel<true><<<R, TRI_WARPS * 32, smem, stream>>>(q_sqrt, M, Gk, nullptr, ldk, nullptr, dq_sqrt, kl);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  tri_grad_kernel<<<R, 256, 0, stream>>>(Gk, ldk, q_sqrt, M, dq_sqrt, kl);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_tri_grad_adam(const double* Gk, long long ldk, double* q_sqrt, int R, int M, double kl,
                                   double* m, double* v, const double* alpha_t, double b1, double b2, double eps,
                                   cudaStream_t stream) {
  if (R <= 0) return 0;
  if (!Gk || !q_sqrt || !m || !v || !alpha_t) return FCEST_ERR_ARGUMENT;
  if (!(M <= 200 && (M & 1) == 0)) return FCEST_ERR_UNSUPPORTED;   // caller falls back to tri_grad + adam_step
  if (tri_grad_use_smem()) {
    const int nb = (M + 7) / 8;
    const size_t smem_r = sizeof(double) * 64 * (size_t)(nb * (nb + 1) / 2);
    cudaError_t e2 = cudaFuncSetAttribute(tri_grad_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
    if (e2 != cudaSuccess) return (int)e2;
    TriAdam ad2;
    ad2.m = m; ad2.v = v; ad2.alpha_dev = alpha_t; ad2.b1 = b1; ad2.b2 = b2; ad2.eps = eps;
    tri_grad_smem_kernel<true><<<R, tsm_threads(M), smem_r, stream>>>(Gk, ldk, q_sqrt, M, nullptr, kl, ad2);
    FCEST_CHECK_LAUNCH();
    return 0;
  }
  const int Mp = (M + 7) / 8 * 8;
  const size_t smem = sizeof(double) * 2 * ((size_t)Mp * TRI_PITCH + (size_t)TRI_KP * (Mp + 4));
  cudaError_t e = cudaFuncSetAttribute(tri_stationary_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  TriAdam ad;
  ad.m = m; ad.v = v; ad.alpha_dev = alpha_t; ad.b1 = b1; ad.b2 = b2; ad.eps = eps;
  tri_stationary_kernel<true, true><<<R, TRI_WARPS * 32, smem, stream>>>(q_sqrt, M, Gk, nullptr, ldk, nullptr, nullptr,
                                                                         kl, ad);
  FCEST_CHECK_LAUNCH();
  return 0;
}

extern "C" int fcest_sym_matvec(const double* Tk, long long ldk, const double* P, long long ldp,
                                int M, int N, const double* g_var, long long ldg, int R,
                                int use_prior, double* dP, long long lddp, double* gs_out,
                                int accumulate, cudaStream_t stream) {
  if (N <= 0) return 0;
  const size_t smem = sizeof(double) * (2 * M + 32 + 8 * M);
  if (smem > 200 * 1024 || M > 416) return FCEST_ERR_UNSUPPORTED;
  if (M <= 224) {   // 7 columns per lane: 56 instead of 104 registers for the four rows in flight
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(sym_matvec_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sym_matvec_kernel<7><<<N, 256, smem, stream>>>(Tk, ldk, P, ldp, M, N, g_var, ldg, R, use_prior, dP, lddp, gs_out,
                                                   accumulate);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(sym_matvec_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sym_matvec_kernel<13><<<N, 256, smem, stream>>>(Tk, ldk, P, ldp, M, N, g_var, ldg, R, use_prior, dP, lddp, gs_out,
                                                    accumulate);
  }
  FCEST_CHECK_LAUNCH();
  return 0;
}
