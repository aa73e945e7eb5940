// (kernel + launch template; instantiated in likelihood_cluster.cu / likelihood_cluster_b.cu / likelihood_cluster_c.cu so that
// the six row-block capacities compile in parallel: each takes ~45 s)
//
// Cluster-tiled Wishart likelihood (forward + analytic backward) for 208 <= D <= 407: one thread-block CLUSTER per
// D x D problem, the Cholesky factor held as packed lower 8 x 8 tiles in the DISTRIBUTED shared memory of the
// cluster's C CTAs (679 KB of tiles at D = 400: 170 - 180 KB per SM in a cluster of four; CTA pairs up to D = 303).
//
// Reference semantics: fcest/models/likelihoods.py:126-206, :253-264 (any D: `tf.linalg.cholesky` at :185); the math
// (bordered y row with unit pivot, Q'' = Sigma^-1 G - alpha (alpha^T G) out of the substitution, the lam_bar identity)
// is the one of likelihood_warp.cu / likelihood_tiles.cu.
//
// Mapping (C CTAs x 8 warps per problem):
//   * tile (ib, jb) of the lower triangle lives in the shared memory of CTA  ib mod C  (block rows dealt cyclically,
//     packed per CTA); every CTA reaches every tile through the cluster's shared-memory window (generic pointers from
//     `mapa`), so all DMMA operands of the O(D^3) phases come from on-chip memory of the cluster;
//   * P0  (separate streaming kernel, per chunk of time steps: wishart_g_fill_kernel in likelihood_tiles.cu) G = A o (fmean +
//         sqrt(fvar) o eps) for all samples of the chunk, written once at HBM speed (a fill inside the persistent kernel was
//         16 % of its time);
//   * P1  Sigma = G G^T: output-stationary 16 x 16 macro tiles in registers, dealt over the cluster's warps; G arrives as
//         4-column panels (one DMMA k-step; dense 32-byte rows are conflict-free for the fragment reads) through a ring
//         of TMA tensor-map copies (`cp.async.bulk.tensor.3d`, SASS UTMALDG; rows >= D zero-filled by the copy engine)
//         with full / empty mbarriers -- no registers and no block barrier in the k loop; finished tiles are stored into
//         their owner's shared memory;
//   * P2  left-looking blocked Cholesky: every CTA updates the tiles of ITS block rows (its own operand rows are local,
//         the tiles of block row jb come from their owner; two partial sums halve the dependent DMMA chain); the owner
//         of row jb factors + inverts the diagonal tile (lik_tiles.cuh); two cluster barriers per block column;
//   * P4  per 8-column block of G one warp: right-looking block forward substitution, row D negated, right-looking
//         back substitution with all row tiles of the column block in registers; operand fragments are requested a
//         chunk ahead of their DMMAs (explicit software pipeline over the cluster window); the per-sample cotangents
//         accumulate in the cluster's global slot (U1, U2);
//   * P5  lam_bar, ve on CTA 0 (row sums and alpha arrive through the cluster window).
// The noise of the NEXT sample is pulled into L2 with `cp.async.bulk.prefetch.L2` while the current one is factored.
// What bounds it: the cluster window delivers ~20 B/clk per SM (B300_MICROARCH.md), a quarter-tile DMMA operand is 256 B,
// so with 1/2 (pairs) or 3/4 (clusters of four) of the operands remote the substitution cannot exceed ~40 % / ~30 % of
// the DMMA issue rate; measured 16.5 % (D = 256, pairs) and 11 % (clusters of four) of the fp64 peak.
// Deterministic: every accumulation has a fixed owner and order.
#pragma once
#include <cooperative_groups.h>
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "chol_tiles.cuh"
#include "common.cuh"
#include "fcest_b200.h"
#include "lik_params.cuh"
#include "lik_tiles.cuh"

namespace cg = cooperative_groups;

namespace fcest {

// Warps per CTA by (capacity, cluster size): the column-block phase (one warp per 8-column block of G) is latency bound
// and wants as many warps as blocks -- 42 / 46 blocks on 4 x 11 / 4 x 12 warps in ONE round (D = 320: 35.5 -> 33.9 ms)
// -- where the per-warp row-sum scratch (8 NB doubles each) still fits beside the tile store.  CTA pairs stay at 8 warps:
// 13 warps at 128 registers spill 17 - 24 KB per thread in the unrolled substitution (D = 256: 24.8 -> 30.9 ms).
__host__ __device__ constexpr int cl_warps(int NB, int C) {
  return C == 2 ? 8 : (NB <= 42 ? 11 : NB <= 46 ? 12 : 8);
}
// macro tiles per warp and SYRK pass (16 accumulator registers each)
__host__ __device__ constexpr int cl_mq(int NB, int C) {
  const int nbm = (NB + 1) / 2, nmac = nbm * (nbm + 1) / 2, cw = C * cl_warps(NB, C);
  const int per = (nmac + cw - 1) / cw;
  return cl_warps(NB, C) == 8 ? 6 : (per > 4 ? ((per + 1) / 2 > 4 ? 4 : (per + 1) / 2) : per);
}
constexpr int CL_PC = 4;      // columns of a G panel = one DMMA k-step (pitch 4: fragment reads are contiguous)
constexpr int CL_MAXST = 8;   // at most 8 panel stages
constexpr int CL_PF = 3;      // G tiles in flight ahead of the forward substitution

struct ClLayout {
  int NB, Dp, Dm, maxt, ST;
  int off_pan, off_vec, off_bar, total;  // doubles
};

// packed offset (in tiles) of tile (ib, jb) inside its owner CTA  ib % C
template <int C>
__host__ __device__ __forceinline__ int cl_loff(int ib, int jb) {
  const int r = ib % C, lr = ib / C;
  return C * lr * (lr - 1) / 2 + lr * (r + 1) + jb;
}

template <int C>
__host__ __device__ constexpr ClLayout cl_layout(int NB) {
  ClLayout L = {};
  L.NB = NB;
  L.Dp = 8 * NB;
  L.Dm = (L.Dp + 15) / 16 * 16;
  L.maxt = 0;
  for (int r = 0; r < C && r < NB; ++r) {
    const int nr = (NB - r + C - 1) / C;  // block rows r, r + C, ... < NB
    const int t = C * nr * (nr - 1) / 2 + nr * (r + 1);
    if (t > L.maxt) L.maxt = t;
  }
  L.off_pan = L.maxt * 64;
  const int vec = 4 * L.Dp + 32;                      // dinv, yv, spl, alpha, misc (bad flags of the C CTAs)
  const int stage = L.Dm * CL_PC;                     // one panel: Dm rows x 4 columns
  const int avail = 227 * 1024 / 8 - 16 - L.off_pan - vec - 2 * CL_MAXST;
  int st = avail / stage;
  if (st > CL_MAXST) st = CL_MAXST;
  if (st < 2) st = 2;
  while (st * stage < cl_warps(NB, C) * L.Dp) ++st;   // the ring doubles as per-warp row-sum partials (Dp each) in P4
  L.ST = st;
  L.off_vec = L.off_pan + st * stage;
  L.off_bar = L.off_vec + vec;
  L.total = L.off_bar + 2 * CL_MAXST;
  return L;
}

__device__ __forceinline__ void cl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void cl_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(uint32_t bar, uint32_t parity) {
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
// TMA tensor-map copy of one box (4 columns x BH rows of one cluster's G slot) global -> shared (SASS UTMALDG)
__device__ __forceinline__ void cl_tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void cl_cvt_c_to_b(double c0, double c1, int gid, int tig, double& b0, double& b1) {
  const int src0 = 4 * tig + (gid >> 1), src1 = 4 * (tig + 4) + (gid >> 1);
  const double x0 = __shfl_sync(0xffffffffu, c0, src0), x1 = __shfl_sync(0xffffffffu, c1, src0);
  const double y0 = __shfl_sync(0xffffffffu, c0, src1), y1 = __shfl_sync(0xffffffffu, c1, src1);
  b0 = (gid & 1) ? x1 : x0;
  b1 = (gid & 1) ? y1 : y0;
}

template <int C>
struct ClCtx {
  const double* rb[C];   // tile store of every CTA of the cluster (generic pointers into the cluster window)
  const double* G; const double* ep;   // G slot (row pitch ldgs), noise of this sample
  int D, nu, ldgs, colC, ycol, gid, tig, oA0, oA1, oT0, oT1;
  int kb_start;          // first block row with a non-zero right-hand side (> 0 only for the column block that holds just e_D)
};

template <int C, int IB, int JB>
__device__ __forceinline__ const double* cl_tile(const ClCtx<C>& c) {
  constexpr int r = IB % C, lr = IB / C;
  constexpr int off = C * lr * (lr - 1) / 2 + lr * (r + 1) + JB;
  return c.rb[r] + off * 64;
}

template <int C>
__device__ __forceinline__ void cl_load_g(const ClCtx<C>& c, int ib, double (&g)[2]) {
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

// Operand fragments come from the cluster window (generic loads, a few hundred cycles from a peer SM), so the trailing
// updates run as an explicit software pipeline: volatile loads keep program order, CL_W fragments are requested one
// chunk ahead of the DMMAs that use them (a ring of 2 CL_W registers), and the first chunk is requested before the
// shuffles / diagonal-tile products of the step, which do not depend on it.
template <int NB> constexpr int cl_w() { return NB > 40 ? 6 : 12; }   // fewer registers where 2 NB accumulators are > 80

__device__ __forceinline__ double cl_ldv(const double* ptr) {
  double v;
  asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
  return v;
}

template <int NB, int C, int KB, bool FWD>
struct ClTrail {
  static constexpr int CL_W = cl_w<NB>();
  static constexpr int M = FWD ? NB - KB - 1 : KB;   // tiles below (forward) / left of (backward) the diagonal tile
  static constexpr int n = 2 * M;                    // first all tiles with the first half fragment, then the second
  template <int I>
  static __device__ __forceinline__ const double* ptr(const ClCtx<C>& c) {
    constexpr int row = I % (M > 0 ? M : 1), half = I / (M > 0 ? M : 1);
    if constexpr (FWD) return cl_tile<C, KB + 1 + row, KB>(c) + (half ? c.oA1 : c.oA0);
    else return cl_tile<C, KB, row>(c) + (half ? c.oT1 : c.oT0);
  }
  template <int I, int E>
  static __device__ __forceinline__ void load(double (&buf)[2 * CL_W], const ClCtx<C>& c) {
    if constexpr (I < E && I < n) {
      buf[I % (2 * CL_W)] = cl_ldv(ptr<I>(c));
      load<I + 1, E>(buf, c);
    }
  }
  template <int I, int E>
  static __device__ __forceinline__ void mma(double (&acc)[NB][2], const double (&buf)[2 * CL_W], double b0, double b1) {
    if constexpr (I < E && I < n) {
      constexpr int row = I % (M > 0 ? M : 1), half = I / (M > 0 ? M : 1);
      constexpr int IB = FWD ? KB + 1 + row : row;
      dmma(acc[IB][0], acc[IB][1], buf[I % (2 * CL_W)], half ? b1 : b0);
      mma<I + 1, E>(acc, buf, b0, b1);
    }
  }
  template <int S>
  static __device__ __forceinline__ void run(double (&acc)[NB][2], double (&buf)[2 * CL_W], const ClCtx<C>& c, double b0,
                                             double b1) {
    if constexpr (S < n) {
      load<S + CL_W, S + 2 * CL_W>(buf, c);
      mma<S, S + CL_W>(acc, buf, b0, b1);
      run<S + CL_W>(acc, buf, c, b0, b1);
    }
  }
};

// forward, right-looking: H[kb] = I_kb (G[kb] + acc[kb]); acc[ib] -= L[ib][kb] H[kb] (ib > kb); ik = fragments of the
// diagonal tile of this step (requested one step ahead)
template <int NB, int C, int KB>
__device__ __forceinline__ void cl_fwd(double (&acc)[NB][2], double (&gq)[CL_PF][2], double (&ik)[2], const ClCtx<C>& c) {
  if constexpr (KB < NB) {
    double ikn[2] = {0.0, 0.0};
    if (KB >= c.kb_start) {   // warp-uniform: rows above the first non-zero block of the right-hand side stay zero
      constexpr int CL_W = cl_w<NB>();
      double buf[2 * CL_W];
      ClTrail<NB, C, KB, true>::template load<0, CL_W>(buf, c);
      if constexpr (KB + 1 < NB) {
        ikn[0] = cl_ldv(cl_tile<C, KB + 1, KB + 1>(c) + c.oA0);
        ikn[1] = cl_ldv(cl_tile<C, KB + 1, KB + 1>(c) + c.oA1);
      }
      double b0, b1;
      cl_cvt_c_to_b(acc[KB][0] + gq[KB % CL_PF][0], acc[KB][1] + gq[KB % CL_PF][1], c.gid, c.tig, b0, b1);
      if constexpr (KB + CL_PF < NB) cl_load_g<C>(c, KB + CL_PF, gq[KB % CL_PF]);
      double h0 = 0.0, h1 = 0.0;
      dmma(h0, h1, ik[0], b0);
      dmma(h0, h1, ik[1], b1);
      acc[KB][0] = h0; acc[KB][1] = h1;
      cl_cvt_c_to_b(-h0, -h1, c.gid, c.tig, b0, b1);
      ClTrail<NB, C, KB, true>::template run<0>(acc, buf, c, b0, b1);
    } else {
      if constexpr (KB + CL_PF < NB) cl_load_g<C>(c, KB + CL_PF, gq[KB % CL_PF]);
      if constexpr (KB + 1 < NB) {
        ikn[0] = cl_ldv(cl_tile<C, KB + 1, KB + 1>(c) + c.oA0);
        ikn[1] = cl_ldv(cl_tile<C, KB + 1, KB + 1>(c) + c.oA1);
      }
    }
    cl_fwd<NB, C, KB + 1>(acc, gq, ikn, c);
  }
}
// backward, right-looking: Q[kb] = I_kb^T acc[kb]; acc[ib] -= L[kb][ib]^T Q[kb] (ib < kb)
template <int NB, int C, int KB>
__device__ __forceinline__ void cl_bwd(double (&acc)[NB][2], double (&ik)[2], const ClCtx<C>& c) {
  if constexpr (KB >= 0) {
    constexpr int CL_W = cl_w<NB>();
    double buf[2 * CL_W];
    double ikn[2] = {0.0, 0.0};
    ClTrail<NB, C, KB, false>::template load<0, CL_W>(buf, c);
    if constexpr (KB >= 1) {
      ikn[0] = cl_ldv(cl_tile<C, KB - 1, KB - 1>(c) + c.oT0);
      ikn[1] = cl_ldv(cl_tile<C, KB - 1, KB - 1>(c) + c.oT1);
    }
    double b0, b1;
    cl_cvt_c_to_b(acc[KB][0], acc[KB][1], c.gid, c.tig, b0, b1);
    double q0 = 0.0, q1 = 0.0;
    dmma(q0, q1, ik[0], b0);
    dmma(q0, q1, ik[1], b1);
    acc[KB][0] = q0; acc[KB][1] = q1;
    cl_cvt_c_to_b(-q0, -q1, c.gid, c.tig, b0, b1);
    ClTrail<NB, C, KB, false>::template run<0>(acc, buf, c, b0, b1);
    cl_bwd<NB, C, KB - 1>(acc, ikn, c);
  }
}
// epilogue of a column block, two row tiles at a time with every global load issued before the first use:
// alpha from the e_D column (into CTA 0); row sums of Q'' o G; U1 += Gbar, U2 += Gbar o eps (this lane owns its elements)
template <int NB, int C, int IB>
__device__ __forceinline__ void cl_epilogue(double (&acc)[NB][2], const ClCtx<C>& c, double* U1, double* U2, double* alpha0,
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
        if (d < c.D && k == c.ycol) alpha0[d] = -q;
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
    cl_epilogue<NB, C, IB + T>(acc, c, U1, U2, alpha0, rsw, wS, need_grad, first_sample);
  }
}

// L2 prefetch of `bytes` (multiple of 16) through the bulk-copy engine
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int NB, int C>
__global__ void __launch_bounds__(cl_warps(NB, C) * 32, 1)
wishart_loglik_cluster_kernel(const __grid_constant__ CUtensorMap tmG, const LikParams p, const int ldgs, const int box_rows,
                              const double* Gc, const int n0, const int nc) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) double sm[];
  const int D = p.D, nu = p.nu, R = D * nu, N = p.N;
  constexpr ClLayout L = cl_layout<C>(NB);
  constexpr int Dp = L.Dp, Dm = L.Dm, ST = L.ST;
  constexpr int STAGE = Dm * CL_PC;   // doubles per panel stage
  constexpr int CL_WARPS = cl_warps(NB, C), CL_MQ = cl_mq(NB, C);
  double* tiles = sm;
  double* pan = sm + L.off_pan;
  double* dinv = sm + L.off_vec;
  double* yv = dinv + Dp;
  double* spl = yv + Dp;
  double* alpha = spl + Dp;
  double* badv = alpha + Dp;            // C doubles: first failed pivot seen by each CTA (valid in CTA 0)
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gwarp = rank * CL_WARPS + warp;           // warp index inside the cluster
  constexpr int CW = C * CL_WARPS;
  const int gtid = rank * CL_WARPS * 32 + tid;
  constexpr int CT = C * CL_WARPS * 32;
  const int gid = lane >> 2, tig = lane & 3;
  const int yr = D & 7, yb = D >> 3;  // the bordered y row is row D = 8 yb + yr
  const double wS = 1.0 / p.S;
  const int oA0 = swz(gid, tig), oA1 = swz(gid, 4 + tig);
  const int oT0 = swz(tig, gid), oT1 = swz(4 + tig, gid);
  const int oC = swz(gid, 2 * tig);
  // per-cluster global slot: G (D rows of pitch ldgs: this sample's A o (fmean + sqrt(fvar) o eps)), then U1, U2
  // (sample sums of Gbar, Gbar o eps)
  double* U1 = p.ws + (long long)cid * p.ws_stride;
  double* U2 = U1 + R;
  const uint32_t pan_u32 = smem_u32(pan);
  const uint32_t full_bar = smem_u32(sm + L.off_bar), empty_bar = full_bar + 8 * CL_MAXST;
  uint32_t pbase = 0;   // panels this CTA has consumed so far (kernel lifetime; uniform over the CTA)
  constexpr int nbm = (NB + 1) / 2, nmac = nbm * (nbm + 1) / 2;
  const int npan = (nu + CL_PC - 1) / CL_PC;
  const int ycol = nu, ncb = nu / 8 + 1;  // column `nu` carries e_D (its solution is -alpha)
  double* alpha0 = cluster.map_shared_rank(alpha, 0);
  double* dinv0 = cluster.map_shared_rank(dinv, 0);
  double* badv0 = cluster.map_shared_rank(badv, 0);

  // tile (ib, jb) anywhere in the cluster / in this CTA (local row lr = ib / C)
  auto tile_any = [&](int ib, int jb) -> double* {
    return cluster.map_shared_rank(tiles + cl_loff<C>(ib, jb) * 64, ib % C);
  };
  auto tile_own = [&](int lr, int jb) -> double* {
    return tiles + (C * lr * (lr - 1) / 2 + lr * (rank + 1) + jb) * 64;
  };

  for (int d = tid; d < Dp; d += blockDim.x) spl[d] = d < D ? softplus_d(p.lam[d]) : 1.0;
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { cl_mbar_init(full_bar + 8 * i, 1); cl_mbar_init(empty_bar + 8 * i, CL_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
  }
  // panel g (sequence number over the kernel's lifetime) into its ring stage; issued by thread 0 only
  auto issue_panel = [&](uint32_t g, int k0, int z) {
    const uint32_t stage = g % ST, round = g / ST;
    if (round > 0) cl_mbar_wait(empty_bar + 8 * stage, (round & 1u) ^ 1u);
    cl_mbar_expect_tx(full_bar + 8 * stage, (uint32_t)(STAGE * 8));
    for (int r0 = 0; r0 < Dm; r0 += box_rows)
      cl_tma_load_3d(pan_u32 + (stage * STAGE + r0 * CL_PC) * 8, &tmG, full_bar + 8 * stage, k0, r0, z);
  };
  double glam_acc[2] = {0.0, 0.0};  // CTA 0, thread tid: d = tid and d = tid + 256
  cluster.sync();                   // every CTA of the cluster is resident before any remote access

  for (int n = n0 + cid; n < n0 + nc; n += ncl) {
    const double* fm = p.fmean + (long long)n * p.ldf;
    const double* fv = p.fvar + (long long)n * p.ldf;
    for (int d = tid; d < Dp; d += blockDim.x) yv[d] = d < D ? p.y[(long long)n * D + d] : 0.0;
    if (tid < C) badv[tid] = 0.0;
    double ve_acc = 0.0;

    for (int s = 0; s < p.S; ++s) {
      const double* ep = p.eps + ((long long)s * N + n) * R;
      // noise of the next sample (or of this cluster's next time step) into L2 through the bulk-copy engine
      {
        const double* nxt = nullptr;
        if (s + 1 < p.S) nxt = p.eps + ((long long)(s + 1) * N + n) * R;
        else if (n + ncl < N) nxt = p.eps + (long long)(n + ncl) * R;
        if (nxt != nullptr && (R & 1) == 0) {
          const long long bytes = (long long)R * 8;
          for (long long off = (long long)gtid * 4096; off < bytes; off += (long long)CT * 4096) {
            const long long len = bytes - off < 4096 ? bytes - off : 4096;
            bulk_prefetch_l2(reinterpret_cast<const char*>(nxt) + off, (unsigned)(len & ~15LL));
          }
        }
      }
      // G of this (sample, time step): slab zs of the chunk filled by wishart_g_fill_kernel
      const int zs = s * nc + (n - n0);
      const double* Gs = Gc + (long long)zs * D * ldgs;
      asm volatile("fence.proxy.async;" ::: "memory");   // generic writes to the ring (row-sum scratch) before the TMA writes
      cluster.sync();
      // ---------------- P1. Sigma tiles = G G^T ----------------
      for (int mbase = 0; mbase < nmac; mbase += CW * CL_MQ) {
        int Iq[CL_MQ], Jq[CL_MQ];
        bool on[CL_MQ];
        double acc[CL_MQ][2][2][2];
#pragma unroll
        for (int q = 0; q < CL_MQ; ++q) {
          const int m = mbase + gwarp + q * CW;
          on[q] = m < nmac;
          const int mm = on[q] ? m : 0;   // macro index -> (I, J), J <= I
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
          for (int j = 0; j < ST - 1 && j < npan; ++j) issue_panel(pbase + j, CL_PC * j, zs);
        }
        for (int pi = 0; pi < npan; ++pi) {
          const uint32_t g = pbase + pi, stage = g % ST;
          cl_mbar_wait(full_bar + 8 * stage, (g / ST) & 1u);
          // refill the stage every warp released one k-step ago
          if (tid == 0 && pi + ST - 1 < npan) issue_panel(g + ST - 1, CL_PC * (pi + ST - 1), zs);
          const double* P = pan + stage * STAGE;
#pragma unroll
          for (int q = 0; q < CL_MQ; ++q) {
            if (!on[q]) continue;
            const double* pa = P + (16 * Iq[q] + gid) * CL_PC + tig;
            const double* pb = P + (16 * Jq[q] + gid) * CL_PC + tig;
            const double a0 = pa[0], a1 = pa[8 * CL_PC], b0 = pb[0], b1 = pb[8 * CL_PC];
            dmma(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
            dmma(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
            dmma(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
            dmma(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
          }
          __syncwarp();
          if (lane == 0) cl_mbar_arrive(empty_bar + 8 * stage);
        }
        pbase += (uint32_t)npan;
        // + diag(softplus(lam)), bordered y row, identity padding; store the lower tiles into their owners
#pragma unroll
        for (int q = 0; q < CL_MQ; ++q) {
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
              *reinterpret_cast<double2*>(tile_any(ib, jb) + oC) = make_double2(c[0], c[1]);
            }
          }
        }
      }
      cluster.sync();

      // ---------------- P2. left-looking blocked Cholesky on the distributed tiles ----------------
      int bad = 0;
      constexpr int SL = ((NB + C - 1) / C + CL_WARPS - 1) / CL_WARPS;  // own block rows per warp
      for (int jb = 0; jb < NB; ++jb) {
        // (a) column jb of this CTA's block rows: S[ib][jb] -= sum_{kb < jb} L[ib][kb] L[jb][kb]^T
        {
          const int lr0 = jb > rank ? (jb - rank + C - 1) / C : 0;
          double c[SL][2], c2[SL][2];   // two partial sums (even / odd kb): half the dependent DMMA chain
          int lrs[SL];
          bool on[SL];
#pragma unroll
          for (int t = 0; t < SL; ++t) {
            lrs[t] = lr0 + warp + t * CL_WARPS;
            on[t] = lrs[t] * C + rank < NB;
            c2[t][0] = c2[t][1] = 0.0;
            if (on[t]) {
              const double2 v = *reinterpret_cast<const double2*>(tile_own(lrs[t], jb) + oC);
              c[t][0] = v.x; c[t][1] = v.y;
            } else { c[t][0] = c[t][1] = 0.0; }
          }
          if (on[0]) {
            const double* Tj = tile_any(jb, 0);   // block row jb is contiguous in its owner: tile (jb, kb) = Tj + 64 kb
            int kb = 0;
#pragma unroll 2
            for (; kb + 1 < jb; kb += 2) {
              const double nb0 = -Tj[kb * 64 + oA0], nb1 = -Tj[kb * 64 + oA1];
              const double nd0 = -Tj[kb * 64 + 64 + oA0], nd1 = -Tj[kb * 64 + 64 + oA1];
#pragma unroll
              for (int t = 0; t < SL; ++t) {
                if (on[t]) {
                  const double* t0 = tile_own(lrs[t], kb);
                  const double x0 = t0[oA0], x1 = t0[oA1], y0 = t0[64 + oA0], y1 = t0[64 + oA1];
                  dmma(c[t][0], c[t][1], x0, nb0);
                  dmma(c2[t][0], c2[t][1], y0, nd0);
                  dmma(c[t][0], c[t][1], x1, nb1);
                  dmma(c2[t][0], c2[t][1], y1, nd1);
                }
              }
            }
            if (kb < jb) {
              const double nb0 = -Tj[kb * 64 + oA0], nb1 = -Tj[kb * 64 + oA1];
#pragma unroll
              for (int t = 0; t < SL; ++t) {
                if (on[t]) {
                  const double* t0 = tile_own(lrs[t], kb);
                  dmma(c[t][0], c[t][1], t0[oA0], nb0);
                  dmma(c[t][0], c[t][1], t0[oA1], nb1);
                }
              }
            }
#pragma unroll
            for (int t = 0; t < SL; ++t)
              if (on[t])
                *reinterpret_cast<double2*>(tile_own(lrs[t], jb) + oC) = make_double2(c[t][0] + c2[t][0], c[t][1] + c2[t][1]);
          }
        }
        __syncthreads();
        // (b) the owner of block row jb: diagonal tile -> I_jj = L_jj^-1 (in place), dinv (kept in CTA 0)
        if (rank == jb % C && warp == 0) {
          bad = factor_diag_swz(tile_own(jb / C, jb), dinv + 8 * jb, lane, (jb == yb) ? yr : -1, 8 * jb, D, bad);
          __syncwarp();
          if (rank != 0 && lane < 8) dinv0[8 * jb + lane] = dinv[8 * jb + lane];
        }
        cluster.sync();
        // (c) panel: L[ib][jb] = S[ib][jb] I_jj^T for this CTA's block rows ib > jb
        {
          const int lr0 = jb + 1 > rank ? (jb + 1 - rank + C - 1) / C : 0;
          const double* Ij = tile_any(jb, jb);
          double r[SL][2];
          int lrs[SL];
          bool on[SL];
          bool any = false;
#pragma unroll
          for (int t = 0; t < SL; ++t) {
            lrs[t] = lr0 + warp + t * CL_WARPS;
            on[t] = lrs[t] * C + rank < NB;
            any = any || on[t];
            r[t][0] = r[t][1] = 0.0;
          }
          if (any) {
            const double b0 = Ij[oA0], b1 = Ij[oA1];
#pragma unroll
            for (int t = 0; t < SL; ++t) {
              if (on[t]) {
                const double* X = tile_own(lrs[t], jb);
                const double x0 = X[oA0], x1 = X[oA1];
                dmma(r[t][0], r[t][1], x0, b0);
                dmma(r[t][0], r[t][1], x1, b1);
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < SL; ++t)
            if (on[t]) *reinterpret_cast<double2*>(tile_own(lrs[t], jb) + oC) = make_double2(r[t][0], r[t][1]);
        }
        cluster.sync();
      }
      if (warp == 0 && lane == 0 && bad) badv0[rank] = (double)bad;

      // ---------------- P4. column blocks: Q'' = L^-T (L^-1 G with row D negated) ----------------
      double* rsw = pan + warp * Dp;  // this warp's row sums of Q'' o G (the panel buffers are free now)
      for (int d = lane; d < Dp; d += 32) rsw[d] = 0.0;
      __syncwarp();
      // forward only: just the e_D column block (alpha -> quadratic form), on warp 0 of CTA 0
      const int cb_first = p.need_grad ? gwarp : (gwarp == 0 ? ncb - 1 : ncb);
      for (int cb = cb_first; cb < ncb; cb += CW) {
        ClCtx<C> c;
#pragma unroll
        for (int r = 0; r < C; ++r) c.rb[r] = cluster.map_shared_rank(tiles, r);
        c.G = Gs; c.ep = ep; c.D = D; c.nu = nu; c.ldgs = ldgs; c.colC = 8 * cb + 2 * tig;
        c.ycol = ycol; c.gid = gid; c.tig = tig; c.oA0 = oA0; c.oA1 = oA1; c.oT0 = oT0; c.oT1 = oT1;
        c.kb_start = 8 * cb >= nu ? yb : 0;   // the block past the last column of G holds only e_D (row D = block yb)
        double acc[NB][2], gq[CL_PF][2];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) acc[ib][0] = acc[ib][1] = 0.0;
#pragma unroll
        for (int t = 0; t < CL_PF; ++t) {
          gq[t][0] = gq[t][1] = 0.0;
          if (t < NB) cl_load_g<C>(c, t, gq[t]);
        }
        double ik[2];
        ik[0] = cl_ldv(cl_tile<C, 0, 0>(c) + oA0);
        ik[1] = cl_ldv(cl_tile<C, 0, 0>(c) + oA1);
        cl_fwd<NB, C, 0>(acc, gq, ik, c);
        // row D of H is -alpha^T G: negate it so that the back substitution folds in - alpha (alpha^T G)
        // (the e_D column keeps its sign: its solution is column D of L^-T = (-alpha; 1))
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
          if (ib == yb && gid == yr) {
            if (c.colC != ycol) acc[ib][0] = -acc[ib][0];
            if (c.colC + 1 != ycol) acc[ib][1] = -acc[ib][1];
          }
        }
        ik[0] = cl_ldv(cl_tile<C, NB - 1, NB - 1>(c) + oT0);
        ik[1] = cl_ldv(cl_tile<C, NB - 1, NB - 1>(c) + oT1);
        cl_bwd<NB, C, NB - 1>(acc, ik, c);
        cl_epilogue<NB, C, 0>(acc, c, U1, U2, alpha0, rsw, wS, p.need_grad != 0, s == 0);
      }
      cluster.sync();

      // ---------------- P5. lam_bar, ve of this sample (CTA 0) ----------------
      if (rank == 0) {
        if (p.need_grad) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int d = tid + h * CL_WARPS * 32;
            if (d < D) {
              double sr = 0.0;
              for (int r2 = 0; r2 < C; ++r2) {
                const double* pr = cluster.map_shared_rank(pan, r2);
                for (int w2 = 0; w2 < CL_WARPS; ++w2) sr += pr[w2 * Dp + d];
              }
              glam_acc[h] += sigmoid_d(p.lam[d]) * wS * (-0.5) * (1.0 - alpha[d] * yv[d] - sr) / spl[d];
            }
          }
        }
        if (warp == 0) {
          double q = 0.0, ld = 0.0;
          for (int j = lane; j < D; j += 32) { q += yv[j] * alpha[j]; ld += log(dinv[j]); }
          q = warp_sum(q);
          ld = warp_sum(ld);
          ve_acc += wS * (-0.5 * D * 1.8378770664093453 + ld - 0.5 * q);
        }
      }
      __threadfence();
      cluster.sync();
    }  // samples

    // mu_bar = a U1, v_bar = a U2 / (2 sd), A_bar partial (per cluster) += fm U1 + sd U2
    if (p.need_grad) {
      const bool first = (n == cid);
      for (int r = gtid; r < R; r += CT) {
        const double a = p.a_mode == 0 ? __ldg(p.A + r) : __ldg(p.A + r % nu);
        const double f = fm[r], v = fv[r], rsq = rsqrt(v), sd = v * rsq;
        const double u1 = U1[r], u2 = U2[r];
        p.g_fmean[(long long)n * p.ldg + r] = a * u1;
        p.g_fvar[(long long)n * p.ldg + r] = a * u2 * (0.5 * rsq);
        const long long oa = (long long)cid * R + r;
        p.gA_part[oa] = (first ? 0.0 : p.gA_part[oa]) + f * u1 + sd * u2;
      }
    }
    if (rank == 0 && tid == 0) {
      int b = 0;
      for (int r2 = 0; r2 < C; ++r2) {
        const int br = (int)badv[r2];
        if (br && (!b || br < b)) b = br;
      }
      p.ve[n] = b ? nan("") : ve_acc;
      p.info[n] = b;
    }
    __threadfence();
    cluster.sync();   // the slot (G, U1, U2) and the flags are re-used by the next time step
  }
  if (p.need_grad && rank == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int d = tid + h * CL_WARPS * 32;
      if (d < D) {   // per-cluster partial, accumulated over the chunks in launch order
        const long long o = (long long)cid * D + d;
        p.glam_part[o] = (n0 == 0 ? 0.0 : p.glam_part[o]) + glam_acc[h];
      }
    }
  }
}

// Scratch layout: [U1, U2 slots: 2 R doubles per cluster, at most sms / 2 clusters][G of one chunk of time steps: (S, nc, D)
// rows of pitch ldgs = nu rounded up to 4 (32-byte aligned rows for the tensor map)], filled by wishart_g_fill_kernel.  The
// chunk budget is 8 samples x one time step per cluster of two; other sample counts get proportionally more / fewer.
inline long long cluster_slot_doubles(int D, int nu) { return 2LL * D * nu; }
inline long long cluster_gchunk_doubles(int D, int nu, int sms) { return 8LL * (sms / 2) * D * ((nu + 3) & ~3); }




typedef CUresult (*ClEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline ClEncodeTiledFn cl_encode_fn() {
  static ClEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (ClEncodeTiledFn)ptr;
  }
  return fn;
}

template <int NB, int C>
int launch_cluster_nb(const LikParams& p, cudaStream_t stream, int* part_rows) {
  constexpr ClLayout L = cl_layout<C>(NB);
  static_assert((long long)L.total * 8 + 64 <= 227 * 1024, "tile store + ring + vectors exceed the shared memory of an SM");
  const long long smem = (long long)L.total * 8;
  auto kern = wishart_loglik_cluster_kernel<NB, C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(sms / C * C), 1, 1);
  cfg.blockDim = dim3(cl_warps(NB, C) * 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int ncl = 0;
  e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
  if (e != cudaSuccess) return (int)e;
  if (ncl < 1) return FCEST_ERR_UNSUPPORTED;
  if (ncl > sms / C) ncl = sms / C;
  if (ncl > p.N) ncl = p.N;
  cfg.gridDim = dim3((unsigned)(ncl * C), 1, 1);
  // chunks of time steps: G of a chunk is filled by the streaming kernel, then the persistent cluster kernel walks it
  const int ldgs = (p.nu + 3) & ~3;
  const long long slab = (long long)p.D * ldgs;
  long long NC = cluster_gchunk_doubles(p.D, p.nu, sms) / (slab * p.S);
  if (NC < 1) return FCEST_ERR_UNSUPPORTED;
  if (NC > p.N) NC = p.N;
  if (ncl > NC) ncl = (int)NC;                  // the same clusters (partial rows) in every chunk
  NC = NC / ncl * ncl;                          // whole rounds of the persistent clusters per chunk (no tail inside a chunk)
  cfg.gridDim = dim3((unsigned)(ncl * C), 1, 1);
  double* Gc = p.ws + 2LL * p.D * p.nu * (sms / 2);
  // tensor map over the chunk's G: (nu columns, D rows, S x NC slabs); a box is 4 columns x box_rows rows of one slab,
  // rows >= D and columns >= nu are zero-filled by the copy engine
  ClEncodeTiledFn enc = cl_encode_fn();
  if (!enc) return FCEST_ERR_UNSUPPORTED;
  const int box_rows = L.Dm > 256 ? L.Dm / 2 : L.Dm;
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)p.nu, (cuuint64_t)p.D, (cuuint64_t)(p.S * NC)};
  const cuuint64_t strides[2] = {(cuuint64_t)ldgs * 8, (cuuint64_t)slab * 8};
  const cuuint32_t box[3] = {(cuuint32_t)CL_PC, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)Gc, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return FCEST_ERR_ARGUMENT;
  for (int n0 = 0; n0 < p.N; n0 += (int)NC) {
    const int nc = p.N - n0 < NC ? p.N - n0 : (int)NC;
    const int rc = launch_loglik_g_fill(p, n0, nc, ldgs, Gc, stream);
    if (rc) return rc;
    const double* Gcc = Gc;
    e = cudaLaunchKernelEx(&cfg, kern, tm, p, ldgs, box_rows, Gcc, n0, nc);
    if (e != cudaSuccess) return (int)e;
    if (n0 + nc < p.N) FCEST_CHECK_LAUNCH();   // the caller counts / checks the last launch
  }
  *part_rows = ncl;
  return 0;
}


}  // namespace fcest
