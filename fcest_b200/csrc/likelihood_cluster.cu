// Cluster-tiled Wishart likelihood for 208 <= D <= 407: dispatch by row-block capacity and the instantiations for
// capacities 30 and 34 (the kernel lives in likelihood_cluster_impl.cuh; capacities 38 / 42 and 46 / 51 are instantiated in
// likelihood_cluster_b.cu / likelihood_cluster_c.cu).
#include "likelihood_cluster_impl.cuh"

namespace fcest {

int launch_cluster_38(const LikParams& p, cudaStream_t stream, int* part_rows);
int launch_cluster_42(const LikParams& p, cudaStream_t stream, int* part_rows);
int launch_cluster_46(const LikParams& p, cudaStream_t stream, int* part_rows);
int launch_cluster_51(const LikParams& p, cudaStream_t stream, int* part_rows);

static constexpr long long kSmemLimitC = 227 * 1024;

// Row-block capacities instantiated (the padded system has 8 NB rows; rows > D are identity) and the cluster size of
// each: pairs while half of the tiles fit one SM (D <= 303: half of the operand reads stay local and all 148 SMs are used;
// 24.8 ms against 34.9 ms for clusters of four at D = 256, N = 296, S = 8), clusters of four above (132 SMs can host them).
static const int kCaps[] = {30, 34, 38, 42, 46, 51};
static int cluster_capacity(int D) {
  const int need = D / 8 + 1;
  for (int c : kCaps) if (need <= c) return c;
  return 0;
}
static int cluster_size(int NB) { return NB <= 38 ? 2 : 4; }

bool loglik_cluster_supported(int D, int nu) {
  if (D < 8 || nu < 1 || nu > 407) return false;
  const int NB = cluster_capacity(D);
  if (!NB) return false;
  const long long total = cluster_size(NB) == 2 ? cl_layout<2>(NB).total : cl_layout<4>(NB).total;
  return total * 8 + 64 <= kSmemLimitC;
}

long long loglik_cluster_scratch_bytes(int D, int nu) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return (long long)sizeof(double) * (cluster_slot_doubles(D, nu) * (sms / 2) + cluster_gchunk_doubles(D, nu, sms));
}

int launch_loglik_cluster(LikParams p, cudaStream_t stream, int* part_rows) {
  if (!loglik_cluster_supported(p.D, p.nu)) return -1;
  if (p.ws == nullptr) return FCEST_ERR_ARGUMENT;
  p.ws_stride = cluster_slot_doubles(p.D, p.nu);
  switch (cluster_capacity(p.D)) {
    case 30: return launch_cluster_nb<30, 2>(p, stream, part_rows);
    case 34: return launch_cluster_nb<34, 2>(p, stream, part_rows);
    case 38: return launch_cluster_38(p, stream, part_rows);
    case 42: return launch_cluster_42(p, stream, part_rows);
    case 46: return launch_cluster_46(p, stream, part_rows);
    case 51: return launch_cluster_51(p, stream, part_rows);
    default: return -1;
  }
}

}  // namespace fcest
