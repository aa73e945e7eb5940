// Cluster-tiled Wishart likelihood: instantiations for row-block capacities 46 and 51 (clusters of four).
#include "likelihood_cluster_impl.cuh"

namespace fcest {
int launch_cluster_46(const LikParams& p, cudaStream_t stream, int* part_rows) { return launch_cluster_nb<46, 4>(p, stream, part_rows); }
int launch_cluster_51(const LikParams& p, cudaStream_t stream, int* part_rows) { return launch_cluster_nb<51, 4>(p, stream, part_rows); }
}  // namespace fcest
