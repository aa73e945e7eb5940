// Cluster-tiled Wishart likelihood: instantiations for row-block capacities 38 (CTA pairs) and 42 (clusters of four).
#include "likelihood_cluster_impl.cuh"

namespace fcest {
int launch_cluster_38(const LikParams& p, cudaStream_t stream, int* part_rows) { return launch_cluster_nb<38, 2>(p, stream, part_rows); }
int launch_cluster_42(const LikParams& p, cudaStream_t stream, int* part_rows) { return launch_cluster_nb<42, 4>(p, stream, part_rows); }
}  // namespace fcest
