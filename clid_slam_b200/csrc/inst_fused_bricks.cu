// train_fused_l1_kernel instantiations, search kind: bricks (launchers.cuh).
#include "launchers.cuh"

namespace clid {
int dispatch_train_fused_bricks(const TrainFusedParams& p, cudaStream_t stream) { return dispatch_train_fused_t<kSearchBricks>(p, stream); }
}  // namespace clid
