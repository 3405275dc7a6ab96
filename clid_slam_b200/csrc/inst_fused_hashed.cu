// train_fused_l1_kernel instantiations, search kind: hashed (launchers.cuh).
#include "launchers.cuh"

namespace clid {
int dispatch_train_fused_hashed(const TrainFusedParams& p, cudaStream_t stream) { return dispatch_train_fused_t<kSearchHashed>(p, stream); }
}  // namespace clid
