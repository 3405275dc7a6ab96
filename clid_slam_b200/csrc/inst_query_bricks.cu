// query_forward_kernel instantiations, search kind: bricks (launchers.cuh).
#include "launchers.cuh"

namespace clid {
int dispatch_query_bricks(const QueryParams& p, bool has_dec, cudaStream_t stream) { return dispatch_query_t<kSearchBricks>(p, has_dec, stream); }
}  // namespace clid
