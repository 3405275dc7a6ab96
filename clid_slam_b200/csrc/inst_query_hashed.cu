// query_forward_kernel instantiations, search kind: hashed (launchers.cuh).
#include "launchers.cuh"

namespace clid {
int dispatch_query_hashed(const QueryParams& p, bool has_dec, cudaStream_t stream) { return dispatch_query_t<kSearchHashed>(p, has_dec, stream); }
}  // namespace clid
