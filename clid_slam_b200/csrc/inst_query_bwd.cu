// query_backward_kernel instantiations (autograd path of the unfused query).
#include "launch.h"
#include "query_bwd.cuh"

namespace clid {

template <bool kSecond>
static int launch(const QueryBwdParams& p, int grid, cudaStream_t stream, const char* what) {
  if (p.map.knn <= 6) query_backward_kernel<6, kSecond><<<grid, 128, 0, stream>>>(p);
  else query_backward_kernel<8, kSecond><<<grid, 128, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return CLID_OK;
}

int launch_query_backward_first(const QueryBwdParams& p, int grid, cudaStream_t stream) {
  return launch<false>(p, grid, stream, "clid_query_backward");
}
int launch_query_backward_second(const QueryBwdParams& p, int grid, cudaStream_t stream) {
  return launch<true>(p, grid, stream, "clid_query_backward_backward");
}

}  // namespace clid
