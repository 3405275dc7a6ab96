// query_forward_tc_kernel instantiations (query_fwd_tc.cuh): brick-index search, 64 x 1 decoder on tcgen05.
#include "launch.h"
#include "search.cuh"
#if CLID_QUERY_THREADS == 128
#include "query_fwd_tc.cuh"

namespace clid {

template <int K>
static int launch_query_tc(const QueryParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr size_t smem = query_tc_smem_bytes();
  auto kern = query_forward_tc_kernel<K>;
  // The occupancy API reports ONE resident CTA per SM for any kernel that uses tcgen05 (it cannot see how many TMEM
  // columns the kernel will allocate); the hardware does co-schedule CTAs by registers / shared memory and
  // tcgen05.alloc hands each its columns.  The kernel is built for 4 CTAs per SM: __launch_bounds__(128, 4) caps the
  // registers, 4 x 43 KB of shared memory fit, and 4 x 128 TMEM columns are exactly the SM's 512 (measured: 592 CTAs
  // resident at once, same duration as the 4-CTA FMA kernel; a 148-CTA grid takes twice as long).
  constexpr int blocks_per_sm = 512 / tc::kTmemCols;
  static_assert(blocks_per_sm >= CLID_QUERY_MIN_BLOCKS, "the TMEM budget (4 x 128 columns) must cover the register-budgeted CTAs per SM");
  static_assert(smem * blocks_per_sm <= 220 * 1024, "shared memory of the resident CTAs");
  const int64_t want = (p.n + kQueryThreads - 1) / kQueryThreads;
  const int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  kern<<<(int)(want < cap ? want : cap), kQueryThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "query_forward_tc_kernel launch");
  return CLID_OK;
}

int dispatch_query_tc(const QueryParams& p, cudaStream_t stream) {
  if (p.map.knn <= 6) return launch_query_tc<6>(p, stream);
  return launch_query_tc<8>(p, stream);
}

}  // namespace clid
#else  // experiment builds with another CTA size: the tensor-core kernel maps one TMEM lane per thread of a 128-thread CTA
namespace clid {
int dispatch_query_tc(const QueryParams&, cudaStream_t) {
  return set_error(CLID_EUNSUPPORTED, "query_forward_tc_kernel is built for 128-thread CTAs");
}
}  // namespace clid
#endif
