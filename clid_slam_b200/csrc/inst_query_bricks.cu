// query_forward_kernel instantiations, brick index search.
#include "launch.h"
#include "query_fwd.cuh"

namespace clid {

template <int H, int L, int K, bool kBricks>
static int launch_query(const QueryParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kThreads = kQueryThreads;
  constexpr int kDecFloats = H > 0 ? MlpLayout<(H > 0 ? H : 4), (H > 0 ? L : 1)>::kFloats : 0;
  size_t smem = kDecFloats * sizeof(float) +
                (kBricks ? 64 * kBrickSlots * sizeof(uint64_t) + sizeof(BrickScratch) : CLID_MAX_KC * sizeof(int64_t));
  auto kern = query_forward_kernel<H, L, K, kBricks>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  int64_t want = (p.n + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, kThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "query_forward_kernel launch");
  return CLID_OK;
}


template <int H, int L>
static int dispatch_k(const QueryParams& p, cudaStream_t stream) {
  if (p.map.knn <= 6) return launch_query<H, L, 6, true>(p, stream);
  return launch_query<H, L, 8, true>(p, stream);
}

int dispatch_query_bricks(const QueryParams& p, bool has_dec, cudaStream_t stream) {
  if (!has_dec) return dispatch_k<0, 1>(p, stream);
  const int H = p.dec.hidden_dim, L = p.dec.levels;
  if (L == 1 && H == 64) return dispatch_k<64, 1>(p, stream);
  if (L == 1 && H == 32) return dispatch_k<32, 1>(p, stream);
  if (L == 1 && H == 128) return dispatch_k<128, 1>(p, stream);
  if (L == 2 && H == 32) return dispatch_k<32, 2>(p, stream);
  if (L == 2 && H == 64) return dispatch_k<64, 2>(p, stream);
  return set_error(CLID_EUNSUPPORTED,
                   "decoder %d x %d not in the fused kernel set {64x1, 32x1, 128x1, 32x2, 64x2}; "
                   "use the unfused query + torch decoder path", H, L);
}

}  // namespace clid
