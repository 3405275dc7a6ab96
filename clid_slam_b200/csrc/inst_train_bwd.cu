// train_backward_l1_kernel instantiations (split backward of the mapping iteration).
#include "launch.h"
#include "train.cuh"

namespace clid {

template <int H, int K>
static int launch_train_backward(const TrainBwdParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kWarps = kBwdThreads / 32;
  size_t smem = (MlpLayout<H, 1>::kFloats + kWarps * 32 * kInPad + kWarps * 32 * (H / 32) + kWarps * H * kInPad) * sizeof(float);
  auto kern = train_backward_l1_kernel<H, K>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kBwdThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  int64_t want = (p.n + kBwdThreads - 1) / kBwdThreads;
  int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, kBwdThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "train_backward_l1_kernel launch");
  return CLID_OK;
}


int dispatch_train_backward(const TrainBwdParams& p, cudaStream_t stream) {
  const int H = p.dec.hidden_dim;
  if (p.dec.levels != 1 || (H != 32 && H != 64 && H != 128))
    return set_error(CLID_EUNSUPPORTED, "fused backward is compiled for one hidden level with H in {32,64,128}; got %d x %d",
                     H, p.dec.levels);
  const bool k6 = p.map.knn <= 6;
  if (H == 64) return k6 ? launch_train_backward<64, 6>(p, stream) : launch_train_backward<64, 8>(p, stream);
  if (H == 32) return k6 ? launch_train_backward<32, 6>(p, stream) : launch_train_backward<32, 8>(p, stream);
  return k6 ? launch_train_backward<128, 6>(p, stream) : launch_train_backward<128, 8>(p, stream);
}

}  // namespace clid
