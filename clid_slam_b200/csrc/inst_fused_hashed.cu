// train_fused_l1_kernel instantiations, reference hash table search.
#include "launch.h"
#include "train_fused.cuh"

namespace clid {

template <int H, int K, bool kBricks, bool kNumerical, bool kFoldOut>
static int launch_train_fused(const TrainFusedParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kWarps = kFusedThreads / 32;
  constexpr int kSearchFloats = kBricks ? (2 * 64 * kBrickSlots + (int)(sizeof(BrickScratch) / sizeof(float)))
                                        : 2 * CLID_MAX_KC;
  size_t smem = (MlpLayout<H, 1>::kFloats + kSearchFloats +
                 (kFoldOut ? 0 : kWarps * 32 * kInPad + kWarps * 32 * (H / 32) + kWarps * H * kInPad)) * sizeof(float);
  auto kern = train_fused_l1_kernel<H, K, kBricks, kNumerical, kFoldOut>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kFusedThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  const int64_t per_tile = kNumerical ? kNumTileSamples : 32;  // base samples per 32-lane tile
  const int64_t tiles = (p.n + per_tile - 1) / per_tile;
  int64_t want = (tiles * 32 + kFusedThreads - 1) / kFusedThreads;
  int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, kFusedThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "train_fused_l1_kernel launch");
  return CLID_OK;
}

int dispatch_train_fused_hashed(const TrainFusedParams& p, cudaStream_t stream) {
  const int H = p.dec.hidden_dim;
  if (p.dec.levels != 1 || (H != 32 && H != 64 && H != 128))
    return set_error(CLID_EUNSUPPORTED, "fused training is compiled for one hidden level with H in {32,64,128}; got %d x %d",
                     H, p.dec.levels);
  if (p.map.knn > 6) return set_error(CLID_EUNSUPPORTED, "fused training is compiled for query_nn_k <= 6");
  const bool num = p.num_eps > 0.f;  // set by clid_train_fused only in numerical mode
#define CLID_FUSED(HH) \
  (p.fold_rows ? (num ? launch_train_fused<HH, 6, false, true, true>(p, stream) : launch_train_fused<HH, 6, false, false, true>(p, stream)) \
               : (num ? launch_train_fused<HH, 6, false, true, false>(p, stream) : launch_train_fused<HH, 6, false, false, false>(p, stream)))
  if (H == 64) return CLID_FUSED(64);
  if (H == 32) return CLID_FUSED(32);
  return CLID_FUSED(128);
#undef CLID_FUSED
}

}  // namespace clid
