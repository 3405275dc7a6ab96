// sdf_tile_kernel / decoder_grad_kernel instantiations and launch geometry.
#include "launch.h"
#include "tile_kernel.cuh"

namespace clid {

bool tile_supported(const ClidMap& map, const ClidDecoder& dec, const ClidBricks& bricks) {
  const int H = dec.hidden_dim;
  return dec.levels == 1 && (H == 32 || H == 64 || H == 128) && map.knn <= 6 && bricks.span == 2 && bricks.apron >= 1 &&
         bricks.reach <= 2;
}

template <int H, int kMode>
static int launch_tile_t(TileParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  const size_t smem = (TileDec<H>::kFloats + 2 * 64 * 8) * sizeof(float) + (size_t)kTileWarps * kParkGroups * 32 * sizeof(float4);
  auto kern = sdf_tile_kernel<H, 6, kMode>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(carveout)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kTileThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  const int64_t per_tile = kMode == kTileTrainNumerical ? kNumTile : 32;
  const int64_t tiles = (p.n + per_tile - 1) / per_tile;
  const int64_t want = (tiles + kTileWarps - 1) / kTileWarps;
  const int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid;
  if (want <= cap) {
    grid = (int)want;
    p.work_counter = nullptr;  // one tile per warp: static schedule
  } else {
    grid = (int)cap;
    p.work_counter = p.map.work_counter;  // persistent CTAs, tiles drawn with an atomic (NULL: round-robin)
  }
  kern<<<grid, kTileThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "sdf_tile_kernel launch");
  return CLID_OK;
}

template <int H>
static int launch_tile_h(TileParams& p, int mode, cudaStream_t stream) {
  if (mode == kTileInfer) return launch_tile_t<H, kTileInfer>(p, stream);
  if (mode == kTileTrainAnalytic) return launch_tile_t<H, kTileTrainAnalytic>(p, stream);
  return launch_tile_t<H, kTileTrainNumerical>(p, stream);
}

int launch_tile(TileParams& p, int mode, cudaStream_t stream) {
  const int H = p.dec.hidden_dim;
  if (H == 64) return launch_tile_h<64>(p, mode, stream);
  if (H == 32) return launch_tile_h<32>(p, mode, stream);
  if (H == 128) return launch_tile_h<128>(p, mode, stream);
  return set_error(CLID_EUNSUPPORTED, "tile kernels are compiled for H in {32,64,128}; got %d", H);
}

template <int H>
static int launch_decoder_grad_t(const DecoderGradParams& p, int sm_count, cudaStream_t stream) {
  auto kern = decoder_grad_kernel<H>;
  const int64_t tiles = p.n_rows >> 5;
  const int64_t want = (tiles + DgSmem<H>::kWarps - 1) / DgSmem<H>::kWarps;
  const int grid = (int)(want < sm_count ? want : sm_count);
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DgSmem<H>::kBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decoder_grad_kernel)");
    configured = true;
  }
  kern<<<grid, DgSmem<H>::kWarps * 32, DgSmem<H>::kBytes, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "decoder_grad_kernel launch");
  return CLID_OK;
}

int launch_decoder_grad(const DecoderGradParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  const int64_t tiles = p.n_rows >> 5;
  if (tiles == 0) return CLID_OK;
  const int H = p.dec.hidden_dim;
  if (H == 64) return launch_decoder_grad_t<64>(p, info.sm_count, stream);
  if (H == 32) return launch_decoder_grad_t<32>(p, info.sm_count, stream);
  if (H == 128) return launch_decoder_grad_t<128>(p, info.sm_count, stream);
  return set_error(CLID_EUNSUPPORTED, "decoder_grad_kernel is compiled for H in {32,64,128}; got %d", H);
}

}  // namespace clid
