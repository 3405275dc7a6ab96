// Launch geometry and shape dispatch of the template kernels, shared by the inst_*.cu translation
// units (one per search kind, so nvcc compiles the instantiations in parallel).
#pragma once
#include <cstdlib>
#include "launch.h"
#include "query_fwd.cuh"
#include "train_fused.cuh"

namespace clid {

// The gathers of these kernels live in L1 (every byte of shared memory the driver reserves beyond what the resident
// CTAs use is L1 they lose): ask for the smallest carve-out that holds `blocks` CTAs (+1 KB each the system reserves).
// CLID_CARVEOUT (percent, developer knob) overrides.
static int set_carveout(const void* kern, size_t smem, int blocks) {
  int pct = -1;
  if (const char* env = getenv("CLID_CARVEOUT")) pct = atoi(env);
  if (pct < 0) {
    const size_t need = (smem + 1024) * (size_t)blocks;
    pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(carveout)");
  return CLID_OK;
}

template <int H, int L, int K, int kSearch>
static int launch_query(const QueryParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kThreads = kQueryThreads;
  constexpr int kDecFloats = H > 0 ? MlpLayout<(H > 0 ? H : 4), (H > 0 ? L : 1)>::kFloats : 0;
  const size_t smem = (size_t)(kDecFloats + search_smem_floats<kSearch>()) * sizeof(float);
  auto kern = query_forward_kernel<H, L, K, kSearch>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (int rc = set_carveout(reinterpret_cast<const void*>(kern), smem, blocks_per_sm)) return rc;
  }
  int64_t want = (p.n + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, kThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "query_forward_kernel launch");
  return CLID_OK;
}

template <int H, int L, int kSearch>
static int dispatch_query_k(const QueryParams& p, cudaStream_t stream) {
  if (p.map.knn <= 6) return launch_query<H, L, 6, kSearch>(p, stream);
  return launch_query<H, L, 8, kSearch>(p, stream);
}

template <int kSearch>
static int dispatch_query_t(const QueryParams& p, bool has_dec, cudaStream_t stream) {
  if (!has_dec) return dispatch_query_k<0, 1, kSearch>(p, stream);
  const int H = p.dec.hidden_dim, L = p.dec.levels;
  if (L == 1 && H == 64) return dispatch_query_k<64, 1, kSearch>(p, stream);
  if (L == 1 && H == 32) return dispatch_query_k<32, 1, kSearch>(p, stream);
  if (L == 1 && H == 128) return dispatch_query_k<128, 1, kSearch>(p, stream);
  if (L == 2 && H == 32) return dispatch_query_k<32, 2, kSearch>(p, stream);
  if (L == 2 && H == 64) return dispatch_query_k<64, 2, kSearch>(p, stream);
  return set_error(CLID_EUNSUPPORTED,
                   "decoder %d x %d not in the fused kernel set {64x1, 32x1, 128x1, 32x2, 64x2}; "
                   "use the unfused query + torch decoder path", H, L);
}

template <int H, int L, int K, int kSearch, bool kNumerical, bool kFoldOut>
static int launch_train_fused(const TrainFusedParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kWarps = kFusedThreads / 32;
  const size_t smem = (size_t)(MlpLayout<H, L>::kFloats + search_smem_floats<kSearch>() +
                               (kFoldOut ? 0 : kWarps * 32 * kInPad + kWarps * 32 * (H / 32) + kWarps * H * kInPad)) * sizeof(float);
  auto kern = train_fused_kernel<H, L, K, kSearch, kNumerical, kFoldOut>;
  static thread_local int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kFusedThreads, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (int rc = set_carveout(reinterpret_cast<const void*>(kern), smem, blocks_per_sm)) return rc;
  }
  const int64_t per_tile = kNumerical ? kNumTileSamples : 32;  // base samples per 32-lane tile
  const int64_t tiles = (p.n + per_tile - 1) / per_tile;
  int64_t want = (tiles * 32 + kFusedThreads - 1) / kFusedThreads;
  int64_t cap = (int64_t)info.sm_count * blocks_per_sm;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, kFusedThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "train_fused_kernel launch");
  return CLID_OK;
}

template <int kSearch>
static int dispatch_train_fused_t(const TrainFusedParams& p, cudaStream_t stream) {
  const int H = p.dec.hidden_dim;
  if (p.map.knn > 6) return set_error(CLID_EUNSUPPORTED, "fused training is compiled for query_nn_k <= 6");
  const bool num = p.num_eps > 0.f;  // set by clid_train_fused only in numerical mode
  if (p.dec.levels == 2 && H == 32) {
    // two hidden levels: the decoder gradient always leaves the kernel as rows (a frozen decoder still gets its scratch:
    // the rows park h1 / h2 between the forward and the finish)
    if (!p.fold_rows) return set_error(CLID_EINVAL, "a two-level decoder needs ClidTrainFusedArgs.scratch");
    return num ? launch_train_fused<32, 2, 6, kSearch, true, true>(p, stream) : launch_train_fused<32, 2, 6, kSearch, false, true>(p, stream);
  }
  if (p.dec.levels != 1 || (H != 32 && H != 64 && H != 128))
    return set_error(CLID_EUNSUPPORTED, "fused training is compiled for decoders 32x1, 64x1, 128x1 and 32x2; got %d x %d",
                     H, p.dec.levels);
#define CLID_FUSED(HH) \
  (p.fold_rows ? (num ? launch_train_fused<HH, 1, 6, kSearch, true, true>(p, stream) : launch_train_fused<HH, 1, 6, kSearch, false, true>(p, stream)) \
               : (num ? launch_train_fused<HH, 1, 6, kSearch, true, false>(p, stream) : launch_train_fused<HH, 1, 6, kSearch, false, false>(p, stream)))
  if (H == 64) return CLID_FUSED(64);
  if (H == 32) return CLID_FUSED(32);
  return CLID_FUSED(128);
#undef CLID_FUSED
}

}  // namespace clid
