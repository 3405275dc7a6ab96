// C ABI of libclid_sdf.so: argument validation, kernel selection, launches.
// Declarations and the reference functions each entry point replaces: include/clid_sdf.h
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "query_bwd.cuh"
#include "query_fwd.cuh"
#include "train.cuh"
#include "train_fused.cuh"

namespace clid {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
  return set_error(CLID_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceInfo {
  int sm_count = 0;
};

static int device_info(DeviceInfo* info) {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo cached;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != cached_dev) {
    e = cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    cached_dev = dev;
  }
  *info = cached;
  return CLID_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_decoder(const ClidDecoder* d) {
  if (d->in_dim != kIn) return set_error(CLID_EUNSUPPORTED, "decoder in_dim %d (only %d = feature_dim 8 + 3)", d->in_dim, kIn);
  if (d->levels < 1 || d->levels > CLID_MAX_LEVELS) return set_error(CLID_EINVAL, "decoder levels %d", d->levels);
  for (int l = 0; l < d->levels; ++l)
    if (!d->weight[l]) return set_error(CLID_EINVAL, "decoder weight[%d] is NULL", l);
  if (!d->out_weight) return set_error(CLID_EINVAL, "decoder out_weight is NULL");
  return CLID_OK;
}

static int check_map(const ClidMap* m, uint32_t flags) {
  if (m->feature_dim != kFeat) return set_error(CLID_EUNSUPPORTED, "feature_dim %d (only %d)", m->feature_dim, kFeat);
  if (m->knn < 1 || m->knn > CLID_MAX_KNN) return set_error(CLID_EINVAL, "knn %d outside 1..%d", m->knn, CLID_MAX_KNN);
  if (!(m->resolution > 0.f)) return set_error(CLID_EINVAL, "resolution must be positive");
  if (!m->gather_points || !m->gather_features) return set_error(CLID_EINVAL, "gather arrays are NULL");
  if (!aligned16(m->gather_features)) return set_error(CLID_EINVAL, "gather_features must be 16-byte aligned");
  if (flags & CLID_USE_BRICKS) {
    if (!m->bricks) return set_error(CLID_EINVAL, "CLID_USE_BRICKS without ClidMap.bricks");
    const ClidBricks* b = m->bricks;
    if (!b->headers || !b->records || !b->stencil) return set_error(CLID_EINVAL, "brick index arrays are NULL");
    if (b->span < 1 || b->span > 2) return set_error(CLID_EUNSUPPORTED, "brick span %d outside 1..2", b->span);
    if (!aligned16(b->headers) || !aligned16(b->records)) return set_error(CLID_EINVAL, "brick arrays must be 16-byte aligned");
  } else {
    if (m->kc < 1 || m->kc > CLID_MAX_KC) return set_error(CLID_EINVAL, "kc %d outside 1..%d", m->kc, CLID_MAX_KC);
    if (!m->buffer_pt_index || m->buffer_size <= 0 || !m->neighbor_dx) return set_error(CLID_EINVAL, "hash table arguments are NULL/empty");
    if (!m->neural_points || m->n_global <= 0) return set_error(CLID_EINVAL, "empty neural-point map");
    if ((flags & CLID_QUERY_LOCALLY) && !m->global2local) return set_error(CLID_EINVAL, "CLID_QUERY_LOCALLY without global2local");
    if ((flags & CLID_TIME_FILTER) && (!m->point_ts_create || !m->travel_dist)) return set_error(CLID_EINVAL, "CLID_TIME_FILTER without ts_create/travel_dist");
    if ((flags & CLID_TIME_FILTER) && (m->cur_ts < 0 || m->cur_ts >= m->n_travel)) return set_error(CLID_EINVAL, "cur_ts %d outside travel_dist[%d]", m->cur_ts, m->n_travel);
  }
  if ((flags & CLID_TRAINING_MODE) && !m->certainty_accum) return set_error(CLID_EINVAL, "training mode without certainty_accum");
  return CLID_OK;
}

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
static int dispatch_query_k(const QueryParams& p, cudaStream_t stream) {
  const bool bricks = p.flags & CLID_USE_BRICKS;
  if (p.map.knn <= 6) return bricks ? launch_query<H, L, 6, true>(p, stream) : launch_query<H, L, 6, false>(p, stream);
  return bricks ? launch_query<H, L, 8, true>(p, stream) : launch_query<H, L, 8, false>(p, stream);
}

static int dispatch_query(const QueryParams& p, bool has_dec, cudaStream_t stream) {
  if (!has_dec) return dispatch_query_k<0, 1>(p, stream);
  const int H = p.dec.hidden_dim, L = p.dec.levels;
  if (L == 1 && H == 64) return dispatch_query_k<64, 1>(p, stream);
  if (L == 1 && H == 32) return dispatch_query_k<32, 1>(p, stream);
  if (L == 1 && H == 128) return dispatch_query_k<128, 1>(p, stream);
  if (L == 2 && H == 32) return dispatch_query_k<32, 2>(p, stream);
  if (L == 2 && H == 64) return dispatch_query_k<64, 2>(p, stream);
  return set_error(CLID_EUNSUPPORTED,
                   "decoder %d x %d not in the fused kernel set {64x1, 32x1, 128x1, 32x2, 64x2}; "
                   "use the unfused query + torch decoder path", H, L);
}

template <bool kSecond>
static int launch_query_backward(const ClidMap* map, const float* x, const int32_t* knn_idx, const float* gz,
                                 const float* ggx, int64_t n, uint32_t flags, float* gx, float* g_gz, float* gfeat,
                                 cudaStream_t stream, const char* what) {
  if (!map) return set_error(CLID_EINVAL, "map is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!x || !knn_idx || !gz) return set_error(CLID_EINVAL, "%s: x/knn_idx/gz is NULL", what);
  if (kSecond && (!ggx || !g_gz)) return set_error(CLID_EINVAL, "%s: ggx/g_gz is NULL", what);
  if (map->feature_dim != kFeat) return set_error(CLID_EUNSUPPORTED, "feature_dim %d (only %d)", map->feature_dim, kFeat);
  if (map->knn < 1 || map->knn > CLID_MAX_KNN) return set_error(CLID_EINVAL, "knn %d outside 1..%d", map->knn, CLID_MAX_KNN);
  if (!map->gather_points || !map->gather_features || !aligned16(map->gather_features))
    return set_error(CLID_EINVAL, "gather arrays are NULL or misaligned");
  if (gfeat && !aligned16(gfeat)) return set_error(CLID_EINVAL, "gfeat must be 16-byte aligned");
  QueryBwdParams p;
  memset(&p, 0, sizeof(p));
  p.map = *map;
  p.x = x; p.knn_idx = knn_idx; p.gz = gz; p.ggx = ggx; p.gx = gx; p.g_gz = g_gz; p.gfeat = gfeat;
  p.n = n; p.flags = flags;
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  int64_t want = (n + 127) / 128, cap = (int64_t)info.sm_count * 8;
  int grid = (int)(want < cap ? want : cap);
  if (map->knn <= 6) query_backward_kernel<6, kSecond><<<grid, 128, 0, stream>>>(p);
  else query_backward_kernel<8, kSecond><<<grid, 128, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return CLID_OK;
}

static int elementwise_grid(int64_t work, int threads) {
  DeviceInfo info;
  if (device_info(&info)) return 1;
  int64_t want = (work + threads - 1) / threads;
  int64_t cap = (int64_t)info.sm_count * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

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

template <int H, int K, bool kBricks, bool kNumerical>
static int launch_train_fused(const TrainFusedParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  constexpr int kWarps = kFusedThreads / 32;
  constexpr int kSearchFloats = kBricks ? (2 * 64 * kBrickSlots + (int)(sizeof(BrickScratch) / sizeof(float)))
                                        : 2 * CLID_MAX_KC;
  size_t smem = (MlpLayout<H, 1>::kFloats + kSearchFloats + kWarps * 32 * kInPad + kWarps * 32 * (H / 32) +
                 kWarps * H * kInPad) * sizeof(float);
  auto kern = train_fused_l1_kernel<H, K, kBricks, kNumerical>;
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

static int dispatch_train_fused(const TrainFusedParams& p, cudaStream_t stream) {
  const int H = p.dec.hidden_dim;
  if (p.dec.levels != 1 || (H != 32 && H != 64 && H != 128))
    return set_error(CLID_EUNSUPPORTED, "fused training is compiled for one hidden level with H in {32,64,128}; got %d x %d",
                     H, p.dec.levels);
  if (p.map.knn > 6) return set_error(CLID_EUNSUPPORTED, "fused training is compiled for query_nn_k <= 6");
  const bool bricks = p.flags & CLID_USE_BRICKS;
  const bool num = p.num_eps > 0.f;  // set by clid_train_fused only in numerical mode
#define CLID_FUSED(HH) \
  (bricks ? (num ? launch_train_fused<HH, 6, true, true>(p, stream) : launch_train_fused<HH, 6, true, false>(p, stream)) \
          : (num ? launch_train_fused<HH, 6, false, true>(p, stream) : launch_train_fused<HH, 6, false, false>(p, stream)))
  if (H == 64) return CLID_FUSED(64);
  if (H == 32) return CLID_FUSED(32);
  return CLID_FUSED(128);
#undef CLID_FUSED
}

static int dispatch_train_backward(const ClidMap* map, const ClidDecoder* dec, const float* x, const int32_t* knn_idx,
                                   const float* dlogit, const float* dgrad, int64_t n, int64_t n_r, uint32_t flags,
                                   float* gfeat, uint8_t* touched, float* dec_grad, cudaStream_t stream) {
  TrainBwdParams p;
  memset(&p, 0, sizeof(p));
  p.map = *map; p.dec = *dec;
  p.x = x; p.knn_idx = knn_idx; p.dlogit = dlogit; p.dgrad = dgrad;
  p.gfeat = gfeat; p.touched = touched; p.dec_grad = dec_grad;
  p.n = n; p.n_r = n_r; p.flags = flags;
  const int H = dec->hidden_dim;
  if (dec->levels != 1 || (H != 32 && H != 64 && H != 128))
    return set_error(CLID_EUNSUPPORTED, "fused backward is compiled for one hidden level with H in {32,64,128}; got %d x %d",
                     H, dec->levels);
  const bool k6 = map->knn <= 6;
  if (H == 64) return k6 ? launch_train_backward<64, 6>(p, stream) : launch_train_backward<64, 8>(p, stream);
  if (H == 32) return k6 ? launch_train_backward<32, 6>(p, stream) : launch_train_backward<32, 8>(p, stream);
  return k6 ? launch_train_backward<128, 6>(p, stream) : launch_train_backward<128, 8>(p, stream);
}

}  // namespace clid

using namespace clid;

extern "C" {

int clid_version(void) { return CLID_ABI_VERSION; }

const char* clid_last_error(void) { return g_error; }

int clid_query_forward(const ClidMap* map, const ClidDecoder* dec, const float* x, const int32_t* ts, int64_t n,
                       uint32_t flags, const ClidQueryOut* out, clid_stream_t stream) {
  if (!map || !out) return set_error(CLID_EINVAL, "map/out is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!x) return set_error(CLID_EINVAL, "x is NULL");
  if (int rc = check_map(map, flags)) return rc;
  if (dec) {
    if (int rc = check_decoder(dec)) return rc;
  } else if (out->sdf || out->grad) {
    return set_error(CLID_EINVAL, "sdf/grad requested without a decoder");
  }
  QueryParams p;
  memset(&p, 0, sizeof(p));
  p.map = *map;
  if (dec) p.dec = *dec;
  p.out = *out;
  if (flags & CLID_USE_BRICKS) p.bricks = *map->bricks;
  p.x = x;
  p.ts = ts;
  p.n = n;
  p.flags = flags;
  return dispatch_query(p, dec != nullptr, static_cast<cudaStream_t>(stream));
}

int clid_query_backward(const ClidMap* map, const float* x, const int32_t* knn_idx, const float* gz, int64_t n,
                        uint32_t flags, float* gx, float* gfeat, clid_stream_t stream) {
  return launch_query_backward<false>(map, x, knn_idx, gz, nullptr, n, flags, gx, nullptr, gfeat,
                                      static_cast<cudaStream_t>(stream), "clid_query_backward");
}

int clid_query_backward_backward(const ClidMap* map, const float* x, const int32_t* knn_idx, const float* gz,
                                 const float* ggx, int64_t n, uint32_t flags, float* g_gz, float* gfeat,
                                 clid_stream_t stream) {
  return launch_query_backward<true>(map, x, knn_idx, gz, ggx, n, flags, nullptr, g_gz, gfeat,
                                     static_cast<cudaStream_t>(stream), "clid_query_backward_backward");
}

int clid_sdf_loss(const ClidLossArgs* a, clid_stream_t stream) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->n < 0 || a->nd < 0) return set_error(CLID_EINVAL, "n = %lld, nd = %lld", (long long)a->n, (long long)a->nd);
  if (a->n == 0) return CLID_OK;
  if (!a->sdf || !a->label || !a->dlogit || !a->loss) return set_error(CLID_EINVAL, "sdf/label/dlogit/loss is NULL");
  if (a->grad && a->nd > 0) return set_error(CLID_EINVAL, "analytic grad and numerical nd are mutually exclusive");
  if (a->grad && a->weight_e > 0.f && !a->dgrad) return set_error(CLID_EINVAL, "analytic eikonal needs dgrad");
  if (!(a->sdf_scale > 0.f)) return set_error(CLID_EINVAL, "sdf_scale must be positive");
  if (a->nd > 0 && !(a->num_eps > 0.f)) return set_error(CLID_EINVAL, "num_eps must be positive");
  LossParams p;
  p.sdf = a->sdf; p.grad = a->grad; p.label = a->label; p.weight = a->weight;
  p.dlogit = a->dlogit; p.dgrad = a->dgrad; p.loss = a->loss;
  p.n = a->n; p.nd = a->nd;
  p.n_norm = a->n_norm > 0 ? a->n_norm : a->n;
  p.nd_norm = a->nd_norm > 0 ? a->nd_norm : a->nd;
  p.sdf_scale = a->sdf_scale; p.weight_e = a->weight_e; p.num_eps = a->num_eps;
  p.weighted = a->weighted;
  sdf_loss_kernel<<<elementwise_grid(a->n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "sdf_loss_kernel launch");
  return CLID_OK;
}

int clid_train_backward(const ClidMap* map, const ClidDecoder* dec, const float* x, const int32_t* knn_idx,
                        const float* dlogit, const float* dgrad, int64_t n, int64_t n_r, uint32_t flags, float* gfeat,
                        uint8_t* touched, float* dec_grad, clid_stream_t stream) {
  if (!map || !dec) return set_error(CLID_EINVAL, "map/dec is NULL");
  if (n < 0 || n_r < 0 || n_r > n) return set_error(CLID_EINVAL, "n = %lld, n_r = %lld", (long long)n, (long long)n_r);
  if (n == 0) return CLID_OK;
  if (!x || !knn_idx || !dlogit) return set_error(CLID_EINVAL, "x/knn_idx/dlogit is NULL");
  if (int rc = check_decoder(dec)) return rc;
  if (map->feature_dim != kFeat) return set_error(CLID_EUNSUPPORTED, "feature_dim %d (only %d)", map->feature_dim, kFeat);
  if (map->knn < 1 || map->knn > CLID_MAX_KNN) return set_error(CLID_EINVAL, "knn %d outside 1..%d", map->knn, CLID_MAX_KNN);
  if (!map->gather_points || !map->gather_features || !aligned16(map->gather_features))
    return set_error(CLID_EINVAL, "gather arrays are NULL or misaligned");
  if (gfeat && !aligned16(gfeat)) return set_error(CLID_EINVAL, "gfeat must be 16-byte aligned");
  return dispatch_train_backward(map, dec, x, knn_idx, dlogit, dgrad, n, n_r, flags, gfeat, touched, dec_grad,
                                 static_cast<cudaStream_t>(stream));
}

int clid_train_fused(const ClidMap* map, const ClidDecoder* dec, const ClidTrainFusedArgs* a, uint32_t flags,
                     clid_stream_t stream) {
  if (!map || !dec || !a) return set_error(CLID_EINVAL, "map/dec/args is NULL");
  if (a->n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)a->n);
  if (a->n == 0) return CLID_OK;
  if (!a->x || !a->label || !a->loss) return set_error(CLID_EINVAL, "x/label/loss is NULL");
  flags |= CLID_TRAINING_MODE;
  if (int rc = check_map(map, flags)) return rc;
  if (int rc = check_decoder(dec)) return rc;
  if (a->gfeat && !aligned16(a->gfeat)) return set_error(CLID_EINVAL, "gfeat must be 16-byte aligned");
  if (!(dec->sdf_scale > 0.f)) return set_error(CLID_EINVAL, "sdf_scale must be positive");
  TrainFusedParams p;
  memset(&p, 0, sizeof(p));
  p.map = *map; p.dec = *dec;
  if (flags & CLID_USE_BRICKS) p.bricks = *map->bricks;
  p.x = a->x; p.ts = a->ts; p.label = a->label; p.weight = a->weight;
  p.gfeat = a->gfeat; p.touched = a->touched; p.dec_grad = a->dec_grad; p.loss = a->loss; p.sdf_out = a->sdf_out;
  p.n = a->n; p.n_norm = a->n_norm > 0 ? a->n_norm : a->n;
  p.nd_norm = a->nd_norm > 0 ? a->nd_norm : (a->n + 9) / 10;
  p.weight_e = a->weight_e; p.weighted = a->weighted; p.flags = flags;
  p.num_eps = 0.f;
  if (a->numerical) {
    if (!(a->num_eps > 0.f)) return set_error(CLID_EINVAL, "numerical mode needs num_eps > 0");
    p.num_eps = a->num_eps;
  }
  return dispatch_train_fused(p, static_cast<cudaStream_t>(stream));
}

int clid_adam_step(const ClidAdamArgs* a, clid_stream_t stream) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->rows < 0 || a->step < 1) return set_error(CLID_EINVAL, "rows = %lld, step = %d", (long long)a->rows, a->step);
  if (a->rows > 0 && (!a->feat || !a->feat_grad || !a->feat_m || !a->feat_v))
    return set_error(CLID_EINVAL, "feature buffers are NULL");
  if (a->rows > 0 && (!aligned16(a->feat) || !aligned16(a->feat_grad) || !aligned16(a->feat_m) || !aligned16(a->feat_v)))
    return set_error(CLID_EINVAL, "feature buffers must be 16-byte aligned");
  if (a->weight_decay != 0.f && a->touched) return set_error(CLID_EINVAL, "weight_decay needs the dense step (touched == NULL)");
  if (a->dec_grad && (!a->dec_m || !a->dec_v)) return set_error(CLID_EINVAL, "decoder moment buffers are NULL");
  if (a->dec_tensors < 0 || a->dec_tensors > 2 * CLID_MAX_LEVELS + 2) return set_error(CLID_EINVAL, "dec_tensors %d", a->dec_tensors);
  if (a->rows == 0 && !a->dec_grad) return CLID_OK;
  AdamParams p;
  memset(&p, 0, sizeof(p));
  p.feat = a->feat; p.feat_grad = a->feat_grad; p.feat_m = a->feat_m; p.feat_v = a->feat_v;
  p.touched = a->touched; p.rows = a->rows;
  for (int t = 0; t < a->dec_tensors; ++t) { p.dec_param[t] = a->dec_param[t]; p.dec_numel[t] = a->dec_numel[t]; }
  p.dec_tensors = a->dec_tensors;
  p.dec_grad = a->dec_grad; p.dec_m = a->dec_m; p.dec_v = a->dec_v;
  p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps; p.weight_decay = a->weight_decay;
  // torch evaluates the bias corrections in double on the host (torch/optim/adam.py)
  const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
  const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
  p.step_size = (float)((double)a->lr / bc1);
  p.bc2_sqrt = (float)sqrt(bc2);
  int grid = elementwise_grid(a->rows * 2 > 0 ? a->rows * 2 : 1, 256);
  adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adam_kernel launch");
  return CLID_OK;
}

int clid_radius_search(const ClidMap* map, const float* x, int64_t n, uint32_t flags, float* dist2_out,
                       int64_t* idx_out, clid_stream_t stream) {
  if (!map || !dist2_out || !idx_out) return set_error(CLID_EINVAL, "map/outputs are NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!x) return set_error(CLID_EINVAL, "x is NULL");
  if (map->kc < 1 || map->kc > CLID_MAX_KC || !map->neighbor_dx) return set_error(CLID_EINVAL, "bad neighbourhood table");
  if (!map->buffer_pt_index || map->buffer_size <= 0 || !map->neural_points || map->n_global <= 0)
    return set_error(CLID_EINVAL, "empty neural-point map");
  const bool tf = flags & CLID_TIME_FILTER;
  if (tf && (!map->point_ts_create || !map->travel_dist || map->cur_ts < 0 || map->cur_ts >= map->n_travel))
    return set_error(CLID_EINVAL, "CLID_TIME_FILTER without ts_create/travel_dist/cur_ts");
  radius_search_kernel<<<elementwise_grid(n * map->kc, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *map, x, n, tf, dist2_out, idx_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "radius_search_kernel launch");
  return CLID_OK;
}

int clid_query_certainty(const ClidMap* map, const float* x, int64_t n, const float* point_certainties, float* out,
                         clid_stream_t stream) {
  if (!map || !out || !point_certainties) return set_error(CLID_EINVAL, "map/certainties/out are NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!x) return set_error(CLID_EINVAL, "x is NULL");
  if (map->kc < 1 || map->kc > CLID_MAX_KC || !map->neighbor_dx) return set_error(CLID_EINVAL, "bad neighbourhood table");
  if (!map->buffer_pt_index || map->buffer_size <= 0 || !map->neural_points || map->n_global <= 0)
    return set_error(CLID_EINVAL, "empty neural-point map");
  query_certainty_kernel<<<elementwise_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *map, x, n, point_certainties, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "query_certainty_kernel launch");
  return CLID_OK;
}

}  // extern "C"
