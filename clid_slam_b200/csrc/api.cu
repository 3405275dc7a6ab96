// C ABI of libclid_sdf.so: argument validation, kernel selection, launches of the plain kernels.
// Declarations and the reference functions each entry point replaces: include/clid_sdf.h
// The template kernels are instantiated in inst_*.cu (see launch.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#define CLID_PLAIN_KERNELS 1  // this translation unit owns the non-template __global__ functions
#include "common.cuh"
#include "launch.h"
#include "query_bwd.cuh"
#include "query_fwd.cuh"
#include "train.cuh"
#include "decoder_grad.cuh"
#include "peer.cuh"
#include "mlp_l2.cuh"
#include "feeder.cuh"
#include "mapmaint.cuh"
#include "train_fused.cuh"
#include "decoder_tc.cuh"

namespace clid {

static thread_local char g_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return set_error(CLID_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int device_info(DeviceInfo* info) {
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
  if (reinterpret_cast<uintptr_t>(m->gather_features) & 31u) return set_error(CLID_EINVAL, "gather_features must be 32-byte aligned (rows are read with 256-bit loads)");
  if (flags & CLID_USE_BRICKS) {
    if (!m->bricks) return set_error(CLID_EINVAL, "CLID_USE_BRICKS without ClidMap.bricks");
    const ClidBricks* b = m->bricks;
    if (!b->headers || !b->records || !b->stencil) return set_error(CLID_EINVAL, "brick index arrays are NULL");
    if (b->span != 2 || b->apron < 1 || b->reach > 2)
      return set_error(CLID_EUNSUPPORTED, "brick index must have span 2, a one-brick apron and reach <= 2 (span %d, apron %d, reach %d)",
                       b->span, b->apron, b->reach);
    if (!aligned16(b->headers) || !aligned16(b->records)) return set_error(CLID_EINVAL, "brick arrays must be 16-byte aligned");
    if (b->hood && (reinterpret_cast<uintptr_t>(b->hood) & 127u)) return set_error(CLID_EINVAL, "ClidBricks.hood must be 128-byte aligned");
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
  return kSecond ? launch_query_backward_second(p, grid, stream) : launch_query_backward_first(p, grid, stream);
}

static int elementwise_grid(int64_t work, int threads) {
  DeviceInfo info;
  if (device_info(&info)) return 1;
  int64_t want = (work + threads - 1) / threads;
  int64_t cap = (int64_t)info.sm_count * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}


// ---- per-frame map maintenance (mapmaint.cuh) -------------------------------------------------------------------
namespace {
// workspace layout: result int64[2] | offsets int64[nb] | sums int32[2 nb]
struct ScanSpace {
  int64_t* result;
  int64_t* offsets;
  int32_t* sums;
  int64_t nb;
};
size_t scan_space_bytes(int64_t n) {
  const int64_t nb = scan_blocks(n > 0 ? n : 1);
  return (size_t)(2 + nb) * sizeof(int64_t) + (size_t)(2 * nb) * sizeof(int32_t) + 16;
}
int scan_space(void* ws, size_t bytes, int64_t n, ScanSpace* out) {
  if (!ws || !aligned16(ws)) return set_error(CLID_EINVAL, "workspace is NULL or not 16-byte aligned");
  if (bytes < scan_space_bytes(n)) return set_error(CLID_EINVAL, "workspace of %zu bytes, clid_scan_workspace_bytes(%lld) = %zu", bytes, (long long)n, scan_space_bytes(n));
  out->nb = scan_blocks(n);
  out->result = static_cast<int64_t*>(ws);
  out->offsets = out->result + 2;
  out->sums = reinterpret_cast<int32_t*>(out->offsets + out->nb);
  return CLID_OK;
}
// flags [n] -> rank / selected / mask of the selection the rule picks; three launches
int run_flag_scan(const uint8_t* flags, int64_t n, const ScanSpace& sp, ScanRule rule, int64_t* rank, int64_t* selected,
                  uint8_t* mask, cudaStream_t s) {
  flag_block_sums_kernel<<<(int)sp.nb, kScanThreads, 0, s>>>(flags, n, sp.sums);
  flag_block_offsets_kernel<<<1, kScanThreads, 0, s>>>(sp.sums, sp.nb, rule, sp.offsets, sp.result);
  flag_ranks_kernel<<<(int)sp.nb, kScanThreads, 0, s>>>(flags, n, sp.offsets, sp.result, rank, selected, mask);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "flag scan launch");
  return CLID_OK;
}
__global__ void fill_i32_kernel(int32_t* p, int32_t a0, int32_t a1, int32_t a2, int32_t a3, int32_t a4, int32_t a5, int32_t a6, int32_t a7, int n) {
  const int32_t v[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
  if (threadIdx.x < n) p[threadIdx.x] = v[threadIdx.x];
}
__global__ void window_tail_kernel(int64_t* g2l, uint8_t* mask, int64_t m) {
  g2l[m] = -1;  // the padding row is "local" for the mask and unreachable through the remap
  mask[m] = 1;
}
}  // namespace

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
  if ((flags & CLID_TC_DECODER) && (flags & CLID_USE_BRICKS) && dec && dec->hidden_dim == 64 && dec->levels == 1)
    return dispatch_query_tc(p, static_cast<cudaStream_t>(stream));
  return (flags & CLID_USE_BRICKS) ? dispatch_query_bricks(p, dec != nullptr, static_cast<cudaStream_t>(stream))
                                   : dispatch_query_hashed(p, dec != nullptr, static_cast<cudaStream_t>(stream));
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
  TrainBwdParams p;
  memset(&p, 0, sizeof(p));
  p.map = *map; p.dec = *dec;
  p.x = x; p.knn_idx = knn_idx; p.dlogit = dlogit; p.dgrad = dgrad;
  p.gfeat = gfeat; p.touched = touched; p.dec_grad = dec_grad;
  p.n = n; p.n_r = n_r; p.flags = flags;
  return dispatch_train_backward(p, static_cast<cudaStream_t>(stream));
}

static int64_t fused_row_slots(int64_t n, int32_t numerical) {  // 32 evaluation slots per tile
  const int64_t per_tile = numerical ? kNumTileSamples : 32;
  return (n + per_tile - 1) / per_tile * 32;
}

size_t clid_train_fused_scratch_bytes(int64_t n, int32_t numerical) {
  if (n <= 0) return 0;
  return (size_t)fused_row_slots(n, numerical) * kFoldRow * sizeof(float);
}

size_t clid_train_fused_scratch_bytes_for(const ClidDecoder* dec, int64_t n, int32_t numerical) {
  if (n <= 0 || !dec) return 0;
  const int row = dec->levels == 2 ? L2Row<32>::kFloats : kFoldRow;
  return (size_t)fused_row_slots(n, numerical) * row * sizeof(float);
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
  for (int s = 0; s < 2; ++s) {
    p.peer_grad[s] = a->peer_grad[s];
    p.peer_row[s] = a->peer_row[s];
    if (a->peer_grad[s] && !aligned16(a->peer_grad[s])) return set_error(CLID_EINVAL, "peer_grad must be 16-byte aligned");
  }
  if ((a->peer_grad[0] || a->peer_grad[1]) && (a->peer_axis < 0 || a->peer_axis > 2)) return set_error(CLID_EINVAL, "peer_axis %d", a->peer_axis);
  p.peer_axis = a->peer_axis;
  for (int s = 0; s < 4; ++s) p.peer_band[s] = a->peer_band[s];
  p.num_eps = 0.f;
  if (a->numerical) {
    if (!(a->num_eps > 0.f)) return set_error(CLID_EINVAL, "numerical mode needs num_eps > 0");
    p.num_eps = a->num_eps;
  }
  const bool have_scratch = a->scratch && a->scratch_bytes >= clid_train_fused_scratch_bytes_for(dec, a->n, a->numerical);
  if (a->scratch && (reinterpret_cast<uintptr_t>(a->scratch) & 15u)) return set_error(CLID_EINVAL, "scratch must be 16-byte aligned");
  if ((p.dec_grad || dec->levels == 2) && have_scratch) {
    // decoder-gradient rows go to scratch; a dense reduction kernel folds them afterwards
    p.fold_rows = static_cast<float*>(a->scratch);
  }
  return (flags & CLID_USE_BRICKS) ? dispatch_train_fused_bricks(p, static_cast<cudaStream_t>(stream))
                                   : dispatch_train_fused_hashed(p, static_cast<cudaStream_t>(stream));
}

int clid_decoder_grad_reduce(const ClidDecoder* dec, const void* scratch, int64_t n, int32_t numerical, uint32_t flags,
                             float* dec_grad, clid_stream_t stream) {
  if (!dec || !dec_grad) return set_error(CLID_EINVAL, "dec/dec_grad is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!scratch || (reinterpret_cast<uintptr_t>(scratch) & 15u)) return set_error(CLID_EINVAL, "scratch is NULL or misaligned");
  if (int rc = check_decoder(dec)) return rc;
  if (dec->levels == 2) {
    DecoderGradL2Params g2;
    memset(&g2, 0, sizeof(g2));
    g2.dec = *dec; g2.rows = static_cast<const float*>(scratch); g2.dec_grad = dec_grad; g2.flags = flags;
    g2.n_rows = fused_row_slots(n, numerical);
    return launch_decoder_grad_l2(g2, static_cast<cudaStream_t>(stream));
  }
  DecoderGradParams g;
  memset(&g, 0, sizeof(g));
  g.dec = *dec; g.rows = static_cast<const float*>(scratch); g.dec_grad = dec_grad; g.flags = flags;
  g.n_rows = (int64_t)(clid_train_fused_scratch_bytes(n, numerical) / (kFoldRow * sizeof(float)));
  return launch_decoder_grad(g, static_cast<cudaStream_t>(stream));
}

int clid_adam_step(const ClidAdamArgs* a, clid_stream_t stream) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->rows < 0 || (!a->step_state && a->step < 1))
    return set_error(CLID_EINVAL, "rows = %lld, step = %d", (long long)a->rows, a->step);
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
  p.touched = a->touched; p.rows = a->rows; p.feat_stride = kFeat;
  if (a->rows > 0 && ((reinterpret_cast<uintptr_t>(a->feat) | reinterpret_cast<uintptr_t>(a->feat_grad) | reinterpret_cast<uintptr_t>(a->feat_m) |
                       reinterpret_cast<uintptr_t>(a->feat_v)) & 31u))
    return set_error(CLID_EINVAL, "feature buffers must be 32-byte aligned (rows move with 256-bit loads / stores)");
  for (int t = 0; t < a->dec_tensors; ++t) { p.dec_param[t] = a->dec_param[t]; p.dec_numel[t] = a->dec_numel[t]; }
  p.dec_tensors = a->dec_tensors;
  p.dec_grad = a->dec_grad; p.dec_m = a->dec_m; p.dec_v = a->dec_v;
  p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps; p.weight_decay = a->weight_decay;
  if (a->step_state) {
    // device-resident step counter (graph-capturable): advance it (unless step < 0), then read the scalars
    // in the kernel
    AdamStepState* st = static_cast<AdamStepState*>(a->step_state);
    if (a->step >= 0) adam_advance_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(st, a->lr, a->beta1, a->beta2);
    p.step_scalars = &st->step_size;
  } else {
    // torch evaluates the bias corrections in double on the host (torch/optim/adam.py)
    const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
    const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
    p.step_size = (float)((double)a->lr / bc1);
    p.bc2_sqrt = (float)sqrt(bc2);
  }
#ifndef CLID_ADAM_BLOCKS_PER_SM
#define CLID_ADAM_BLOCKS_PER_SM 16
#endif
  int grid = elementwise_grid(a->rows > 0 ? (a->rows + 1) / 2 : 1, 256);
  {
    // persistent blocks: leave thread slots for the decoder-gradient reduction that may run beside this kernel
    DeviceInfo info;
    if (int rc = device_info(&info)) return rc;
    const int cap = info.sm_count * CLID_ADAM_BLOCKS_PER_SM;
    if (grid > cap) grid = cap;
  }
  adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adam_kernel launch");
  return CLID_OK;
}

int clid_adam_advance(void* step_state, float lr, float beta1, float beta2, clid_stream_t stream) {
  if (!step_state) return set_error(CLID_EINVAL, "step_state is NULL");
  adam_advance_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<AdamStepState*>(step_state), lr, beta1, beta2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adam_advance_kernel launch");
  return CLID_OK;
}

int clid_step_begin(void* step_state, float lr, float beta1, float beta2, float* loss3, clid_stream_t stream) {
  if (!step_state || !loss3) return set_error(CLID_EINVAL, "step_state / loss3 is NULL");
  adam_advance_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<AdamStepState*>(step_state), lr, beta1, beta2, loss3);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adam_advance_kernel launch");
  return CLID_OK;
}

static int check_peer(const ClidPeerArgs* a) {
  if (!a) return set_error(CLID_EINVAL, "peer args are NULL");
  if (a->world < 1 || a->world > 8 || a->rank < 0 || a->rank >= a->world) return set_error(CLID_EINVAL, "rank %d / world %d", a->rank, a->world);
  if (a->n0 < 0 || a->n1 < 0 || a->n0 + a->n1 <= 0 || a->n0 + a->n1 > a->stride) return set_error(CLID_EINVAL, "n0 %d + n1 %d vs stride %d", a->n0, a->n1, a->stride);
  if (!a->epoch) return set_error(CLID_EINVAL, "epoch is NULL");
  for (int r = 0; r < a->world; ++r)
    if (!a->slots_of[r] || !a->flags_of[r]) return set_error(CLID_EINVAL, "slots/flags of rank %d are NULL", r);
  return CLID_OK;
}

int clid_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev == peer_device) return CLID_OK;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceCanAccessPeer");
  if (!can) return set_error(CLID_EUNSUPPORTED, "device %d cannot access device %d (no NVLink / PCIe peer path)", dev, peer_device);
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return CLID_OK; }
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
  return CLID_OK;
}

int clid_peer_alloc(size_t bytes, void** ptr_out, void* handle64_out) {
  if (!ptr_out || !handle64_out || bytes == 0) return set_error(CLID_EINVAL, "ptr_out/handle64_out is NULL or bytes == 0");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaMemset"); }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle64_out, &h, sizeof(h));
  *ptr_out = p;
  return CLID_OK;
}

int clid_peer_free(void* ptr) {
  if (!ptr) return CLID_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
  return CLID_OK;
}

int clid_ipc_open(const void* handle64, void** base_out) {
  if (!handle64 || !base_out) return set_error(CLID_EINVAL, "handle/base_out is NULL");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  return CLID_OK;
}

int clid_ipc_close(void* base) {
  if (!base) return CLID_OK;
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return CLID_OK;
}

int clid_peer_publish(const ClidPeerArgs* a, const float* src0, const float* src1, clid_stream_t stream) {
  if (int rc = check_peer(a)) return rc;
  if ((a->n0 > 0 && !src0) || (a->n1 > 0 && !src1)) return set_error(CLID_EINVAL, "src is NULL");
  peer_publish_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a, src0, src1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "peer_publish_kernel launch");
  return CLID_OK;
}

int clid_peer_reduce(const ClidPeerArgs* a, float* dst0, float* dst1, clid_stream_t stream) {
  if (int rc = check_peer(a)) return rc;
  if ((a->n0 > 0 && !dst0) || (a->n1 > 0 && !dst1)) return set_error(CLID_EINVAL, "dst is NULL");
  peer_reduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a, dst0, dst1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "peer_reduce_kernel launch");
  return CLID_OK;
}

int clid_region_sdf(const ClidLocalCloud* c, const float* points, int64_t n, float* sdf_abs, uint8_t* surface_mask,
                    clid_stream_t stream) {
  if (!c) return set_error(CLID_EINVAL, "cloud is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!points || !sdf_abs || !surface_mask) return set_error(CLID_EINVAL, "points/outputs are NULL");
  if (!c->table || c->buffer_size <= 0 || !c->neighbor_idx || c->kc < 1 || c->kc > CLID_MAX_KC)
    return set_error(CLID_EINVAL, "local point-cloud map: table / neighbourhood missing");
  if (c->n_points > 0 && !c->points) return set_error(CLID_EINVAL, "local point-cloud map: points are NULL");
  if (!(c->resolution > 0.f)) return set_error(CLID_EINVAL, "resolution must be positive");
  RegionSdfParams p;
  memset(&p, 0, sizeof(p));
  p.points = points; p.table = c->table; p.map_points = c->points; p.neighbor_idx = c->neighbor_idx;
  p.n = n; p.buffer_size = c->buffer_size; p.m = c->n_points;
  for (int i = 0; i < 3; ++i) p.primes[i] = c->primes[i];
  p.kc = c->kc; p.resolution = c->resolution; p.max_valid_range = c->max_valid_range;
  p.eta_threshold = 0.2f; p.dist_threshold = 0.1f;  // estimate_plane defaults (local_point_cloud_map.py:155-157)
  p.sdf_abs = sdf_abs; p.surface_mask = surface_mask;
  region_sdf_kernel<<<elementwise_grid(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "region_sdf_kernel launch");
  return CLID_OK;
}

int clid_brick_keep(const ClidMap* m, const float* points, const int64_t* gids, int64_t n, const int32_t* ts_create,
                    int32_t* cells, uint8_t* keep, int32_t* bbox, clid_stream_t stream) {
  if (!m || !points || !cells || !keep || !bbox) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return n < 0 ? set_error(CLID_EINVAL, "n = %lld", (long long)n) : CLID_OK;
  if (!m->buffer_pt_index || m->buffer_size <= 0 || !(m->resolution > 0.f)) return set_error(CLID_EINVAL, "hash table / resolution missing");
  if (ts_create && (!m->travel_dist || m->cur_ts < 0 || m->cur_ts >= m->n_travel)) return set_error(CLID_EINVAL, "time filter without travel_dist / cur_ts");
  BrickKeyParams p;
  memset(&p, 0, sizeof(p));
  p.points = points; p.gids = gids; p.table = m->buffer_pt_index; p.buffer_size = m->buffer_size;
  for (int i = 0; i < 3; ++i) p.primes[i] = m->primes[i];
  p.ts_create = ts_create; p.travel_dist = m->travel_dist; p.cur_ts = m->cur_ts;
  p.diff_travel_dist_local = m->diff_travel_dist_local; p.resolution = m->resolution; p.n = n;
  p.cells = cells; p.keep = keep; p.bbox = bbox;
  brick_keep_kernel<<<elementwise_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "brick_keep_kernel launch");
  return CLID_OK;
}

int clid_brick_keys(const int32_t* cells, const uint8_t* keep, int64_t n, const int32_t* lo3, const int32_t* dims3, int64_t* keys,
                    clid_stream_t stream) {
  if (!cells || !keep || !lo3 || !dims3 || !keys) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return n < 0 ? set_error(CLID_EINVAL, "n = %lld", (long long)n) : CLID_OK;
  brick_key_kernel<<<elementwise_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cells, keep, n, lo3[0], lo3[1], lo3[2],
                                                                                            dims3[0], dims3[1], keys);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "brick_key_kernel launch");
  return CLID_OK;
}

int clid_brick_fill(const int64_t* sorted_keys, const int64_t* order, int64_t n_kept, const float* points, const int32_t* dims3,
                    float* records, ClidBrickHeader* headers, uint32_t* hood, clid_stream_t stream) {
  if (!sorted_keys || !order || !points || !dims3 || !records || !headers) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n_kept <= 0) return set_error(CLID_EINVAL, "n_kept = %lld", (long long)n_kept);
  if (!aligned16(records) || !aligned16(headers) || (hood && (reinterpret_cast<uintptr_t>(hood) & 127u)))
    return set_error(CLID_EINVAL, "records / headers must be 16-byte aligned, hood 128-byte aligned");
  const int64_t nb = (int64_t)dims3[0] * dims3[1] * dims3[2];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // headers start as {mask 0, base INT_MAX, count 0}
  cudaError_t e = cudaMemsetAsync(headers, 0, (size_t)nb * sizeof(ClidBrickHeader), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(headers)");
  brick_header_init_kernel<<<elementwise_grid(nb, 256), 256, 0, s>>>(headers, nb);
  brick_scatter_kernel<<<elementwise_grid(n_kept, 256), 256, 0, s>>>(sorted_keys, order, n_kept, points, nullptr,
                                                                      reinterpret_cast<float4*>(records), headers);
  brick_hood_kernel<<<elementwise_grid(nb * 8, 256), 256, 0, s>>>(headers, dims3[0], dims3[1], dims3[2], hood);
  brick_header_fix_kernel<<<elementwise_grid(nb, 256), 256, 0, s>>>(headers, nb);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "brick fill kernels launch");
  return CLID_OK;
}

size_t clid_scan_workspace_bytes(int64_t n) { return scan_space_bytes(n); }

int clid_voxel_keys(const float* points, const float* value, int64_t n, float voxel_size, int32_t* stats, int64_t* keys,
                    clid_stream_t stream) {
  if (!points || !stats || !keys) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return n < 0 ? set_error(CLID_EINVAL, "n = %lld", (long long)n) : CLID_OK;
  if (!(voxel_size > 0.f)) return set_error(CLID_EINVAL, "voxel_size must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  fill_i32_kernel<<<1, 32, 0, s>>>(stats, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, INT_MIN, 0, 8);
  voxel_stats_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(points, n, voxel_size, stats);
  if (value) {
    // value mode ranks by `value / value.max()`: the maximum replaces the centre-distance maximum in stats[6]
    fill_i32_kernel<<<1, 32, 0, s>>>(stats + 6, INT_MIN, 0, 0, 0, 0, 0, 0, 0, 1);
    ordered_max_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(value, n, stats + 6);
  }
  voxel_keys_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(points, value, stats + 6, n, voxel_size, stats, keys);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "voxel key kernels launch");
  return CLID_OK;
}

int clid_voxel_pick(const int64_t* sorted_keys, const int64_t* order, int64_t n, void* workspace, size_t workspace_bytes,
                    uint8_t* flags, int64_t* selected, int64_t* out, clid_stream_t stream) {
  if (!sorted_keys || !order || !flags || !selected || !out) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  ScanSpace sp;
  if (int rc = scan_space(workspace, workspace_bytes, n, &sp)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  voxel_heads_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(sorted_keys, n, flags);
  if (int rc = run_flag_scan(flags, n, sp, ScanRule{nullptr, 0}, nullptr, selected, nullptr, s)) return rc;
  voxel_pick_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(selected, sp.result, order, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "voxel pick kernels launch");
  return CLID_OK;
}

static int insert_params(const ClidInsertArgs* a, InsertParams* p) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->n <= 0) return set_error(CLID_EINVAL, "n = %lld", (long long)a->n);
  if (!a->cand || !a->buffer_pt_index || a->buffer_size <= 0 || !a->slot || !a->owner || !a->fresh || !a->rank)
    return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (!(a->resolution > 0.f)) return set_error(CLID_EINVAL, "resolution must be positive");
  if (!a->all_fresh && (a->m <= 0 || !a->neural_points)) return set_error(CLID_EINVAL, "a non-empty map needs neural_points");
  if (!a->all_fresh && a->ts_update && (!a->travel_dist || a->cur_ts < 0 || a->cur_ts >= a->n_travel))
    return set_error(CLID_EINVAL, "travel-distance test without travel_dist / cur_ts");
  memset(p, 0, sizeof(*p));
  p->cand = a->cand; p->n = a->n; p->table = a->buffer_pt_index; p->buffer_size = a->buffer_size;
  for (int i = 0; i < 3; ++i) p->primes[i] = a->primes[i];
  p->neural_points = a->neural_points; p->ts_update = a->all_fresh ? nullptr : a->ts_update; p->travel_dist = a->travel_dist;
  p->m = a->m; p->cur_ts = a->cur_ts; p->all_fresh = a->all_fresh; p->resolution = a->resolution; p->far2 = a->far2;
  p->diff_travel_dist_local = a->diff_travel_dist_local;
  p->slot = a->slot; p->owner = a->owner; p->fresh = a->fresh;
  return CLID_OK;
}

int clid_map_insert_probe(const ClidInsertArgs* a, clid_stream_t stream) {
  InsertParams p;
  if (int rc = insert_params(a, &p)) return rc;
  ScanSpace sp;
  if (int rc = scan_space(a->workspace, a->workspace_bytes, a->n, &sp)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  insert_probe_kernel<<<elementwise_grid(a->n, 256), 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "insert_probe_kernel launch");
  return run_flag_scan(a->fresh, a->n, sp, ScanRule{nullptr, 0}, a->rank, nullptr, nullptr, s);
}

int clid_map_insert_commit(const ClidInsertArgs* a, float* new_points, int32_t* new_ts_create, int32_t* new_ts_update,
                           clid_stream_t stream) {
  InsertParams p;
  if (int rc = insert_params(a, &p)) return rc;
  if (!new_points || !new_ts_create || !new_ts_update) return set_error(CLID_EINVAL, "the rows to append are NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  insert_bid_kernel<<<elementwise_grid(a->n, 256), 256, 0, s>>>(a->slot, a->n, a->buffer_pt_index);
  insert_commit_kernel<<<elementwise_grid(a->n, 256), 256, 0, s>>>(p, a->rank, new_points, new_ts_create, new_ts_update);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "insert commit kernels launch");
  return CLID_OK;
}

int clid_local_window_select(const ClidWindowArgs* a, clid_stream_t stream) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->m <= 0) return set_error(CLID_EINVAL, "m = %lld", (long long)a->m);
  if (!a->neural_points || !a->flags || !a->global2local || !a->local_mask || !a->gids)
    return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (a->temporal) {
    if (!a->ts_create || (a->use_mid_ts && !a->ts_update)) return set_error(CLID_EINVAL, "temporal window without time stamps");
    if (a->travel_dist && (a->cur_ts < 0 || a->cur_ts >= a->n_travel)) return set_error(CLID_EINVAL, "cur_ts outside travel_dist");
  }
  ScanSpace sp;
  if (int rc = scan_space(a->workspace, a->workspace_bytes, a->m, &sp)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  WindowParams p;
  memset(&p, 0, sizeof(p));
  p.neural_points = a->neural_points; p.ts_create = a->ts_create; p.ts_update = a->ts_update; p.travel_dist = a->travel_dist;
  p.m = a->m;
  for (int i = 0; i < 3; ++i) p.sensor[i] = a->sensor[i];
  p.radius2 = a->radius2; p.sensor_is_f64 = a->sensor_is_f64; p.temporal = a->temporal; p.use_mid_ts = a->use_mid_ts;
  p.cur_ts = a->cur_ts; p.reboot_ts = a->reboot_test ? a->reboot_ts : INT_MIN; p.diff_ts_local = a->diff_ts_local;
  p.diff_travel_dist_local = a->diff_travel_dist_local; p.flags = a->flags;
  // the in-window count lives in the first sum slot until the block sums overwrite it: keep it in result[1]'s low half
  int32_t* n_in_time = reinterpret_cast<int32_t*>(sp.result + 1);
  fill_i32_kernel<<<1, 32, 0, s>>>(n_in_time, 0, 0, 0, 0, 0, 0, 0, 0, 2);
  p.n_in_time = n_in_time;
  window_flags_kernel<<<elementwise_grid(a->m, 256), 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "window_flags_kernel launch");
  // fewer than 100 points inside the time window: the reference takes every point (in range) instead (:477-481)
  ScanRule rule{a->temporal ? n_in_time : nullptr, 100};
  if (int rc = run_flag_scan(a->flags, a->m, sp, rule, a->global2local, a->gids, a->local_mask, s)) return rc;
  window_tail_kernel<<<1, 1, 0, s>>>(a->global2local, a->local_mask, a->m);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "window_tail_kernel launch");
  return CLID_OK;
}

static int check_rows(const ClidWindowRows* r) {
  if (!r) return set_error(CLID_EINVAL, "rows is NULL");
  if (r->n_local < 0 || r->m < 0) return set_error(CLID_EINVAL, "n_local = %lld, m = %lld", (long long)r->n_local, (long long)r->m);
  if ((r->n_local > 0 && !r->gids) || !r->geo_features || !r->local_features) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (!aligned16(r->geo_features) || !aligned16(r->local_features)) return set_error(CLID_EINVAL, "feature arrays must be 16-byte aligned");
  return CLID_OK;
}

int clid_local_window_gather(const ClidWindowRows* r, clid_stream_t stream) {
  if (int rc = check_rows(r)) return rc;
  if (r->n_local > 0 && (!r->neural_points || !r->point_orientations || !r->point_certainties || !r->point_ts_update ||
                         !r->local_points || !r->local_orientations || !r->local_certainties || !r->local_ts_update))
    return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (r->n_local > 0 && (!aligned16(r->point_orientations) || !aligned16(r->local_orientations)))
    return set_error(CLID_EINVAL, "orientation arrays must be 16-byte aligned");
  WindowGather g;
  g.gids = r->gids; g.n_local = r->n_local; g.neural_points = r->neural_points; g.orientations = r->point_orientations;
  g.certainties = r->point_certainties; g.ts_update = r->point_ts_update; g.geo_features = r->geo_features; g.m = r->m;
  g.local_points = r->local_points; g.local_orientations = r->local_orientations; g.local_certainties = r->local_certainties;
  g.local_ts_update = r->local_ts_update; g.local_features = r->local_features;
  window_gather_kernel<<<elementwise_grid(r->n_local + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "window_gather_kernel launch");
  return CLID_OK;
}

int clid_local_window_scatter(const ClidWindowRows* r, clid_stream_t stream) {
  if (int rc = check_rows(r)) return rc;
  if (r->n_local > 0 && (!r->point_certainties || !r->point_ts_update || !r->local_certainties || !r->local_ts_update))
    return set_error(CLID_EINVAL, "a required pointer is NULL");
  window_scatter_kernel<<<elementwise_grid(r->n_local + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      r->gids, r->n_local, r->m, r->local_features, r->local_certainties, r->local_ts_update, r->geo_features,
      r->point_certainties, r->point_ts_update);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "window_scatter_kernel launch");
  return CLID_OK;
}

int clid_pool_filter_select(const float* global_coord, int64_t n, const double* sensor3, double radius, int32_t sensor_is_f64,
                            int32_t use_norm, uint8_t* flags, int64_t* rank, void* workspace, size_t workspace_bytes,
                            clid_stream_t stream) {
  if (!global_coord || !sensor3 || !flags || !rank) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  ScanSpace sp;
  if (int rc = scan_space(workspace, workspace_bytes, n, &sp)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pool_flags_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(global_coord, n, sensor3[0], sensor3[1], sensor3[2], radius,
                                                            radius * radius, sensor_is_f64, use_norm, flags);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pool_flags_kernel launch");
  return run_flag_scan(flags, n, sp, ScanRule{nullptr, 0}, rank, nullptr, nullptr, s);
}

int clid_compact_rows(const int64_t* rank, int64_t n, const void* const* src, void* const* dst, const int32_t* words,
                      int32_t n_arrays, clid_stream_t stream) {
  if (!rank || !src || !dst || !words) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return n < 0 ? set_error(CLID_EINVAL, "n = %lld", (long long)n) : CLID_OK;
  if (n_arrays < 1 || n_arrays > kCompactArrays) return set_error(CLID_EINVAL, "n_arrays %d outside 1..%d", n_arrays, kCompactArrays);
  CompactParams p;
  memset(&p, 0, sizeof(p));
  p.rank = rank; p.n = n; p.n_arrays = n_arrays;
  for (int a = 0; a < n_arrays; ++a) {
    if (!src[a] || !dst[a] || words[a] < 1) return set_error(CLID_EINVAL, "array %d: NULL pointer or words < 1", a);
    p.src[a] = static_cast<const uint32_t*>(src[a]); p.dst[a] = static_cast<uint32_t*>(dst[a]); p.words[a] = words[a];
  }
  compact_rows_kernel<<<elementwise_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "compact_rows_kernel launch");
  return CLID_OK;
}

int clid_table_store(const int64_t* slot, const int64_t* value, int64_t n, int64_t value_base, int64_t* buffer_pt_index,
                     int64_t buffer_size, clid_stream_t stream) {
  if (!slot || !buffer_pt_index || buffer_size <= 0) return set_error(CLID_EINVAL, "slot / table is NULL or empty");
  if (n <= 0) return n < 0 ? set_error(CLID_EINVAL, "n = %lld", (long long)n) : CLID_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  table_bid_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(slot, n, buffer_size, buffer_pt_index);
  table_commit_kernel<<<elementwise_grid(n, 256), 256, 0, s>>>(slot, value, n, buffer_size, value_base, buffer_pt_index);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "table store kernels launch");
  return CLID_OK;
}

int clid_ray_samples(const ClidRaySampleArgs* a, clid_stream_t stream) {
  if (!a) return set_error(CLID_EINVAL, "args is NULL");
  if (a->n_points <= 0) return a->n_points < 0 ? set_error(CLID_EINVAL, "n_points = %lld", (long long)a->n_points) : CLID_OK;
  if (!a->points || !a->depth || !a->coord || !a->disp || !a->weight) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (a->n_surf < 0 || a->n_front < 0 || a->n_behind < 0) return set_error(CLID_EINVAL, "negative sample counts");
  if ((a->n_surf && !a->randn_surf) || (a->n_front && !a->rand_front) || (a->n_behind && !a->rand_behind))
    return set_error(CLID_EINVAL, "random numbers are NULL");
  RaySampleParams p;
  memset(&p, 0, sizeof(p));
  p.points = a->points; p.depth = a->depth; p.randn_surf = a->randn_surf; p.rand_front = a->rand_front; p.rand_behind = a->rand_behind;
  p.P = a->n_points; p.n_surf = a->n_surf; p.n_front = a->n_front; p.n_behind = a->n_behind;
  p.sigma = a->surface_sample_range_m;
  p.margin_sigma = a->margin;
  p.begin_ratio = a->free_sample_begin_ratio; p.end_dist = a->free_sample_end_dist_m;
  p.weight_top = a->weight_top;
  p.inv_max_range = 1.0f / a->max_range; p.weight_scale = a->dist_weight_scale; p.dist_weight_on = a->dist_weight_on;
  p.coord = a->coord; p.disp = a->disp; p.weight = a->weight;
  ray_samples_kernel<<<elementwise_grid(a->n_points, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ray_samples_kernel launch");
  return CLID_OK;
}

int clid_ray_labels(const float* disp, const float* dist, const uint8_t* reachable, int64_t n_points, int32_t samples_per_ray,
                    int32_t n_surf, float* label, uint8_t* keep, clid_stream_t stream) {
  if (!disp || !label || !keep || (n_surf > 0 && (!dist || !reachable))) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n_points <= 0) return n_points < 0 ? set_error(CLID_EINVAL, "n_points = %lld", (long long)n_points) : CLID_OK;
  if (samples_per_ray < 1 || n_surf < 0 || n_surf >= samples_per_ray) return set_error(CLID_EINVAL, "samples_per_ray %d, n_surf %d", samples_per_ray, n_surf);
  ray_labels_kernel<<<elementwise_grid(n_points * samples_per_ray, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      disp, dist, reachable, n_points, samples_per_ray, n_surf, label, keep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ray_labels_kernel launch");
  return CLID_OK;
}

int clid_flag_ranks(const uint8_t* flags, int64_t n, int64_t* rank, void* workspace, size_t workspace_bytes, clid_stream_t stream) {
  if (!flags || !rank) return set_error(CLID_EINVAL, "a required pointer is NULL");
  if (n <= 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  ScanSpace sp;
  if (int rc = scan_space(workspace, workspace_bytes, n, &sp)) return rc;
  return run_flag_scan(flags, n, sp, ScanRule{nullptr, 0}, rank, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int clid_registration_terms(const float* pc_imu, const float* sdf, const float* grad, const int32_t* nn_count, int64_t n,
                            const float* rot9, int32_t min_nn, float min_grad, float max_grad, double* out28,
                            uint8_t* valid_out, clid_stream_t stream) {
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!pc_imu || !sdf || !grad || !nn_count || !rot9 || !out28) return set_error(CLID_EINVAL, "a required pointer is NULL");
  RegParams p;
  p.pc_imu = pc_imu; p.sdf = sdf; p.grad = grad; p.nn_count = nn_count; p.n = n;
  for (int i = 0; i < 9; ++i) p.rot[i] = rot9[i];  // HOST array: nine floats, row-major
  p.min_nn = min_nn; p.min_grad = min_grad; p.max_grad = max_grad; p.out = out28; p.valid_out = valid_out;
  int grid = elementwise_grid(n, 256);
  if (grid > 592) grid = 592;
  registration_terms_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "registration_terms_kernel launch");
  return CLID_OK;
}

static int check_pool(const ClidReplayPool* pool, int64_t n) {
  if (!pool->coord || !pool->sdf_label || pool->count <= 0) return set_error(CLID_EINVAL, "replay pool is empty or NULL");
  if (pool->bs_new < 0 || pool->bs_new > n) return set_error(CLID_EINVAL, "bs_new %d outside 0..n", pool->bs_new);
  if (pool->bs_new > 0 && (!pool->new_idx || pool->n_new <= 0)) return set_error(CLID_EINVAL, "bs_new without new_idx");
  return CLID_OK;
}

static int launch_draw(const DrawParams& d, cudaStream_t stream) {
  draw_batch_kernel<<<elementwise_grid(d.n, 256), 256, 0, stream>>>(d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "draw_batch_kernel launch");
  return CLID_OK;
}

int clid_draw_batch(const ClidReplayPool* pool, int64_t n, uint64_t seed, uint64_t offset, float* x, float* label,
                    float* weight, int32_t* ts, int64_t* index_out, clid_stream_t stream) {
  if (!pool) return set_error(CLID_EINVAL, "pool is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!x || !label) return set_error(CLID_EINVAL, "x/label is NULL");
  if (int rc = check_pool(pool, n)) return rc;
  if ((weight && !pool->weight) || (ts && !pool->time)) return set_error(CLID_EINVAL, "weight/ts requested but the pool has none");
  DrawParams d;
  memset(&d, 0, sizeof(d));
  d.pool = *pool; d.n = n; d.seed = seed; d.offset = offset;
  d.x = x; d.label = label; d.weight = weight; d.ts = ts; d.index_out = index_out;
  return launch_draw(d, static_cast<cudaStream_t>(stream));
}

int clid_mapping_run(const ClidMap* map, const ClidDecoder* dec, const ClidMappingArgs* a, uint32_t flags,
                     clid_stream_t stream) {
  if (!map || !dec || !a) return set_error(CLID_EINVAL, "map/dec/args is NULL");
  if (a->iters < 0) return set_error(CLID_EINVAL, "iters = %d", a->iters);
  if (a->iters == 0) return CLID_OK;
  const ClidTrainFusedArgs& t = a->train;
  if (t.n <= 0) return set_error(CLID_EINVAL, "train.n = %lld", (long long)t.n);
  if (!t.x || !t.label || !t.loss) return set_error(CLID_EINVAL, "train.x/label/loss scratch is NULL");
  if (!a->adam.step_state) return set_error(CLID_EINVAL, "clid_mapping_run needs the device step counter (adam.step_state)");
  if (int rc = check_pool(&a->pool, t.n)) return rc;
  if ((t.weight && !a->pool.weight) || (t.ts && !a->pool.time)) return set_error(CLID_EINVAL, "weight/ts scratch given but the pool has none");
  if (t.dec_grad && !t.scratch) return set_error(CLID_EINVAL, "clid_mapping_run needs train.scratch when the decoder trains");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DrawParams d;
  memset(&d, 0, sizeof(d));
  d.pool = a->pool; d.n = t.n; d.seed = a->seed;
  d.x = const_cast<float*>(t.x); d.label = const_cast<float*>(t.label);
  d.weight = const_cast<float*>(t.weight); d.ts = const_cast<int32_t*>(t.ts);
  d.loss = t.loss;
  ClidAdamArgs adam = a->adam;
  adam.step = -1;  // the draw kernel of every iteration advances the device counter (one launch less per iteration)
  d.step_state = static_cast<AdamStepState*>(adam.step_state);
  d.lr = adam.lr; d.beta1 = adam.beta1; d.beta2 = adam.beta2;
  for (int it = 0; it < a->iters; ++it) {
    d.offset = a->offset + (uint64_t)it;
    d.loss_prev_out = (it > 0 && a->loss_history) ? a->loss_history + 3 * (it - 1) : nullptr;
    if (int rc = launch_draw(d, s)) return rc;
    if (int rc = clid_train_fused(map, dec, &t, flags, stream)) return rc;
    if (t.dec_grad)
      if (int rc = clid_decoder_grad_reduce(dec, t.scratch, t.n, t.numerical, flags, t.dec_grad, stream)) return rc;
    if (int rc = clid_adam_step(&adam, stream)) return rc;
  }
  if (a->loss_history) {
    copy3_kernel<<<1, 32, 0, s>>>(t.loss, a->loss_history + 3 * (a->iters - 1));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "copy3_kernel launch");
  }
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

int clid_decoder_eval(const ClidDecoder* dec, const float* z, int64_t n, uint32_t flags, float* out, float* a,
                      uint32_t* mask, clid_stream_t stream) {
  if (!dec || !out) return set_error(CLID_EINVAL, "dec/out is NULL");
  if (n < 0) return set_error(CLID_EINVAL, "n = %lld", (long long)n);
  if (n == 0) return CLID_OK;
  if (!z) return set_error(CLID_EINVAL, "z is NULL");
  if (int rc = check_decoder(dec)) return rc;
  if (dec->hidden_dim != tc::kH || dec->levels != 1)
    return set_error(CLID_EUNSUPPORTED, "the tensor-core decoder is compiled for 64 x 1; got %d x %d", dec->hidden_dim, dec->levels);
  DecoderEvalParams p;
  memset(&p, 0, sizeof(p));
  p.dec = *dec; p.z = z; p.out = out; p.a = a; p.mask = mask; p.n = n; p.flags = flags;
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  const size_t smem = ((sizeof(tc::Shared) + 3) / 4 + MlpLayout<tc::kH, 1>::kFloats) * sizeof(float);
  const int64_t want = (n + 127) / 128, cap = (int64_t)info.sm_count * 4;
  decoder_eval_tc_kernel<<<(int)(want < cap ? want : cap), 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "decoder_eval_tc_kernel launch");
  return CLID_OK;
}

}  // extern "C"
