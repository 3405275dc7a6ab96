/*
 * clid_sdf.h -- C ABI of libclid_sdf.so: the B200 (sm_100a) implementation of CLID-SLAM's
 * per-scan neural-SDF training/query hot path.
 *
 * The reference (DUTRobot/CLID-SLAM @ 5c4f9e7) is 100 % Python/PyTorch and has no FFI of
 * its own; every entry point below therefore cites the *Python function(s)* whose
 * arithmetic it replaces (paths relative to the reference tree).  The host-side mirror of
 * the reference's operator interface (model.decoder.Decoder, model.neural_points.
 * NeuralPoints, utils.mapper.Mapper, utils.loss) lives in clid_slam_b200/ and binds these
 * symbols through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - All pointers are DEVICE pointers borrowed for the duration of the call; the library
 *    never allocates, frees or retains device memory.  Layouts are the reference's
 *    (row-major, fp32 values, int64 hash table / remap, int32 timestamps).
 *  - Work is enqueued on the caller's stream; no call synchronises or reads back.
 *  - Return value: 0 (CLID_OK) or a negative CLID_E* code; clid_last_error() returns a
 *    thread-local message for the last failure.  n == 0 is a successful no-op.
 */
#ifndef CLID_SDF_H_
#define CLID_SDF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLID_ABI_VERSION 5
#define CLID_MAX_LEVELS 3   /* hidden layers of the decoder MLP */
#define CLID_MAX_KNN 8      /* query_nn_k */
#define CLID_MAX_KC 256     /* probed cells per query */

#define CLID_API __attribute__((visibility("default")))

typedef void* clid_stream_t; /* cudaStream_t */

enum ClidStatus {
  CLID_OK = 0,
  CLID_EINVAL = -1,       /* null / misaligned / out-of-range argument */
  CLID_EUNSUPPORTED = -2, /* shape outside the compiled kernel set */
  CLID_ECUDA = -3         /* CUDA launch or runtime failure */
};

enum ClidFlags {
  CLID_TRAINING_MODE = 1 << 0, /* certainty scatter-add and, with ts, ts_update amax (neural_points.py:708-733) */
  CLID_QUERY_LOCALLY = 1 << 1, /* remap through global2local, gather from the local window (:595-598) */
  CLID_TIME_FILTER = 1 << 2,   /* travel-distance window on point_ts_create (:1003-1009) */
  CLID_LAYER_NORM = 1 << 3,    /* F.layer_norm over the feature dim, no affine, eps 1e-5 (:632-633) */
  CLID_LEAKY_RELU = 1 << 4,    /* decoder.py:66-74, slope 0.01 */
  CLID_USE_BRICKS = 1 << 5,    /* probe through ClidMap.bricks instead of the hash table */
  CLID_TC_DECODER = 1 << 6     /* evaluate a 64 x 1 decoder on the tensor cores (tcgen05.mma kind::tf32 with split
                                * operands, accumulators in TMEM; csrc/decoder_tc.cuh).  Ignored (fp32 FMA path)
                                * for the other decoder shapes and for the hash-table probe. */
};

/* Brick index: a compact, per-frame restatement of "which neural point does the voxel hash
 * return for cell C, and does it survive the time / locality filters" (DESIGN.md section 3).
 * Space is cut into bricks of 4x4x4 voxels laid out densely over the bounding box of the
 * indexed points.  Each brick header holds a 64-bit occupancy mask (bit = x + 4 y + 16 z) and
 * the index of its first record; records are sorted by (brick, bit) so the record of an
 * occupied cell is base + popcount(mask below its bit).  A query ANDs the masks of the
 * span^3 bricks around it with a precomputed stencil of its neighbourhood and visits only
 * occupied cells.  Candidate sets are identical to the hashed probe (the builder verifies
 * the no-near-collision condition that makes this exact, else the hashed path is used). */
typedef struct ClidBrickHeader {
  uint64_t mask;
  int32_t base;
  int32_t count;
} ClidBrickHeader;

typedef struct ClidBricks {
  const ClidBrickHeader* headers; /* [dims[0]*dims[1]*dims[2]], x fastest                      */
  const uint32_t* hood;           /* [dims[0]*dims[1]*dims[2]][32] or NULL: entry b packs the headers of the
                                     2x2x2 bricks whose lower corner is b into ONE 128-byte line --
                                     words 0..15 occupancy masks (lo, hi) of bricks s = dx + 2 dy + 4 dz,
                                     words 16..23 their first-record indices, 24..31 unused -- so a query
                                     reads its whole neighbourhood directory from a single line
                                     (used when not NULL, else the eight headers are read)       */
  const float* records;           /* [n_records,4] = (px, py, pz, bit-cast int32 gather row)   */
  const uint64_t* stencil;        /* [64][span^3] neighbourhood masks by in-brick position of
                                     the neighbourhood's lower corner, then brick offset      */
  int32_t origin[3];              /* cell coordinate of the first cell of brick (0,0,0)        */
  int32_t dims[3];                /* bricks per axis                                           */
  int32_t span;                   /* bricks per axis one neighbourhood can touch               */
  int32_t reach;                  /* num_nei_cells: the neighbourhood is [-reach, reach]^3     */
  int32_t n_records;
  int32_t apron;                  /* empty bricks surrounding the indexed bricks on every side; the
                                     kernels need 1 (a query is range-tested once, not per brick) */
} ClidBricks;

/* Neural-point map state read by a query.  model/neural_points.py:79-133 */
typedef struct ClidMap {
  const int64_t* buffer_pt_index; /* [buffer_size] voxel hash -> global point id, -1 empty  */
  int64_t buffer_size;
  int64_t primes[3];
  const float* neural_points;     /* [n_global,3]                                           */
  const int32_t* point_ts_create; /* [n_global]  (read only with CLID_TIME_FILTER)          */
  int64_t n_global;
  const float* travel_dist;       /* [n_travel]  (CLID_TIME_FILTER)                         */
  int32_t n_travel;
  int32_t cur_ts;
  float diff_travel_dist_local;
  float resolution;               /* voxel_size_m                                           */
  float max_valid_dist2;
  int32_t kc;                     /* rows of neighbor_dx                                    */
  const int64_t* neighbor_dx;     /* [kc,3] cell offsets (set_search_neighborhood, :931-969) */
  const int64_t* global2local;    /* [n_global+1] (CLID_QUERY_LOCALLY)                      */
  /* arrays the k nearest neighbours are gathered from: local_* with CLID_QUERY_LOCALLY,
   * else the global ones */
  const float* gather_points;     /* [n_gather,3]                                           */
  const float* gather_features;   /* [n_gather+1,feature_dim]                               */
  const float* gather_certainties;/* [n_gather]                                             */
  float* certainty_accum;         /* [n_gather] += weights in training mode; may alias
                                     gather_certainties (then queried certainty races, as
                                     nobody reads it in training) or be a separate buffer   */
  int32_t* gather_ts_update;      /* [n_gather] amax target, or NULL                        */
  int64_t n_gather;
  int32_t feature_dim;            /* 8                                                      */
  int32_t knn;                    /* query_nn_k <= CLID_MAX_KNN                             */
  const ClidBricks* bricks;       /* HOST pointer to a ClidBricks or NULL                   */
  int32_t* work_counter;          /* NULL (recommended): tiles of 32 queries are dealt round-robin to
                                     the resident warps.  Else 4 bytes of zero-initialised device scratch
                                     for a dynamic tile scheduler (each warp draws its next tile with an
                                     atomic; the kernel leaves it zero again; one per stream of concurrent
                                     launches): evens out very uneven tiles, but the ticket's round trip
                                     queues behind the training kernel's own reductions (slower there).  */
} ClidMap;

/* model/decoder.py:13-82: Linear(in->H) act [Linear(H->H) act]^(levels-1) Linear(H->1), x sdf_scale */
typedef struct ClidDecoder {
  const float* weight[CLID_MAX_LEVELS]; /* [H,in] then [H,H], row-major like nn.Linear.weight */
  const float* bias[CLID_MAX_LEVELS];   /* [H] or NULL (mlp_bias_on = False)                  */
  const float* out_weight;              /* [1,H]                                              */
  const float* out_bias;                /* [1] or NULL                                        */
  int32_t in_dim;                       /* feature_dim + 3                                    */
  int32_t hidden_dim;
  int32_t levels;                       /* 1..CLID_MAX_LEVELS                                 */
  float sdf_scale;
} ClidDecoder;

/* Outputs of a forward query; any pointer may be NULL (not produced). */
typedef struct ClidQueryOut {
  float* sdf;        /* [n]        Decoder.sdf(z)                       (needs a decoder)   */
  float* grad;       /* [n,3]      d sdf / d x, closed form of get_gradient (needs decoder) */
  float* z;          /* [n,F+3]    weighted_first geo_features_vector                       */
  float* weights;    /* [n,knn]    weight_vector                                             */
  int32_t* knn_idx;  /* [n,knn]    gather row of each neighbour, ascending distance, -1 pad  */
  int32_t* nn_count; /* [n]        valid candidates among the kc cells (not clamped to knn)  */
  float* certainty;  /* [n]        queried_certainty                                        */
} ClidQueryOut;

CLID_API int clid_version(void);
CLID_API const char* clid_last_error(void);

/* Fused forward of the path
 *   NeuralPoints.radius_neighborhood_search (model/neural_points.py:971-1030)
 *   NeuralPoints.query_feature              (model/neural_points.py:553-769, weighted_first)
 *   Decoder.sdf                             (model/decoder.py:58-82)
 *   get_gradient                            (utils/tools.py:298-311, closed form)
 * x: [n,3] query points; ts: [n] int32 query timestamps or NULL; dec may be NULL when only
 * z / weights / counts are wanted. */
CLID_API int clid_query_forward(const ClidMap* map, const ClidDecoder* dec, const float* x,
                       const int32_t* ts, int64_t n, uint32_t flags, const ClidQueryOut* out,
                       clid_stream_t stream);

/* Backward of query_feature for callers that differentiate through it with torch autograd
 * (what autograd does through index / sort / div / sum at model/neural_points.py:585-749 when
 * utils/tools.py:298-311 get_gradient or loss.backward() runs).  knn_idx is the forward's
 * ClidQueryOut.knn_idx; only the gather_* arrays, knn and feature_dim of `map` are read.
 *   gx    [n,3]   = (d z / d x)^T gz                       (may be NULL)
 *   gfeat [n_gather+1,F] += w_k * gz[:F] per neighbour     (may be NULL; caller zero-fills) */
CLID_API int clid_query_backward(const ClidMap* map, const float* x, const int32_t* knn_idx,
                                 const float* gz, int64_t n, uint32_t flags, float* gx,
                                 float* gfeat, clid_stream_t stream);

/* Backward of the above with respect to gz and the features (grad-of-grad: the analytic
 * eikonal loss differentiates d sdf / d x, utils/mapper.py:695-696,780-798,835).
 *   g_gz  [n,F+3] = (d z / d x) ggx
 *   gfeat [n_gather+1,F] += (d w_k / d x . ggx) * gz[:F] per neighbour   (may be NULL)
 * Second derivatives with respect to x itself are not produced (no caller needs them). */
CLID_API int clid_query_backward_backward(const ClidMap* map, const float* x,
                                          const int32_t* knn_idx, const float* gz,
                                          const float* ggx, int64_t n, uint32_t flags,
                                          float* g_gz, float* gfeat, clid_stream_t stream);

/* ---- training step (utils/mapper.py:642-836) ------------------------------------------------ */

/* sdf_bce_loss (utils/loss.py:44-62) + eikonal term (utils/mapper.py:780-798).
 * sdf holds the n batch predictions followed, in numerical-gradient mode, by the 6 nd
 * central-difference predictions in get_numerical_gradient's order (utils/mapper.py:985-1034:
 * x+ex, x-ex, x+ey, x-ey, x+ez, x-ez, nd rows each).  Exactly one of {grad != NULL, nd > 0}
 * selects the eikonal input; weight_e == 0 disables the term.
 *   dlogit [n + 6 nd]  d L / d (Decoder.mlp output) per evaluated point
 *   dgrad  [n,3]       d L / d (d sdf / d x), analytic mode only
 *   loss   [3] +=      total, bce, eikonal (caller zero-fills) */
typedef struct ClidLossArgs {
  const float* sdf;
  const float* grad;
  const float* label;
  const float* weight;   /* signed sample weights, |.| is applied (mapper.py:747-749); NULL = 1 */
  float* dlogit;
  float* dgrad;
  float* loss;
  int64_t n;
  int64_t nd;
  int64_t n_norm;        /* denominators of the two means; 0 = n / nd.  A rank that holds a shard of the */
  int64_t nd_norm;       /* batch passes the global sizes so that summing over ranks gives the global loss */
  float sdf_scale;
  float weight_e;
  float num_eps;
  int32_t weighted;      /* loss_weight_on */
} ClidLossArgs;
CLID_API int clid_sdf_loss(const ClidLossArgs* args, clid_stream_t stream);

/* Closed-form backward of loss -> decoder -> interpolation (what cur_loss.backward() computes,
 * utils/mapper.py:834-835, including the double backward through get_gradient in analytic
 * mode).  x / knn_idx / dlogit cover n points; dgrad covers the first n_r of them (0 in
 * numerical mode).  Only the gather_* arrays, knn, feature_dim of `map` are read.
 *   gfeat    [n_gather+1,F] +=  d L / d local_geo_features
 *   touched  [n_gather+1]   = 1 for rows that received a contribution (may be NULL)
 *   dec_grad flat [W0 (HxD), b0 (H), wout (H), bout (1)] +=, NULL when the decoder is frozen
 * Compiled for one hidden level with H in {32, 64, 128}; other decoders return
 * CLID_EUNSUPPORTED (clid_train_fused also covers 32 x 2; otherwise the host differentiates through
 * clid_query_backward instead). */
CLID_API int clid_train_backward(const ClidMap* map, const ClidDecoder* dec, const float* x,
                                 const int32_t* knn_idx, const float* dlogit, const float* dgrad,
                                 int64_t n, int64_t n_r, uint32_t flags, float* gfeat,
                                 uint8_t* touched, float* dec_grad, clid_stream_t stream);

/* One-kernel mapping iteration: clid_query_forward (training mode) + clid_sdf_loss +
 * clid_train_backward fused per evaluated point, i.e. utils/mapper.py:660-835 from query_feature to
 * cur_loss.backward() in a single launch.  Same outputs and accumulation rules as the three
 * calls, except that with `scratch` the decoder gradient is left as per-point rows for
 * clid_decoder_grad_reduce (below); sdf_out [n] is optional.  `map` needs everything
 * clid_query_forward needs (CLID_USE_BRICKS honoured) plus certainty_accum.
 *   numerical == 0: analytic eikonal gradient (loss.numerical_grad_on: False); d L / d logit and
 *                   d L / d grad of a sample depend on that sample alone.
 *   numerical != 0: get_numerical_gradient (utils/mapper.py:985-1034) on the samples i % 10 == 0
 *                   (gradient_decimation is fixed at 10 here); the six shifted copies of such a
 *                   sample are evaluated by neighbouring lanes of the same warp and exchanged with
 *                   shuffles.  Other decimations use the three-call path. */
typedef struct ClidTrainFusedArgs {
  const float* x;        /* [n,3] */
  const int32_t* ts;     /* [n] or NULL */
  const float* label;    /* [n] */
  const float* weight;   /* [n] or NULL */
  int64_t n;
  int64_t n_norm;        /* mean denominator of the bce term; 0 = n */
  int64_t nd_norm;       /* mean denominator of the numerical eikonal term; 0 = ceil(n / 10) */
  float weight_e;        /* 0 disables the eikonal term */
  float num_eps;         /* central-difference step, voxel_size_m * num_grad_step_ratio (numerical) */
  int32_t weighted;      /* loss_weight_on */
  int32_t numerical;     /* 0 analytic, 1 numerical eikonal gradient */
  float* gfeat;          /* [n_gather+1,F] += or NULL */
  uint8_t* touched;      /* [n_gather+1] or NULL */
  float* dec_grad;       /* flat [W0,b0,wout,bout] += or NULL (frozen decoder) */
  float* loss;           /* [3] += total, bce, eikonal */
  float* sdf_out;        /* [n] or NULL */
  /* Multi-GPU, samples and neural points partitioned into slabs along one axis (clid_slam_b200/dist.py): the
   * feature rows within the boundary band a rank shares with a slab neighbour must receive the gradients of BOTH
   * ranks.  With peer_grad set the kernel adds every contribution to such a row a second time, straight into the
   * neighbour's gfeat through its peer mapping (red.global.add.v4.f32 over NVLink) -- no pack / send / recv / unpack
   * pass, the exchange rides inside the compute kernel. */
  float* peer_grad[2];   /* gfeat of the lower / upper slab neighbour, peer-mapped device pointers, or NULL         */
  int32_t peer_axis;     /* 0..2: axis the slabs are cut along                                                   */
  int32_t peer_band[4];  /* inclusive cell ranges [lo0, lo1] / [hi0, hi1] of the band shared with the lower /
                            upper neighbour, in voxel cells floor(p[axis] / resolution) of the NEURAL POINT       */
  const int32_t* peer_row[2]; /* PARTITIONED map (every rank holds only its slab plus the neighbours' halves of its
                            bands, so the same neural point has a different row on either side): [n_gather+1]
                            row of each local row in the lower / upper neighbour's table, -1 where the neighbour
                            does not hold it.  NULL: the tables are replicated, rows are numbered alike.      */
  void* scratch;         /* clid_train_fused_scratch_bytes(n, numerical) bytes of device scratch, or NULL.
                            With it every evaluated point writes its 64-byte decoder-gradient row
                            [delta z + s tau ; delta | activation bits] there and the caller reduces the
                            rows into dec_grad with clid_decoder_grad_reduce afterwards (dec_grad is not
                            touched by this call); without it the warps fold the rows themselves inside
                            the one kernel (slower).  Unused when dec_grad == NULL.                    */
  size_t scratch_bytes;
} ClidTrainFusedArgs;
/* Device scratch clid_train_fused wants for n samples (64 bytes per evaluated point, 16-byte aligned). */
CLID_API size_t clid_train_fused_scratch_bytes(int64_t n, int32_t numerical);
/* The same for a given decoder: two-level decoders (32 x 2) write 448-byte rows
 * [delta z + tau0 ; delta | beta1 | delta h1 + tau1 | delta h2 + tau2 | activation bits] and need the scratch even
 * when the decoder is frozen (csrc/mlp_l2.cuh). */
CLID_API size_t clid_train_fused_scratch_bytes_for(const ClidDecoder* dec, int64_t n, int32_t numerical);
CLID_API int clid_train_fused(const ClidMap* map, const ClidDecoder* dec, const ClidTrainFusedArgs* args,
                              uint32_t flags, clid_stream_t stream);

/* Second half of clid_train_fused when it was given scratch: the dense reduction of the per-point rows
 *   Gd[j][i'] = sum_n act'(pre_nj) c'_ni'  ->  dW0 = wout (.) Gd, db0, dwout, dbout
 * (the decoder part of cur_loss.backward(), utils/mapper.py:834-835) accumulated into the flat
 * dec_grad [W0 (HxD), b0 (H), wout (H), bout (1)].  n / numerical as passed to clid_train_fused. */
CLID_API int clid_decoder_grad_reduce(const ClidDecoder* dec, const void* scratch, int64_t n, int32_t numerical,
                                      uint32_t flags, float* dec_grad, clid_stream_t stream);

/* torch.optim.Adam step (utils/tools.py:205-255: betas (0.9, 0.99), eps adam_eps) on the touched
 * neural-point feature rows and on the decoder tensors; applied gradients are re-zeroed.
 * Rows never touched since the optimiser was created have m = v = g = 0, for which dense Adam
 * is the identity, so skipping them is exact (weight_decay != 0 needs touched == NULL). */
typedef struct ClidAdamArgs {
  float* feat;             /* [rows,F] */
  float* feat_grad;
  float* feat_m;
  float* feat_v;
  const uint8_t* touched;  /* [rows] or NULL = all rows */
  int64_t rows;
  float* dec_param[2 * CLID_MAX_LEVELS + 2]; /* W0,b0,(W1,b1,..),wout,bout; NULL entries skipped */
  int32_t dec_numel[2 * CLID_MAX_LEVELS + 2];
  int32_t dec_tensors;
  float* dec_grad;         /* flat in the same order, or NULL (decoder frozen) */
  float* dec_m;
  float* dec_v;
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;            /* 1-based, shared by every parameter like torch's per-call optimiser.  With
                              step_state: >= 0 advances the device counter first, < 0 uses it as it is
                              (a second call for the same optimiser step, e.g. decoder and features
                              updated by two concurrent launches after one clid_adam_advance)        */
  void* step_state;        /* NULL, or 16 bytes of device memory {int32 step; float step_size; float
                              bc2_sqrt; pad} owned by the caller and zero-initialised when the optimiser
                              is created.  With it the step counter lives on the device: the call first
                              advances it (a one-thread kernel evaluates the bias corrections in double,
                              like torch does on the host) and `step` is ignored -- so a whole iteration
                              can be captured once in a CUDA graph and replayed.                       */
} ClidAdamArgs;
CLID_API int clid_adam_step(const ClidAdamArgs* args, clid_stream_t stream);
/* step_state.step += 1 and the bias-correction scalars of the new step (what clid_adam_step does first when
 * step >= 0).  For callers that split one optimiser step over several clid_adam_step(step < 0) launches. */
CLID_API int clid_adam_advance(void* step_state, float lr, float beta1, float beta2, clid_stream_t stream);
/* The same advance at the HEAD of an iteration, clearing the iteration's loss accumulators loss3 [3] in the same launch
 * (replaces a memset and keeps the advance out of the dependency chain fused kernel -> Adam): follow it with
 * clid_train_fused and clid_adam_step(step < 0) launches. */
CLID_API int clid_step_begin(void* step_state, float lr, float beta1, float beta2, float* loss3, clid_stream_t stream);

/* ---- per-frame feeders ---------------------------------------------------------------------- */

/* LocalPointCloudMap.region_specific_sdf_estimation + estimate_plane (model/local_point_cloud_map.py:98-201): for
 * every surface sample, probe the kc cells (neighbor_idx [kc,3]) of the raw-point voxel hash `table`
 * (primes (73856093, 19349663, 83492791), -1 empty), take the four nearest stored points, fit a least-squares
 * plane through them and return the point-to-plane distance when the neighbourhood is flat
 * (sigma_min / (sigma_mid + 1e-6) <= 0.2) and all four points lie within 0.1 m of the plane, else the distance to the
 * nearest stored point; samples with fewer than four neighbours use the nearest point; surface_mask = a stored point
 * was found at all.  One launch instead of ~30 eager ops and a batched cuSOLVER SVD per scan. */
typedef struct ClidLocalCloud {
  const int64_t* table;         /* [buffer_size]                      */
  int64_t buffer_size;
  int64_t primes[3];
  const float* points;          /* [n_points,3] local_point_cloud_map */
  int64_t n_points;
  const int64_t* neighbor_idx;  /* [kc,3]                             */
  int32_t kc;
  float resolution;             /* local_voxel_size_m                 */
  float max_valid_range;
} ClidLocalCloud;
CLID_API int clid_region_sdf(const ClidLocalCloud* cloud, const float* points, int64_t n, float* sdf_abs,
                             uint8_t* surface_mask, clid_stream_t stream);

/* Brick index build (ClidBricks; replaces the ~20 eager torch ops of a per-frame rebuild), three calls around the
 * caller's sort:
 *   clid_brick_keep   per candidate point (the local window with its global ids `gids`, or the whole map with
 *                     gids == NULL): voxel cell, "owns its hash slot" and travel-distance predicates
 *                     (model/neural_points.py:1003-1009 when ts_create != NULL) -> cells [n,3] i32, keep [n] u8, and
 *                     bbox [7] i32 (caller presets {INT_MAX x3, INT_MIN x3, 0}) = cell bounding box + count of the kept
 *   clid_brick_keys   keys [n] i64 = brick * 64 + cell bit for kept points (INT64_MAX otherwise), for the grid with
 *                     first cell lo[3] and dims[0..1] bricks per row / plane
 *   clid_brick_fill   after sorting the keys (order [n] = source index of every sorted position): records
 *                     [n_kept,4], headers [dims product] and, if hood != NULL, the 128-byte neighbourhood lines */
CLID_API int clid_brick_keep(const ClidMap* map, const float* points, const int64_t* gids, int64_t n,
                             const int32_t* ts_create, int32_t* cells, uint8_t* keep, int32_t* bbox, clid_stream_t stream);
CLID_API int clid_brick_keys(const int32_t* cells, const uint8_t* keep, int64_t n, const int32_t* lo3, const int32_t* dims3,
                             int64_t* keys, clid_stream_t stream);
CLID_API int clid_brick_fill(const int64_t* sorted_keys, const int64_t* order, int64_t n_kept, const float* points,
                             const int32_t* dims3, float* records, ClidBrickHeader* headers, uint32_t* hood,
                             clid_stream_t stream);

/* ---- per-frame map maintenance --------------------------------------------------------------
 * Native bodies of NeuralPoints.update / reset_local_map / assign_local_to_global and of voxel_down_sample_torch.
 * Every order-dependent result follows the reference's sequential CPU semantics (the numbering of new and local
 * points is element order, the last duplicate wins a hash slot, ties in a voxel go to the smaller index).
 * The compactions share one workspace: clid_scan_workspace_bytes(n) bytes, 16-byte aligned, contents opaque except
 * for its FIRST TWO int64: {number of selected elements, selection used}, valid once the call has completed. */
CLID_API size_t clid_scan_workspace_bytes(int64_t n);

/* utils/tools.py:639-682 voxel_down_sample_torch (value == NULL) and :685-724 voxel_down_sample_min_value_torch,
 * two calls around the caller's STABLE ascending sort of `keys`:
 *   clid_voxel_keys  keys [n] i64 = voxel key * 1024 + level, level = trunc(d / max(d) * 999) of the distance to the
 *                    voxel centre (or of `value`); stats [8] i32 scratch, stats[7] != 0 afterwards means the voxel
 *                    key does not fit (span >= 2^17 cells): use another path
 *   clid_voxel_pick  out[0 .. count) = source index (`order` of the sort) of the first element of every voxel run,
 *                    in ascending voxel order = the reference's return value; count = workspace int64[0];
 *                    flags [n] u8 and selected [n] i64 are scratch */
CLID_API int clid_voxel_keys(const float* points, const float* value, int64_t n, float voxel_size, int32_t* stats,
                             int64_t* keys, clid_stream_t stream);
CLID_API int clid_voxel_pick(const int64_t* sorted_keys, const int64_t* order, int64_t n, void* workspace,
                             size_t workspace_bytes, uint8_t* flags, int64_t* selected, int64_t* out,
                             clid_stream_t stream);

/* model/neural_points.py:340-385 NeuralPoints.update between the down-sampling and the torch.cat growth.
 *   clid_map_insert_probe   per candidate: hash slot, current owner, "fresh" = slot empty | owner farther than
 *                           sqrt(far2) | owner last updated more than diff_travel_dist_local ago (ts_update != NULL),
 *                           or every candidate when all_fresh; rank [n] = position among the fresh candidates
 *                           (-1 otherwise); workspace int64[0] = number of new points
 *   clid_map_insert_commit  new_points [n_new,3], new_ts_create / new_ts_update [n_new] = the rows the caller
 *                           appends; buffer_pt_index[slot] = m + rank (fresh) or the old owner, the LAST candidate
 *                           of a repeated slot winning (the reference's sequential index_put) */
typedef struct ClidInsertArgs {
  const float* cand;            /* [n,3] down-sampled scan points                  */
  int64_t n;
  int64_t* buffer_pt_index;     /* [buffer_size], updated by the commit            */
  int64_t buffer_size;
  int64_t primes[3];
  const float* neural_points;   /* [m,3]                                           */
  const int32_t* ts_update;     /* [m] point_ts_update, NULL without temporal map  */
  const float* travel_dist;     /* [n_travel]                                      */
  int64_t m;                    /* points in the map before the insert             */
  int64_t n_travel;
  int32_t cur_ts;
  int32_t all_fresh;            /* empty map or cur_ts == reboot_ts                */
  float resolution;
  float far2;                   /* 3 * resolution^2                                */
  float diff_travel_dist_local;
  int64_t* slot;                /* [n] scratch, probe -> commit                    */
  int64_t* owner;               /* [n] scratch, probe -> commit                    */
  uint8_t* fresh;               /* [n] out                                         */
  int64_t* rank;                /* [n] out                                         */
  void* workspace;
  size_t workspace_bytes;
} ClidInsertArgs;
CLID_API int clid_map_insert_probe(const ClidInsertArgs* args, clid_stream_t stream);
CLID_API int clid_map_insert_commit(const ClidInsertArgs* args, float* new_points, int32_t* new_ts_create,
                                    int32_t* new_ts_update, clid_stream_t stream);

/* model/neural_points.py:439-536 NeuralPoints.reset_local_map.
 *   clid_local_window_select  window predicate per point (travel-distance or time-stamp window on ts_create or the
 *                             mid stamp, optional reboot test, fewer than 100 points in the window -> all points;
 *                             then within sqrt(radius2) of the sensor) -> global2local [m+1] (-1 = not local, entry m
 *                             = -1), local_mask [m+1] u8 (entry m = 1), gids [count] ascending global ids;
 *                             count = workspace int64[0]
 *   clid_local_window_gather  local_* copies of the selected rows (local_features gets count + 1 rows: the padding row)
 *   clid_local_window_scatter NeuralPoints.assign_local_to_global (:538-549): the inverse copy of features (incl.
 *                             the padding row), certainties and ts_update */
typedef struct ClidWindowArgs {
  const float* neural_points;   /* [m,3]                                           */
  const int32_t* ts_create;     /* [m]                                             */
  const int32_t* ts_update;     /* [m], read when use_mid_ts                       */
  const float* travel_dist;     /* NULL: |cur_ts - stamp| < diff_ts_local          */
  int64_t m;
  int64_t n_travel;
  double sensor[3];
  double radius2;               /* local_map_radius^2                              */
  int32_t sensor_is_f64;        /* the position tensor was float64 (the distance is then taken in float64) */
  int32_t temporal;             /* temporal_local_map_on                           */
  int32_t use_mid_ts;
  int32_t cur_ts;
  int32_t reboot_test;          /* reboot_map                                      */
  int32_t reboot_ts;
  int32_t diff_ts_local;
  float diff_travel_dist_local;
  uint8_t* flags;               /* [m] scratch                                     */
  int64_t* global2local;        /* [m+1] out                                       */
  uint8_t* local_mask;          /* [m+1] out                                       */
  int64_t* gids;                /* [m] out, first count entries                    */
  void* workspace;
  size_t workspace_bytes;
} ClidWindowArgs;
CLID_API int clid_local_window_select(const ClidWindowArgs* args, clid_stream_t stream);
typedef struct ClidWindowRows {
  const int64_t* gids;          /* [n_local]                                       */
  int64_t n_local;
  int64_t m;
  float* neural_points;         /* [m,3]   global arrays (read by gather, written by scatter where noted) */
  float* point_orientations;    /* [m,4]                                           */
  float* point_certainties;     /* [m]     scatter target                          */
  int32_t* point_ts_update;     /* [m]     scatter target                          */
  float* geo_features;          /* [m+1,8] scatter target                          */
  float* local_points;          /* [n_local,3]   gather targets                    */
  float* local_orientations;    /* [n_local,4]                                     */
  float* local_certainties;     /* [n_local]                                       */
  int32_t* local_ts_update;     /* [n_local]                                       */
  float* local_features;        /* [n_local+1,8]                                   */
} ClidWindowRows;
CLID_API int clid_local_window_gather(const ClidWindowRows* rows, clid_stream_t stream);
CLID_API int clid_local_window_scatter(const ClidWindowRows* rows, clid_stream_t stream);

/* Replay-pool filter of Mapper.process_frame (utils/mapper.py:420-459): keep the samples within sqrt(radius2) of the
 * sensor, in pool order.
 *   clid_pool_filter_select  flags [n] u8, rank [n] i64 (position among the kept, -1 otherwise); kept count =
 *                            workspace int64[0].  use_norm == 0: sum of squares < radius^2 (the pool filter);
 *                            use_norm != 0: torch.norm(p - sensor) < radius with the CPU kernel's rounding
 *                            (LocalPointCloudMap.update_map, model/local_point_cloud_map.py:58-72)
 *   clid_compact_rows        dst[a][rank[i]] = src[a][i] for every kept row of n_arrays (<= 8) arrays whose rows are
 *                            words[a] 32-bit words (coord / global_coord: 3, labels, weights, time stamps: 1) */
CLID_API int clid_pool_filter_select(const float* global_coord, int64_t n, const double* sensor3, double radius,
                                     int32_t sensor_is_f64, int32_t use_norm, uint8_t* flags, int64_t* rank,
                                     void* workspace, size_t workspace_bytes, clid_stream_t stream);
CLID_API int clid_compact_rows(const int64_t* rank, int64_t n, const void* const* src, void* const* dst,
                               const int32_t* words, int32_t n_arrays, clid_stream_t stream);

/* Ray samples of DataSampler.sample_pin / sample (utils/data_sampler.py:35-140, :283-345): for every scan point its
 * 1 + n_surf + n_front + n_behind samples, ray-major.  The caller draws the random numbers in the reference's order
 * (randn for the surface samples, rand for the front and the behind free-space samples; element j * P + i = sample j
 * of point i) and passes depth = |point|.  coord [P*S,3], disp [P*S] (label = -disp), weight [P*S] (negative =
 * free space). */
typedef struct ClidRaySampleArgs {
  const float* points;          /* [P,3] sensor frame                              */
  const float* depth;           /* [P]                                             */
  const float* randn_surf;      /* [n_surf * P]                                    */
  const float* rand_front;      /* [n_front * P]                                   */
  const float* rand_behind;     /* [n_behind * P]                                  */
  int64_t n_points;
  int32_t n_surf, n_front, n_behind;
  float surface_sample_range_m;
  float margin;                 /* 2 * surface_sample_range_m: free-space samples keep this far from the surface */
  float free_sample_begin_ratio;
  float free_sample_end_dist_m;
  float weight_top;             /* 1 + dist_weight_scale / 2                       */
  float dist_weight_scale;
  float max_range;
  int32_t dist_weight_on;
  float* coord;
  float* disp;
  float* weight;
} ClidRaySampleArgs;
CLID_API int clid_ray_samples(const ClidRaySampleArgs* args, clid_stream_t stream);
/* Labels of DataSampler.sample (:347-377): near-surface samples (slots 1..n_surf of every ray) take +-dist (row
 * i * n_surf + j of the region-specific distances, sign = side of the surface) and keep = reachable; the others
 * label = -disp, keep = 1.  label / keep [P*S]. */
CLID_API int clid_ray_labels(const float* disp, const float* dist, const uint8_t* reachable, int64_t n_points,
                             int32_t samples_per_ray, int32_t n_surf, float* label, uint8_t* keep, clid_stream_t stream);
/* rank [n] (position among the elements with flags[i] & 1, -1 otherwise) and their count in workspace int64[0]:
 * the compaction index for clid_compact_rows */
CLID_API int clid_flag_ranks(const uint8_t* flags, int64_t n, int64_t* rank, void* workspace, size_t workspace_bytes,
                             clid_stream_t stream);

/* buffer_pt_index[slot[i]] = value[i] (value == NULL: value_base + i) where the LAST element of a repeated slot wins:
 * what the reference's sequential CPU index_put leaves behind (model/neural_points.py:392, :911-925 recreate_hash,
 * model/local_point_cloud_map.py:52-56, :66-72); CUDA index_put is unordered.  slot may hold torch.fmod's negative
 * remainders.  The touched table entries must not hold values below -1 on entry. */
CLID_API int clid_table_store(const int64_t* slot, const int64_t* value, int64_t n, int64_t value_base,
                              int64_t* buffer_pt_index, int64_t buffer_size, clid_stream_t stream);

/* ---- registration epilogue (utils/error_state_iekf.py:176-264 h_model, :303-309 update_iterated) ---------- */

/* What IEKFOM.update_iterated needs from h_model, reduced on the device: with, per scan point i,
 *   valid_i = nn_count_i >= min_nn  and  min_grad < |g_i| < max_grad          (:230-247; sdf_std == 0 for weighted_first)
 *   h_i     = [ -(g_i^T R [p_i]x) , g_i ]   (fp32, like H[:, 0:3] / H[:, 3:6], :250-255; p_i in the imu frame)
 *   w_i     = 1000 / (1 + (|g_i| - 1)^2) * 0.4 / (0.4 + sdf_i^2)              (fp64, R_inv, :258-262)
 * out[0..20] = upper triangle (row-major) of sum_i w_i h_i h_i^T  (= H^T R^-1 H, its 6 x 6 non-zero block),
 * out[21..26] = sum_i w_i h_i sdf_i (= H^T R^-1 z), out[27] = number of valid points, all accumulated in fp64 (+=;
 * the caller zero-fills).  sdf / grad / nn_count are the outputs of clid_query_forward at the transformed points;
 * valid_out [n] (uint8, optional) receives the mask; rot9 is a HOST array (row-major R).  One launch instead of ~25 eager ops and a boolean-mask
 * compaction (host synchronisation) per registration iteration. */
CLID_API int clid_registration_terms(const float* pc_imu, const float* sdf, const float* grad, const int32_t* nn_count,
                                     int64_t n, const float* rot9, int32_t min_nn, float min_grad, float max_grad,
                                     double* out28, uint8_t* valid_out, clid_stream_t stream);

/* ---- the mapping loop (utils/mapper.py:473-523 get_batch + :642-836 loop body) ----------------------- */

/* Replay pool of the mapper (utils/mapper.py:84-97, 297-333): device pointers, borrowed. */
typedef struct ClidReplayPool {
  const float* coord;      /* [count,3] global_coord_pool (global_coord=True) or coord_pool               */
  const float* sdf_label;  /* [count]                                                                      */
  const float* weight;     /* [count] signed sample weight                                                 */
  const int32_t* time;     /* [count] frame stamps                                                         */
  int64_t count;           /* pool_sample_count                                                            */
  const int64_t* new_idx;  /* [n_new] pool rows of this frame's newly observed samples, or NULL            */
  int64_t n_new;
  int32_t bs_new;          /* how many of the batch come from new_idx (min(n_new, bs_new_sample), 0 = none)  */
} ClidReplayPool;

/* Mapper.get_batch (utils/mapper.py:473-523): n uniformly drawn pool rows -- the last bs_new of them from
 * new_idx -- gathered into x [n,3], label [n], weight [n], ts [n]; index_out [n] int64 (optional) receives the
 * drawn rows.  The draw is a counter-based generator (Philox4x32-10, key = seed, counter = (offset, sample)),
 * i.e. reproducible for (seed, offset) but NOT the sequence torch.randint would produce: callers that need
 * the reference's exact batches feed them through clid_train_fused themselves. */
CLID_API int clid_draw_batch(const ClidReplayPool* pool, int64_t n, uint64_t seed, uint64_t offset, float* x,
                             float* label, float* weight, int32_t* ts, int64_t* index_out, clid_stream_t stream);

/* `iters` iterations of the loop body of Mapper.mapping (utils/mapper.py:642-836) enqueued back to back by ONE
 * call: [clid_draw_batch -> clid_train_fused -> clid_decoder_grad_reduce -> clid_adam_step] x iters, no host
 * work in between (the reference runs ~600 eager launches and ~10 host synchronisations per iteration; the
 * Python driver of this library ~10 ctypes calls).  train.x / label / weight / ts must point at caller-owned
 * [train.n] scratch batches that the loop fills; adam.step_state is required (device-side step counter);
 * train.loss [3] is used as the per-iteration accumulator and loss_history [iters,3] receives every
 * iteration's (total, bce, eikonal). */
typedef struct ClidMappingArgs {
  ClidReplayPool pool;
  ClidTrainFusedArgs train;
  ClidAdamArgs adam;
  int32_t iters;
  uint64_t seed;
  uint64_t offset;         /* iteration i draws with counter offset + i                                    */
  float* loss_history;     /* [iters,3]                                                                    */
} ClidMappingArgs;
CLID_API int clid_mapping_run(const ClidMap* map, const ClidDecoder* dec, const ClidMappingArgs* args,
                              uint32_t flags, clid_stream_t stream);

/* ---- one-shot all-reduce of [decoder gradients | loss] over peer memory (multi-GPU, one box) ---------------
 * Every rank owns `slots` [world][n_floats] and `flags` [world] (uint32, zero-initialised) in peer-accessible
 * device memory.  clid_peer_publish copies the rank's [src0 | src1] into slot `rank` of EVERY rank (its own
 * included) through the peer mappings, then stores a new epoch into flag `rank` of every rank (system-scope
 * release).  clid_peer_reduce waits until all `world` flags of the local rank carry the current epoch -- which also
 * orders the peers' remote gradient adds of clid_train_fused before whatever follows -- and writes the sum over
 * ranks, taken in rank order so every rank gets bit-identical values, back into dst0 / dst1.  Both are single
 * kernel launches that can be captured in a CUDA graph: the epoch lives in `epoch` (device, uint32, zero-initialised,
 * local).  No NCCL, no host involvement. */
typedef struct ClidPeerArgs {
  float* slots_of[8];       /* slots base of every rank (peer-mapped; own = local pointer)                        */
  uint32_t* flags_of[8];    /* flags base of every rank                                                           */
  uint32_t* epoch;          /* local                                                                              */
  int32_t rank, world;      /* world <= 8                                                                         */
  int32_t n0, n1;           /* floats of src0 / src1 (n0 + n1 <= slot stride)                                      */
  int32_t stride;           /* floats per slot                                                                    */
  int32_t timeout_ms;       /* clid_peer_reduce gives up waiting after this long (0 = 2000) and sets *error       */
  int32_t* error;           /* local device int: set to 1 on a wait time-out (a peer died); may be NULL           */
} ClidPeerArgs;
/* cudaDeviceEnablePeerAccess(peer_device) for the current device (idempotent): the kernels of this device may then
 * store / red into memory of peer_device mapped into this process (CUDA IPC). */
CLID_API int clid_enable_peer_access(int32_t peer_device);
/* cudaIpcOpenMemHandle / cudaIpcCloseMemHandle with the CURRENT device as the importing device (lazy peer access to
 * the exporting one): handle64 = the 64-byte cudaIpcMemHandle_t of another process's allocation; *base_out = the base
 * of its mapping in this process. */
/* The one place where the library owns device memory: a zero-filled cudaMalloc allocation that can be exported to
 * the other ranks of the box (its base pointer and its 64-byte cudaIpcMemHandle_t), and its release. */
CLID_API int clid_peer_alloc(size_t bytes, void** ptr_out, void* handle64_out);
CLID_API int clid_peer_free(void* ptr);
CLID_API int clid_ipc_open(const void* handle64, void** base_out);
CLID_API int clid_ipc_close(void* base);
CLID_API int clid_peer_publish(const ClidPeerArgs* args, const float* src0, const float* src1, clid_stream_t stream);
CLID_API int clid_peer_reduce(const ClidPeerArgs* args, float* dst0, float* dst1, clid_stream_t stream);

/* NeuralPoints.radius_neighborhood_search (model/neural_points.py:971-1030): the raw candidate
 * table.  dist2_out [n,kc] f32, idx_out [n,kc] int64 global ids (-1 invalid).  Only
 * CLID_TIME_FILTER is read from flags. */
CLID_API int clid_radius_search(const ClidMap* map, const float* x, int64_t n, uint32_t flags,
                                float* dist2_out, int64_t* idx_out, clid_stream_t stream);

/* NeuralPoints.query_certainty (model/neural_points.py:1032-1051): max of the GLOBAL
 * point_certainties over the probed cells, invalid cells counting as 0.  out [n]. */
CLID_API int clid_query_certainty(const ClidMap* map, const float* x, int64_t n,
                                  const float* point_certainties, float* out, clid_stream_t stream);

/* Decoder.mlp on caller-supplied inputs (model/decoder.py:58-82) and its input gradient (what
 * utils/tools.py:298-311 get_gradient returns for d out / d z), evaluated on the tensor cores:
 * 64 x 1 decoders only.  z [n,11]; out [n] the un-scaled logit (sdf = sdf_scale * out); a [n,11] or NULL;
 * mask [n,2] or NULL: activation pattern, unit j -> bit j % 32 of word j / 32.  CLID_LEAKY_RELU is read from flags. */
CLID_API int clid_decoder_eval(const ClidDecoder* dec, const float* z, int64_t n, uint32_t flags,
                               float* out, float* a, uint32_t* mask, clid_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLID_SDF_H_ */
