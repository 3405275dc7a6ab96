// Per-frame maintenance of the neural-point map as native kernels.  Replaces the eager-torch
// bodies of
//   utils/tools.py:639-682           voxel_down_sample_torch   (keys -> caller's stable sort -> segment heads)
//   model/neural_points.py:324-437   NeuralPoints.update       (probe, "fresh" predicate, order-preserving
//                                                               numbering of the new points, last-writer-wins table store)
//   model/neural_points.py:439-536   NeuralPoints.reset_local_map (window predicate, global->local numbering, gathers)
//   model/neural_points.py:538-549   NeuralPoints.assign_local_to_global (scatter of the trained window)
// Every order-dependent result (which point of a voxel survives, the numbering of new / local points, which
// duplicate wins a hash slot) follows the reference's sequential CPU semantics, so the map is identical to the
// fixtures produced by the reference on CPU.
//
// All compactions share one three-launch exclusive scan over per-element flags: block sums (2048 elements per
// block) -> one-block scan of the sums -> per-block scan that writes the ranks.  Streams of 1-4 M elements: each
// launch reads its flags once, fully coalesced.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace clid {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // elements per block

__host__ __device__ inline int64_t scan_blocks(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, total in `total`
__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
  __shared__ int warp_sum[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();  // warp_sum may still be read by the previous call
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const int s = warp_sum[w];
    if (w < warp) base += s;
    tot += s;
  }
  total = tot;
  return base + inc - v;
}

// ---- flags: bit 0 and bit 1 of every element are two candidate selections; `sums[2 * block + s]` counts
// selection s in a block.  The scan of the sums picks ONE of the two selections from a device-side rule
// (ScanRule) so that "too few points in the time window -> take everything" needs no host round trip.
struct ScanRule {
  const int32_t* decide;  // NULL: always selection 0; else use selection 1 when *decide < threshold
  int32_t threshold;
};

__global__ void __launch_bounds__(kScanThreads) flag_block_sums_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                      int32_t* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int c0 = 0, c1 = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k * kScanThreads + threadIdx.x;
    if (i < n) { const uint8_t f = flags[i]; c0 += f & 1; c1 += (f >> 1) & 1; }
  }
  c0 = __reduce_add_sync(0xffffffffu, c0);
  c1 = __reduce_add_sync(0xffffffffu, c1);
  __shared__ int s0[kScanThreads / 32], s1[kScanThreads / 32];
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = c0; s1[threadIdx.x >> 5] = c1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) { a += s0[w]; b += s1[w]; }
    sums[2 * blockIdx.x] = a;
    sums[2 * blockIdx.x + 1] = b;
  }
}

// one block: exclusive scan of the chosen selection's block sums -> offsets [n_blocks] i64;
// result[0] = total, result[1] = selection used (0 / 1)
__global__ void __launch_bounds__(kScanThreads) flag_block_offsets_kernel(const int32_t* __restrict__ sums, int64_t n_blocks,
                                                                         ScanRule rule, int64_t* __restrict__ offsets,
                                                                         int64_t* __restrict__ result) {
  const int sel = (rule.decide != nullptr && *rule.decide < rule.threshold) ? 1 : 0;
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
    const int64_t b = b0 + threadIdx.x;
    const int v = b < n_blocks ? sums[2 * b + sel] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, tot);
    if (b < n_blocks) offsets[b] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) { result[0] = carry; result[1] = sel; }
}

// rank of every selected element (its position among the selected, in element order), -1 for the others;
// optionally the list of selected element indices
__global__ void __launch_bounds__(kScanThreads) flag_ranks_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                 const int64_t* __restrict__ offsets,
                                                                 const int64_t* __restrict__ result,
                                                                 int64_t* __restrict__ rank, int64_t* __restrict__ selected,
                                                                 uint8_t* __restrict__ mask) {
  const int sel = (int)result[1];
  // a thread owns kScanItems CONSECUTIVE elements so that one block scan orders the whole tile
  const int64_t first = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int f[kScanItems], c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = first + k;
    f[k] = i < n ? (flags[i] >> sel) & 1 : 0;
    c += f[k];
  }
  int tot;
  int64_t pos = offsets[blockIdx.x] + block_exclusive_scan(c, tot);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = first + k;
    if (i < n) {
      if (rank) rank[i] = f[k] ? pos : -1;
      if (mask) mask[i] = (uint8_t)f[k];
      if (f[k]) {
        if (selected) selected[pos] = i;
        ++pos;
      }
    }
  }
}

// ---- voxel down-sampling (utils/tools.py:639-682) ------------------------------------------------------------
// pass 1: per-axis min / max of the points and the largest distance to a voxel centre.  Floats are reduced through
// their order-preserving integer image so that plain integer atomics do it.
__device__ __forceinline__ int32_t float_order(float v) {
  const int32_t b = __float_as_int(v);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__host__ __device__ inline float float_unorder(int32_t o) {
  const int32_t b = o >= 0 ? o : o ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

__device__ __forceinline__ float centre_distance(float x, float y, float z, float vs) {
  // centre = (floor(p / vs) + 0.5) * vs ; dist = sqrt(((p - centre) ** 2).sum())   (fp32, torch's rounding)
  const float cx = __fmul_rn(__fadd_rn(floorf(__fdiv_rn(x, vs)), 0.5f), vs);
  const float cy = __fmul_rn(__fadd_rn(floorf(__fdiv_rn(y, vs)), 0.5f), vs);
  const float cz = __fmul_rn(__fadd_rn(floorf(__fdiv_rn(z, vs)), 0.5f), vs);
  return __fsqrt_rn(dist2_torch(x - cx, y - cy, z - cz));
}

// stats [8] i32 (caller presets {INT_MAX x3, INT_MIN x3, INT_MIN, 0}): ordered min xyz, max xyz, max distance
__global__ void __launch_bounds__(256) voxel_stats_kernel(const float* __restrict__ pts, int64_t n, float vs,
                                                          int32_t* __restrict__ stats) {
  int32_t lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN}, dm = INT_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int32_t ox = float_order(x), oy = float_order(y), oz = float_order(z);
    lo[0] = min(lo[0], ox); lo[1] = min(lo[1], oy); lo[2] = min(lo[2], oz);
    hi[0] = max(hi[0], ox); hi[1] = max(hi[1], oy); hi[2] = max(hi[2], oz);
    dm = max(dm, float_order(centre_distance(x, y, z, vs)));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  dm = __reduce_max_sync(0xffffffffu, dm);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(stats + a, lo[a]); atomicMax(stats + 3 + a, hi[a]); }
    atomicMax(stats + 6, dm);
  }
}

// pass 2: sort key = voxel key * 1024 + quantised distance (value mode: quantised `value`).  A STABLE ascending sort
// of these keys puts, at the head of every voxel's run, the point with the smallest (level, index) -- what the
// reference's scatter_reduce(amin) over `index + level * 10^digits` selects.  stats[7] is set when the key does
// not fit 63 bits (the caller then uses the eager path).
__global__ void __launch_bounds__(256) voxel_keys_kernel(const float* __restrict__ pts, const float* __restrict__ value,
                                                         const int32_t* __restrict__ value_max_ordered, int64_t n, float vs,
                                                         int32_t* __restrict__ stats, int64_t* __restrict__ keys) {
  const float minx = float_unorder(stats[0]), miny = float_unorder(stats[1]), minz = float_unorder(stats[2]);
  const float maxx = float_unorder(stats[3]), maxy = float_unorder(stats[4]), maxz = float_unorder(stats[5]);
  const int64_t lx = (int64_t)floorf(__fdiv_rn(minx, vs)), ly = (int64_t)floorf(__fdiv_rn(miny, vs)),
                lz = (int64_t)floorf(__fdiv_rn(minz, vs));
  // the reference uses ONE span for all axes: the largest shifted cell coordinate (tools.py:662-663)
  int64_t span = (int64_t)floorf(__fdiv_rn(maxx, vs)) - lx;
  span = max(span, (int64_t)floorf(__fdiv_rn(maxy, vs)) - ly);
  span = max(span, (int64_t)floorf(__fdiv_rn(maxz, vs)) - lz);
  if (blockIdx.x == 0 && threadIdx.x == 0 && span >= (int64_t)1 << 17) stats[7] = 1;  // span^3 * 1024 must stay below 2^63
  const float top = value ? float_unorder(*value_max_ordered) : float_unorder(stats[6]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int64_t cx = (int64_t)floorf(__fdiv_rn(x, vs)) - lx, cy = (int64_t)floorf(__fdiv_rn(y, vs)) - ly,
                  cz = (int64_t)floorf(__fdiv_rn(z, vs)) - lz;
    const float v = value ? value[i] : centre_distance(x, y, z, vs);
    // (v / v.max() * 999).long(): IEEE division, fp32 product, truncation
    const int64_t level = (int64_t)__fmul_rn(__fdiv_rn(v, top), 999.0f);
    keys[i] = (cx + cy * span + cz * span * span) * 1024 + level;
  }
}

// ordered maximum of a float array (the `value.max()` of voxel_down_sample_min_value_torch)
__global__ void __launch_bounds__(256) ordered_max_kernel(const float* __restrict__ v, int64_t n, int32_t* __restrict__ out) {
  int32_t m = INT_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, float_order(v[i]));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// pass 3 (after the stable sort): a sorted position heads a voxel run when its voxel key differs from its predecessor's
__global__ void __launch_bounds__(256) voxel_heads_kernel(const int64_t* __restrict__ sorted_keys, int64_t n,
                                                          uint8_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flags[i] = (i == 0 || (sorted_keys[i] >> 10) != (sorted_keys[i - 1] >> 10)) ? 1 : 0;
}

// out[rank] = order[position] for every head: the kept source indices in ascending voxel-key order
__global__ void __launch_bounds__(256) voxel_pick_kernel(const int64_t* __restrict__ selected, const int64_t* __restrict__ result,
                                                         const int64_t* __restrict__ order, int64_t* __restrict__ out) {
  const int64_t m = result[0];
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x)
    out[r] = order[selected[r]];
}

// ---- NeuralPoints.update (model/neural_points.py:340-385) --------------------------------------------------------
struct InsertParams {
  const float* cand;             // [n,3] down-sampled scan points
  int64_t n;
  int64_t* table;                // [buffer_size] buffer_pt_index
  int64_t buffer_size;
  int64_t primes[3];
  const float* neural_points;    // [m,3]
  const int32_t* ts_update;      // [m] point_ts_update (NULL: no travel-distance test)
  const float* travel_dist;
  int64_t m;                     // points in the map before the insert
  int32_t cur_ts;
  int32_t all_fresh;             // empty map or cur_ts == reboot_ts: every candidate becomes a point
  float resolution;
  float far2;                    // 3 * resolution^2 rounded to fp32
  float diff_travel_dist_local;
  int64_t* slot;                 // [n] out: hash slot of every candidate (non-negative)
  int64_t* owner;                // [n] out: table entry found there
  uint8_t* fresh;                // [n] out
};

__global__ void __launch_bounds__(256) insert_probe_kernel(const InsertParams p) {
  float td_cur = 0.f;
  if (p.ts_update) td_cur = p.travel_dist[p.cur_ts];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = p.cand[3 * i], y = p.cand[3 * i + 1], z = p.cand[3 * i + 2];
    const int64_t cx = cell_of(x, p.resolution), cy = cell_of(y, p.resolution), cz = cell_of(z, p.resolution);
    // torch.fmod keeps the sign and a negative index wraps once: the same entry as the non-negative remainder
    const int64_t s = floor_mod(cx * p.primes[0] + cy * p.primes[1] + cz * p.primes[2], p.buffer_size);
    const int64_t own = p.table[s];
    bool fresh = true;
    if (!p.all_fresh && own >= 0) {
      const float* q = p.neural_points + 3 * own;
      const float d2 = dist2_torch(q[0] - x, q[1] - y, q[2] - z);
      fresh = d2 > p.far2;
      // a voxel whose point was last seen more than the local window ago gets a new point (the stale one stays in
      // the arrays, unreachable through the table)
      if (p.ts_update) fresh = fresh || (td_cur - p.travel_dist[p.ts_update[own]] > p.diff_travel_dist_local);
    }
    p.slot[i] = s;
    p.owner[i] = own;
    p.fresh[i] = fresh ? 1 : 0;
  }
}

// table[slot] = value where the LAST candidate of a repeated slot wins (sequential index_put): every candidate
// bids -2 - i with atomicMin (the largest i leaves the smallest bid), then the candidate whose bid stands writes.
__global__ void __launch_bounds__(256) insert_bid_kernel(const int64_t* __restrict__ slot, int64_t n, int64_t* __restrict__ table) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicMin(reinterpret_cast<long long*>(table + slot[i]), (long long)(-2 - i));
}

// also appends the new points: neural_points_out / ts rows [m + rank]
__global__ void __launch_bounds__(256) insert_commit_kernel(const InsertParams p, const int64_t* __restrict__ rank,
                                                            float* __restrict__ new_points, int32_t* __restrict__ new_ts_create,
                                                            int32_t* __restrict__ new_ts_update) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rank[i];
    if (r >= 0) {
      new_points[3 * r] = p.cand[3 * i]; new_points[3 * r + 1] = p.cand[3 * i + 1]; new_points[3 * r + 2] = p.cand[3 * i + 2];
      new_ts_create[r] = p.cur_ts;
      new_ts_update[r] = p.cur_ts;
    }
    const int64_t s = p.slot[i];
    if (p.table[s] == -2 - i) p.table[s] = r >= 0 ? p.m + r : p.owner[i];
  }
}

// ---- NeuralPoints.reset_local_map (model/neural_points.py:439-536) -------------------------------------------------
struct WindowParams {
  const float* neural_points;    // [m,3]
  const int32_t* ts_create;      // [m]
  const int32_t* ts_update;      // [m] (use_mid_ts)
  const float* travel_dist;      // NULL: compare time stamps (use_travel_dist False)
  int64_t m;
  double sensor[3];
  double radius2;
  int32_t sensor_is_f64;         // the reference subtracts a float64 position from float32 points: promoted to float64
  int32_t temporal;              // temporal_local_map_on
  int32_t use_mid_ts;
  int32_t cur_ts;
  int32_t reboot_ts;             // INT_MIN: no reboot test
  int32_t diff_ts_local;
  float diff_travel_dist_local;
  uint8_t* flags;                // [m] out: bit 0 = in time window AND in range, bit 1 = in range
  int32_t* n_in_time;            // [1] in/out (caller zeroes): points inside the time window
};

__global__ void __launch_bounds__(256) window_flags_kernel(const WindowParams p) {
  int cnt = 0;
  float td_cur = 0.f;
  if (p.temporal && p.travel_dist) td_cur = p.travel_dist[p.cur_ts];
  const float sx = (float)p.sensor[0], sy = (float)p.sensor[1], sz = (float)p.sensor[2];
  const float r2f = (float)p.radius2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.m; i += (int64_t)gridDim.x * blockDim.x) {
    bool in_time = true;
    if (p.temporal) {
      int32_t stamp = p.ts_create[i];
      if (p.use_mid_ts) stamp = (int32_t)((float)(stamp + p.ts_update[i]) / 2.0f);  // ((a + b) / 2).int(): fp32 quotient, truncated
      if (p.travel_dist) in_time = fabsf(td_cur - p.travel_dist[stamp]) < p.diff_travel_dist_local;
      else in_time = abs(p.cur_ts - stamp) < p.diff_ts_local;
      if (p.reboot_ts != INT_MIN) in_time = in_time && stamp >= p.reboot_ts;
    }
    const float x = p.neural_points[3 * i], y = p.neural_points[3 * i + 1], z = p.neural_points[3 * i + 2];
    bool in_range;
    if (p.sensor_is_f64) {
      const double dx = (double)x - p.sensor[0], dy = (double)y - p.sensor[1], dz = (double)z - p.sensor[2];
      in_range = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) < p.radius2;
    } else {
      in_range = dist2_torch(x - sx, y - sy, z - sz) < r2f;
    }
    p.flags[i] = (uint8_t)(((in_time && in_range) ? 1 : 0) | (in_range ? 2 : 0));
    cnt += in_time ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) atomicAdd(p.n_in_time, cnt);
}

// gathers of the window: one thread per local row
struct WindowGather {
  const int64_t* gids;           // [n_local] ascending global ids
  int64_t n_local;
  const float* neural_points;    // [m,3]
  const float* orientations;     // [m,4]
  const float* certainties;      // [m]
  const int32_t* ts_update;      // [m]
  const float* geo_features;     // [m+1,8]
  int64_t m;
  float* local_points;           // [n_local,3]
  float* local_orientations;     // [n_local,4]
  float* local_certainties;      // [n_local]
  int32_t* local_ts_update;      // [n_local]
  float* local_features;         // [n_local+1,8] (last row = padding row m)
};

__global__ void __launch_bounds__(256) window_gather_kernel(const WindowGather p) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= p.n_local; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = r < p.n_local ? p.gids[r] : p.m;
    const float4* src = reinterpret_cast<const float4*>(p.geo_features + g * kFeat);
    float4* dst = reinterpret_cast<float4*>(p.local_features + r * kFeat);
    dst[0] = src[0];
    dst[1] = src[1];
    if (r < p.n_local) {
      p.local_points[3 * r] = p.neural_points[3 * g];
      p.local_points[3 * r + 1] = p.neural_points[3 * g + 1];
      p.local_points[3 * r + 2] = p.neural_points[3 * g + 2];
      reinterpret_cast<float4*>(p.local_orientations)[r] = reinterpret_cast<const float4*>(p.orientations)[g];
      p.local_certainties[r] = p.certainties[g];
      p.local_ts_update[r] = p.ts_update[g];
    }
  }
}

// assign_local_to_global: the inverse scatter (features incl. the padding row, certainties, ts_update)
__global__ void __launch_bounds__(256) window_scatter_kernel(const int64_t* __restrict__ gids, int64_t n_local, int64_t m,
                                                             const float* __restrict__ local_features,
                                                             const float* __restrict__ local_certainties,
                                                             const int32_t* __restrict__ local_ts_update,
                                                             float* __restrict__ geo_features, float* __restrict__ certainties,
                                                             int32_t* __restrict__ ts_update) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_local; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = r < n_local ? gids[r] : m;
    const float4* src = reinterpret_cast<const float4*>(local_features + r * kFeat);
    float4* dst = reinterpret_cast<float4*>(geo_features + g * kFeat);
    dst[0] = src[0];
    dst[1] = src[1];
    if (r < n_local) {
      certainties[g] = local_certainties[r];
      ts_update[g] = local_ts_update[r];
    }
  }
}

// ---- replay-pool filter (utils/mapper.py:420-459) -----------------------------------------------------------------
// flag = the sample lies within sqrt(radius2) of the sensor (window_radius); fp64 when the pose tensor was float64
// use_norm: compare the distance itself, torch.norm(p - sensor) < radius (LocalPointCloudMap.update_map)
__global__ void __launch_bounds__(256) pool_flags_kernel(const float* __restrict__ coord, int64_t n, double sx, double sy, double sz,
                                                         double radius, double radius2, int is_f64, int use_norm,
                                                         uint8_t* __restrict__ flags) {
  const float fx = (float)sx, fy = (float)sy, fz = (float)sz, r2f = (float)radius2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = coord[3 * i], y = coord[3 * i + 1], z = coord[3 * i + 2];
    bool keep;
    if (is_f64) {
      const double dx = (double)x - sx, dy = (double)y - sy, dz = (double)z - sz;
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      keep = use_norm ? __dsqrt_rn(d2) < radius : d2 < radius2;
    } else if (use_norm) {
      // torch.norm of an fp32 tensor on the reference's device (CPU) accumulates in double and rounds once
      const double dx = (double)(x - fx), dy = (double)(y - fy), dz = (double)(z - fz);
      keep = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) < (float)radius;
    } else {
      keep = dist2_torch(x - fx, y - fy, z - fz) < r2f;
    }
    flags[i] = keep ? 1 : 0;
  }
}

// dst[rank[i]] = src[i] for the selected rows of up to kCompactArrays arrays whose rows are `words` 32-bit words
constexpr int kCompactArrays = 8;
struct CompactParams {
  const int64_t* rank;
  int64_t n;
  const uint32_t* src[kCompactArrays];
  uint32_t* dst[kCompactArrays];
  int32_t words[kCompactArrays];
  int32_t n_arrays;
};
__global__ void __launch_bounds__(256) compact_rows_kernel(const CompactParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = p.rank[i];
    if (r < 0) continue;
    for (int a = 0; a < p.n_arrays; ++a) {
      const int w = p.words[a];
      for (int k = 0; k < w; ++k) p.dst[a][r * w + k] = p.src[a][i * w + k];
    }
  }
}

// ---- DataSampler.sample / sample_pin: the ray samples (utils/data_sampler.py:35-140, :283-345) ------------------------
// One thread per scan point: its 1 + n_surf + n_front + n_behind samples, written ray-major (the reference builds
// them block by block with ~30 eager ops and transposes at the end).  The random numbers are drawn by the caller
// (torch.randn / torch.rand in the reference's order: surface, front, behind; element j * P + i belongs to sample j of
// point i) and every arithmetic step keeps torch's fp32 rounding (separate multiply / add, the divisions torch does).
struct RaySampleParams {
  const float* points;     // [P,3] sensor frame
  const float* depth;      // [P]   |point| (torch.linalg.norm by the caller)
  const float* randn_surf; // [n_surf * P]
  const float* rand_front; // [n_front * P]
  const float* rand_behind;// [n_behind * P]
  int64_t P;
  int32_t n_surf, n_front, n_behind;
  float sigma;             // surface_sample_range_m
  float margin_sigma;      // 2 sigma: free-space samples keep this far from the surface
  float begin_ratio;       // free_sample_begin_ratio
  float end_dist;          // free_sample_end_dist_m
  float weight_top;        // 1 + dist_weight_scale / 2
  float inv_max_range;     // 1 / max_range (torch's CUDA division by a python scalar multiplies by the reciprocal)
  float weight_scale;      // dist_weight_scale
  int32_t dist_weight_on;
  float* coord;            // [P * S, 3] ray-major
  float* disp;             // [P * S]    displacement along the ray (label = -disp)
  float* weight;           // [P * S]
};

__global__ void __launch_bounds__(256) ray_samples_kernel(const RaySampleParams p) {
  const int S = 1 + p.n_surf + p.n_front + p.n_behind;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.P; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = p.points[3 * i], y = p.points[3 * i + 1], z = p.points[3 * i + 2];
    const float d = p.depth[i];
    float w_near = 1.0f;
    if (p.dist_weight_on) w_near = __fsub_rn(p.weight_top, __fmul_rn(__fmul_rn(d, p.inv_max_range), p.weight_scale));
    float* c = p.coord + (i * S) * 3;
    float* dp = p.disp + i * S;
    float* w = p.weight + i * S;
    int s = 0;
    c[0] = x; c[1] = y; c[2] = z; dp[0] = 0.f; w[0] = w_near;  // the end point itself (ratio 1)
    ++s;
    for (int j = 0; j < p.n_surf; ++j, ++s) {
      const float ds = __fmul_rn(p.randn_surf[(int64_t)j * p.P + i], p.sigma);
      const float r = __fadd_rn(__fdiv_rn(ds, d), 1.0f);
      c[3 * s] = __fmul_rn(x, r); c[3 * s + 1] = __fmul_rn(y, r); c[3 * s + 2] = __fmul_rn(z, r);
      dp[s] = ds; w[s] = w_near;
    }
    {
      // python_scalar / tensor is tensor.reciprocal() * scalar in torch
      const float hi = __fsub_rn(1.0f, __fmul_rn(__frcp_rn(d), p.margin_sigma));
      const float span = __fsub_rn(hi, p.begin_ratio);
      for (int j = 0; j < p.n_front; ++j, ++s) {
        const float r = __fadd_rn(__fmul_rn(p.rand_front[(int64_t)j * p.P + i], span), p.begin_ratio);
        c[3 * s] = __fmul_rn(x, r); c[3 * s + 1] = __fmul_rn(y, r); c[3 * s + 2] = __fmul_rn(z, r);
        dp[s] = __fmul_rn(__fsub_rn(r, 1.0f), d); w[s] = -1.0f;
      }
    }
    {
      const float hi = __fadd_rn(__fmul_rn(__frcp_rn(d), p.end_dist), 1.0f);
      const float lo = __fadd_rn(1.0f, __fmul_rn(__frcp_rn(d), p.margin_sigma));
      const float span = __fsub_rn(hi, lo);
      for (int j = 0; j < p.n_behind; ++j, ++s) {
        const float r = __fadd_rn(__fmul_rn(p.rand_behind[(int64_t)j * p.P + i], span), lo);
        c[3 * s] = __fmul_rn(x, r); c[3 * s + 1] = __fmul_rn(y, r); c[3 * s + 2] = __fmul_rn(z, r);
        dp[s] = __fmul_rn(__fsub_rn(r, 1.0f), d); w[s] = -1.0f;
      }
    }
  }
}

// labels and keep flags of DataSampler.sample (utils/data_sampler.py:347-377): a near-surface sample takes the
// region-specific distance (row i * n_surf + j of `dist` / `reachable`), signed by the side it was drawn on, and is
// dropped when no stored point is in reach; every other sample keeps label = -disp
__global__ void __launch_bounds__(256) ray_labels_kernel(const float* __restrict__ disp, const float* __restrict__ dist,
                                                         const uint8_t* __restrict__ reachable, int64_t P, int S, int n_surf,
                                                         float* __restrict__ label, uint8_t* __restrict__ keep) {
  const int64_t total = P * S;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / S;
    const int s = (int)(t - i * S);
    const float ds = disp[t];
    float lab = -ds;
    uint8_t k = 1;
    if (s >= 1 && s <= n_surf) {
      const int64_t r = i * n_surf + (s - 1);
      lab = ds < 0.f ? dist[r] : -dist[r];
      k = reachable[r];
    }
    label[t] = lab;
    keep[t] = k;
  }
}

// ---- table[slot] = value with the LAST element of a repeated slot winning (sequential index_put) ------------------
// slots may be negative (torch.fmod keeps the sign; a negative index wraps once)
__global__ void __launch_bounds__(256) table_bid_kernel(const int64_t* __restrict__ slot, int64_t n, int64_t buffer_size,
                                                        int64_t* __restrict__ table) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = slot[i];
    s = s < 0 ? s + buffer_size : s;
    atomicMin(reinterpret_cast<long long*>(table + s), (long long)(-2 - i));
  }
}
__global__ void __launch_bounds__(256) table_commit_kernel(const int64_t* __restrict__ slot, const int64_t* __restrict__ value,
                                                           int64_t n, int64_t buffer_size, int64_t value_base,
                                                           int64_t* __restrict__ table) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = slot[i];
    s = s < 0 ? s + buffer_size : s;
    if (table[s] == -2 - i) table[s] = value ? value[i] : value_base + i;
  }
}

}  // namespace clid
