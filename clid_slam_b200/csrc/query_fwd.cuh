// Fused forward: voxel kNN search -> IDW feature blend -> decoder MLP -> closed-form
// d sdf / d x.  One thread per query; everything between the probe and the outputs lives
// in registers.  Replaces (reference, CPU/GPU eager torch):
//   model/neural_points.py:971-1030 radius_neighborhood_search
//   model/neural_points.py:553-769  query_feature (weighted_first)
//   model/decoder.py:58-82          Decoder.mlp / sdf
//   utils/tools.py:298-311          get_gradient
#pragma once
#include "common.cuh"

namespace clid {

#ifndef CLID_QUERY_MIN_BLOCKS
#define CLID_QUERY_MIN_BLOCKS 4  // resident CTAs per SM the forward kernel is register-budgeted for
#endif
#ifndef CLID_QUERY_THREADS
#define CLID_QUERY_THREADS 128
#endif
constexpr int kQueryThreads = CLID_QUERY_THREADS;
constexpr int kBrickSlots = 8;  // span 2: a neighbourhood touches at most 2x2x2 bricks

struct QueryParams {
  ClidMap map;
  ClidDecoder dec;
  ClidQueryOut out;
  ClidBricks bricks;
  const float* x;
  const int32_t* ts;
  int64_t n;
  uint32_t flags;
};

// Ascending top-K by squared distance with an integer payload; ties keep the earlier candidate.
template <int K>
struct TopK {
  float d[K];
  int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = __int_as_float(0x7f800000); id[k] = -1; }
  }
  __device__ __forceinline__ void insert(float dc, int ic) {
    if (!(dc < d[K - 1])) return;
    // one pass of compare-exchange from the front: the carried element is always the larger one
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool lt = dc < d[k];
      const float dk = d[k];
      const int ik = id[k];
      d[k] = lt ? dc : dk;
      id[k] = lt ? ic : ik;
      dc = lt ? dk : dc;
      ic = lt ? ik : ic;
    }
  }
};

// ---- candidate enumeration through the reference's hash table ----------------------------
// Payload of the top-K: gather row (local row with CLID_QUERY_LOCALLY, else global id).
template <int K>
__device__ __forceinline__ int search_hashed(const ClidMap& m, const int64_t* __restrict__ cell_mod, float px,
                                             float py, float pz, const bool local, const bool time_filter,
                                             TopK<K>& top) {
  constexpr int U = 9;
  const int gx = cell_of(px, m.resolution), gy = cell_of(py, m.resolution), gz = cell_of(pz, m.resolution);
  const int64_t B = m.buffer_size;
  const int64_t m0 = floor_mod((int64_t)gx * m.primes[0] + (int64_t)gy * m.primes[1] + (int64_t)gz * m.primes[2], B);
  float td_cur = 0.f;
  if (time_filter) td_cur = m.travel_dist[m.cur_ts];
  int count = 0;
  for (int c0 = 0; c0 < m.kc; c0 += U) {
    int gi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int c = c0 + u;
      int64_t v = -1;
      if (c < m.kc) {
        int64_t slot = m0 + cell_mod[c];
        slot = slot >= B ? slot - B : slot;
        v = __ldg(m.buffer_pt_index + slot);
      }
      gi[u] = (int)v;
    }
    float cx[U], cy[U], cz[U];
    int li[U];
    int tsc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int g = gi[u] < 0 ? 0 : gi[u];  // invalid lanes read row 0 (always mapped); result discarded
      const float* p = m.neural_points + 3 * (int64_t)g;
      cx[u] = __ldg(p); cy[u] = __ldg(p + 1); cz[u] = __ldg(p + 2);
      li[u] = local ? (int)__ldg(m.global2local + g) : g;
      tsc[u] = time_filter ? __ldg(m.point_ts_create + g) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool ok = gi[u] >= 0;
      if (time_filter) {
        float gap = fabsf(td_cur - __ldg(m.travel_dist + tsc[u]));
        ok = ok && (gap < m.diff_travel_dist_local);
      }
      float d2 = dist2_torch(cx[u] - px, cy[u] - py, cz[u] - pz);  // neighbour - query, as the reference
      ok = ok && !(d2 > m.max_valid_dist2) && li[u] >= 0;
      if (ok) {
        ++count;
        top.insert(d2, li[u]);
      }
    }
  }
  return count;
}

// ---- candidate enumeration through the brick index ---------------------------------------
// Payload of the top-K: record index.  A 64-cell brick is walked as two 32-cell halves (z < 2,
// z >= 2) so every bit operation is a single 32-bit instruction.  The grid carries a one-brick
// empty apron (ClidBricks.apron) and a neighbourhood spans 2 x 2 x 2 bricks (span == 2): a query is
// range-tested once, the eight header addresses follow by constant strides.
// Phase 1: load the 8 brick headers, AND with the stencil, compact the non-empty (want, occupancy,
// first-record) half-brick triples into this lane's scratch column (word s of the column is
// col[s * kStride]: want in slots 0..11, occupancy in 12..23, first record in 24..35).
// Phase 2 (warp-converged): every lane pops up to kWalkBatch candidates, issues their record loads
// together, then ranks them; a warp iterates ceil(max-over-lanes(candidates) / kWalkBatch) times
// instead of diverging inside nested loops.
// A neighbourhood is at most 5 cells wide (reach <= 2): 5 consecutive z cells touch at most 3 of the
// 2-cell z halves, so at most 3 x 2 x 2 half-bricks can be non-empty.
constexpr int kHalfSlots = 12;
#ifndef CLID_WALK_BATCH
#define CLID_WALK_BATCH 6   // measured 2 / 4 / 6 / 8: 61.5 / 58.4 / 57.4 / 57.4 us (forward, 131072 queries, cold L2)
#endif
constexpr int kWalkBatch = CLID_WALK_BATCH;
#ifndef CLID_PF_RECORDS
#define CLID_PF_RECORDS 0   // L2 prefetch of the record lines of every non-empty half-brick: paid off with
                            // 16-B-per-iteration walks, no longer with 6 record loads in flight per lane
#endif
#ifndef CLID_PF_FEATURES
#define CLID_PF_FEATURES 1  // L2 prefetch of the feature row of every candidate that enters the top-K
#endif

struct BrickScratch {  // [slot][thread] columns of a 128-thread CTA
  uint32_t want[kHalfSlots][kQueryThreads];
  uint32_t occ[kHalfSlots][kQueryThreads];
  int base[kHalfSlots][kQueryThreads];
};

template <int K, int kStride>
__device__ __forceinline__ int search_bricks(const ClidMap& m, const ClidBricks& b, const uint64_t* stencil,
                                           uint32_t* col, bool live, float px, float py, float pz,
                                           TopK<K>& top) {
  const int rx = cell_of(px, m.resolution) - b.origin[0] - b.reach;
  const int ry = cell_of(py, m.resolution) - b.origin[1] - b.reach;
  const int rz = cell_of(pz, m.resolution) - b.origin[2] - b.reach;
  const int bx0 = rx >> 2, by0 = ry >> 2, bz0 = rz >> 2;
  const int D0 = b.dims[0], D1 = b.dims[1];
  const bool in = live && (unsigned)bx0 < (unsigned)(D0 - 1) && (unsigned)by0 < (unsigned)(D1 - 1) &&
                  (unsigned)bz0 < (unsigned)(b.dims[2] - 1);
  const float4* records = reinterpret_cast<const float4*>(b.records);
  int nfill = 0;
  if (in) {
    const uint2* st = reinterpret_cast<const uint2*>(stencil) + (((rz & 3) * 4 + (ry & 3)) * 4 + (rx & 3)) * 8;
    const uint4* h0 = reinterpret_cast<const uint4*>(b.headers) + ((int64_t)bz0 * D1 + by0) * D0 + bx0;
    const int sy = D0, sz = D0 * D1;
    uint4 h[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) h[s] = __ldg(h0 + (s & 1) + ((s >> 1) & 1) * sy + (s >> 2) * sz);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const uint2 sten = st[s];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t occ = half ? h[s].y : h[s].x;
        const uint32_t want = occ & (half ? sten.y : sten.x);
        if (want) {
          const int base = (int)h[s].z + (half ? __popc(h[s].x) : 0);
          col[nfill * kStride] = want;
          col[(kHalfSlots + nfill) * kStride] = occ;
          col[(2 * kHalfSlots + nfill) * kStride] = (uint32_t)base;
          ++nfill;
          // the records of a half-brick are contiguous: pull their first and last line towards L2
#if CLID_PF_RECORDS
          prefetch_l2(records + base);
          prefetch_l2(records + base + __popc(occ) - 1);
#endif
        }
      }
    }
  }
  // cursor over the filled slots: a pointer into the lane's column and the number of slots left
  int count = 0, left = nfill;
  const uint32_t* sp = col;
  uint32_t w = 0, occ = 0;
  int base = 0;
  if (nfill > 0) { w = sp[0]; occ = sp[kHalfSlots * kStride]; base = (int)sp[2 * kHalfSlots * kStride]; }
  while (__any_sync(0xffffffffu, w != 0)) {
    int rec[kWalkBatch];
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) {
      rec[j] = -1;
      if (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        rec[j] = base + __popc(occ & ((1u << bit) - 1u));
        if (w == 0 && left > 1) {
          --left;
          sp += kStride;
          w = sp[0]; occ = sp[kHalfSlots * kStride]; base = (int)sp[2 * kHalfSlots * kStride];
        }
      }
    }
    float4 r[kWalkBatch];
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) r[j] = __ldg(records + (rec[j] < 0 ? 0 : rec[j]));
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) {
      const float d2 = dist2_torch(r[j].x - px, r[j].y - py, r[j].z - pz);
      if (rec[j] >= 0 && !(d2 > m.max_valid_dist2)) {
        ++count;
        if (d2 < top.d[K - 1]) {
#if CLID_PF_FEATURES
          prefetch_l2(m.gather_features + (int64_t)__float_as_int(r[j].w) * kFeat);  // likely neighbour
#endif
          top.insert(d2, rec[j]);
        }
      }
    }
  }
  return count;
}

template <int H, int L, int K, bool kBricks>
__global__ void __launch_bounds__(kQueryThreads, CLID_QUERY_MIN_BLOCKS) query_forward_kernel(const __grid_constant__ QueryParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sm_dec = smem;
  constexpr int kDecFloats = H > 0 ? MlpLayout<(H > 0 ? H : 4), (H > 0 ? L : 1)>::kFloats : 0;
  int64_t* cell_mod = reinterpret_cast<int64_t*>(smem + kDecFloats);  // hashed: per-cell hash residues
  uint64_t* stencil = reinterpret_cast<uint64_t*>(smem + kDecFloats);  // bricks: 64 x 8 neighbourhood stencils
  BrickScratch& scratch = *reinterpret_cast<BrickScratch*>(smem + kDecFloats + 2 * 64 * kBrickSlots);
  const ClidMap& m = p.map;

  if constexpr (H > 0) stage_decoder<H, L>(sm_dec, p.dec);
  if constexpr (kBricks) {
    const uint4* st_src = reinterpret_cast<const uint4*>(p.bricks.stencil);
    uint4* st_dst = reinterpret_cast<uint4*>(stencil);
#pragma unroll
    for (int i = threadIdx.x; i < 64 * kBrickSlots / 2; i += kQueryThreads) st_dst[i] = __ldg(st_src + i);
  } else {
    for (int c = threadIdx.x; c < m.kc; c += blockDim.x) {
      int64_t h = m.neighbor_dx[3 * c] * m.primes[0] + m.neighbor_dx[3 * c + 1] * m.primes[1] +
                  m.neighbor_dx[3 * c + 2] * m.primes[2];
      cell_mod[c] = floor_mod(h, m.buffer_size);
    }
  }
  __syncthreads();

  const bool training = p.flags & CLID_TRAINING_MODE;
  const bool local = p.flags & CLID_QUERY_LOCALLY;
  const bool time_filter = p.flags & CLID_TIME_FILTER;
  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const int knn = m.knn;

  // warp-uniform trip count: every lane stays in the loop so warp votes see full warps
  TileScheduler sched(p.map.work_counter, p.n);
  for (int64_t tile = sched.next(); tile >= 0; tile = sched.next()) {
    const int64_t q = tile * 32 + (threadIdx.x & 31);
    const bool live = q < p.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = p.x[3 * q]; py = p.x[3 * q + 1]; pz = p.x[3 * q + 2]; }
    TopK<K> top;
    top.init();
    int count = 0;
    if constexpr (kBricks) count = search_bricks<K, kQueryThreads>(m, p.bricks, stencil, &scratch.want[0][threadIdx.x], live, px, py, pz, top);
    else if (live) count = search_hashed<K>(m, cell_mod, px, py, pz, local, time_filter, top);
    if (!live) continue;

    // ---- neighbour rows, offsets and inverse-distance weights (neural_points.py:653-706)
    int row[K];
    float vx[K], vy[K], vz[K], w[K], u[K];
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool valid = k < knn && top.id[k] >= 0;
      row[k] = -1;
      vx[k] = vy[k] = vz[k] = 0.f;
      u[k] = 0.f;
      if (valid) {
        float qx, qy, qz;
        if constexpr (kBricks) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(p.bricks.records) + top.id[k]);
          qx = r.x; qy = r.y; qz = r.z;
          row[k] = __float_as_int(r.w);
        } else {
          row[k] = top.id[k];
          const float* g = m.gather_points + 3 * (int64_t)row[k];
          qx = __ldg(g); qy = __ldg(g + 1); qz = __ldg(g + 2);
        }
        vx[k] = px - qx; vy[k] = py - qy; vz[k] = pz - qz;
        u[k] = 1.0f / (top.d[k] + kIdwEps);
        S += u[k];
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = row[k] >= 0 ? u[k] / S : 0.f;

    // ---- gather + blend
    float z[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) z[i] = 0.f;
    float cert = 0.f;
    const bool want_grad = H > 0 && p.out.grad != nullptr;
    Moments mom;
    mom.clear();
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (row[k] >= 0) {
        float f[kFeat];
        load_feature_row(m.gather_features, row[k], f);
        if (layer_norm) { float mu, rs; layer_norm8(f, mu, rs); }
        if (p.out.certainty) cert = fmaf(__ldg(m.gather_certainties + row[k]), w[k], cert);
#pragma unroll
        for (int i = 0; i < kFeat; ++i) z[i] = fmaf(w[k], f[i], z[i]);
        z[8] = fmaf(w[k], vx[k], z[8]);
        z[9] = fmaf(w[k], vy[k], z[9]);
        z[10] = fmaf(w[k], vz[k], z[10]);
        if (want_grad) mom.add(f, u[k], vx[k], vy[k], vz[k]);  // the only pass over the feature rows
      }
    }

    // ---- side effects (neural_points.py:708-733)
    if (training) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (row[k] >= 0) {
          atomicAdd(m.certainty_accum + row[k], w[k]);
          if (p.ts && m.gather_ts_update) atomicMax(m.gather_ts_update + row[k], p.ts[q]);
        }
      }
    }

    // ---- outputs of the query
    if (p.out.nn_count) p.out.nn_count[q] = count;
    if (p.out.certainty) p.out.certainty[q] = cert;
    if (p.out.z) {
#pragma unroll
      for (int i = 0; i < kIn; ++i) p.out.z[q * kIn + i] = z[i];
    }
    if (p.out.weights) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < knn) p.out.weights[q * knn + k] = w[k];
    }
    if (p.out.knn_idx) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < knn) p.out.knn_idx[q * knn + k] = row[k];
    }

    // ---- decoder + closed-form spatial gradient
    if constexpr (H > 0) {
      float o, a[kIn];
      mlp_value_and_input_grad<H, L>(sm_dec, z, slope, o, a);
      const float s = p.dec.sdf_scale;
      if (p.out.sdf) p.out.sdf[q] = o * s;
      if (p.out.grad) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (count > 0) {
          float cbar = 0.f;
#pragma unroll
          for (int i = 0; i < kIn; ++i) cbar = fmaf(z[i], a[i], cbar);
          // d u_k / d x = -2 u_k^2 v_k ; sum_k c_k d w_k / d x = (1/S) sum_k (c_k - cbar) d u_k / d x,
          // evaluated from the neighbourhood moments; + a_p because sum_k w_k == 1
          mom.logit_gradient(a, cbar, 1.0f / S, gx, gy, gz);
        }
        p.out.grad[3 * q] = gx * s;
        p.out.grad[3 * q + 1] = gy * s;
        p.out.grad[3 * q + 2] = gz * s;
      }
    }
  }
}

// ---- API-compat kernels: the raw neighbourhood table and the per-query max certainty ------
// model/neural_points.py:971-1030 radius_neighborhood_search -> dist2 [n,kc] f32, idx [n,kc] i64
// model/neural_points.py:1032-1051 query_certainty           -> max over cells of certainty
// One thread per (query, cell) so both outputs are written fully coalesced.
#ifdef CLID_PLAIN_KERNELS  // defined once, in api.cu
__global__ void __launch_bounds__(256) radius_search_kernel(const ClidMap m, const float* __restrict__ x, int64_t n,
                                                            bool time_filter, float* __restrict__ dist2_out,
                                                            int64_t* __restrict__ idx_out) {
  const int64_t total = n * m.kc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / m.kc;
    const int c = (int)(t - q * m.kc);
    const float px = x[3 * q], py = x[3 * q + 1], pz = x[3 * q + 2];
    const int64_t gx = cell_of(px, m.resolution) + m.neighbor_dx[3 * c];
    const int64_t gy = cell_of(py, m.resolution) + m.neighbor_dx[3 * c + 1];
    const int64_t gz = cell_of(pz, m.resolution) + m.neighbor_dx[3 * c + 2];
    const int64_t slot = floor_mod(gx * m.primes[0] + gy * m.primes[1] + gz * m.primes[2], m.buffer_size);
    int64_t gi = m.buffer_pt_index[slot];
    if (time_filter && gi >= 0) {
      float gap = fabsf(m.travel_dist[m.cur_ts] - m.travel_dist[m.point_ts_create[gi]]);
      if (!(gap < m.diff_travel_dist_local)) gi = -1;
    }
    float d2 = m.max_valid_dist2;
    if (gi >= 0) {
      const float* p = m.neural_points + 3 * gi;
      d2 = dist2_torch(p[0] - px, p[1] - py, p[2] - pz);
      if (d2 > m.max_valid_dist2) gi = -1;
    }
    dist2_out[t] = d2;
    idx_out[t] = gi;
  }
}

__global__ void __launch_bounds__(256) query_certainty_kernel(const ClidMap m, const float* __restrict__ x, int64_t n,
                                                              const float* __restrict__ certainties,
                                                              float* __restrict__ out) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const float px = x[3 * q], py = x[3 * q + 1], pz = x[3 * q + 2];
    const int64_t cx = cell_of(px, m.resolution), cy = cell_of(py, m.resolution), cz = cell_of(pz, m.resolution);
    float best = -__int_as_float(0x7f800000);
    for (int c = 0; c < m.kc; ++c) {
      const int64_t gx = cx + m.neighbor_dx[3 * c], gy = cy + m.neighbor_dx[3 * c + 1], gz = cz + m.neighbor_dx[3 * c + 2];
      const int64_t slot = floor_mod(gx * m.primes[0] + gy * m.primes[1] + gz * m.primes[2], m.buffer_size);
      const int64_t gi = m.buffer_pt_index[slot];
      float v = 0.f;  // invalid candidates count as certainty 0 (neural_points.py:1045)
      if (gi >= 0) {
        const float* p = m.neural_points + 3 * gi;
        float d2 = dist2_torch(p[0] - px, p[1] - py, p[2] - pz);
        if (!(d2 > m.max_valid_dist2)) v = certainties[gi];
      }
      best = fmaxf(best, v);
    }
    out[q] = best;
  }
}

#endif  // CLID_PLAIN_KERNELS

}  // namespace clid
