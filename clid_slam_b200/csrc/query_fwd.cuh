// Fused forward: voxel kNN search -> IDW feature blend -> decoder MLP -> closed-form
// d sdf / d x.  One thread per query; everything between the probe and the outputs lives
// in registers.  Replaces (reference, CPU/GPU eager torch):
//   model/neural_points.py:971-1030 radius_neighborhood_search
//   model/neural_points.py:553-769  query_feature (weighted_first)
//   model/decoder.py:58-82          Decoder.mlp / sdf
//   utils/tools.py:298-311          get_gradient
#pragma once
#include "common.cuh"
#include "search.cuh"

namespace clid {

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



// kSearch: how the candidates of a query are enumerated (search.cuh)
enum SearchKind { kSearchHashed = 0, kSearchBricks = 1 };

// dynamic shared memory of the search phase behind the decoder weights, in floats
template <int kSearch>
constexpr int search_smem_floats() {
  return kSearch == kSearchHashed ? 2 * CLID_MAX_KC : 2 * 64 * kBrickSlots + (int)(sizeof(BrickScratch) / sizeof(float)) + kStage * 4 * kQueryThreads;
}

template <int H, int L, int K, int kSearch>
__global__ void __launch_bounds__(kQueryThreads, CLID_QUERY_MIN_BLOCKS) query_forward_kernel(const __grid_constant__ QueryParams p) {
  constexpr bool kBricks = kSearch != kSearchHashed;  // the top-K payload is a record index
  extern __shared__ __align__(16) float smem[];
  float* sm_dec = smem;
  constexpr int kDecFloats = H > 0 ? MlpLayout<(H > 0 ? H : 4), (H > 0 ? L : 1)>::kFloats : 0;
  int64_t* cell_mod = reinterpret_cast<int64_t*>(smem + kDecFloats);  // hashed: per-cell hash residues
  uint64_t* stencil = reinterpret_cast<uint64_t*>(smem + kDecFloats);  // bricks: 64 x 8 neighbourhood stencils
  BrickScratch& scratch = *reinterpret_cast<BrickScratch*>(smem + kDecFloats + 2 * 64 * kBrickSlots);
  float4* stage_col = reinterpret_cast<float4*>(&scratch + 1) + threadIdx.x;  // kStage record slots per lane (search.cuh)
  const ClidMap& m = p.map;

  // the first tile's coordinates are requested before anything else: their cold miss overlaps the prologue
  TileScheduler sched(p.map.work_counter, p.n);
  const int64_t tile0 = sched.next();
  float x0 = 0.f, y0 = 0.f, z0 = 0.f;
  if (tile0 >= 0 && tile0 * 32 + (threadIdx.x & 31) < p.n) {
    const float* xp = p.x + 3 * (tile0 * 32 + (threadIdx.x & 31));
    x0 = xp[0]; y0 = xp[1]; z0 = xp[2];
  }
  // asynchronous prologue (common.cuh): stencil by one TMA bulk copy, decoder by cp.async element copies, both
  // completing on mbarriers that are only waited on where the data is first used
  __shared__ StageBarriers stage;
  stage_barriers_init(stage);
  if constexpr (kBricks) stage_stencil_async(stencil, p.bricks.stencil, stage);
  if constexpr (H > 0) stage_decoder_async<H, L>(sm_dec, p.dec, stage);
  if constexpr (!kBricks) {
    for (int c = threadIdx.x; c < m.kc; c += blockDim.x) {
      int64_t h = m.neighbor_dx[3 * c] * m.primes[0] + m.neighbor_dx[3 * c + 1] * m.primes[1] +
                  m.neighbor_dx[3 * c + 2] * m.primes[2];
      cell_mod[c] = floor_mod(h, m.buffer_size);
    }
    __syncthreads();
  }
  bool stencil_ready = !kBricks, decoder_ready = H == 0;

  const bool training = p.flags & CLID_TRAINING_MODE;
  const bool local = p.flags & CLID_QUERY_LOCALLY;
  const bool time_filter = p.flags & CLID_TIME_FILTER;
  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const int knn = m.knn;

  // warp-uniform trip count: every lane stays in the loop so warp votes see full warps
  bool first_tile = true;
  stagger_start();
  for (int64_t tile = tile0; tile >= 0; tile = sched.next()) {
    const int64_t q = tile * 32 + (threadIdx.x & 31);
    const bool live = q < p.n;
    float px = x0, py = y0, pz = z0;
    if (!first_tile) {
      px = py = pz = 0.f;
      if (live) { px = p.x[3 * q]; py = p.x[3 * q + 1]; pz = p.x[3 * q + 2]; }
    }
    first_tile = false;
    TopK<K> top;
    top.init();
    int count = 0;
    if (!stencil_ready) { mbar_wait(&stage.stencil, 0); stencil_ready = true; }
    if constexpr (kSearch == kSearchBricks) count = search_bricks<K, kQueryThreads>(m, p.bricks, stencil, &scratch.want[0][threadIdx.x], stage_col, live, px, py, pz, top);
    else if (live) count = search_hashed<K>(m, cell_mod, px, py, pz, local, time_filter, top);
#if CLID_PF_NEXT_TILE
    {  // the next tile's coordinates travel towards L2 during the blend and the decoder (their load is otherwise a cold
       // miss at the head of the tile); the ticket was drawn at the head of this tile, its round trip is long over
      const int64_t nt = sched.peek();
      if (nt >= 0 && nt * 32 + (threadIdx.x & 31) < p.n) prefetch_l2(p.x + 3 * (nt * 32 + (threadIdx.x & 31)));
    }
#endif
    if (!live) continue;

    // ---- neighbour rows, offsets and inverse-distance weights (neural_points.py:653-706)
    // All K record loads are issued before the first one is used, then the feature rows in batches of
    // three: per-neighbour `if (valid) { load; use }` blocks compile to one global round trip per
    // neighbour (12 exposed round trips per tile, profiles/r2a_*), these to three.
    int row[K];
    float vx[K], vy[K], vz[K], w[K], u[K];
    float S = 0.f;
    if constexpr (kBricks) {
      float4 rec[K];
#pragma unroll
      for (int k = 0; k < K; ++k) rec[k] = __ldg(reinterpret_cast<const float4*>(p.bricks.records) + (top.id[k] < 0 ? 0 : top.id[k]));
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool valid = k < knn && top.id[k] >= 0;
        row[k] = valid ? __float_as_int(rec[k].w) : -1;
        vx[k] = valid ? px - rec[k].x : 0.f; vy[k] = valid ? py - rec[k].y : 0.f; vz[k] = valid ? pz - rec[k].z : 0.f;
        u[k] = valid ? 1.0f / (top.d[k] + kIdwEps) : 0.f;
        S += u[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool valid = k < knn && top.id[k] >= 0;
        row[k] = valid ? top.id[k] : -1;
        const float* g = m.gather_points + 3 * (int64_t)(valid ? row[k] : 0);
        const float qx = __ldg(g), qy = __ldg(g + 1), qz = __ldg(g + 2);
        vx[k] = valid ? px - qx : 0.f; vy[k] = valid ? py - qy : 0.f; vz[k] = valid ? pz - qz : 0.f;
        u[k] = valid ? 1.0f / (top.d[k] + kIdwEps) : 0.f;
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
    for (int k0 = 0; k0 < K; k0 += kFeatBatch) {
      float fb[kFeatBatch][kFeat], cb[kFeatBatch];
#pragma unroll
      for (int j = 0; j < kFeatBatch; ++j) {
        if (k0 + j < K) {
          const int rr = row[k0 + j] < 0 ? 0 : row[k0 + j];  // invalid neighbours read row 0; the result is discarded
          load_feature_row256(m.gather_features, rr, fb[j]);
          cb[j] = p.out.certainty ? __ldg(m.gather_certainties + rr) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < kFeatBatch; ++j) {
        const int k = k0 + j;
        if (k < K && row[k] >= 0) {
          float (&f)[kFeat] = fb[j];
          if (layer_norm) { float mu, rs; layer_norm8(f, mu, rs); }
          cert = fmaf(cb[j], w[k], cert);
#pragma unroll
          for (int i = 0; i < kFeat; ++i) z[i] = fmaf(w[k], f[i], z[i]);
          z[8] = fmaf(w[k], vx[k], z[8]);
          z[9] = fmaf(w[k], vy[k], z[9]);
          z[10] = fmaf(w[k], vz[k], z[10]);
          if (want_grad) mom.add(f, u[k], vx[k], vy[k], vz[k]);  // the only pass over the feature rows
        }
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

      if (!decoder_ready) { mbar_wait(&stage.decoder, 0); decoder_ready = true; }
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
