// One-kernel mapping iteration (utils/mapper.py:642-835): per evaluated point, in one thread and
// without leaving registers,
//   kNN search -> IDW blend -> decoder (-> closed-form d sdf/dx)   (as query_forward_kernel)
//   bce + eikonal loss terms and their derivatives                  (utils/loss.py:44-62, mapper.py:780-798)
//   closed-form backward: feature-gradient scatter, decoder-gradient fold
// Replaces clid_query_forward + clid_sdf_loss + clid_train_backward and the re-gather / MLP
// re-evaluation of the split backward.
//
// Analytic eikonal gradient (numerical_grad_on: False): d L / d logit and d L / d grad of a sample
// depend on that sample alone, so everything the backward needs is still live when they are known.
//
// Numerical eikonal gradient (the default of every shipped run file; mapper.py:985-1034): the
// gradient of sample i (i % 10 == 0) needs the SDF at six shifted copies of it.  The evaluation
// slots are arranged so that those seven evaluations sit in the same warp: a tile covers 20
// consecutive samples,
//     lanes  0.. 6   sample 20t      and its +x -x +y -y +z -z copies
//     lanes  7..13   sample 20t + 10 and its six copies
//     lanes 14..31   the other 18 samples of the tile
// (exactly the reference's x[::10] subset), the six values are exchanged with warp shuffles after
// the decoder, and each shifted lane continues with its own d L / d logit = +- s r_axis / (2 eps).
#pragma once
#include "common.cuh"
#include "query_bwd.cuh"
#include "query_fwd.cuh"
#include "train.cuh"
#include "mlp_l2.cuh"

namespace clid {

struct TrainFusedParams {
  ClidMap map;
  ClidDecoder dec;
  ClidBricks bricks;
  const float* x;        // [n,3]
  const int32_t* ts;     // [n] or NULL
  const float* label;    // [n]
  const float* weight;   // [n] or NULL
  float* gfeat;          // [n_gather+1,8] += or NULL
  uint8_t* touched;      // [n_gather+1] or NULL
  float* dec_grad;       // flat [W0,b0,wout,bout] += or NULL (frozen decoder)
  float* loss;           // [3] += total, bce, eikonal
  float* sdf_out;        // [n] or NULL (diagnostics)
  float* fold_rows;      // kFoldOut: [tiles*32][16] rows (c'[12], activation bits) for decoder_grad_kernel
  float* peer_grad[2];   // gfeat of the lower / upper slab neighbour (peer-mapped) or NULL
  int peer_axis;
  int peer_band[4];      // inclusive cell ranges of the bands shared with the lower / upper neighbour
  const int32_t* peer_row[2];  // partitioned map: local row -> the neighbour's row (-1: absent), or NULL (same numbering)
  int64_t n;
  int64_t n_norm;        // mean denominator of the bce term (global batch size when sharded)
  int64_t nd_norm;       // mean denominator of the numerical eikonal term (global decimated count)
  float weight_e;
  float num_eps;         // central-difference step (numerical mode)
  int weighted;
  uint32_t flags;
};

constexpr int kFusedThreads = kQueryThreads;

// which boundary band (1 = shared with the lower slab neighbour, 2 = with the upper one, 0 = private) the neural
// point at coordinate q of the slab axis lies in; the cell rounding is the host's (dist.SpatialShards)
__device__ __forceinline__ uint32_t band_of(const TrainFusedParams& p, float q, float res) {
  const int c = cell_of(q, res);
  uint32_t b = 0u;
  if (p.peer_grad[0] != nullptr && c >= p.peer_band[0] && c <= p.peer_band[1]) b |= 1u;
  if (p.peer_grad[1] != nullptr && c >= p.peer_band[2] && c <= p.peer_band[3]) b |= 2u;
  return b;
}
constexpr int kNumTileSamples = 20;  // base samples per warp tile in numerical mode

// kFoldOut: the decoder-gradient fold (Gd += d c') is not done by the warp itself; every lane writes
// its 64-byte row [c'(12) | activation bits | pad] to global memory and decoder_grad_kernel
// (tile_kernel.cuh) reduces all rows afterwards.  The warp-serial fold costs ~28 % of this kernel's
// time (profiles/), as a separate dense reduction it is a few microseconds.
// L: hidden levels of the decoder.  L == 2 (H == 32) always hands its decoder gradient to the reduction kernel as
// per-sample rows (mlp_l2.cuh); L == 1 can also fold it inside this kernel (kFoldOut false).
template <int H, int L, int K, int kSearch, bool kNumerical, bool kFoldOut>
__global__ void __launch_bounds__(kFusedThreads, CLID_QUERY_MIN_BLOCKS) train_fused_kernel(const __grid_constant__ TrainFusedParams p) {
  static_assert(L == 1 || (L == 2 && kFoldOut), "two-level decoders write rows for decoder_grad_l2_kernel");
  using Lay = MlpLayout<H, L>;
  constexpr int kRows = H / 32;
  constexpr int kMaskWords = H / 32;
  constexpr int kWarps = kFusedThreads / 32;
  static_assert(kFusedThreads == kQueryThreads, "BrickScratch is sized for kQueryThreads");
  extern __shared__ __align__(16) float smem[];
  float* sm_dec = smem;
  // search scratch (hash residues or stencil + brick cursor columns), then the fold staging
  constexpr bool kBricks = kSearch != kSearchHashed;
  constexpr int kSearchFloats = search_smem_floats<kSearch>();
  int64_t* cell_mod = reinterpret_cast<int64_t*>(smem + Lay::kFloats);
  uint64_t* stencil = reinterpret_cast<uint64_t*>(smem + Lay::kFloats);
  BrickScratch& scratch = *reinterpret_cast<BrickScratch*>(smem + Lay::kFloats + 2 * 64 * kBrickSlots);
  float4* stage_col = reinterpret_cast<float4*>(&scratch + 1) + threadIdx.x;  // kStage record slots per lane (search.cuh)
  float* sm_c = smem + Lay::kFloats + kSearchFloats;                            // [warps][32][12]
  uint32_t* sm_m = reinterpret_cast<uint32_t*>(sm_c + kWarps * 32 * kInPad);     // [warps][32][words]
  float* sm_red = reinterpret_cast<float*>(sm_m + kWarps * 32 * kMaskWords);     // [warps][H][12] partial Gd
  __shared__ float sm_scalar[3][kWarps];
  __shared__ float sm_sample[3][kFusedThreads];  // label, weight, ts of this lane's sample: staged by cp.async at the head of the tile

  const ClidMap& m = p.map;
#if CLID_PF_NEXT_TILE
  {  // the first tile's coordinates start travelling towards L2 before the prologue (static first round: warp w takes tile w)
    const int64_t q0 = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * (kNumerical ? kNumTileSamples : 32) + (threadIdx.x & 31);
    if (q0 < p.n) prefetch_l2(p.x + 3 * q0);
  }
#endif
  // asynchronous prologue (common.cuh): TMA bulk copy of the stencil, cp.async copies of the decoder, mbarriers
  __shared__ StageBarriers stage;
  stage_barriers_init(stage);
  if constexpr (kBricks) stage_stencil_async(stencil, p.bricks.stencil, stage);
  stage_decoder_async<H, L>(sm_dec, p.dec, stage);
  if constexpr (!kBricks) {
    for (int c = threadIdx.x; c < m.kc; c += blockDim.x) {
      int64_t h = m.neighbor_dx[3 * c] * m.primes[0] + m.neighbor_dx[3 * c + 1] * m.primes[1] +
                  m.neighbor_dx[3 * c + 2] * m.primes[2];
      cell_mod[c] = floor_mod(h, m.buffer_size);
    }
    __syncthreads();
  }
  bool stencil_ready = !kBricks, decoder_ready = false;

  const bool local = p.flags & CLID_QUERY_LOCALLY;
  const bool time_filter = p.flags & CLID_TIME_FILTER;
  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const float s = p.dec.sdf_scale;
  const float inv_n = 1.0f / (float)p.n_norm;
  const int knn = m.knn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my_c = sm_c + (warp * 32) * kInPad;
  uint32_t* my_m = sm_m + (warp * 32) * kMaskWords;

  // Gd partial sums of this warp live in shared memory between folds (rows lane, lane + 32, ...),
  // so their registers are free while a sample is being evaluated
  float* my_gd = sm_red + warp * H * kInPad;
  if (!kFoldOut && p.dec_grad) {
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int i = 0; i < kInPad; i += 4)
        *reinterpret_cast<float4*>(my_gd + (lane + 32 * r) * kInPad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float delta_sum = 0.f, bce_sum = 0.f, eik_sum = 0.f;

  // lane role (numerical mode): which sample of the tile, and which of its 7 evaluations
  int role_sample = lane, role_variant = 0;
  if constexpr (kNumerical) {
    if (lane < 14) {
      role_sample = lane < 7 ? 0 : 10;
      role_variant = lane < 7 ? lane : lane - 7;
    } else {
      const int r = lane - 14;             // 0..17 -> samples 1..9 and 11..19
      role_sample = r < 9 ? r + 1 : r + 2;
    }
  }
  const int64_t tile_samples = kNumerical ? kNumTileSamples : 32;
  const int64_t n_tiles_work = (p.n + tile_samples - 1) / tile_samples;

  TileScheduler sched(p.map.work_counter, n_tiles_work * 32);  // the scheduler counts 32-slot tiles
  stagger_start();
  for (int64_t tile = sched.next(); tile >= 0; tile = sched.next()) {
    const int64_t q = tile * tile_samples + role_sample;
    const bool live = q < p.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) {
      px = p.x[3 * q]; py = p.x[3 * q + 1]; pz = p.x[3 * q + 2];
      if constexpr (kNumerical) {
        // x + eps e_a for odd variants 1,3,5 ; x - eps e_a for 2,4,6 (mapper.py:991-1003)
        const float sh = (role_variant & 1) ? p.num_eps : -p.num_eps;
        if (role_variant == 1 || role_variant == 2) px += sh;
        if (role_variant == 3 || role_variant == 4) py += sh;
        if (role_variant == 5 || role_variant == 6) pz += sh;
      }
    }
#if CLID_PF_NEXT_TILE
    if (live && role_variant == 0) {  // read after the decoder, ~10 us from here: cold misses otherwise
      // asynchronous copies (no destination registers) instead of loads after the decoder: ~7 % of this kernel's
      // stall samples were waits on these three cold 4-byte loads (profiles/r2_train_fused_regions.txt)
      cp_async4(&sm_sample[0][threadIdx.x], p.label + q);
      if (p.weight) cp_async4(&sm_sample[1][threadIdx.x], p.weight + q);
      if (p.ts) cp_async4(&sm_sample[2][threadIdx.x], reinterpret_cast<const float*>(p.ts + q));
    }
#endif
    TopK<K> top;
    top.init();
    int count = 0;
    if (!stencil_ready) { mbar_wait(&stage.stencil, 0); stencil_ready = true; }
    if constexpr (kSearch == kSearchBricks) count = search_bricks<K, kQueryThreads>(m, p.bricks, stencil, &scratch.want[0][threadIdx.x], stage_col, live, px, py, pz, top);
    else if (live) count = search_hashed<K>(m, cell_mod, px, py, pz, local, time_filter, top);

#if CLID_PF_NEXT_TILE
    {  // the next tile's coordinates travel towards L2 during the blend and the decoder
      const int64_t nt = sched.peek();
      if (nt >= 0 && nt * tile_samples + role_sample < p.n) prefetch_l2(p.x + 3 * (nt * tile_samples + role_sample));
    }
#endif
    float c[kInPad];
#pragma unroll
    for (int i = 0; i < kInPad; ++i) c[i] = 0.f;
    uint32_t mask[kMaskWords];
#pragma unroll
    for (int w = 0; w < kMaskWords; ++w) mask[w] = 0u;

    // ---- part 1 (per lane): neighbours, blend, side effects, decoder
    int row[K];
    float vx[K], vy[K], vz[K], w[K], u[K];
    float S = 0.f, sdf = 0.f, cbar = 0.f;
    uint32_t m1_bits = 0u, m2_bits = 0u;  // L == 2: activation patterns of the two hidden levels
    float stau[kIn];                      // s tau0: the tangent input of the analytic eikonal term
#pragma unroll
    for (int i = 0; i < kIn; ++i) stau[i] = 0.f;
    uint32_t peer_bits = 0u;  // two bits per neighbour: its row is shared with the lower (1) / upper (2) slab neighbour
    const bool peers = p.peer_grad[0] != nullptr || p.peer_grad[1] != nullptr;
    float z[kIn], a[kIn];
    Moments mom;
#pragma unroll
    for (int i = 0; i < kIn; ++i) { z[i] = 0.f; a[i] = 0.f; }
#pragma unroll
    for (int k = 0; k < K; ++k) { row[k] = -1; vx[k] = vy[k] = vz[k] = 0.f; u[k] = 0.f; w[k] = 0.f; }
    if (live) {
      // all K record loads before the first use, feature rows in batches of three (see query_forward_kernel)
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
          if (peers && valid) peer_bits |= band_of(p, p.peer_axis == 0 ? rec[k].x : (p.peer_axis == 1 ? rec[k].y : rec[k].z), m.resolution) << (2 * k);
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
          if (peers && valid) peer_bits |= band_of(p, p.peer_axis == 0 ? qx : (p.peer_axis == 1 ? qy : qz), m.resolution) << (2 * k);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) w[k] = row[k] >= 0 ? u[k] / S : 0.f;

      // the feature rows pass through registers exactly once: blend (+ neighbourhood moments, which
      // the analytic spatial gradient and the tangent input tau0 = s J r are linear in)
      if constexpr (!kNumerical) mom.clear();
#pragma unroll
      for (int k0 = 0; k0 < K; k0 += kFeatBatch) {
        float fb[kFeatBatch][kFeat];
#pragma unroll
        for (int j = 0; j < kFeatBatch; ++j)
          if (k0 + j < K) load_feature_row256(m.gather_features, row[k0 + j] < 0 ? 0 : row[k0 + j], fb[j]);
#pragma unroll
        for (int j = 0; j < kFeatBatch; ++j) {
          const int k = k0 + j;
          if (k < K && row[k] >= 0) {
            float (&f)[kFeat] = fb[j];
            if (layer_norm) { float mu, rs; layer_norm8(f, mu, rs); }
#pragma unroll
            for (int i = 0; i < kFeat; ++i) z[i] = fmaf(w[k], f[i], z[i]);
            z[8] = fmaf(w[k], vx[k], z[8]); z[9] = fmaf(w[k], vy[k], z[9]); z[10] = fmaf(w[k], vz[k], z[10]);
            if constexpr (!kNumerical) mom.add(f, u[k], vx[k], vy[k], vz[k]);
          }
        }
      }

      // side effects (neural_points.py:708-733); the shifted copies carry no timestamp
      // (Mapper.sdf queries without ts, mapper.py:968-969)
      int32_t ts_q = 0;
      const bool stamp = role_variant == 0 && p.ts && m.gather_ts_update;
#if CLID_PF_NEXT_TILE
      if (stamp) { cp_async_wait_all(); ts_q = __float_as_int(sm_sample[2][threadIdx.x]); }
#else
      if (stamp) ts_q = p.ts[q];
#endif
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (row[k] >= 0) {
          atomicAdd(m.certainty_accum + row[k], w[k]);
          if (stamp) atomicMax(m.gather_ts_update + row[k], ts_q);
        }
      }

      float out;
      if (!decoder_ready) { mbar_wait(&stage.decoder, 0); decoder_ready = true; }
      if constexpr (L == 2) mlp_l2_forward_train<H>(sm_dec, z, slope, out, a, m1_bits, m2_bits, p.fold_rows + (tile * 32 + lane) * L2Row<H>::kFloats);
      else mlp_l1_pairs<H, true>(sm_dec, z, slope, out, a, mask);
      sdf = out * s;
      if (p.sdf_out && role_variant == 0) p.sdf_out[q] = sdf;
#pragma unroll
      for (int i = 0; i < kIn; ++i) cbar = fmaf(z[i], a[i], cbar);
    }

    // ---- part 2: loss terms and d L / d logit (numerical mode: warp exchange of the shifted values)
    float delta = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;  // r = d L / d (d sdf/dx), analytic mode only
    const float invS = (live && count > 0) ? 1.0f / S : 0.f;
    if constexpr (kNumerical) {
      const int g0 = lane < 7 ? 0 : 7;  // first lane of this lane's group (meaningful for lane < 14)
      const float s_xp = __shfl_sync(0xffffffffu, sdf, g0 + 1), s_xn = __shfl_sync(0xffffffffu, sdf, g0 + 2);
      const float s_yp = __shfl_sync(0xffffffffu, sdf, g0 + 3), s_yn = __shfl_sync(0xffffffffu, sdf, g0 + 4);
      const float s_zp = __shfl_sync(0xffffffffu, sdf, g0 + 5), s_zn = __shfl_sync(0xffffffffu, sdf, g0 + 6);
      if (live && lane < 14 && p.weight_e > 0.f) {
        const float two_eps = 2.0f * p.num_eps;
        const float gx = (s_xp - s_xn) / two_eps, gy = (s_yp - s_yn) / two_eps, gz = (s_zp - s_zn) / two_eps;
        const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
        const float dev = gn - 1.0f;
        if (role_variant == 0) eik_sum += dev * dev;
        else {
          // d L / d sdf(x +- eps e_a) = +- r_a / (2 eps), r = weight_e 2 (|g|-1)/nd g/|g| ; times s for the logit
          const float kk = gn > 0.f ? p.weight_e * 2.0f * dev / (float)p.nd_norm / gn * s / two_eps : 0.f;
          const float ga = role_variant <= 2 ? gx : (role_variant <= 4 ? gy : gz);
          delta = (role_variant & 1) ? kk * ga : -kk * ga;
        }
      }
    }
    if (live && role_variant == 0) {
      const float l = sdf / s;  // BCEWithLogits(pred / sigma, sigmoid(label / sigma))
#if CLID_PF_NEXT_TILE
      cp_async_wait_all();
      const float label_q = sm_sample[0][threadIdx.x];
      const float weight_q = (p.weighted && p.weight) ? sm_sample[1][threadIdx.x] : 1.0f;
#else
      const float label_q = p.label[q];
      const float weight_q = (p.weighted && p.weight) ? p.weight[q] : 1.0f;
#endif
      const float t = 1.0f / (1.0f + expf(-(label_q / s)));
      const float wgt = (p.weighted && p.weight) ? fabsf(weight_q) : 1.0f;
      bce_sum += wgt * ((1.0f - t) * l + fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l))));
      delta = wgt * (1.0f / (1.0f + expf(-l)) - t) * inv_n;
      if constexpr (!kNumerical) {
        if (p.weight_e > 0.f) {
          float gx = 0.f, gy = 0.f, gz = 0.f;
          if (count > 0) mom.logit_gradient(a, cbar, invS, gx, gy, gz);  // g_j = s (invS (a.M_j + a.P_j - cbar qv_j) + a_pj)
          gx *= s; gy *= s; gz *= s;
          const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
          const float dev = gn - 1.0f;
          eik_sum += dev * dev;
          const float kk = gn > 0.f ? p.weight_e * 2.0f * dev * inv_n / gn : 0.f;
          rx = kk * gx; ry = kk * gy; rz = kk * gz;
        }
      }
    }

    // ---- part 3 (per lane): c' = [delta z + s tau0 ; delta] and the feature-gradient scatter
    if (live) {
      float dusum = 0.f;
      if constexpr (!kNumerical) {
        // tangent input: sum_k e_k q_k = invS (sum_j r_j [M_j; P_j] - dusum z),  dusum = r . qv
        const float (&M)[3][kFeat] = mom.M;
        const float (&P)[6] = mom.P;
        dusum = rx * mom.qv[0] + ry * mom.qv[1] + rz * mom.qv[2];
        float tau[kIn];
#pragma unroll
        for (int i = 0; i < kFeat; ++i) tau[i] = invS * (rx * M[0][i] + ry * M[1][i] + rz * M[2][i] - dusum * z[i]);
        tau[8] = invS * (rx * P[0] + ry * P[1] + rz * P[2] - dusum * z[8]);
        tau[9] = invS * (rx * P[1] + ry * P[3] + rz * P[4] - dusum * z[9]);
        tau[10] = invS * (rx * P[2] + ry * P[4] + rz * P[5] - dusum * z[10]);
        if (count > 0) { tau[8] += rx; tau[9] += ry; tau[10] += rz; }
#pragma unroll
        for (int i = 0; i < kIn; ++i) { stau[i] = s * tau[i]; c[i] = fmaf(delta, z[i], stau[i]); }
      } else {
#pragma unroll
        for (int i = 0; i < kIn; ++i) c[i] = delta * z[i];
      }
      c[kIn] = delta;
      delta_sum += delta;

      // neural-point feature gradients: dL/df_k = a_f (delta w_k + s e_k), e_k = d w_k/d x . r
      if (p.gfeat) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (row[k] >= 0) {
            float coef = delta * w[k];
            if constexpr (!kNumerical) {
              const float du = -2.f * u[k] * u[k] * (vx[k] * rx + vy[k] * ry + vz[k] * rz);
              coef = fmaf(s, (du - w[k] * dusum) * invS, coef);
            }
            float tt[kFeat];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) tt[i] = coef * a[i];
            if (layer_norm) {
              float f[kFeat], mu, rs;
              load_feature_row(m.gather_features, row[k], f);
              layer_norm8(f, mu, rs);
              layer_norm8_vjp(f, rs, tt);
            }
            red_add_row(p.gfeat, row[k], tt);
            if (p.touched) p.touched[row[k]] = 1;
            // band rows: the same contribution straight into the slab neighbour's gradient table (NVLink)
#pragma unroll
            for (int side = 0; side < 2; ++side) {
              if (peer_bits & ((1u << side) << (2 * k))) {
                const int prow = p.peer_row[side] ? __ldg(p.peer_row[side] + row[k]) : row[k];
                if (prow >= 0) red_add_row(p.peer_grad[side], prow, tt);
              }
            }
          }
        }
      }
    }

    if constexpr (L == 2) {
      // ---- two-level decoder: finish this sample's row (tangent products, c1, e2) or clear it
      if (p.fold_rows) {
        float* l2row = p.fold_rows + (tile * 32 + lane) * L2Row<H>::kFloats;
        if (live) mlp_l2_finish_row<H, !kNumerical>(sm_dec, c, stau, c[kIn], slope, m1_bits, m2_bits, l2row);
        else l2_zero_row<H>(l2row);
      }
    } else if constexpr (kFoldOut) {
      // ---- row for the decoder-gradient reduction (dead lanes write zeros: c stays 0 for them)
      if (p.fold_rows) {
        float4* dst = reinterpret_cast<float4*>(p.fold_rows + (tile * 32 + lane) * 16);
        dst[0] = make_float4(c[0], c[1], c[2], c[3]);
        dst[1] = make_float4(c[4], c[5], c[6], c[7]);
        dst[2] = make_float4(c[8], c[9], c[10], c[11]);
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        mk.x = __uint_as_float(mask[0]);
        if constexpr (kMaskWords > 1) mk.y = __uint_as_float(mask[1]);
        if constexpr (kMaskWords > 2) { mk.z = __uint_as_float(mask[2]); mk.w = __uint_as_float(mask[3]); }
        dst[3] = mk;
      }
    }
    // ---- decoder-gradient fold (see train_backward_l1_kernel)
    if (!kFoldOut && p.dec_grad) {
      float4* dst = reinterpret_cast<float4*>(my_c + lane * kInPad);
      dst[0] = make_float4(c[0], c[1], c[2], c[3]);
      dst[1] = make_float4(c[4], c[5], c[6], c[7]);
      dst[2] = make_float4(c[8], c[9], c[10], c[11]);
#pragma unroll
      for (int ww = 0; ww < kMaskWords; ++ww) my_m[lane * kMaskWords + ww] = mask[ww];
      __syncwarp();
      float Gd[kRows][kInPad];
#pragma unroll
      for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int i = 0; i < kInPad; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(my_gd + (lane + 32 * r) * kInPad + i);
          Gd[r][i] = v.x; Gd[r][i + 1] = v.y; Gd[r][i + 2] = v.z; Gd[r][i + 3] = v.w;
        }
#pragma unroll 4
      for (int nn = 0; nn < 32; ++nn) {
        const float4* src = reinterpret_cast<const float4*>(my_c + nn * kInPad);
        const float4 c0 = src[0], c1 = src[1], c2 = src[2];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const float d = ((my_m[nn * kMaskWords + r] >> lane) & 1u) ? 1.f : slope;
          Gd[r][0] = fmaf(d, c0.x, Gd[r][0]); Gd[r][1] = fmaf(d, c0.y, Gd[r][1]);
          Gd[r][2] = fmaf(d, c0.z, Gd[r][2]); Gd[r][3] = fmaf(d, c0.w, Gd[r][3]);
          Gd[r][4] = fmaf(d, c1.x, Gd[r][4]); Gd[r][5] = fmaf(d, c1.y, Gd[r][5]);
          Gd[r][6] = fmaf(d, c1.z, Gd[r][6]); Gd[r][7] = fmaf(d, c1.w, Gd[r][7]);
          Gd[r][8] = fmaf(d, c2.x, Gd[r][8]); Gd[r][9] = fmaf(d, c2.y, Gd[r][9]);
          Gd[r][10] = fmaf(d, c2.z, Gd[r][10]); Gd[r][11] = fmaf(d, c2.w, Gd[r][11]);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int i = 0; i < kInPad; i += 4)
          *reinterpret_cast<float4*>(my_gd + (lane + 32 * r) * kInPad + i) =
              make_float4(Gd[r][i], Gd[r][i + 1], Gd[r][i + 2], Gd[r][i + 3]);
      __syncwarp();
    }
  }

  // ---- block epilogue: loss scalars, then decoder gradients
  bce_sum = warp_sum(bce_sum);
  eik_sum = warp_sum(eik_sum);
  delta_sum = warp_sum(delta_sum);
  if (lane == 0) { sm_scalar[0][warp] = bce_sum; sm_scalar[1][warp] = eik_sum; sm_scalar[2][warp] = delta_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f, e = 0.f, d = 0.f;
    for (int w = 0; w < kWarps; ++w) { b += sm_scalar[0][w]; e += sm_scalar[1][w]; d += sm_scalar[2][w]; }
    const float bce = b * inv_n;
    const float eik = p.weight_e > 0.f ? e / (float)(kNumerical ? p.nd_norm : p.n_norm) : 0.f;
    atomicAdd(p.loss + 1, bce);
    atomicAdd(p.loss + 2, eik);
    atomicAdd(p.loss + 0, bce + p.weight_e * eik);
    if (!kFoldOut && p.dec_grad && p.dec.out_bias) atomicAdd(p.dec_grad + H * kIn + 2 * H, d);
  }
  if (kFoldOut || !p.dec_grad) return;
  if (!decoder_ready) mbar_wait(&stage.decoder, 0);  // a warp without tiles reads the weights below
  float* gW0 = p.dec_grad;
  float* gb0 = gW0 + H * kIn;
  float* gwout = gb0 + H;
  for (int j0 = threadIdx.x; j0 < H; j0 += blockDim.x) {
    const int j = (j0 + blockIdx.x) % H;  // blocks start at different rows: spreads the same-address atomics in time
    float g[kInPad];
#pragma unroll
    for (int i = 0; i < kInPad; ++i) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += sm_red[(w * H + j) * kInPad + i];
      g[i] = v;
    }
    const float wout = sm_dec[Lay::kWout + j];
    float dw = sm_dec[Lay::kB0 + j] * g[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) {
      atomicAdd(gW0 + j * kIn + i, wout * g[i]);
      dw = fmaf(sm_dec[Lay::kW0 + Lay::w0_index(j, i)], g[i], dw);
    }
    if (p.dec.bias[0]) atomicAdd(gb0 + j, wout * g[kIn]);
    atomicAdd(gwout + j, dw);
  }
}

}  // namespace clid
