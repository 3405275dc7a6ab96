// Phase-parked tile kernel: the fused neural-SDF evaluation (kNN search -> IDW blend -> decoder ->
// closed-form d sdf/dx [-> loss -> backward]) restructured so that every phase fits in 72 registers
// and one SM holds 28 warps (4 CTAs x 7 warps).  At 131072 samples that is ONE wave of 32-sample
// tiles (4096 tiles <= 148 x 28 warps), against 1.7 waves of 16 warps/SM for the register-resident
// kernels in query_fwd.cuh / train_fused.cuh, which it replaces on the brick-index fast path.
//
// Replaces (reference, eager torch):
//   model/neural_points.py:971-1030 radius_neighborhood_search, :553-769 query_feature (weighted_first)
//   model/decoder.py:58-82 Decoder.mlp / sdf,  utils/tools.py:298-311 get_gradient
//   utils/loss.py:44-62 sdf_bce_loss,  utils/mapper.py:780-798 eikonal term, :985-1034 numerical gradient
//   utils/mapper.py:834-835 backward (closed form)
//
// A warp owns one tile; its lanes own one evaluation point each.  State that has to survive a phase
// it is not used in is parked in the warp's private slice of shared memory ([group][lane] float4,
// conflict-free 512-byte rows) instead of registers:
//   search   scratch columns (want / occupancy / first-record of the <= 12 non-empty half-bricks)
//   blend A  neighbour records -> offsets, weights, certainty/ts side effects; positional moments;
//            parks the neighbour list (row, w, t v) for pass B
//   blend B  feature rows pass through registers once: z_f and the feature moments M
//   decoder  packed-FMA MLP (bias folded into the padded input slot); moments parked meanwhile
//   loss     bce + eikonal from the parked moments; c' = [delta z + s tau ; delta] written as a 64-byte
//            row to global memory for the decoder-gradient reduction kernel (decoder_grad_kernel)
//   scatter  feature-gradient vector atomics (neighbour records re-read from L2)
#pragma once
#include "common.cuh"
#include "query_bwd.cuh"
#include "query_fwd.cuh"
#include "train.cuh"

namespace clid {

constexpr int kTileWarps = 7;
constexpr int kTileThreads = kTileWarps * 32;
#ifndef CLID_TILE_BLOCKS
#define CLID_TILE_BLOCKS 4
#endif
constexpr int kTileBlocksPerSm = CLID_TILE_BLOCKS;
constexpr int kParkGroups = 11;   // float4 groups per lane in the warp's park slice
constexpr int kFoldRow = 16;      // floats per row handed to decoder_grad_kernel: c'[12], mask words, pad
constexpr int kNumTile = 20;      // base samples per tile in numerical mode (see train_fused.cuh)

enum TileMode { kTileInfer = 0, kTileTrainAnalytic = 1, kTileTrainNumerical = 2 };

struct TileParams {
  ClidMap map;
  ClidDecoder dec;
  ClidBricks bricks;
  const float* x;        // [n,3]
  const int32_t* ts;     // [n] or NULL
  const float* label;    // [n]          (training)
  const float* weight;   // [n] or NULL  (training)
  float* sdf_out;        // [n] or NULL
  float* grad_out;       // [n,3] or NULL (inference)
  int32_t* nn_count;     // [n] or NULL  (inference)
  float* certainty;      // [n] or NULL  (inference)
  float* gfeat;          // [n_gather+1,8] += or NULL
  uint8_t* touched;      // [n_gather+1] or NULL
  float* fold_rows;      // [tiles*32][16] rows for decoder_grad_kernel, or NULL (frozen decoder)
  float* loss;           // [3] += total, bce, eikonal
  int32_t* work_counter; // dynamic tile scheduler counter or NULL (static: one tile per warp)
  int64_t n;
  int64_t n_norm;
  int64_t nd_norm;
  float weight_e;
  float num_eps;
  int weighted;
  uint32_t flags;
};

// decoder image in shared memory: W0[H][12] with b0 in column 11 | wout[H] | bout, pad
template <int H>
struct TileDec {
  static constexpr int kW0 = 0;
  static constexpr int kWout = H * kInPad;
  static constexpr int kBout = kWout + H;
  static constexpr int kFloats = ((kBout + 1 + 3) / 4) * 4;
};

template <int H>
__device__ __forceinline__ void stage_tile_decoder(float* sm, const ClidDecoder& dec) {
  using Lay = TileDec<H>;
  constexpr int kW = H * kIn;
  // loads first, stores after: one global round trip instead of one per element
  constexpr int kPer = (kW + kTileThreads - 1) / kTileThreads;
  float v[kPer];
#pragma unroll
  for (int r = 0; r < kPer; ++r) {
    const int i = threadIdx.x + r * kTileThreads;
    v[r] = i < kW ? __ldg(dec.weight[0] + i) : 0.f;
  }
  float b = 0.f, wo = 0.f;
  if (threadIdx.x < H) {
    b = dec.bias[0] ? __ldg(dec.bias[0] + threadIdx.x) : 0.f;
    wo = __ldg(dec.out_weight + threadIdx.x);
  }
#pragma unroll
  for (int r = 0; r < kPer; ++r) {
    const int i = threadIdx.x + r * kTileThreads;
    if (i < kW) {
      const int j = i / kIn, c = i - j * kIn;
      sm[Lay::kW0 + j * kInPad + c] = v[r];
    }
  }
  if (threadIdx.x < H) {
    sm[Lay::kW0 + threadIdx.x * kInPad + kIn] = b;
    sm[Lay::kWout + threadIdx.x] = wo;
  }
  if (threadIdx.x == 0) sm[Lay::kBout] = dec.out_bias ? __ldg(dec.out_bias) : 0.f;
}

// One-hidden-level decoder, packed fp32 FMAs, bias folded: zp[5] = (z10, 1).  Returns the logit,
// a = d logit / d z (a[11] is scratch) and the activation bits (unit j -> bit j % 32 of word j / 32).
template <int H, bool kMask>
__device__ __forceinline__ void tile_mlp(const float* __restrict__ sm, const float2 (&zp)[6], float slope, float& out,
                                         float2 (&ap)[6], uint32_t* __restrict__ mask) {
  using Lay = TileDec<H>;
  const float4* w0 = reinterpret_cast<const float4*>(sm + Lay::kW0);
#pragma unroll
  for (int i = 0; i < 6; ++i) ap[i] = make_float2(0.f, 0.f);
  out = sm[Lay::kBout];
#pragma unroll
  for (int jw = 0; jw < H / 32; ++jw) {
    uint32_t bits = 0u;
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const int j = jw * 32 + jj;
      const float4 r0 = w0[j * 3 + 0], r1 = w0[j * 3 + 1], r2 = w0[j * 3 + 2];
      float2 p0 = __fmul2_rn(make_float2(r0.x, r0.y), zp[0]);
      float2 p1 = __fmul2_rn(make_float2(r0.z, r0.w), zp[1]);
      p0 = __ffma2_rn(make_float2(r1.x, r1.y), zp[2], p0);
      p1 = __ffma2_rn(make_float2(r1.z, r1.w), zp[3], p1);
      p0 = __ffma2_rn(make_float2(r2.x, r2.y), zp[4], p0);
      p1 = __ffma2_rn(make_float2(r2.z, r2.w), zp[5], p1);  // r2.w = b0[j], zp[5].y = 1
      const float2 ps = __fadd2_rn(p0, p1);
      const float pre = ps.x + ps.y;
      const bool on = pre > 0.f;
      if (kMask) bits = (bits >> 1) | (on ? 0x80000000u : 0u);  // after 32 steps unit jj sits at bit jj
      const float wo = sm[Lay::kWout + j];
      const float cj = on ? wo : wo * slope;
      out = fmaf(cj, pre, out);
      const float2 cc = make_float2(cj, cj);
      ap[0] = __ffma2_rn(make_float2(r0.x, r0.y), cc, ap[0]);
      ap[1] = __ffma2_rn(make_float2(r0.z, r0.w), cc, ap[1]);
      ap[2] = __ffma2_rn(make_float2(r1.x, r1.y), cc, ap[2]);
      ap[3] = __ffma2_rn(make_float2(r1.z, r1.w), cc, ap[3]);
      ap[4] = __ffma2_rn(make_float2(r2.x, r2.y), cc, ap[4]);
      ap[5] = __ffma2_rn(make_float2(r2.z, r2.w), cc, ap[5]);
    }
    if (kMask) mask[jw] = bits;
  }
}

// one 32-byte feature row with a single 256-bit load (LDG.E.256, sm_100+): one L1 wavefront per
// lane instead of two
__device__ __forceinline__ void load_feature_row256(const float* __restrict__ feats, int id, float (&f)[kFeat]) {
  const float* p = feats + (int64_t)id * kFeat;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(p));
}
template <int H, int K, int kMode>
__global__ void __launch_bounds__(kTileThreads, kTileBlocksPerSm) sdf_tile_kernel(const __grid_constant__ TileParams p) {
  using Lay = TileDec<H>;
  static_assert(K <= 6, "the park slice holds six neighbour entries");
  constexpr bool kTrain = kMode != kTileInfer;
  constexpr bool kNumerical = kMode == kTileTrainNumerical;
  constexpr bool kMoments = kMode != kTileTrainNumerical;  // analytic d sdf/dx is evaluated in the kernel
  constexpr int kMaskWords = H / 32;
  extern __shared__ __align__(16) float smem[];
  float* sm_dec = smem;
  uint64_t* stencil = reinterpret_cast<uint64_t*>(smem + Lay::kFloats);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* park = reinterpret_cast<float4*>(smem + Lay::kFloats + 2 * 64 * 8) + warp * (kParkGroups * 32) + lane;
  uint32_t* col = reinterpret_cast<uint32_t*>(smem + Lay::kFloats + 2 * 64 * 8) + warp * (kParkGroups * 128) + lane;
  __shared__ float sm_scalar[3][kTileWarps];

  const ClidMap& m = p.map;
  stage_tile_decoder<H>(sm_dec, p.dec);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bricks.stencil);
    uint4* dst = reinterpret_cast<uint4*>(stencil);
    for (int i = threadIdx.x; i < 64 * 8 / 2; i += kTileThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const float s = p.dec.sdf_scale;
  const int knn = m.knn;
  const float4* records = reinterpret_cast<const float4*>(p.bricks.records);

  int role_sample = lane, role_variant = 0;
  if constexpr (kNumerical) {
    if (lane < 14) {
      role_sample = lane < 7 ? 0 : 10;
      role_variant = lane < 7 ? lane : lane - 7;
    } else {
      const int r = lane - 14;  // 0..17 -> samples 1..9 and 11..19
      role_sample = r < 9 ? r + 1 : r + 2;
    }
  }
  const int64_t tile_samples = kNumerical ? kNumTile : 32;
  const int64_t n_tiles = (p.n + tile_samples - 1) / tile_samples;
  float bce_sum = 0.f, eik_sum = 0.f;

  TileScheduler sched(p.work_counter, n_tiles * 32);
  for (int64_t tile = sched.next(); tile >= 0; tile = sched.next()) {
    const int64_t q = tile * tile_samples + role_sample;
    const bool live = q < p.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) {
      px = p.x[3 * q]; py = p.x[3 * q + 1]; pz = p.x[3 * q + 2];
      if constexpr (kNumerical) {
        const float sh = (role_variant & 1) ? p.num_eps : -p.num_eps;
        if (role_variant == 1 || role_variant == 2) px += sh;
        if (role_variant == 3 || role_variant == 4) py += sh;
        if (role_variant == 5 || role_variant == 6) pz += sh;
      }
    }

    // ---- search
    TopK<K> top;
    top.init();
    const int count = search_bricks<K, 32>(m, p.bricks, stencil, col, live, px, py, pz, top);
    __syncwarp();  // the scratch columns are re-used as the park slice below

    // ---- blend, pass A: neighbour records -> offsets, weights, side effects, positional moments
    int rid[K];
    float S = 0.f, cert = 0.f;
    float zp8 = 0.f, zp9 = 0.f, zp10 = 0.f;
    {
      float4 rk[K];
      float u[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        rid[k] = (k < knn) ? top.id[k] : -1;
        rk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rid[k] >= 0) rk[k] = __ldg(records + rid[k]);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = rid[k] >= 0 ? 1.0f / (top.d[k] + kIdwEps) : 0.f;
        S += u[k];
      }
      float P[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, qv[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float4 e0 = make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
        float tz = 0.f;
        if (rid[k] >= 0) {
          const int row = __float_as_int(rk[k].w);
          const float vx = px - rk[k].x, vy = py - rk[k].y, vz = pz - rk[k].z;
          const float w = u[k] / S;
          const float t2 = -2.f * u[k] * u[k];
          const float tx = t2 * vx, ty = t2 * vy;
          tz = t2 * vz;
          zp8 = fmaf(w, vx, zp8); zp9 = fmaf(w, vy, zp9); zp10 = fmaf(w, vz, zp10);
          if constexpr (kMoments) {
            P[0] = fmaf(tx, vx, P[0]); P[1] = fmaf(tx, vy, P[1]); P[2] = fmaf(tx, vz, P[2]);
            P[3] = fmaf(ty, vy, P[3]); P[4] = fmaf(ty, vz, P[4]); P[5] = fmaf(tz, vz, P[5]);
            qv[0] += tx; qv[1] += ty; qv[2] += tz;
          }
          if constexpr (kTrain) {
            // side effects (neural_points.py:708-733); the shifted copies carry no timestamp
            atomicAdd(m.certainty_accum + row, w);
            if (role_variant == 0 && p.ts && m.gather_ts_update) atomicMax(m.gather_ts_update + row, p.ts[q]);
          } else if (p.certainty) {
            cert = fmaf(__ldg(m.gather_certainties + row), w, cert);
          }
          e0 = make_float4(rk[k].w, w, tx, ty);
        }
        park[k * 32] = e0;                                   // groups 0..5: (row, w, tx, ty)
        reinterpret_cast<float*>(park + (6 + (k >> 2)) * 32)[k & 3] = tz;  // groups 6..7: tz
      }
      if constexpr (kMoments) {
        park[8 * 32] = make_float4(P[0], P[1], P[2], P[3]);
        park[9 * 32] = make_float4(P[4], P[5], qv[0], qv[1]);
        park[10 * 32] = make_float4(qv[2], 0.f, 0.f, 0.f);
      }
    }

    // ---- blend, pass B: the feature rows pass through registers exactly once
    float2 zp[6];
    {
      float zf[kFeat];
      float M[3][kFeat];
#pragma unroll
      for (int i = 0; i < kFeat; ++i) { zf[i] = 0.f; M[0][i] = M[1][i] = M[2][i] = 0.f; }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 e0 = park[k * 32];
        const int row = __float_as_int(e0.x);
        if (row >= 0) {
          float f[kFeat];
          load_feature_row256(m.gather_features, row, f);
          if (layer_norm) { float mu, rs; layer_norm8(f, mu, rs); }
#pragma unroll
          for (int i = 0; i < kFeat; ++i) zf[i] = fmaf(e0.y, f[i], zf[i]);
          if constexpr (kMoments) {
            const float tz = reinterpret_cast<const float*>(park + (6 + (k >> 2)) * 32)[k & 3];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) {
              M[0][i] = fmaf(e0.z, f[i], M[0][i]); M[1][i] = fmaf(e0.w, f[i], M[1][i]); M[2][i] = fmaf(tz, f[i], M[2][i]);
            }
          }
        }
      }
      if constexpr (kMoments) {
        // the neighbour list is dead: its groups take the feature moments across the decoder
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          park[(2 * j) * 32] = make_float4(M[j][0], M[j][1], M[j][2], M[j][3]);
          park[(2 * j + 1) * 32] = make_float4(M[j][4], M[j][5], M[j][6], M[j][7]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) zp[i] = make_float2(zf[2 * i], zf[2 * i + 1]);
      zp[4] = make_float2(zp8, zp9);
      zp[5] = make_float2(zp10, 1.0f);
    }

    // ---- decoder
    float out;
    float2 ap[6];
    uint32_t mask[kMaskWords];
    tile_mlp<H, kTrain>(sm_dec, zp, slope, out, ap, mask);
    const float sdf = out * s;
    float cbar = zp[5].x * ap[5].x;
#pragma unroll
    for (int i = 0; i < 5; ++i) cbar = fmaf(zp[i].x, ap[i].x, fmaf(zp[i].y, ap[i].y, cbar));
    const float invS = (live && count > 0) ? 1.0f / S : 0.f;

    if constexpr (!kTrain) {
      // ---- inference outputs
      if (live) {
        if (p.sdf_out) p.sdf_out[q] = sdf;
        if (p.nn_count) p.nn_count[q] = count;
        if (p.certainty) p.certainty[q] = cert;
        if (p.grad_out) {
          float gx = 0.f, gy = 0.f, gz = 0.f;
          if (count > 0) {
            const float4 P0 = park[8 * 32], P1 = park[9 * 32], P2 = park[10 * 32];
            float sx = ap[4].x * P0.x + ap[4].y * P0.y + ap[5].x * P0.z - cbar * P1.z;
            float sy = ap[4].x * P0.y + ap[4].y * P0.w + ap[5].x * P1.x - cbar * P1.w;
            float sz = ap[4].x * P0.z + ap[4].y * P1.x + ap[5].x * P1.y - cbar * P2.x;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 mx = park[(0 + h2) * 32], my = park[(2 + h2) * 32], mz = park[(4 + h2) * 32];
              const float2 a0 = ap[2 * h2], a1 = ap[2 * h2 + 1];
              sx = fmaf(a0.x, mx.x, fmaf(a0.y, mx.y, fmaf(a1.x, mx.z, fmaf(a1.y, mx.w, sx))));
              sy = fmaf(a0.x, my.x, fmaf(a0.y, my.y, fmaf(a1.x, my.z, fmaf(a1.y, my.w, sy))));
              sz = fmaf(a0.x, mz.x, fmaf(a0.y, mz.y, fmaf(a1.x, mz.z, fmaf(a1.y, mz.w, sz))));
            }
            gx = fmaf(invS, sx, ap[4].x); gy = fmaf(invS, sy, ap[4].y); gz = fmaf(invS, sz, ap[5].x);
          }
          p.grad_out[3 * q] = gx * s; p.grad_out[3 * q + 1] = gy * s; p.grad_out[3 * q + 2] = gz * s;
        }
      }
      __syncwarp();
      continue;
    } else {
      // ---- loss terms and d L / d logit
      const float inv_n = 1.0f / (float)p.n_norm;
      float delta = 0.f, rx = 0.f, ry = 0.f, rz = 0.f;
      if (p.sdf_out && live && role_variant == 0) p.sdf_out[q] = sdf;
      if constexpr (kNumerical) {
        const int g0 = lane < 7 ? 0 : 7;
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
            const float kk = gn > 0.f ? p.weight_e * 2.0f * dev / (float)p.nd_norm / gn * s / two_eps : 0.f;
            const float ga = role_variant <= 2 ? gx : (role_variant <= 4 ? gy : gz);
            delta = (role_variant & 1) ? kk * ga : -kk * ga;
          }
        }
      }
      float c[kInPad];
#pragma unroll
      for (int i = 0; i < kInPad; ++i) c[i] = 0.f;
      float dusum = 0.f;
      if (live) {
        if (role_variant == 0) {
          const float l = sdf / s;  // BCEWithLogits(pred / sigma, sigmoid(label / sigma))
          const float t = 1.0f / (1.0f + expf(-(p.label[q] / s)));
          const float wgt = (p.weighted && p.weight) ? fabsf(p.weight[q]) : 1.0f;
          bce_sum += wgt * ((1.0f - t) * l + fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l))));
          delta = wgt * (1.0f / (1.0f + expf(-l)) - t) * inv_n;
        }
        if constexpr (!kNumerical) {
          if (p.weight_e > 0.f) {
            // moments back from the park slice: M_j in groups 2j, 2j+1; P, qv in groups 8..10
            const float4 P0 = park[8 * 32], P1 = park[9 * 32], P2 = park[10 * 32];
            float4 Mx0 = park[0 * 32], Mx1 = park[1 * 32], My0 = park[2 * 32], My1 = park[3 * 32];
            float4 Mz0 = park[4 * 32], Mz1 = park[5 * 32];
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (count > 0) {
              float sx = ap[4].x * P0.x + ap[4].y * P0.y + ap[5].x * P0.z - cbar * P1.z;
              float sy = ap[4].x * P0.y + ap[4].y * P0.w + ap[5].x * P1.x - cbar * P1.w;
              float sz = ap[4].x * P0.z + ap[4].y * P1.x + ap[5].x * P1.y - cbar * P2.x;
              sx = fmaf(ap[0].x, Mx0.x, fmaf(ap[0].y, Mx0.y, fmaf(ap[1].x, Mx0.z, fmaf(ap[1].y, Mx0.w, sx))));
              sx = fmaf(ap[2].x, Mx1.x, fmaf(ap[2].y, Mx1.y, fmaf(ap[3].x, Mx1.z, fmaf(ap[3].y, Mx1.w, sx))));
              sy = fmaf(ap[0].x, My0.x, fmaf(ap[0].y, My0.y, fmaf(ap[1].x, My0.z, fmaf(ap[1].y, My0.w, sy))));
              sy = fmaf(ap[2].x, My1.x, fmaf(ap[2].y, My1.y, fmaf(ap[3].x, My1.z, fmaf(ap[3].y, My1.w, sy))));
              sz = fmaf(ap[0].x, Mz0.x, fmaf(ap[0].y, Mz0.y, fmaf(ap[1].x, Mz0.z, fmaf(ap[1].y, Mz0.w, sz))));
              sz = fmaf(ap[2].x, Mz1.x, fmaf(ap[2].y, Mz1.y, fmaf(ap[3].x, Mz1.z, fmaf(ap[3].y, Mz1.w, sz))));
              gx = fmaf(invS, sx, ap[4].x); gy = fmaf(invS, sy, ap[4].y); gz = fmaf(invS, sz, ap[5].x);
            }
            gx *= s; gy *= s; gz *= s;
            const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
            const float dev = gn - 1.0f;
            eik_sum += dev * dev;
            const float kk = gn > 0.f ? p.weight_e * 2.0f * dev * inv_n / gn : 0.f;
            rx = kk * gx; ry = kk * gy; rz = kk * gz;
            // tangent input tau0 = invS (sum_j r_j [M_j; P_j] - dusum z) (+ r on the positional part)
            dusum = rx * P1.z + ry * P1.w + rz * P2.x;
            const float sr = s * invS;
            c[0] = sr * (rx * Mx0.x + ry * My0.x + rz * Mz0.x - dusum * zp[0].x);
            c[1] = sr * (rx * Mx0.y + ry * My0.y + rz * Mz0.y - dusum * zp[0].y);
            c[2] = sr * (rx * Mx0.z + ry * My0.z + rz * Mz0.z - dusum * zp[1].x);
            c[3] = sr * (rx * Mx0.w + ry * My0.w + rz * Mz0.w - dusum * zp[1].y);
            c[4] = sr * (rx * Mx1.x + ry * My1.x + rz * Mz1.x - dusum * zp[2].x);
            c[5] = sr * (rx * Mx1.y + ry * My1.y + rz * Mz1.y - dusum * zp[2].y);
            c[6] = sr * (rx * Mx1.z + ry * My1.z + rz * Mz1.z - dusum * zp[3].x);
            c[7] = sr * (rx * Mx1.w + ry * My1.w + rz * Mz1.w - dusum * zp[3].y);
            c[8] = sr * (rx * P0.x + ry * P0.y + rz * P0.z - dusum * zp[4].x);
            c[9] = sr * (rx * P0.y + ry * P0.w + rz * P1.x - dusum * zp[4].y);
            c[10] = sr * (rx * P0.z + ry * P1.x + rz * P1.y - dusum * zp[5].x);
            if (count > 0) { c[8] = fmaf(s, rx, c[8]); c[9] = fmaf(s, ry, c[9]); c[10] = fmaf(s, rz, c[10]); }
          }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) { c[2 * i] = fmaf(delta, zp[i].x, c[2 * i]); c[2 * i + 1] = fmaf(delta, zp[i].y, c[2 * i + 1]); }
        c[10] = fmaf(delta, zp[5].x, c[10]);
        c[kIn] = delta;
      }
      // ---- row for the decoder-gradient reduction (dead lanes write zeros)
      if (p.fold_rows) {
        float4* dst = reinterpret_cast<float4*>(p.fold_rows + (tile * 32 + lane) * kFoldRow);
        dst[0] = make_float4(c[0], c[1], c[2], c[3]);
        dst[1] = make_float4(c[4], c[5], c[6], c[7]);
        dst[2] = make_float4(c[8], c[9], c[10], c[11]);
        float4 mk = make_float4(0.f, 0.f, 0.f, 0.f);
        mk.x = __uint_as_float(mask[0]);
        if constexpr (kMaskWords > 1) mk.y = __uint_as_float(mask[1]);
        if constexpr (kMaskWords > 2) { mk.z = __uint_as_float(mask[2]); mk.w = __uint_as_float(mask[3]); }
        dst[3] = mk;
      }

      // ---- neural-point feature gradients: dL/df_k = a_f (delta w_k + s e_k), e_k = d w_k/d x . r
      if (live && p.gfeat) {
        float4 rk[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          rk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rid[k] >= 0) rk[k] = __ldg(records + rid[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (rid[k] >= 0) {
            const int row = __float_as_int(rk[k].w);
            const float ex = rk[k].x - px, ey = rk[k].y - py, ez = rk[k].z - pz;
            const float u = 1.0f / (dist2_torch(ex, ey, ez) + kIdwEps);
            const float w = u / S;
            float coef = delta * w;
            if constexpr (!kNumerical) {
              const float du = 2.f * u * u * (ex * rx + ey * ry + ez * rz);  // -2 u^2 (v . r), v = -e
              coef = fmaf(s, (du - w * dusum) * invS, coef);
            }
            float tt[kFeat];
            tt[0] = coef * ap[0].x; tt[1] = coef * ap[0].y; tt[2] = coef * ap[1].x; tt[3] = coef * ap[1].y;
            tt[4] = coef * ap[2].x; tt[5] = coef * ap[2].y; tt[6] = coef * ap[3].x; tt[7] = coef * ap[3].y;
            if (layer_norm) {
              float f[kFeat], mu, rs;
              load_feature_row(m.gather_features, row, f);
              layer_norm8(f, mu, rs);
              layer_norm8_vjp(f, rs, tt);
            }
            red_add_row(p.gfeat, row, tt);
            if (p.touched) p.touched[row] = 1;
          }
        }
      }
      __syncwarp();
    }
  }

  if constexpr (kTrain) {
    bce_sum = warp_sum(bce_sum);
    eik_sum = warp_sum(eik_sum);
    if (lane == 0) { sm_scalar[0][warp] = bce_sum; sm_scalar[1][warp] = eik_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = 0.f, e = 0.f;
      for (int w = 0; w < kTileWarps; ++w) { b += sm_scalar[0][w]; e += sm_scalar[1][w]; }
      const float bce = b / (float)p.n_norm;
      const float eik = p.weight_e > 0.f ? e / (float)(kNumerical ? p.nd_norm : p.n_norm) : 0.f;
      atomicAdd(p.loss + 1, bce);
      atomicAdd(p.loss + 2, eik);
      atomicAdd(p.loss + 0, bce + p.weight_e * eik);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Decoder gradients from the rows written by sdf_tile_kernel:
//   Gd[j][i'] = sum_n d_nj c'_ni',  d_nj = act'(pre_nj) from the activation bits
//   dW0[j][i] = wout_j Gd[j][i],  db0[j] = wout_j Gd[j][11],
//   dwout[j]  = sum_i' [W0 | b0][j][i'] Gd[j][i'],  dbout = sum_n delta_n
// A warp stages 32 rows in shared memory and every lane folds them into the hidden rows it owns
// (lane, lane + 32, ...); one block reduction and 12 H + H + 1 atomics per CTA at the end.
// ------------------------------------------------------------------------------------------
struct DecoderGradParams {
  ClidDecoder dec;
  const float* rows;  // [n_rows][16]
  float* dec_grad;    // flat [W0 (H x 11), b0 (H), wout (H), bout (1)] +=
  int64_t n_rows;     // multiple of 32
  uint32_t flags;
};

// one fat CTA per SM: ~one 32-row tile per warp at 131072 rows (H = 128: half the warps, its
// per-warp partial Gd is twice as large)
template <int H>
struct DgSmem {
  static constexpr int kWarps = H > 64 ? 16 : 32;
  static constexpr int kRows = kWarps * 32 * kFoldRow;      // staged rows
  static constexpr int kGd = kWarps * H * kInPad;           // per-warp partial Gd
  static constexpr size_t kBytes = (size_t)(kRows + kGd + kWarps) * sizeof(float);
};

template <int H>
__global__ void __launch_bounds__(DgSmem<H>::kWarps * 32, 1) decoder_grad_kernel(const __grid_constant__ DecoderGradParams p) {
  constexpr int kR = H / 32;
  constexpr int kDgWarps = DgSmem<H>::kWarps;
  extern __shared__ __align__(16) float dg_smem[];
  float* sm_rows = dg_smem;                          // [warps][32][16]
  float* sm_gd = dg_smem + DgSmem<H>::kRows;         // [warps][H][12]
  float* sm_delta = sm_gd + DgSmem<H>::kGd;          // [warps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  float2 Gd[kR][6];
#pragma unroll
  for (int r = 0; r < kR; ++r)
#pragma unroll
    for (int i = 0; i < 6; ++i) Gd[r][i] = make_float2(0.f, 0.f);
  float dsum = 0.f;
  float* my_rows = sm_rows + warp * 32 * kFoldRow;
  const int64_t n_tiles = p.n_rows >> 5;
  for (int64_t t = (int64_t)blockIdx.x * kDgWarps + warp; t < n_tiles; t += (int64_t)gridDim.x * kDgWarps) {
    const float4* src = reinterpret_cast<const float4*>(p.rows + (t * 32 + lane) * kFoldRow);
    const float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
    float4* dst = reinterpret_cast<float4*>(my_rows + lane * kFoldRow);
    dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3;
    dsum += v2.w;
    __syncwarp();
#pragma unroll 4
    for (int nn = 0; nn < 32; ++nn) {
      const float4* rr = reinterpret_cast<const float4*>(my_rows + nn * kFoldRow);
      const float4 c0 = rr[0], c1 = rr[1], c2 = rr[2];
      const uint4 mk = *reinterpret_cast<const uint4*>(rr + 3);
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const uint32_t word = r == 0 ? mk.x : (r == 1 ? mk.y : (r == 2 ? mk.z : mk.w));
        const float d = ((word >> lane) & 1u) ? 1.f : slope;
        const float2 dd = make_float2(d, d);
        Gd[r][0] = __ffma2_rn(dd, make_float2(c0.x, c0.y), Gd[r][0]);
        Gd[r][1] = __ffma2_rn(dd, make_float2(c0.z, c0.w), Gd[r][1]);
        Gd[r][2] = __ffma2_rn(dd, make_float2(c1.x, c1.y), Gd[r][2]);
        Gd[r][3] = __ffma2_rn(dd, make_float2(c1.z, c1.w), Gd[r][3]);
        Gd[r][4] = __ffma2_rn(dd, make_float2(c2.x, c2.y), Gd[r][4]);
        Gd[r][5] = __ffma2_rn(dd, make_float2(c2.z, c2.w), Gd[r][5]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    float4* dst = reinterpret_cast<float4*>(sm_gd + (warp * H + lane + 32 * r) * kInPad);
    dst[0] = make_float4(Gd[r][0].x, Gd[r][0].y, Gd[r][1].x, Gd[r][1].y);
    dst[1] = make_float4(Gd[r][2].x, Gd[r][2].y, Gd[r][3].x, Gd[r][3].y);
    dst[2] = make_float4(Gd[r][4].x, Gd[r][4].y, Gd[r][5].x, Gd[r][5].y);
  }
  dsum = warp_sum(dsum);
  if (lane == 0) sm_delta[warp] = dsum;
  __syncthreads();
  // element (j, i) of the CTA's Gd: one thread each sums the kDgWarps partials, then forms its outputs
  for (int e = threadIdx.x; e < H * kInPad; e += blockDim.x) {
    float v = 0.f;
#pragma unroll 8
    for (int w = 0; w < kDgWarps; ++w) v += sm_gd[w * H * kInPad + e];
    sm_gd[e] = v;  // only this thread reads and writes column e of the partials
  }
  __syncthreads();
  // outputs in the flat dec_grad order, staged in shared memory so the atomics below are coalesced
  // (32 consecutive floats per instruction = 4 sectors, instead of 32 sectors with the row stride)
  float* sm_out = sm_rows;  // the staged rows are dead
  for (int e = threadIdx.x; e < H * kIn; e += blockDim.x) {
    const int j = e / kIn, i = e - j * kIn;
    sm_out[e] = __ldg(p.dec.out_weight + j) * sm_gd[j * kInPad + i];
  }
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const float* g = sm_gd + j * kInPad;
    float dw = p.dec.bias[0] ? __ldg(p.dec.bias[0] + j) * g[kIn] : 0.f;
#pragma unroll
    for (int i = 0; i < kIn; ++i) dw = fmaf(__ldg(p.dec.weight[0] + j * kIn + i), g[i], dw);
    sm_out[H * kIn + j] = p.dec.bias[0] ? __ldg(p.dec.out_weight + j) * g[kIn] : 0.f;
    sm_out[H * kIn + H + j] = dw;
  }
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int w = 0; w < kDgWarps; ++w) d += sm_delta[w];
    sm_out[H * kIn + 2 * H] = p.dec.out_bias ? d : 0.f;
  }
  __syncthreads();
  constexpr int kOut = H * kIn + 2 * H + 1;
  for (int e = threadIdx.x; e < kOut; e += blockDim.x) atomicAdd(p.dec_grad + e, sm_out[e]);
}

}  // namespace clid
