// Shared device helpers for libclid_sdf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "clid_sdf.h"

namespace clid {

constexpr int kFeat = 8;          // feature_dim (utils/config.py:127)
constexpr int kIn = kFeat + 3;    // decoder input: features + relative position
constexpr int kInPad = 12;        // row stride of first-layer weights in shared memory
constexpr float kIdwEps = 1e-15f; // neural_points.py:688
constexpr float kLnEps = 1e-5f;   // F.layer_norm default
constexpr float kLeakySlope = 0.01f;

int set_error(int code, const char* fmt, ...);

__device__ __forceinline__ int64_t floor_mod(int64_t a, int64_t b) {
  int64_t r = a % b;
  return r < 0 ? r + b : r;
}

// floor(x / res) exactly as torch does on fp32: IEEE division, then floor.
__device__ __forceinline__ int cell_of(float x, float res) {
  return (int)floorf(__fdiv_rn(x, res));
}

// (dx^2 + dy^2) + dz^2 with no FMA contraction: the rounding of torch's (v**2).sum(-1).
__device__ __forceinline__ float dist2_torch(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- decoder weights staged in shared memory --------------------------------------
// layout (floats): W0[H][kInPad] | b0[H] | {Wl[H][H] | bl[H]} (levels-1) | wout[H] | bout | pad
template <int H, int L>
struct MlpLayout {
  static constexpr int kW0 = 0;
  static constexpr int kB0 = kW0 + H * kInPad;
  static constexpr int kHidden = kB0 + H;  // start of level-1.. blocks
  static constexpr int kHiddenStride = H * H + H;
  static constexpr int kWout = kHidden + (L - 1) * kHiddenStride;
  static constexpr int kBout = kWout + H;
  static constexpr int kFloats = ((kBout + 1 + 3) / 4) * 4;
};

template <int H, int L>
__device__ __forceinline__ void stage_decoder(float* sm, const ClidDecoder& dec) {
  using Lay = MlpLayout<H, L>;
  for (int i = threadIdx.x; i < H * kInPad; i += blockDim.x) {
    int j = i / kInPad, c = i - j * kInPad;
    sm[Lay::kW0 + i] = c < kIn ? dec.weight[0][j * kIn + c] : 0.f;
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    sm[Lay::kB0 + i] = dec.bias[0] ? dec.bias[0][i] : 0.f;
    sm[Lay::kWout + i] = dec.out_weight[i];
  }
#pragma unroll
  for (int l = 1; l < L; ++l) {
    float* blk = sm + Lay::kHidden + (l - 1) * Lay::kHiddenStride;
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) blk[i] = dec.weight[l][i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) blk[H * H + i] = dec.bias[l] ? dec.bias[l][i] : 0.f;
  }
  if (threadIdx.x == 0) sm[Lay::kBout] = dec.out_bias ? dec.out_bias[0] : 0.f;
}

// Forward of the MLP plus a = d out / d z (back-propagated through the activation masks).
// out is the un-scaled logit (Decoder.mlp); sdf = sdf_scale * out.
template <int H, int L>
__device__ __forceinline__ void mlp_value_and_input_grad(const float* __restrict__ sm, const float (&z)[kIn],
                                                         float slope, float& out, float (&a)[kIn]) {
  using Lay = MlpLayout<H, L>;
  const float4* w0 = reinterpret_cast<const float4*>(sm + Lay::kW0);
  out = sm[Lay::kBout];
#pragma unroll
  for (int i = 0; i < kIn; ++i) a[i] = 0.f;
  if constexpr (L == 1) {
#pragma unroll 8
    for (int j = 0; j < H; ++j) {
      float4 r0 = w0[j * 3 + 0], r1 = w0[j * 3 + 1], r2 = w0[j * 3 + 2];
      float pre = sm[Lay::kB0 + j];
      pre = fmaf(r0.x, z[0], pre); pre = fmaf(r0.y, z[1], pre); pre = fmaf(r0.z, z[2], pre); pre = fmaf(r0.w, z[3], pre);
      pre = fmaf(r1.x, z[4], pre); pre = fmaf(r1.y, z[5], pre); pre = fmaf(r1.z, z[6], pre); pre = fmaf(r1.w, z[7], pre);
      pre = fmaf(r2.x, z[8], pre); pre = fmaf(r2.y, z[9], pre); pre = fmaf(r2.z, z[10], pre);
      float d = pre > 0.f ? 1.f : slope;
      float c = sm[Lay::kWout + j] * d;
      out = fmaf(c, pre, out);
      a[0] = fmaf(c, r0.x, a[0]); a[1] = fmaf(c, r0.y, a[1]); a[2] = fmaf(c, r0.z, a[2]); a[3] = fmaf(c, r0.w, a[3]);
      a[4] = fmaf(c, r1.x, a[4]); a[5] = fmaf(c, r1.y, a[5]); a[6] = fmaf(c, r1.z, a[6]); a[7] = fmaf(c, r1.w, a[7]);
      a[8] = fmaf(c, r2.x, a[8]); a[9] = fmaf(c, r2.y, a[9]); a[10] = fmaf(c, r2.z, a[10]);
    }
  } else {
    // level 0
    float h[H], dact[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float4 r0 = w0[j * 3 + 0], r1 = w0[j * 3 + 1], r2 = w0[j * 3 + 2];
      float pre = sm[Lay::kB0 + j];
      pre = fmaf(r0.x, z[0], pre); pre = fmaf(r0.y, z[1], pre); pre = fmaf(r0.z, z[2], pre); pre = fmaf(r0.w, z[3], pre);
      pre = fmaf(r1.x, z[4], pre); pre = fmaf(r1.y, z[5], pre); pre = fmaf(r1.z, z[6], pre); pre = fmaf(r1.w, z[7], pre);
      pre = fmaf(r2.x, z[8], pre); pre = fmaf(r2.y, z[9], pre); pre = fmaf(r2.z, z[10], pre);
      dact[j] = pre > 0.f ? 1.f : slope;
      h[j] = pre * dact[j];
    }
    // levels 1..L-1: forward, remembering masks; then pull d out / d h back level by level.
    // L is at most 3; for L == 2 this is a single hidden->hidden layer.
    float t[H];  // d out / d h_prev accumulated
    if constexpr (L == 2) {
      const float* w1 = sm + Lay::kHidden;
      const float* b1 = w1 + H * H;
#pragma unroll
      for (int i = 0; i < H; ++i) t[i] = 0.f;
#pragma unroll 4
      for (int j = 0; j < H; ++j) {
        float pre = b1[j];
        const float4* row = reinterpret_cast<const float4*>(w1 + j * H);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
          float4 r = row[q];
          pre = fmaf(r.x, h[4 * q + 0], pre); pre = fmaf(r.y, h[4 * q + 1], pre);
          pre = fmaf(r.z, h[4 * q + 2], pre); pre = fmaf(r.w, h[4 * q + 3], pre);
        }
        float d = pre > 0.f ? 1.f : slope;
        float c = sm[Lay::kWout + j] * d;
        out = fmaf(c, pre, out);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
          float4 r = row[q];
          t[4 * q + 0] = fmaf(c, r.x, t[4 * q + 0]); t[4 * q + 1] = fmaf(c, r.y, t[4 * q + 1]);
          t[4 * q + 2] = fmaf(c, r.z, t[4 * q + 2]); t[4 * q + 3] = fmaf(c, r.w, t[4 * q + 3]);
        }
      }
    } else {
      static_assert(L == 2, "only 1 or 2 hidden levels are compiled into the fused kernels");
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float c = t[j] * dact[j];
      float4 r0 = w0[j * 3 + 0], r1 = w0[j * 3 + 1], r2 = w0[j * 3 + 2];
      a[0] = fmaf(c, r0.x, a[0]); a[1] = fmaf(c, r0.y, a[1]); a[2] = fmaf(c, r0.z, a[2]); a[3] = fmaf(c, r0.w, a[3]);
      a[4] = fmaf(c, r1.x, a[4]); a[5] = fmaf(c, r1.y, a[5]); a[6] = fmaf(c, r1.z, a[6]); a[7] = fmaf(c, r1.w, a[7]);
      a[8] = fmaf(c, r2.x, a[8]); a[9] = fmaf(c, r2.y, a[9]); a[10] = fmaf(c, r2.z, a[10]);
    }
  }
}

__device__ __forceinline__ void load_feature_row(const float* __restrict__ feats, int id, float (&f)[kFeat]) {
  const float4* row = reinterpret_cast<const float4*>(feats + (int64_t)id * kFeat);
  float4 a = __ldg(row), b = __ldg(row + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// LayerNorm over the 8 feature channels, no affine (F.layer_norm(x, [8])).
__device__ __forceinline__ void layer_norm8(float (&f)[kFeat], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kFeat; ++i) s += f[i];
  mean = s * (1.f / kFeat);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kFeat; ++i) {
    float d = f[i] - mean;
    v = fmaf(d, d, v);
  }
  rstd = 1.0f / sqrtf(v * (1.f / kFeat) + kLnEps);
#pragma unroll
  for (int i = 0; i < kFeat; ++i) f[i] = (f[i] - mean) * rstd;
}

}  // namespace clid
