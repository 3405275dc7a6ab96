// Dense reduction of the per-point decoder-gradient rows written by train_fused_l1_kernel
// (ClidTrainFusedArgs.scratch) into the flat decoder gradient: the decoder part of
// cur_loss.backward() (utils/mapper.py:834-835 of the reference).
#pragma once
#include "common.cuh"
#include "train.cuh"

namespace clid {

constexpr int kFoldRow = 16;  // floats per row: c'[12], activation-mask words[4]

// ------------------------------------------------------------------------------------------
// Decoder gradients from the rows written by train_fused_l1_kernel:
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
