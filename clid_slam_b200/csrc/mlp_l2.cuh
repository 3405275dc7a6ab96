// Two-hidden-level decoder inside the fused training kernel (model/decoder.py:40-82 with hidden_level = 2;
// BASELINE configs[0]: 2 layers of 32).  Closed-form backward of SURVEY.md 8a-G2: with the logit o, delta = dL/do,
// the tangent input tau0 = s J r of the analytic eikonal term, activations h_l, derivative masks d_l,
//   beta2 = d2 (.) wout,  beta1 = d1 (.) (W1^T beta2),  a = do/dz = W0^T beta1
//   tau1 = d1 (.) (W0 tau0),  tau2 = d2 (.) (W1 tau1)
//   dW0 = sum beta1 (x) (delta z + tau0)   db0 = sum delta beta1
//   dW1 = sum beta2 (x) (delta h1 + tau1)  db1 = sum delta beta2
//   dwout = sum (delta h2 + tau2)          dbout = sum delta
// A sample cannot hold three H-vectors in registers next to the search / blend state, and the [H, H] gradient is a
// real [H, N] x [N, H] contraction over the batch: every sample writes ONE row
//   [ c0 = delta z + tau0 ; delta (12) | beta1 (H) | c1 = delta h1 + tau1 (H) | e2 = delta h2 + tau2 (H) | d2 bits ]
// to scratch (h1 / h2 are parked there between the forward and the finish), and decoder_grad_l2_kernel contracts
// the rows over the batch.  H = 32 (one mask word, one hidden unit per lane in the reduction).
#pragma once
#include "common.cuh"

namespace clid {

template <int H>
struct L2Row {
  static_assert(H == 32, "the fused two-level trainer is compiled for 32 hidden units");
  static constexpr int kC0 = 0, kB1 = kInPad, kC1 = kInPad + H, kE2 = kInPad + 2 * H, kBits = kInPad + 3 * H;
  static constexpr int kFloats = ((kBits + 1 + 3) / 4) * 4;  // 112 for H = 32
};

// forward: out (un-scaled logit), a = d out / d z, activation masks; parks h1 (at kC1), beta1, h2 (at kE2) in the row
template <int H>
__device__ __forceinline__ void mlp_l2_forward_train(const float* __restrict__ sm, const float (&z)[kIn], float slope, float& out,
                                                     float (&a)[kIn], uint32_t& m1, uint32_t& m2, float* __restrict__ row) {
  using Lay = MlpLayout<H, 2>;
  using R = L2Row<H>;
  const float4* w0 = reinterpret_cast<const float4*>(sm + Lay::kW0);  // row-major [H][12]
  const float* w1 = sm + Lay::kHidden;
  const float* b1 = w1 + H * H;
  float h1[H];
  m1 = 0u;
#pragma unroll
  for (int j = 0; j < H; ++j) {
    const float4 r0 = w0[j * 3 + 0], r1 = w0[j * 3 + 1], r2 = w0[j * 3 + 2];
    float pre = sm[Lay::kB0 + j];
    pre = fmaf(r0.x, z[0], pre); pre = fmaf(r0.y, z[1], pre); pre = fmaf(r0.z, z[2], pre); pre = fmaf(r0.w, z[3], pre);
    pre = fmaf(r1.x, z[4], pre); pre = fmaf(r1.y, z[5], pre); pre = fmaf(r1.z, z[6], pre); pre = fmaf(r1.w, z[7], pre);
    pre = fmaf(r2.x, z[8], pre); pre = fmaf(r2.y, z[9], pre); pre = fmaf(r2.z, z[10], pre);
    const bool on = pre > 0.f;
    m1 |= on ? (1u << j) : 0u;
    h1[j] = pre * (on ? 1.f : slope);
  }
#pragma unroll
  for (int q = 0; q < H / 4; ++q)
    *reinterpret_cast<float4*>(row + R::kC1 + 4 * q) = make_float4(h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
  float t[H];
#pragma unroll
  for (int k = 0; k < H; ++k) t[k] = 0.f;
  m2 = 0u;
  out = sm[Lay::kBout];
#pragma unroll 4
  for (int j = 0; j < H; ++j) {
    const float4* wr = reinterpret_cast<const float4*>(w1 + j * H);
    float pre = b1[j];
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 r = wr[q];
      pre = fmaf(r.x, h1[4 * q], pre); pre = fmaf(r.y, h1[4 * q + 1], pre);
      pre = fmaf(r.z, h1[4 * q + 2], pre); pre = fmaf(r.w, h1[4 * q + 3], pre);
    }
    const bool on = pre > 0.f;
    m2 |= on ? (1u << j) : 0u;
    const float d = on ? 1.f : slope;
    row[R::kE2 + j] = pre * d;  // h2_j
    const float c = sm[Lay::kWout + j] * d;  // beta2_j
    out = fmaf(c, pre, out);
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 r = wr[q];
      t[4 * q] = fmaf(c, r.x, t[4 * q]); t[4 * q + 1] = fmaf(c, r.y, t[4 * q + 1]);
      t[4 * q + 2] = fmaf(c, r.z, t[4 * q + 2]); t[4 * q + 3] = fmaf(c, r.w, t[4 * q + 3]);
    }
  }
#pragma unroll
  for (int i = 0; i < kIn; ++i) a[i] = 0.f;
#pragma unroll
  for (int k = 0; k < H; ++k) {
    const float bk = t[k] * (((m1 >> k) & 1u) ? 1.f : slope);  // beta1_k
    t[k] = bk;
    const float4 r0 = w0[k * 3 + 0], r1 = w0[k * 3 + 1], r2 = w0[k * 3 + 2];
    a[0] = fmaf(bk, r0.x, a[0]); a[1] = fmaf(bk, r0.y, a[1]); a[2] = fmaf(bk, r0.z, a[2]); a[3] = fmaf(bk, r0.w, a[3]);
    a[4] = fmaf(bk, r1.x, a[4]); a[5] = fmaf(bk, r1.y, a[5]); a[6] = fmaf(bk, r1.z, a[6]); a[7] = fmaf(bk, r1.w, a[7]);
    a[8] = fmaf(bk, r2.x, a[8]); a[9] = fmaf(bk, r2.y, a[9]); a[10] = fmaf(bk, r2.z, a[10]);
  }
#pragma unroll
  for (int q = 0; q < H / 4; ++q)
    *reinterpret_cast<float4*>(row + R::kB1 + 4 * q) = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
}

// finish: the row's c0, c1 = delta h1 + tau1, e2 = delta h2 + tau2 and the mask word.  tau0 = s J r (zero in
// numerical mode and for samples without an eikonal term: kTangent false skips the two tangent products)
template <int H, bool kTangent>
__device__ __forceinline__ void mlp_l2_finish_row(const float* __restrict__ sm, const float (&c0)[kInPad], const float (&tau0)[kIn],
                                                  float delta, float slope, uint32_t m1, uint32_t m2, float* __restrict__ row) {
  using Lay = MlpLayout<H, 2>;
  using R = L2Row<H>;
  const float4* w0 = reinterpret_cast<const float4*>(sm + Lay::kW0);
  const float* w1 = sm + Lay::kHidden;
  float tau1[H];
#pragma unroll
  for (int k = 0; k < H; ++k) {
    float v = 0.f;
    if constexpr (kTangent) {
      const float4 r0 = w0[k * 3 + 0], r1 = w0[k * 3 + 1], r2 = w0[k * 3 + 2];
      v = fmaf(r0.x, tau0[0], v); v = fmaf(r0.y, tau0[1], v); v = fmaf(r0.z, tau0[2], v); v = fmaf(r0.w, tau0[3], v);
      v = fmaf(r1.x, tau0[4], v); v = fmaf(r1.y, tau0[5], v); v = fmaf(r1.z, tau0[6], v); v = fmaf(r1.w, tau0[7], v);
      v = fmaf(r2.x, tau0[8], v); v = fmaf(r2.y, tau0[9], v); v = fmaf(r2.z, tau0[10], v);
      v *= ((m1 >> k) & 1u) ? 1.f : slope;
    }
    tau1[k] = v;
  }
#pragma unroll
  for (int q = 0; q < H / 4; ++q) {
    float4 h = *reinterpret_cast<float4*>(row + R::kC1 + 4 * q);
    h.x = fmaf(delta, h.x, tau1[4 * q]); h.y = fmaf(delta, h.y, tau1[4 * q + 1]);
    h.z = fmaf(delta, h.z, tau1[4 * q + 2]); h.w = fmaf(delta, h.w, tau1[4 * q + 3]);
    *reinterpret_cast<float4*>(row + R::kC1 + 4 * q) = h;
  }
#pragma unroll 4
  for (int j = 0; j < H; ++j) {
    float v = 0.f;
    if constexpr (kTangent) {
      const float4* wr = reinterpret_cast<const float4*>(w1 + j * H);
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        const float4 r = wr[q];
        v = fmaf(r.x, tau1[4 * q], v); v = fmaf(r.y, tau1[4 * q + 1], v);
        v = fmaf(r.z, tau1[4 * q + 2], v); v = fmaf(r.w, tau1[4 * q + 3], v);
      }
      v *= ((m2 >> j) & 1u) ? 1.f : slope;
    }
    row[R::kE2 + j] = fmaf(delta, row[R::kE2 + j], v);
  }
  *reinterpret_cast<float4*>(row + R::kC0) = make_float4(c0[0], c0[1], c0[2], c0[3]);
  *reinterpret_cast<float4*>(row + R::kC0 + 4) = make_float4(c0[4], c0[5], c0[6], c0[7]);
  *reinterpret_cast<float4*>(row + R::kC0 + 8) = make_float4(c0[8], c0[9], c0[10], c0[11]);
  row[R::kBits] = __uint_as_float(m2);
}

template <int H>
__device__ __forceinline__ void l2_zero_row(float* __restrict__ row) {
#pragma unroll
  for (int q = 0; q < L2Row<H>::kFloats / 4; ++q) *reinterpret_cast<float4*>(row + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------
// contraction of the rows over the batch: lane j of a warp owns hidden unit j
//   dW0[j][i] = sum_n beta1_nj c0_ni   db0[j] = sum_n beta1_nj delta_n
//   dW1[j][k] = wout_j sum_n d2_nj c1_nk   db1[j] = wout_j sum_n d2_nj delta_n
//   dwout[j] = sum_n e2_nj   dbout = sum_n delta_n
// flat dec_grad order: [W0 (H x 11), b0 (H), W1 (H x H), b1 (H), wout (H), bout (1)]
// ------------------------------------------------------------------------------------------
struct DecoderGradL2Params {
  ClidDecoder dec;
  const float* rows;
  float* dec_grad;
  int64_t n_rows;
  uint32_t flags;
};

template <int H>
__global__ void __launch_bounds__(256, 1) decoder_grad_l2_kernel(const __grid_constant__ DecoderGradL2Params p) {
  using R = L2Row<H>;
  constexpr int kOut = H * kIn + H + H * H + H + H + 1;
  __shared__ float sm_out[kOut];
  for (int e = threadIdx.x; e < kOut; e += blockDim.x) sm_out[e] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  float g0[kInPad], g1[H], gw = 0.f, gb1 = 0.f, gbout = 0.f;
#pragma unroll
  for (int i = 0; i < kInPad; ++i) g0[i] = 0.f;
#pragma unroll
  for (int k = 0; k < H; ++k) g1[k] = 0.f;
  for (int64_t n = (int64_t)blockIdx.x * warps + warp; n < p.n_rows; n += (int64_t)gridDim.x * warps) {
    const float* row = p.rows + n * R::kFloats;
    const float b1 = __ldg(row + R::kB1 + lane);
    const float e2 = __ldg(row + R::kE2 + lane);
    const uint32_t bits = __float_as_uint(__ldg(row + R::kBits));
    const float d2 = ((bits >> lane) & 1u) ? 1.f : slope;
    const float4 c00 = __ldg(reinterpret_cast<const float4*>(row)), c01 = __ldg(reinterpret_cast<const float4*>(row) + 1),
                 c02 = __ldg(reinterpret_cast<const float4*>(row) + 2);
    g0[0] = fmaf(b1, c00.x, g0[0]); g0[1] = fmaf(b1, c00.y, g0[1]); g0[2] = fmaf(b1, c00.z, g0[2]); g0[3] = fmaf(b1, c00.w, g0[3]);
    g0[4] = fmaf(b1, c01.x, g0[4]); g0[5] = fmaf(b1, c01.y, g0[5]); g0[6] = fmaf(b1, c01.z, g0[6]); g0[7] = fmaf(b1, c01.w, g0[7]);
    g0[8] = fmaf(b1, c02.x, g0[8]); g0[9] = fmaf(b1, c02.y, g0[9]); g0[10] = fmaf(b1, c02.z, g0[10]); g0[11] = fmaf(b1, c02.w, g0[11]);
    const float delta = c02.w;
    gw += e2;
    gb1 = fmaf(d2, delta, gb1);
    gbout += delta;
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(row + R::kC1) + q);
      g1[4 * q] = fmaf(d2, c.x, g1[4 * q]); g1[4 * q + 1] = fmaf(d2, c.y, g1[4 * q + 1]);
      g1[4 * q + 2] = fmaf(d2, c.z, g1[4 * q + 2]); g1[4 * q + 3] = fmaf(d2, c.w, g1[4 * q + 3]);
    }
  }
  // fold the warps of the CTA in shared memory, then one atomic per output and CTA
  const float wout = __ldg(p.dec.out_weight + lane);
  float* o_w0 = sm_out;
  float* o_b0 = o_w0 + H * kIn;
  float* o_w1 = o_b0 + H;
  float* o_b1 = o_w1 + H * H;
  float* o_wo = o_b1 + H;
  float* o_bo = o_wo + H;
#pragma unroll
  for (int i = 0; i < kIn; ++i) atomicAdd(o_w0 + lane * kIn + i, g0[i]);
  atomicAdd(o_b0 + lane, g0[kIn]);
#pragma unroll
  for (int k = 0; k < H; ++k) atomicAdd(o_w1 + lane * H + k, wout * g1[k]);
  atomicAdd(o_b1 + lane, wout * gb1);
  atomicAdd(o_wo + lane, gw);
  if (lane == 0) atomicAdd(o_bo, gbout);
  __syncthreads();
  for (int e = threadIdx.x; e < kOut; e += blockDim.x) {
    const bool is_b0 = e >= H * kIn && e < H * kIn + H;
    const bool is_b1 = e >= H * kIn + H + H * H && e < H * kIn + H + H * H + H;
    const bool is_bo = e == kOut - 1;
    if ((is_b0 && !p.dec.bias[0]) || (is_b1 && !p.dec.bias[1]) || (is_bo && !p.dec.out_bias)) continue;
    atomicAdd(p.dec_grad + e, sm_out[e]);
  }
}

}  // namespace clid
