// Training-side kernels of one mapping iteration (utils/mapper.py:642-836 of the reference):
//   sdf_loss_kernel      sdf_bce_loss (utils/loss.py:44-62) + eikonal term (mapper.py:780-798,
//                        with get_numerical_gradient's central differences, mapper.py:985-1034)
//                        -> per-sample d L / d logit, d L / d grad, and the three loss scalars
//   train_backward_kernel closed-form backward: decoder gradients reduced in
//                        the block, neural-point feature gradients scattered with vector atomics
//   adam_kernel          torch.optim.Adam(betas .9/.99, eps adam_eps) step (utils/tools.py:205-255)
//                        on the touched feature rows and the decoder tensors
#pragma once
#include "common.cuh"
#include "query_bwd.cuh"

namespace clid {

// ------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------
struct LossParams {
  const float* sdf;      // [n + 6 nd]  predictions: batch, then x+ex, x-ex, x+ey, x-ey, x+ez, x-ez blocks of nd
  const float* grad;     // [n,3] analytic d sdf / d x or NULL
  const float* label;    // [n]
  const float* weight;   // [n] signed sample weight (|.| is used) or NULL
  float* dlogit;         // [n + 6 nd]  d L / d (mlp output)
  float* dgrad;          // [n,3] d L / d grad (analytic mode) or NULL
  float* loss;           // [3] += total, bce, eikonal (un-weighted eikonal mean, like the reference logs)
  int64_t n;
  int64_t nd;            // decimated count for numerical mode, 0 otherwise
  int64_t n_norm;        // mean denominators: the GLOBAL batch / decimated sizes when the batch is
  int64_t nd_norm;       // sharded over ranks (== n / nd on one GPU)
  float sdf_scale;       // sigma of the BCE (= decoder sdf_scale)
  float weight_e;        // eikonal weight, 0 disables
  float num_eps;         // central-difference step
  int weighted;          // loss_weight_on
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#ifdef CLID_PLAIN_KERNELS  // defined once, in api.cu
__global__ void __launch_bounds__(256) sdf_loss_kernel(const LossParams p) {
  __shared__ float red[2][8];
  float bce_sum = 0.f, eik_sum = 0.f;
  const float inv_n = 1.0f / (float)p.n_norm;
  const float s = p.sdf_scale;
  const bool analytic = p.grad != nullptr && p.weight_e > 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    // BCEWithLogits(pred / sigma, sigmoid(label / sigma), weight=|w|), mean reduction
    const float l = p.sdf[i] / s;
    const float t = 1.0f / (1.0f + expf(-(p.label[i] / s)));
    const float wgt = (p.weighted && p.weight) ? fabsf(p.weight[i]) : 1.0f;
    const float softplus_neg = fmaxf(-l, 0.f) + log1pf(expf(-fabsf(l)));  // log(1 + exp(-l))
    bce_sum += wgt * ((1.0f - t) * l + softplus_neg);
    const float sig = 1.0f / (1.0f + expf(-l));
    p.dlogit[i] = wgt * (sig - t) * inv_n;
    if (analytic) {
      const float gx = p.grad[3 * i], gy = p.grad[3 * i + 1], gz = p.grad[3 * i + 2];
      const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
      const float dev = gn - 1.0f;
      eik_sum += dev * dev;
      const float k = gn > 0.f ? p.weight_e * 2.0f * dev * inv_n / gn : 0.f;
      p.dgrad[3 * i] = k * gx; p.dgrad[3 * i + 1] = k * gy; p.dgrad[3 * i + 2] = k * gz;
    }
  }
  if (p.nd > 0 && p.weight_e > 0.f) {
    const float inv_nd = 1.0f / (float)p.nd_norm;
    const float two_eps = 2.0f * p.num_eps;
    const float* sh = p.sdf + p.n;
    float* dsh = p.dlogit + p.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nd; i += (int64_t)gridDim.x * blockDim.x) {
      const float gx = (sh[i] - sh[p.nd + i]) / two_eps;
      const float gy = (sh[2 * p.nd + i] - sh[3 * p.nd + i]) / two_eps;
      const float gz = (sh[4 * p.nd + i] - sh[5 * p.nd + i]) / two_eps;
      const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
      const float dev = gn - 1.0f;
      eik_sum += dev * dev;
      // d L / d sdf(x +- eps e_a) = +- r_a / (2 eps);  d L / d logit = s * that
      const float k = gn > 0.f ? p.weight_e * 2.0f * dev * inv_nd / gn * s / two_eps : 0.f;
      dsh[i] = k * gx; dsh[p.nd + i] = -k * gx;
      dsh[2 * p.nd + i] = k * gy; dsh[3 * p.nd + i] = -k * gy;
      dsh[4 * p.nd + i] = k * gz; dsh[5 * p.nd + i] = -k * gz;
    }
  }
  bce_sum = warp_sum(bce_sum);
  eik_sum = warp_sum(eik_sum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = bce_sum; red[1][warp] = eik_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f, e = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { b += red[0][w]; e += red[1][w]; }
    const int64_t n_e = analytic ? p.n_norm : p.nd_norm;
    const float bce = b * inv_n;
    const float eik = n_e > 0 ? e / (float)n_e : 0.f;
    atomicAdd(p.loss + 1, bce);
    atomicAdd(p.loss + 2, eik);
    atomicAdd(p.loss + 0, bce + p.weight_e * eik);
  }
}
#endif  // CLID_PLAIN_KERNELS

// ------------------------------------------------------------------------------------------
// backward (one hidden level): thread per sample, decoder gradients through the masked sums
//   Gd[j][i'] = sum_n d_nj c'_ni',  c' = [delta z + tau0 ; delta],  d_nj = act'(pre_nj)
//   dW0[j][i] = wout_j Gd[j][i],  db0[j] = wout_j Gd[j][11],
//   dwout[j]  = sum_i' [W0 | b0][j][i'] Gd[j][i'],  dbout = sum_n delta_n
// ------------------------------------------------------------------------------------------
struct TrainBwdParams {
  ClidMap map;
  ClidDecoder dec;
  const float* x;          // [n,3]
  const int32_t* knn_idx;  // [n,knn] from the forward
  const float* dlogit;     // [n]
  const float* dgrad;      // [n_r,3] or NULL: samples >= n_r have no gradient term
  float* gfeat;            // [n_gather+1,8] += (caller zero-fills once; Adam re-zeroes touched rows)
  uint8_t* touched;        // [n_gather+1] set to 1 for rows that received a gradient, or NULL
  float* dec_grad;         // flat [W0 (H x 11), b0 (H), wout (H), bout (1)] +=, or NULL (frozen decoder)
  int64_t n;
  int64_t n_r;
  uint32_t flags;
};

constexpr int kBwdThreads = 128;

template <int H, int K>
__global__ void __launch_bounds__(kBwdThreads) train_backward_l1_kernel(const __grid_constant__ TrainBwdParams p) {
  using Lay = MlpLayout<H, 1>;
  constexpr int kRows = H / 32;       // hidden rows owned by a lane
  constexpr int kMaskWords = H / 32;
  constexpr int kWarps = kBwdThreads / 32;
  extern __shared__ __align__(16) float smem[];
  float* sm_dec = smem;
  float* sm_c = smem + Lay::kFloats;                                   // [warps][32][12]
  uint32_t* sm_m = reinterpret_cast<uint32_t*>(sm_c + kWarps * 32 * kInPad);  // [warps][32][kMaskWords]
  float* sm_red = reinterpret_cast<float*>(sm_m + kWarps * 32 * kMaskWords);   // [warps][H][12] epilogue

  const ClidMap& m = p.map;
  stage_decoder<H, 1>(sm_dec, p.dec);
  __syncthreads();

  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const float s = p.dec.sdf_scale;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my_c = sm_c + (warp * 32) * kInPad;
  uint32_t* my_m = sm_m + (warp * 32) * kMaskWords;
  const float4* w0 = reinterpret_cast<const float4*>(sm_dec + Lay::kW0);

  float Gd[kRows][kInPad];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int i = 0; i < kInPad; ++i) Gd[r][i] = 0.f;
  float delta_sum = 0.f;

  TileScheduler sched(p.map.work_counter, p.n);
  for (int64_t tile = sched.next(); tile >= 0; tile = sched.next()) {
    const int64_t q = tile * 32 + (threadIdx.x & 31);
    const bool live = q < p.n;
    float c[kInPad];
#pragma unroll
    for (int i = 0; i < kInPad; ++i) c[i] = 0.f;
    uint32_t mask[kMaskWords];
#pragma unroll
    for (int w = 0; w < kMaskWords; ++w) mask[w] = 0u;

    if (live) {
      const float px = p.x[3 * q], py = p.x[3 * q + 1], pz = p.x[3 * q + 2];
      const float delta = p.dlogit[q];
      const bool has_r = p.dgrad != nullptr && q < p.n_r;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (has_r) { rx = p.dgrad[3 * q]; ry = p.dgrad[3 * q + 1]; rz = p.dgrad[3 * q + 2]; }
      Neighbors<K> nb;
      load_neighbors<K>(m, p.knn_idx, q, px, py, pz, nb);

      // e_k = d w_k / d x . r  (zero without a gradient term)
      float e[K];
      if (has_r && nb.any) {
        float du[K], dusum = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          du[k] = nb.id[k] >= 0 ? -2.f * nb.u[k] * nb.u[k] * (nb.vx[k] * rx + nb.vy[k] * ry + nb.vz[k] * rz) : 0.f;
          dusum += du[k];
        }
        const float invS = 1.0f / nb.S;
#pragma unroll
        for (int k = 0; k < K; ++k) e[k] = nb.id[k] >= 0 ? (du[k] - nb.w[k] * dusum) * invS : 0.f;
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) e[k] = 0.f;
      }

      // z and the tangent input tau0 = s (sum_k e_k q_k + [0; r] sum_k w_k)
      float z[kIn], tau[kIn];
#pragma unroll
      for (int i = 0; i < kIn; ++i) { z[i] = 0.f; tau[i] = 0.f; }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (nb.id[k] >= 0) {
          float f[kFeat];
          load_feature_row(m.gather_features, nb.id[k], f);
          if (layer_norm) { float mu, rs; layer_norm8(f, mu, rs); }
#pragma unroll
          for (int i = 0; i < kFeat; ++i) { z[i] = fmaf(nb.w[k], f[i], z[i]); tau[i] = fmaf(e[k], f[i], tau[i]); }
          z[8] = fmaf(nb.w[k], nb.vx[k], z[8]); z[9] = fmaf(nb.w[k], nb.vy[k], z[9]); z[10] = fmaf(nb.w[k], nb.vz[k], z[10]);
          tau[8] = fmaf(e[k], nb.vx[k], tau[8]); tau[9] = fmaf(e[k], nb.vy[k], tau[9]); tau[10] = fmaf(e[k], nb.vz[k], tau[10]);
        }
      }
      if (has_r && nb.any) { tau[8] += rx; tau[9] += ry; tau[10] += rz; }

      // decoder forward for the activation pattern and a = d logit / d z
      float logit_unused, a[kIn];
      mlp_l1_pairs<H, true>(sm_dec, z, slope, logit_unused, a, mask);

      // c' = [delta z + s tau ; delta]
#pragma unroll
      for (int i = 0; i < kIn; ++i) c[i] = fmaf(delta, z[i], s * tau[i]);
      c[kIn] = delta;
      delta_sum += delta;

      // neural-point feature gradients: dL/df_k = a_f (delta w_k + s e_k), through LN if on
      if (p.gfeat) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (nb.id[k] >= 0) {
            const float coef = fmaf(s, e[k], delta * nb.w[k]);
            float t[kFeat];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) t[i] = coef * a[i];
            if (layer_norm) {
              float f[kFeat], mu, rs;
              load_feature_row(m.gather_features, nb.id[k], f);
              layer_norm8(f, mu, rs);
              layer_norm8_vjp(f, rs, t);
            }
            red_add_row(p.gfeat, nb.id[k], t);
            if (p.touched) p.touched[nb.id[k]] = 1;
          }
        }
      }
    }

    if (p.dec_grad) {
      // stage this warp's 32 samples, then every lane folds them into the hidden rows it owns
      float4* dst = reinterpret_cast<float4*>(my_c + lane * kInPad);
      dst[0] = make_float4(c[0], c[1], c[2], c[3]);
      dst[1] = make_float4(c[4], c[5], c[6], c[7]);
      dst[2] = make_float4(c[8], c[9], c[10], c[11]);
#pragma unroll
      for (int w = 0; w < kMaskWords; ++w) my_m[lane * kMaskWords + w] = mask[w];
      __syncwarp();
#pragma unroll 4
      for (int nn = 0; nn < 32; ++nn) {
        const float4* src = reinterpret_cast<const float4*>(my_c + nn * kInPad);
        const float4 c0 = src[0], c1 = src[1], c2 = src[2];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          // row j = lane + 32 r lives in mask word r, bit `lane`
          const float d = ((my_m[nn * kMaskWords + r] >> lane) & 1u) ? 1.f : slope;
          Gd[r][0] = fmaf(d, c0.x, Gd[r][0]); Gd[r][1] = fmaf(d, c0.y, Gd[r][1]);
          Gd[r][2] = fmaf(d, c0.z, Gd[r][2]); Gd[r][3] = fmaf(d, c0.w, Gd[r][3]);
          Gd[r][4] = fmaf(d, c1.x, Gd[r][4]); Gd[r][5] = fmaf(d, c1.y, Gd[r][5]);
          Gd[r][6] = fmaf(d, c1.z, Gd[r][6]); Gd[r][7] = fmaf(d, c1.w, Gd[r][7]);
          Gd[r][8] = fmaf(d, c2.x, Gd[r][8]); Gd[r][9] = fmaf(d, c2.y, Gd[r][9]);
          Gd[r][10] = fmaf(d, c2.z, Gd[r][10]); Gd[r][11] = fmaf(d, c2.w, Gd[r][11]);
        }
      }
      __syncwarp();
    }
  }

  if (!p.dec_grad) return;
  // ---- block epilogue: sum the warps' partial Gd, turn them into parameter gradients
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int i = 0; i < kInPad; ++i) sm_red[(warp * H + lane + 32 * r) * kInPad + i] = Gd[r][i];
  delta_sum = warp_sum(delta_sum);
  __shared__ float sm_delta[kWarps];
  if (lane == 0) sm_delta[warp] = delta_sum;
  __syncthreads();
  float* gW0 = p.dec_grad;
  float* gb0 = gW0 + H * kIn;
  float* gwout = gb0 + H;
  float* gbout = gwout + H;
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
  if (threadIdx.x == 0 && p.dec.out_bias) {
    float d = 0.f;
    for (int w = 0; w < kWarps; ++w) d += sm_delta[w];
    atomicAdd(gbout, d);
  }
}

// ------------------------------------------------------------------------------------------
// Adam
// ------------------------------------------------------------------------------------------
struct AdamParams {
  // neural-point features
  float* feat;            // [rows,8] parameters (local_geo_features)
  float* feat_grad;       // [rows,8] gradient, re-zeroed for the rows it was applied to
  float* feat_m;          // [rows,8]
  float* feat_v;          // [rows,8]
  const uint8_t* touched; // [rows] or NULL = every row
  int64_t rows;
  int32_t feat_stride;    // floats between consecutive rows of `feat` (8)
  // decoder tensors, in flat order [W0, b0, (W1, b1, ...), wout, bout]
  float* dec_param[2 * CLID_MAX_LEVELS + 2];
  int32_t dec_numel[2 * CLID_MAX_LEVELS + 2];
  int32_t dec_tensors;
  float* dec_grad;        // flat, re-zeroed
  float* dec_m;           // flat
  float* dec_v;           // flat
  // hyper-parameters; the bias corrections are evaluated on the host in double like torch does
  float beta1, beta2, eps, weight_decay;
  float step_size;        // lr / (1 - beta1^t)
  float bc2_sqrt;         // sqrt(1 - beta2^t)
  const float* step_scalars;  // device {step_size, bc2_sqrt} (ClidAdamArgs.step_state + 4) or NULL
};

// ClidAdamArgs.step_state: the optimiser's step counter and the two scalars derived from it
struct AdamStepState {
  int32_t step;
  float step_size;
  float bc2_sqrt;
  int32_t pad;
};

// `inv_bc2` = 1 / sqrt(1 - beta2^t), formed once per thread.  The two IEEE divisions torch's formula implies
// (sqrt(v) / bc2_sqrt, m / denom) compile to ~20-instruction subroutines each and made the step ISSUE-bound
// (ncu r2: issue active 66 %, 4.25 TB/s); a multiplication and a MUFU-based division (<= 2 ulp, far inside the
// 1e-4 parity bound on post-Adam parameters) leave ~15 instructions per element and the kernel on the HBM roofline.
__device__ __forceinline__ float adam_update(float p, float g, float& mm, float& vv, const AdamParams& a,
                                             float step_size, float inv_bc2) {
  mm = mm + (g - mm) * (1.0f - a.beta1);                     // exp_avg.lerp_(grad, 1 - beta1)
  vv = fmaf(1.0f - a.beta2, g * g, vv * a.beta2);            // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = fmaf(__fsqrt_rn(vv), inv_bc2, a.eps);  // sqrt(v) / bc2_sqrt + eps
  return fmaf(-step_size, __fdividef(mm, denom), p);         // p - step_size * m / denom
}

#ifdef CLID_PLAIN_KERNELS
// step += 1; bias corrections in double, as torch/optim/adam.py evaluates them on the host
// zero3 (optional): the iteration's loss accumulators [3], cleared by the same launch (clid_step_begin)
__global__ void adam_advance_kernel(AdamStepState* st, float lr, float beta1, float beta2, float* zero3 = nullptr) {
  if (zero3 != nullptr && blockIdx.x == 0 && threadIdx.x >= 1 && threadIdx.x <= 3) zero3[threadIdx.x - 1] = 0.f;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int t = st->step + 1;
    st->step = t;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    st->step_size = (float)((double)lr / bc1);
    st->bc2_sqrt = (float)sqrt(bc2);
  }
}
#endif

#ifdef CLID_PLAIN_KERNELS
// ------------------------------------------------------------------------------------------
// replay-pool draw (Mapper.get_batch, utils/mapper.py:473-523)
// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator torch / cuRAND use as well
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

struct DrawParams {
  ClidReplayPool pool;
  int64_t n;
  uint64_t seed, offset;
  float* x;
  float* label;
  float* weight;
  int32_t* ts;
  int64_t* index_out;
  // bookkeeping of clid_mapping_run folded into the draw: loss of the previous iteration -> history, clear
  float* loss;
  float* loss_prev_out;
  // ... and the optimiser's step-counter advance of the coming step (adam_advance_kernel), NULL = not here
  AdamStepState* step_state;
  float lr, beta1, beta2;
};

__global__ void __launch_bounds__(256) draw_batch_kernel(const DrawParams p) {
  if (p.loss != nullptr && blockIdx.x == 0 && threadIdx.x < 3) {
    if (p.loss_prev_out != nullptr) p.loss_prev_out[threadIdx.x] = p.loss[threadIdx.x];
    p.loss[threadIdx.x] = 0.f;
  }
  if (p.step_state != nullptr && blockIdx.x == 0 && threadIdx.x == 32) {
    const int t = p.step_state->step + 1;
    p.step_state->step = t;
    const double bc1 = 1.0 - pow((double)p.beta1, (double)t);
    const double bc2 = 1.0 - pow((double)p.beta2, (double)t);
    p.step_state->step_size = (float)((double)p.lr / bc1);
    p.step_state->bc2_sqrt = (float)sqrt(bc2);
  }
  const int64_t n_hist = p.n - p.pool.bs_new;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)p.offset, (uint32_t)(p.offset >> 32)),
                                  make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
    // 64 random bits scaled to [0, range): floor(u * range / 2^64)
    const uint64_t u = ((uint64_t)r.x << 32) | r.y;
    int64_t row;
    if (i < n_hist) row = (int64_t)__umul64hi(u, (uint64_t)p.pool.count);
    else row = p.pool.new_idx[__umul64hi(u, (uint64_t)p.pool.n_new)];
    p.x[3 * i] = p.pool.coord[3 * row]; p.x[3 * i + 1] = p.pool.coord[3 * row + 1]; p.x[3 * i + 2] = p.pool.coord[3 * row + 2];
    p.label[i] = p.pool.sdf_label[row];
    if (p.weight) p.weight[i] = p.pool.weight[row];
    if (p.ts) p.ts[i] = p.pool.time[row];
    if (p.index_out) p.index_out[i] = row;
  }
}

// registration epilogue (include/clid_sdf.h clid_registration_terms)
struct RegParams {
  const float* pc_imu;
  const float* sdf;
  const float* grad;
  const int32_t* nn_count;
  int64_t n;
  float rot[9];
  int32_t min_nn;
  float min_grad, max_grad;
  double* out;
  uint8_t* valid_out;
};

__global__ void __launch_bounds__(256) registration_terms_kernel(const RegParams p) {
  double acc[28];
#pragma unroll
  for (int i = 0; i < 28; ++i) acc[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gx = p.grad[3 * i], gy = p.grad[3 * i + 1], gz = p.grad[3 * i + 2];
    const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
    const bool valid = p.nn_count[i] >= p.min_nn && gn < p.max_grad && gn > p.min_grad;
    if (p.valid_out) p.valid_out[i] = valid ? 1 : 0;
    if (!valid) continue;
    const float x = p.pc_imu[3 * i], y = p.pc_imu[3 * i + 1], z = p.pc_imu[3 * i + 2];
    // A = R [p]x ; [p]x = [[0,-z,y],[z,0,-x],[-y,x,0]]  ->  column c of A = R * column c of [p]x
    const float* R = p.rot;
    float h[6];
    {
      float u0 = gx * R[0] + gy * R[3] + gz * R[6];  // g^T R
      float u1 = gx * R[1] + gy * R[4] + gz * R[7];
      float u2 = gx * R[2] + gy * R[5] + gz * R[8];
      h[0] = -(u1 * z - u2 * y);   // -(g^T R [p]x)_0 = -(u . col0), col0 = (0, z, -y)
      h[1] = -(-u0 * z + u2 * x);  // col1 = (-z, 0, x)
      h[2] = -(u0 * y - u1 * x);   // col2 = (y, -x, 0)
    }
    h[3] = gx; h[4] = gy; h[5] = gz;
    const double s = (double)p.sdf[i];
    const double an = (double)gn - 1.0;
    const double w = (1.0 / (1.0 + an * an)) * (0.4 / (0.4 + s * s)) * 1000.0;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double wa = w * (double)h[a];
#pragma unroll
      for (int b = a; b < 6; ++b) acc[k++] += wa * (double)h[b];
      acc[21 + a] += wa * s;
    }
    acc[27] += 1.0;
  }
  __shared__ double red[8][28];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 28; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    if (v != 0.0) atomicAdd(p.out + threadIdx.x, v);
  }
}

__global__ void copy3_kernel(const float* src, float* dst) {
  if (threadIdx.x < 3) dst[threadIdx.x] = src[threadIdx.x];
}
#endif

#ifdef CLID_PLAIN_KERNELS
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamParams a) {
  // the parameter block stays in the constant bank: the two scalars are the only per-step values
  const float step_size = a.step_scalars ? __ldg(a.step_scalars) : a.step_size;
  const float bc2_sqrt = 1.0f / (a.step_scalars ? __ldg(a.step_scalars + 1) : a.bc2_sqrt);  // used as a factor below
  // feature rows: one thread per 32-byte row, 256-bit loads / stores (LDG.256 / STG.256, sm_100+): four read and four
  // write instructions per row instead of eight each, and two rows in flight per thread
  auto ld8 = [](const float* ptr, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(ptr));
  };
  auto st8 = [](float* ptr, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < a.rows; r0 += 2 * stride) {
    float g[2][8], p[2][8], mm[2][8], vv[2][8];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t row = r0 + u * stride;
      on[u] = row < a.rows && !(a.touched && !a.touched[row]);
      if (on[u]) {
        ld8(a.feat_grad + row * 8, g[u]);
        ld8(a.feat + row * a.feat_stride, p[u]);
        ld8(a.feat_m + row * 8, mm[u]);
        ld8(a.feat_v + row * 8, vv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!on[u]) continue;
      const int64_t row = r0 + u * stride;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (a.weight_decay != 0.f) g[u][i] = fmaf(a.weight_decay, p[u][i], g[u][i]);
        p[u][i] = adam_update(p[u][i], g[u][i], mm[u][i], vv[u][i], a, step_size, bc2_sqrt);
        g[u][i] = 0.f;
      }
      st8(a.feat + row * a.feat_stride, p[u]);
      st8(a.feat_m + row * 8, mm[u]);
      st8(a.feat_v + row * 8, vv[u]);
      st8(a.feat_grad + row * 8, g[u]);
    }
  }
  // decoder: block 0 walks the small tensors
  if (blockIdx.x == 0 && a.dec_grad) {
    int base = 0;
    for (int t = 0; t < a.dec_tensors; ++t) {
      float* prm = a.dec_param[t];
      const int numel = a.dec_numel[t];
      if (prm) {
        for (int i = threadIdx.x; i < numel; i += blockDim.x) {
          float mm = a.dec_m[base + i], vv = a.dec_v[base + i];
          prm[i] = adam_update(prm[i], a.dec_grad[base + i], mm, vv, a, step_size, bc2_sqrt);
          a.dec_m[base + i] = mm; a.dec_v[base + i] = vv;
          a.dec_grad[base + i] = 0.f;
        }
      }
      base += numel;
    }
  }
}

#endif  // CLID_PLAIN_KERNELS

}  // namespace clid
