// Backward and double-backward of the (unfused) feature query, used when a caller
// differentiates through NeuralPoints.query_feature with torch autograd:
//   z = sum_k w_k(x) [LN(f_k); x - p_k]              (model/neural_points.py:620-749)
// backward:          gx = J_x^T gz,  gfeat[idx_k] += LN'(w_k gz_f)
// double backward:   g_gz = J_x ggx, gfeat[idx_k] += LN'(e_k gz_f),  e_k = d w_k/d x . ggx
// (what torch.autograd.grad(..., create_graph=True) + backward computes through
//  index/sort/div/sum in the reference; utils/tools.py:298-311, utils/mapper.py:695-835)
#pragma once
#include "common.cuh"

namespace clid {

template <int K>
struct Neighbors {
  int id[K];
  float vx[K], vy[K], vz[K];
  float u[K], w[K];
  float S;
  bool any;
};

// Re-derive the interpolation state of one query from the neighbour rows saved by the forward.
template <int K>
__device__ __forceinline__ void load_neighbors(const ClidMap& m, const int32_t* __restrict__ knn_idx, int64_t q,
                                               float px, float py, float pz, Neighbors<K>& nb) {
  nb.S = 0.f;
  nb.any = false;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int id = k < m.knn ? knn_idx[q * m.knn + k] : -1;
    nb.id[k] = id;
    if (id >= 0) {
      const float* p = m.gather_points + 3 * (int64_t)id;
      float ex = __ldg(p) - px, ey = __ldg(p + 1) - py, ez = __ldg(p + 2) - pz;
      float d2 = dist2_torch(ex, ey, ez);
      nb.vx[k] = -ex; nb.vy[k] = -ey; nb.vz[k] = -ez;
      nb.u[k] = 1.0f / (d2 + kIdwEps);
      nb.S += nb.u[k];
      nb.any = true;
    } else {
      nb.vx[k] = nb.vy[k] = nb.vz[k] = 0.f;
      nb.u[k] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) nb.w[k] = nb.id[k] >= 0 ? nb.u[k] / nb.S : 0.f;
}

// vector-Jacobian product of y = LN(raw) (no affine): t -> rstd * (t - mean(t) - y * mean(t * y))
__device__ __forceinline__ void layer_norm8_vjp(const float (&y)[kFeat], float rstd, float (&t)[kFeat]) {
  float mt = 0.f, mty = 0.f;
#pragma unroll
  for (int i = 0; i < kFeat; ++i) { mt += t[i]; mty = fmaf(t[i], y[i], mty); }
  mt *= (1.f / kFeat);
  mty *= (1.f / kFeat);
#pragma unroll
  for (int i = 0; i < kFeat; ++i) t[i] = rstd * (t[i] - mt - y[i] * mty);
}

// gfeat[row] += t  (8 floats) with two 16-byte vector reductions (red.global.add.v4.f32, sm_90+)
__device__ __forceinline__ void red_add_row(float* __restrict__ gfeat, int id, const float (&t)[kFeat]) {
  float* dst = gfeat + (int64_t)id * kFeat;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(t[4]), "f"(t[5]), "f"(t[6]), "f"(t[7]) : "memory");
}

struct QueryBwdParams {
  ClidMap map;
  const float* x;
  const int32_t* knn_idx;
  const float* gz;   // [n,11]
  const float* ggx;  // [n,3]   (double backward only)
  float* gx;         // [n,3]   (backward)
  float* g_gz;       // [n,11]  (double backward)
  float* gfeat;      // [n_gather+1,8] accumulated, may be NULL
  int64_t n;
  uint32_t flags;
};

// kSecond == false: backward.  kSecond == true: double backward.
template <int K, bool kSecond>
__global__ void __launch_bounds__(128) query_backward_kernel(const __grid_constant__ QueryBwdParams p) {
  const ClidMap& m = p.map;
  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < p.n; q += (int64_t)gridDim.x * blockDim.x) {
    const float px = p.x[3 * q], py = p.x[3 * q + 1], pz = p.x[3 * q + 2];
    Neighbors<K> nb;
    load_neighbors<K>(m, p.knn_idx, q, px, py, pz, nb);
    float gz[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) gz[i] = p.gz[q * kIn + i];

    if (!nb.any) {
      if constexpr (kSecond) {
#pragma unroll
        for (int i = 0; i < kIn; ++i) p.g_gz[q * kIn + i] = 0.f;
      } else if (p.gx) {
        p.gx[3 * q] = p.gx[3 * q + 1] = p.gx[3 * q + 2] = 0.f;
      }
      continue;
    }
    const float invS = 1.0f / nb.S;

    if constexpr (!kSecond) {
      // ---- gx = (1/S) sum_k (c_k - cbar) du_k + gz_p,  c_k = q_k . gz,  du_k = -2 u_k^2 v_k
      float c[K], cbar = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        c[k] = 0.f;
        if (nb.id[k] >= 0) {
          float f[kFeat];
          load_feature_row(m.gather_features, nb.id[k], f);
          float mu, rs = 1.f;
          if (layer_norm) layer_norm8(f, mu, rs);
          float ck = gz[8] * nb.vx[k] + gz[9] * nb.vy[k] + gz[10] * nb.vz[k];
#pragma unroll
          for (int i = 0; i < kFeat; ++i) ck = fmaf(f[i], gz[i], ck);
          c[k] = ck;
          cbar = fmaf(nb.w[k], ck, cbar);
          if (p.gfeat) {
            float t[kFeat];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) t[i] = nb.w[k] * gz[i];
            if (layer_norm) layer_norm8_vjp(f, rs, t);
            red_add_row(p.gfeat, nb.id[k], t);
          }
        }
      }
      if (p.gx) {
        float gx = 0.f, gy = 0.f, gzz = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (nb.id[k] >= 0) {
            float coef = (c[k] - cbar) * (-2.f * nb.u[k] * nb.u[k]) * invS;
            gx = fmaf(coef, nb.vx[k], gx); gy = fmaf(coef, nb.vy[k], gy); gzz = fmaf(coef, nb.vz[k], gzz);
          }
        }
        p.gx[3 * q] = gx + gz[8];
        p.gx[3 * q + 1] = gy + gz[9];
        p.gx[3 * q + 2] = gzz + gz[10];
      }
    } else {
      // ---- e_k = d w_k / d x . ggx = (1/S)(du_k.ggx - w_k sum_j du_j.ggx)
      const float rx = p.ggx[3 * q], ry = p.ggx[3 * q + 1], rz = p.ggx[3 * q + 2];
      float du[K], dusum = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        du[k] = nb.id[k] >= 0 ? -2.f * nb.u[k] * nb.u[k] * (nb.vx[k] * rx + nb.vy[k] * ry + nb.vz[k] * rz) : 0.f;
        dusum += du[k];
      }
      float out[kIn];
#pragma unroll
      for (int i = 0; i < kIn; ++i) out[i] = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (nb.id[k] >= 0) {
          const float e = (du[k] - nb.w[k] * dusum) * invS;
          float f[kFeat];
          load_feature_row(m.gather_features, nb.id[k], f);
          float mu, rs = 1.f;
          if (layer_norm) layer_norm8(f, mu, rs);
#pragma unroll
          for (int i = 0; i < kFeat; ++i) out[i] = fmaf(e, f[i], out[i]);
          out[8] = fmaf(e, nb.vx[k], out[8]);
          out[9] = fmaf(e, nb.vy[k], out[9]);
          out[10] = fmaf(e, nb.vz[k], out[10]);
          if (p.gfeat) {
            float t[kFeat];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) t[i] = e * gz[i];
            if (layer_norm) layer_norm8_vjp(f, rs, t);
            red_add_row(p.gfeat, nb.id[k], t);
          }
        }
      }
      out[8] += rx; out[9] += ry; out[10] += rz;  // sum_k w_k == 1
#pragma unroll
      for (int i = 0; i < kIn; ++i) p.g_gz[q * kIn + i] = out[i];
    }
  }
}

}  // namespace clid
