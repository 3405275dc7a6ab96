// Fused forward with the decoder on the tensor cores (CLID_TC_DECODER): the same search -> blend -> decoder -> closed-form
// gradient as query_forward_kernel<64, 1, K, bricks> (query_fwd.cuh; model/neural_points.py:553-769, :971-1030,
// model/decoder.py:58-82, utils/tools.py:298-311), but a CTA works on 128 consecutive queries at a time:
//   every thread: search, blend, moments -> z in registers -> its TMEM lane (split hi / lo)
//   thread 0:     layer-1 MMAs (tcgen05.mma kind::tf32, A in TMEM)          [other threads: side effects, outputs]
//   every thread: pre-activations out of TMEM -> logit, masks back into TMEM
//   thread 0:     layer-2 MMAs                                              [other threads: sdf store]
//   every thread: a = d out / d z out of TMEM -> spatial gradient
// (decoder_tc.cuh).  Four such CTAs share an SM, so the tensor-core round trips of one are covered by the others.
#pragma once
#include "common.cuh"
#include "decoder_tc.cuh"
#include "query_fwd.cuh"
#include "search.cuh"

namespace clid {

// dynamic shared memory: tc::Shared | fp32 decoder (MlpLayout<64,1>, sign safeguard) | stencil | brick cursor columns
constexpr int kTcSharedFloats = (int)((sizeof(tc::Shared) + 15) / 16) * 4;
constexpr size_t query_tc_smem_bytes() {
  return (size_t)(kTcSharedFloats + MlpLayout<tc::kH, 1>::kFloats + search_smem_floats<kSearchBricks>() - kStage * 4 * kQueryThreads) * sizeof(float);  // this kernel never stages records
}

// CTA-level tile draw: static first round (tile = blockIdx.x), then tickets from the map's work counter, drawn by
// thread 0 one tile ahead and published through shared memory across the CTA barriers of the tile.  Every CTA draws
// exactly one ticket past the end; the CTA that draws the last one resets the counter for the next launch.
struct CtaTiles {
  int32_t* counter;
  int64_t n_tiles, remaining, tile;
  int it;
  __device__ __forceinline__ CtaTiles(int32_t* counter_, int64_t n_tiles_) : counter(counter_), n_tiles(n_tiles_) {
    remaining = n_tiles - gridDim.x;
    tile = blockIdx.x;
    it = 0;
  }
  __device__ __forceinline__ bool valid() const { return tile < n_tiles; }
  // thread 0, before the first CTA barrier of the tile
  __device__ __forceinline__ void draw(int32_t* slots) {
    if (counter != nullptr && remaining > 0) {
      const int t = atomicAdd(counter, 1);
      if ((int64_t)t == remaining + gridDim.x - 1) atomicExch(counter, 0);
      slots[it & 1] = t;
    }
  }
  // every thread, after the last CTA barrier of the tile
  __device__ __forceinline__ void advance(const int32_t* slots) {
    if (counter == nullptr) tile += gridDim.x;
    else if (remaining <= 0) tile = n_tiles;
    else {
      const int t = slots[it & 1];
      tile = (int64_t)t >= remaining ? n_tiles : (int64_t)gridDim.x + t;
    }
    ++it;
  }
};

template <int K>
__global__ void __launch_bounds__(kQueryThreads, CLID_QUERY_MIN_BLOCKS) query_forward_tc_kernel(const __grid_constant__ QueryParams p) {
  static_assert(kQueryThreads == 128, "one TMEM lane per thread");
  constexpr int H = tc::kH;
  extern __shared__ __align__(16) float smem[];
  tc::Shared& sh = *reinterpret_cast<tc::Shared*>(smem);
  float* sm_dec = smem + kTcSharedFloats;
  uint64_t* stencil = reinterpret_cast<uint64_t*>(sm_dec + MlpLayout<H, 1>::kFloats);
  BrickScratch& scratch = *reinterpret_cast<BrickScratch*>(sm_dec + MlpLayout<H, 1>::kFloats + 2 * 64 * kBrickSlots);
  const ClidMap& m = p.map;

  __shared__ StageBarriers stage;
  stage_barriers_init(stage);
  stage_stencil_async(stencil, p.bricks.stencil, stage);
  stage_decoder_async<H, 1>(sm_dec, p.dec, stage);
  mbar_wait(&stage.decoder, 0);                          // the tensor-core operands are built from the staged weights
  const uint32_t tmem = tc::prologue(sh, sm_dec);        // operands, TMEM, barriers; ends with a CTA barrier
  const uint32_t tlane = tc::lane_base(tmem);
  bool stencil_ready = false;

  const bool training = p.flags & CLID_TRAINING_MODE;
  const bool layer_norm = p.flags & CLID_LAYER_NORM;
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const int knn = m.knn;

  CtaTiles tiles(p.map.work_counter, (p.n + kQueryThreads - 1) / kQueryThreads);
  for (; tiles.valid(); tiles.advance(sh.next_tile)) {
    if (threadIdx.x == 0) tiles.draw(sh.next_tile);
    const uint32_t parity = tiles.it & 1;
    const int64_t q = tiles.tile * kQueryThreads + threadIdx.x;
    const bool live = q < p.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = p.x[3 * q]; py = p.x[3 * q + 1]; pz = p.x[3 * q + 2]; }
    TopK<K> top;
    top.init();
    if (!stencil_ready) { mbar_wait(&stage.stencil, 0); stencil_ready = true; }
    const int count = search_bricks<K, kQueryThreads, false>(m, p.bricks, stencil, &scratch.want[0][threadIdx.x], nullptr, live, px, py, pz, top);

    // ---- neighbour rows, offsets and inverse-distance weights (as query_forward_kernel; dead lanes have no neighbours)
    int row[K];
    float vx[K], vy[K], vz[K], w[K], u[K];
    float S = 0.f;
    {
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
    }
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = row[k] >= 0 ? u[k] / S : 0.f;

    float z[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) z[i] = 0.f;
    float cert = 0.f;
    const bool want_grad = p.out.grad != nullptr;
    Moments mom;
    mom.clear();
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 3) {
      float fb[3][kFeat], cb[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (k0 + j < K) {
          const int rr = row[k0 + j] < 0 ? 0 : row[k0 + j];
          load_feature_row256(m.gather_features, rr, fb[j]);
          cb[j] = p.out.certainty ? __ldg(m.gather_certainties + rr) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
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
          if (want_grad) mom.add(f, u[k], vx[k], vy[k], vz[k]);
        }
      }
    }

    // ---- decoder, layer 1: inputs into TMEM, one thread issues the MMAs
    tc::store_inputs(tlane, z);
    __syncthreads();
    if (threadIdx.x == 0) tc::issue_layer1(sh, tmem);

    // ---- side effects and the outputs that do not need the decoder, behind the tensor core's back
    if (training) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (row[k] >= 0) {
          atomicAdd(m.certainty_accum + row[k], w[k]);
          if (p.ts && m.gather_ts_update) atomicMax(m.gather_ts_update + row[k], p.ts[q]);
        }
      }
    }
    if (live) {
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
    }

    // ---- hidden layer: logit and activation masks
    tc::mbar_wait_bounded(&sh.bar[0], parity);
    float o;
    uint32_t m0, m1;
    tc::hidden_epilogue<false>(sh, sm_dec, tlane, z, slope, o, m0, m1);
    __syncthreads();
    if (threadIdx.x == 0) tc::issue_layer2(sh, tmem);
    const float s = p.dec.sdf_scale;
    if (live && p.out.sdf) p.out.sdf[q] = o * s;

    // ---- a = d out / d z, closed-form spatial gradient
    tc::mbar_wait_bounded(&sh.bar[1], parity);
    float a[kIn];
    tc::load_input_grad(sh, tlane, slope, a);
    if (live && p.out.grad) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (count > 0) {
        float cbar = 0.f;
#pragma unroll
        for (int i = 0; i < kIn; ++i) cbar = fmaf(z[i], a[i], cbar);
        mom.logit_gradient(a, cbar, 1.0f / S, gx, gy, gz);
      }
      p.out.grad[3 * q] = gx * s;
      p.out.grad[3 * q + 1] = gy * s;
      p.out.grad[3 * q + 2] = gz * s;
    }
  }
  tc::epilogue_free(sh, tmem);
}

}  // namespace clid
