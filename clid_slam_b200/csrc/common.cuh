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

// Work distribution of the thread-per-sample kernels: a tile is 32 consecutive samples, owned by
// one warp.  The first round is static (warp w takes tile w, no atomic on the launch path); with a
// counter the remaining tiles are drawn dynamically (an atomic per tile, fetched one tile ahead so
// its round trip overlaps the tile's work), which evens out per-sample cost differences and the
// tail of a static schedule.  Every warp draws exactly one ticket past the end; the warp that draws
// the last one (remaining + n_warps - 1) resets the counter for the next launch.  Without a
// counter the tiles are dealt round-robin.
#ifndef CLID_DYNAMIC_TILES
#define CLID_DYNAMIC_TILES 0  // 1: honour ClidMap.work_counter (ticket scheduler); 0: the round-robin deal is compiled in and the
                              // ticket path (its registers, its atomics) disappears from the kernels -- see ops/query.py
#endif
struct TileScheduler {
  int32_t* counter;
  int64_t n_tiles;
  int64_t next_static;
  int64_t stride;     // warps in the grid
  int64_t remaining;  // tiles beyond the static first round
  int lane;
  int ticket;         // dynamic: the ticket drawn for the NEXT call (valid in lane 0)
  bool first;
  __device__ __forceinline__ TileScheduler(int32_t* counter_, int64_t n) : counter(CLID_DYNAMIC_TILES ? counter_ : nullptr) {
    n_tiles = (n + 31) >> 5;
    lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    next_static = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    stride = (int64_t)gridDim.x * warps_per_block;
    remaining = n_tiles - stride;
    ticket = 0;
    first = true;
    if (counter != nullptr && remaining > 0) draw();
  }
  __device__ __forceinline__ void draw() {
    if (lane == 0) {
      ticket = atomicAdd(counter, 1);
      if ((int64_t)ticket == remaining + stride - 1) atomicExch(counter, 0);  // last ticket of the launch
    }
  }
  // the tile the NEXT call of next() will return (-1: none), without consuming it; reads the ticket drawn one tile
  // ahead, so call it well after next() (the atomic's round trip) -- used to prefetch the next tile's inputs
  __device__ __forceinline__ int64_t peek() const {
    if (counter == nullptr) return next_static < n_tiles ? next_static : -1;
    if (remaining <= 0) return -1;
    const int t = __shfl_sync(0xffffffffu, ticket, 0);
    return (int64_t)t >= remaining ? -1 : stride + (int64_t)t;
  }
  // returns the tile index for this warp or -1 when the work is exhausted (warp-uniform)
  __device__ __forceinline__ int64_t next() {
    if (counter == nullptr || first) {
      first = false;
      const int64_t t = next_static;
      next_static += stride;
      return t < n_tiles ? t : -1;
    }
    if (remaining <= 0) return -1;
    const int t = __shfl_sync(0xffffffffu, ticket, 0);
    if ((int64_t)t >= remaining) return -1;
    draw();
    return stride + (int64_t)t;
  }
};

// Graded start stagger of the persistent kernels: when the grid fills the GPU (k resident CTAs on every SM), the CTAs of
// resident slot s = blockIdx / #SMs start s * CLID_STAGGER_NS later, so that the four warps of a scheduler do not run
// their search (L1TEX-heavy) and decoder (FMA-heavy) phases in lockstep in the first tile round.  Measured (trimmed
// means of 200 launches, 131072 queries, cold L2): 0 / 0.5 / 1 / 2 us per slot -> 50.4 / 49.8 / 49.6 / 50.3 us.
// Smaller grids (one round, not every SM full) are left alone: the delay would only add to their single tile latency.
#ifndef CLID_STAGGER_NS
#define CLID_STAGGER_NS 1000
#endif
__device__ __forceinline__ void stagger_start() {
#if CLID_STAGGER_NS > 0
  uint32_t n_sm;
  asm("mov.u32 %0, %%nsmid;" : "=r"(n_sm));
  if (gridDim.x < 2 * n_sm) return;
  const uint64_t wait_ns = (uint64_t)(blockIdx.x / n_sm) * CLID_STAGGER_NS;
  if (wait_ns) {
    uint64_t t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(128);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < wait_ns);
  }
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- decoder weights staged in shared memory --------------------------------------
// layout (floats): W0 (H * kInPad) | b0[H] | {Wl[H][H] | bl[H]} (levels-1) | wout[H] | bout | pad
// W0 of a one-level decoder is stored UNIT-PAIR interleaved for mlp_l1_pairs: element (j, i) sits at
// (j / 2) * 2 kInPad + 2 i + (j & 1), i.e. pair p holds (W[2p][i], W[2p+1][i]) for i = 0..11 (i = 11 is
// zero padding); deeper decoders keep row-major rows [H][kInPad].
template <int H, int L>
struct MlpLayout {
  __host__ __device__ static constexpr int w0_index(int j, int i) {
    return L == 1 ? (j >> 1) * (2 * kInPad) + 2 * i + (j & 1) : j * kInPad + i;
  }
  static constexpr int kW0 = 0;
  static constexpr int kB0 = kW0 + H * kInPad;
  static constexpr int kHidden = kB0 + H;  // start of level-1.. blocks
  static constexpr int kHiddenStride = H * H + H;
  static constexpr int kWout = kHidden + (L - 1) * kHiddenStride;
  static constexpr int kBout = kWout + H;
  static constexpr int kFloats = ((kBout + 1 + 3) / 4) * 4;
};

// ---- asynchronous staging (sm_90+ async proxy: TMA bulk copies, cp.async, mbarrier) ---------------------
// The CTA prologue used to be: load decoder + stencil through registers, store to shared memory,
// __syncthreads -- one exposed global round trip (~4 % of the forward kernel's stall samples) before the
// first query could even load its coordinates.  Now thread 0 hands the 4 KB stencil to the TMA engine as ONE
// bulk copy that completes on an mbarrier, every thread issues its share of the decoder as cp.async element
// copies (global -> shared, no registers, scattered into the padded / pair-interleaved MlpLayout) that arrive
// on a second mbarrier, and the warps only wait where the data is first needed: the stencil before the first
// search, the decoder before the first MLP -- i.e. behind the first tile's whole search.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {  // make the initialised barriers visible to the async proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned); completes on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
// the mbarrier receives one arrival of this thread once all its cp.async copies issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// Barriers of the asynchronous prologue (static shared memory of the kernels that use it).  Thread 0 initialises
// them, a __syncthreads publishes them; `decoder` expects two arrivals per thread (its cp.async copies, and an
// explicit one that releases its plain padding stores).
struct StageBarriers {
  uint64_t stencil;
  uint64_t decoder;
};

__device__ __forceinline__ void stage_barriers_init(StageBarriers& sb) {
  if (threadIdx.x == 0) {
    mbar_init(&sb.stencil, 1);
    mbar_init(&sb.decoder, 2 * blockDim.x);
    mbar_fence_init();
  }
  __syncthreads();
}

// stencil: ONE TMA bulk copy (UBLKCP) issued by thread 0
__device__ __forceinline__ void stage_stencil_async(uint64_t* sm_stencil, const uint64_t* stencil, StageBarriers& sb) {
  if (threadIdx.x == 0) {
    constexpr uint32_t kBytes = 64 * 8 * sizeof(uint64_t);
    mbar_expect_tx(&sb.stencil, kBytes);
    bulk_copy_g2s(sm_stencil, stencil, kBytes, &sb.stencil);
  }
}

// decoder: every thread issues its element copies and arrives; nobody waits here
template <int H, int L>
__device__ __forceinline__ void stage_decoder_async(float* sm, const ClidDecoder& dec, StageBarriers& sb) {
  using Lay = MlpLayout<H, L>;
  constexpr int kW = H * kIn;
  for (int i = threadIdx.x; i < kW; i += blockDim.x) {
    const int j = i / kIn, c = i - j * kIn;
    cp_async4(sm + Lay::kW0 + Lay::w0_index(j, c), dec.weight[0] + i);
  }
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    sm[Lay::kW0 + Lay::w0_index(j, kIn)] = 0.f;  // padding column
    if (dec.bias[0]) cp_async4(sm + Lay::kB0 + j, dec.bias[0] + j);
    else sm[Lay::kB0 + j] = 0.f;
    cp_async4(sm + Lay::kWout + j, dec.out_weight + j);
  }
#pragma unroll
  for (int l = 1; l < L; ++l) {
    float* blk = sm + Lay::kHidden + (l - 1) * Lay::kHiddenStride;
    if ((reinterpret_cast<uintptr_t>(dec.weight[l]) & 15u) == 0) {
      for (int i = threadIdx.x * 4; i < H * H; i += blockDim.x * 4) cp_async16(blk + i, dec.weight[l] + i);
    } else {
      for (int i = threadIdx.x; i < H * H; i += blockDim.x) cp_async4(blk + i, dec.weight[l] + i);
    }
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
      if (dec.bias[l]) cp_async4(blk + H * H + i, dec.bias[l] + i);
      else blk[H * H + i] = 0.f;
    }
  }
  if (threadIdx.x == 0) {
    if (dec.out_bias) cp_async4(sm + Lay::kBout, dec.out_bias);
    else sm[Lay::kBout] = 0.f;
  }
  cp_async_arrive(&sb.decoder);  // async arrival: when this thread's copies have landed
  mbar_arrive(&sb.decoder);      // release of this thread's plain stores above
}

// synchronous variant (kernels whose first use of the weights is immediate)
template <int H, int L>
__device__ __forceinline__ void stage_decoder(float* sm, const ClidDecoder& dec) {
  using Lay = MlpLayout<H, L>;
  constexpr int kW = H * kIn;
  for (int i = threadIdx.x; i < kW; i += blockDim.x) {
    const int j = i / kIn, c = i - j * kIn;
    sm[Lay::kW0 + Lay::w0_index(j, c)] = __ldg(dec.weight[0] + i);
  }
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    sm[Lay::kW0 + Lay::w0_index(j, kIn)] = 0.f;  // padding column
    sm[Lay::kB0 + j] = dec.bias[0] ? __ldg(dec.bias[0] + j) : 0.f;
    sm[Lay::kWout + j] = __ldg(dec.out_weight + j);
  }
#pragma unroll
  for (int l = 1; l < L; ++l) {
    float* blk = sm + Lay::kHidden + (l - 1) * Lay::kHiddenStride;
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) blk[i] = dec.weight[l][i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) blk[H * H + i] = dec.bias[l] ? dec.bias[l][i] : 0.f;
  }
  if (threadIdx.x == 0) sm[Lay::kBout] = dec.out_bias ? dec.out_bias[0] : 0.f;
}

// One-hidden-level decoder with Blackwell's packed fp32 FMA (FFMA2, `fma.rn.f32x2`, sm_100+), two hidden
// units per instruction: with the unit-pair interleaved W0 (MlpLayout) the pre-activations of units
// (2p, 2p+1) are one chain of packed FMAs over the 11 inputs against (z_i, z_i) -- no horizontal adds, no
// register shuffling, the LDS.128 results are the FFMA2 operands -- and a = d out / d z accumulates as
// (even-unit, odd-unit) partial sums, 11 packed FMAs per pair, folded once at the end.  Per unit: 11 FFMA2
// instead of 12 FFMA2 + 3 FADD + pairing MOVs, and 4 LDS instead of 5.  Optionally records the activation
// pattern as bit masks (unit j -> bit j % 32 of word j / 32) for the backward's decoder-gradient fold.
#ifndef CLID_MLP_UNROLL
#define CLID_MLP_UNROLL 4  // unit pairs per iteration of the rolled decoder loop
#endif
constexpr int kMlpUnroll = CLID_MLP_UNROLL;
template <int H, bool kMask>
__device__ __forceinline__ void mlp_l1_pairs(const float* __restrict__ sm, const float (&z)[kIn], float slope,
                                             float& out, float (&a)[kIn], uint32_t* __restrict__ mask) {
  using Lay = MlpLayout<H, 1>;
  const float4* w = reinterpret_cast<const float4*>(sm + Lay::kW0);  // pair p: float4 6p .. 6p+5
  const float2* b2 = reinterpret_cast<const float2*>(sm + Lay::kB0);
  const float2* wo2 = reinterpret_cast<const float2*>(sm + Lay::kWout);
  float2 zz[kIn], acc[kIn];
#pragma unroll
  for (int i = 0; i < kIn; ++i) { zz[i] = make_float2(z[i], z[i]); acc[i] = make_float2(0.f, 0.f); }
  float2 o2 = make_float2(sm[Lay::kBout], 0.f);
  const float2 sl2 = make_float2(slope, slope);
#pragma unroll
  for (int jw = 0; jw < H / 32; ++jw) {
    uint32_t bits = 0u;
#pragma unroll kMlpUnroll
    for (int pp = 0; pp < 16; ++pp) {
      const int p = jw * 16 + pp;
      const float4 r0 = w[p * 6 + 0], r1 = w[p * 6 + 1], r2 = w[p * 6 + 2];
      const float4 r3 = w[p * 6 + 3], r4 = w[p * 6 + 4], r5 = w[p * 6 + 5];
      float2 e = b2[p], o = make_float2(0.f, 0.f);  // two independent chains over even / odd inputs
      e = __ffma2_rn(make_float2(r0.x, r0.y), zz[0], e);  o = __ffma2_rn(make_float2(r0.z, r0.w), zz[1], o);
      e = __ffma2_rn(make_float2(r1.x, r1.y), zz[2], e);  o = __ffma2_rn(make_float2(r1.z, r1.w), zz[3], o);
      e = __ffma2_rn(make_float2(r2.x, r2.y), zz[4], e);  o = __ffma2_rn(make_float2(r2.z, r2.w), zz[5], o);
      e = __ffma2_rn(make_float2(r3.x, r3.y), zz[6], e);  o = __ffma2_rn(make_float2(r3.z, r3.w), zz[7], o);
      e = __ffma2_rn(make_float2(r4.x, r4.y), zz[8], e);  o = __ffma2_rn(make_float2(r4.z, r4.w), zz[9], o);
      e = __ffma2_rn(make_float2(r5.x, r5.y), zz[10], e);
      const float2 pre = __fadd2_rn(e, o);
      const bool on0 = pre.x > 0.f, on1 = pre.y > 0.f;
      if (kMask) bits = (bits >> 2) | (on0 ? 0x40000000u : 0u) | (on1 ? 0x80000000u : 0u);  // unit jj ends at bit jj
      const float2 wo = wo2[p];
      const float2 ws = __fmul2_rn(wo, sl2);
      const float2 c = make_float2(on0 ? wo.x : ws.x, on1 ? wo.y : ws.y);
      o2 = __ffma2_rn(c, pre, o2);
      acc[0] = __ffma2_rn(make_float2(r0.x, r0.y), c, acc[0]);  acc[1] = __ffma2_rn(make_float2(r0.z, r0.w), c, acc[1]);
      acc[2] = __ffma2_rn(make_float2(r1.x, r1.y), c, acc[2]);  acc[3] = __ffma2_rn(make_float2(r1.z, r1.w), c, acc[3]);
      acc[4] = __ffma2_rn(make_float2(r2.x, r2.y), c, acc[4]);  acc[5] = __ffma2_rn(make_float2(r2.z, r2.w), c, acc[5]);
      acc[6] = __ffma2_rn(make_float2(r3.x, r3.y), c, acc[6]);  acc[7] = __ffma2_rn(make_float2(r3.z, r3.w), c, acc[7]);
      acc[8] = __ffma2_rn(make_float2(r4.x, r4.y), c, acc[8]);  acc[9] = __ffma2_rn(make_float2(r4.z, r4.w), c, acc[9]);
      acc[10] = __ffma2_rn(make_float2(r5.x, r5.y), c, acc[10]);
    }
    if (kMask) mask[jw] = bits;
  }
  out = o2.x + o2.y;
#pragma unroll
  for (int i = 0; i < kIn; ++i) a[i] = acc[i].x + acc[i].y;
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
    mlp_l1_pairs<H, false>(sm, z, slope, out, a, nullptr);
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

// one 32-byte feature row with a single 256-bit load (LDG.E.256, sm_100+)
__device__ __forceinline__ void load_feature_row256(const float* __restrict__ feats, int id, float (&f)[kFeat]) {
  const float* row = feats + (int64_t)id * kFeat;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(row));
}

// Second moments of a query's neighbourhood.  With t_k = -2 u_k^2 (d u_k / d x = t_k v_k):
//   M_j = sum_k t_k v_kj f_k   (3 x 8),   P_jl = sum_k t_k v_kj v_kl   (symmetric 3 x 3),   qv_j = sum_k t_k v_kj
// The spatial gradient and the tangent input of the analytic eikonal term are linear in them:
//   sum_k c_k d w_k/d x = (1/S) (a_f . M_j + a_p . P_j - cbar qv_j)_j,   c_k = [f_k; v_k] . a,  cbar = z . a
// so they are accumulated while the feature rows pass through registers for the blend, and no
// row has to be read a second time (the gathers are L1-wavefront bound).
struct Moments {
  float M[3][kFeat];
  float P[6];  // xx xy xz yy yz zz
  float qv[3];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      qv[j] = 0.f;
#pragma unroll
      for (int i = 0; i < kFeat; ++i) M[j][i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) P[i] = 0.f;
  }
  __device__ __forceinline__ void add(const float (&f)[kFeat], float u, float vx, float vy, float vz) {
    const float t2 = -2.f * u * u;
    const float tx = t2 * vx, ty = t2 * vy, tz = t2 * vz;
#pragma unroll
    for (int i = 0; i < kFeat; ++i) {
      M[0][i] = fmaf(tx, f[i], M[0][i]); M[1][i] = fmaf(ty, f[i], M[1][i]); M[2][i] = fmaf(tz, f[i], M[2][i]);
    }
    P[0] = fmaf(tx, vx, P[0]); P[1] = fmaf(tx, vy, P[1]); P[2] = fmaf(tx, vz, P[2]);
    P[3] = fmaf(ty, vy, P[3]); P[4] = fmaf(ty, vz, P[4]); P[5] = fmaf(tz, vz, P[5]);
    qv[0] += tx; qv[1] += ty; qv[2] += tz;
  }
  // (1/S) sum_k (c_k - cbar) d u_k/d x + a_p  (un-scaled d logit / d x for a query with neighbours)
  __device__ __forceinline__ void logit_gradient(const float (&a)[kIn], float cbar, float invS, float& gx, float& gy,
                                                 float& gz) const {
    float sx = a[8] * P[0] + a[9] * P[1] + a[10] * P[2] - cbar * qv[0];
    float sy = a[8] * P[1] + a[9] * P[3] + a[10] * P[4] - cbar * qv[1];
    float sz = a[8] * P[2] + a[9] * P[4] + a[10] * P[5] - cbar * qv[2];
#pragma unroll
    for (int i = 0; i < kFeat; ++i) { sx = fmaf(a[i], M[0][i], sx); sy = fmaf(a[i], M[1][i], sy); sz = fmaf(a[i], M[2][i], sz); }
    gx = fmaf(invS, sx, a[8]); gy = fmaf(invS, sy, a[9]); gz = fmaf(invS, sz, a[10]);
  }
};

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
