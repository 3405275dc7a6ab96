// One-hidden-level decoder (11 -> 64 -> 1) on the 5th-generation tensor cores (tcgen05 + TMEM), for the kernels
// whose CTA evaluates 128 samples at a time (thread t = sample t = TMEM lane t).
//   model/decoder.py:58-82  Decoder.mlp / sdf        out_t = b_out + sum_j c_tj pre_tj
//   utils/tools.py:298-311  get_gradient (d / d z)   a_t   = sum_j c_tj W0[j,:]
// with pre = W0 z + b0 and c_tj = w_out[j] (pre_tj > 0 ? 1 : slope).
//
// Both contractions are issued by ONE thread as tcgen05.mma.kind::tf32 with the A operand in TMEM (no shared-memory
// staging of per-sample data at all) and the static weights as B operands in shared memory:
//   layer 1   D1[128 x 64] = Zext[128 x 16] . W0ext^T          Zext = [z | 1 | 0 0 0 0], W0ext = [W0 | b0 | 0]
//   layer 2   D2[128 x 16] = Mask[128 x 64] . (w_out (.) W0ext)  Mask_tj = [pre_tj > 0]  (0 / 1: exact in tf32)
// fp32 accuracy comes from operand splitting (x = hi + lo, hi = tf32(x), lo = tf32(x - hi)):
//   layer 1: Zhi.Whi + Zlo.Whi + Zhi.Wlo (3 x 2 K-steps), layer 2: Mask.Bhi + Mask.Blo (2 x 8 K-steps, A exact);
// the tensor core accumulates in fp32.  What a split product cannot guarantee is the SIGN of a pre-activation
// that is ~1e-6 of its own terms, and a flipped ReLU mask changes `a` discontinuously, so every unit whose |pre|
// is below an error bound of the split product is recomputed with the fp32 FMA chain of mlp_l1_pairs
// (common.cuh): masks are those of the fp32 path (a few units per 10^5 samples take the slow branch).
//
// TMEM columns of a CTA (128 allocated; 4 CTAs/SM use the whole 512):
//   [0,16) Zhi   [16,32) Zlo   [32,96) D1, overwritten in place by Mask   [96,112) D2
#pragma once
#include "common.cuh"

namespace clid {
namespace tc {

constexpr int kH = 64;
constexpr int kK1 = 16;  // layer-1 K: 11 inputs, the bias column, 4 zero columns (two K = 8 steps)
constexpr int kTmemCols = 128;
constexpr int kColZhi = 0, kColZlo = 16, kColD1 = 32, kColD2 = 96;
constexpr uint32_t kOneBits = 0x3f800000u;

// Shared-memory operands.  b1 / b2 are "K-major, no swizzle" UMMA operands: 8 x 16-byte core matrices,
//   b1 element (n = unit j, k = input i): chunk i/4 (stride 8 core matrices = LBO 1024 B), row group j/8 (SBO 128 B)
//   b2 element (n = input i, k = unit j): chunk j/4 (stride 2 core matrices = LBO 256 B),  row group i/8 (SBO 128 B)
struct __align__(16) Shared {
  float b1[2][kH * kK1];   // hi, lo of W0ext
  float b2[2][kK1 * kH];   // hi, lo of w_out[j] W0ext[j][i]
  float wout[kH];
  float a_all[kK1];        // sum_j w_out[j] W0ext[j][i]: the leaky-ReLU part of a
  float bout;
  uint32_t w1max_bits;     // max_j ||W0[j,:]||_1 (float bits; non-negative floats order like integers)
  uint32_t bmax_bits;      // max_j |b0[j]|
  uint32_t tmem_base;
  uint64_t bar[2];         // completion of the layer-1 / layer-2 MMAs (tcgen05.commit)
  int32_t next_tile[2];
};

__host__ __device__ constexpr int b1_index(int j, int i) { return (i >> 2) * 256 + (j >> 3) * 32 + (j & 7) * 4 + (i & 3); }
__host__ __device__ constexpr int b2_index(int i, int j) { return (j >> 2) * 64 + (i >> 3) * 32 + (i & 7) * 4 + (j & 3); }

// x = hi + lo with hi = x rounded to tf32 (10 mantissa bits, round to nearest by an integer add: `cvt.rna.tf32.f32`
// has no single SASS instruction and costs ~6) and lo = x - hi exactly; the tensor core ignores the 13 low mantissa
// bits of lo, a relative error of 2^-21 of x.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// ---- TMEM management (one warp allocates and frees) ------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(kTmemCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kTmemCols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_smem_to_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive columns of this thread's lane (32x32b: lane = 32 (warp % 4) + lane id, one register per column)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MMA issue (one thread) ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(const void* p, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint64_t addr = smem_addr(p);
  return ((addr & 0x3FFFFull) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// bounded wait: a descriptor / protocol mistake must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// ---- CTA prologue: operands, barriers, TMEM ---------------------------------------------------------------------
// All threads of the CTA call this (blockDim.x = 128) once the fp32 decoder is staged in sm_dec (MlpLayout<64,1>): the
// caller has waited on its cp.async barrier, or has staged it synchronously (the first barrier below publishes it).
// Everything is built from shared memory: no global round trip.  Ends with a CTA barrier.  Returns the TMEM base.
__device__ __forceinline__ uint32_t prologue(Shared& sh, const float* __restrict__ sm_dec) {
  using Lay = MlpLayout<kH, 1>;
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid < 32) tmem_alloc(&sh.tmem_base);
  if (tid == 0) {
    mbar_init(&sh.bar[0], 1);
    mbar_init(&sh.bar[1], 1);
    mbar_fence_init();
    sh.w1max_bits = 0u;
    sh.bmax_bits = 0u;
  }
  if (tid < kK1) sh.a_all[tid] = 0.f;
  __syncthreads();
  auto w0ext = [&](int j, int i) -> float {
    return i < kIn ? sm_dec[Lay::kW0 + Lay::w0_index(j, i)] : (i == kIn ? sm_dec[Lay::kB0 + j] : 0.f);
  };
  // one 16-byte K chunk per item: b1 chunk = 4 inputs of one unit, b2 chunk = 4 units of one input
  for (int e = tid; e < kH * (kK1 / 4); e += nthr) {
    const int j = e >> 2, i0 = (e & 3) * 4;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) split_tf32(w0ext(j, i0 + v), hi[v], lo[v]);
    *reinterpret_cast<uint4*>(&sh.b1[0][b1_index(j, i0)]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(&sh.b1[1][b1_index(j, i0)]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int e = tid; e < kK1 * (kH / 4); e += nthr) {
    const int i = e & 15, j0 = (e >> 4) * 4;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) split_tf32(sm_dec[Lay::kWout + j0 + v] * w0ext(j0 + v, i), hi[v], lo[v]);
    *reinterpret_cast<uint4*>(&sh.b2[0][b2_index(i, j0)]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(&sh.b2[1][b2_index(i, j0)]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int j = tid; j < kH; j += nthr) {
    sh.wout[j] = sm_dec[Lay::kWout + j];
    float l1 = 0.f;
#pragma unroll
    for (int i = 0; i < kIn; ++i) l1 += fabsf(w0ext(j, i));
    atomicMax(&sh.w1max_bits, __float_as_uint(l1));
    atomicMax(&sh.bmax_bits, __float_as_uint(fabsf(sm_dec[Lay::kB0 + j])));
  }
  {  // a_all[i] = sum_j w_out[j] W0ext[j][i]: 8 threads per column, 8 units each
    const int i = tid & 15, g = (tid >> 4) & 7;
    float s = 0.f;
#pragma unroll
    for (int jj = 0; jj < kH / 8; ++jj) s = fmaf(sm_dec[Lay::kWout + 8 * g + jj], w0ext(8 * g + jj, i), s);
    if (tid < 128) atomicAdd(&sh.a_all[i], s);
  }
  if (tid == 0) sh.bout = sm_dec[Lay::kBout];
  fence_smem_to_async_proxy();  // the tensor core reads b1 / b2 through the async proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return sh.tmem_base;
}

__device__ __forceinline__ void epilogue_free(Shared& sh, uint32_t tmem) {
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem);
}

// per-sample bound on |pre(split product) - pre(fp32 chain)|: each of the 12 products is off by < 2^-20 of its
// magnitude (dropped lo.lo term, tf32 rounding of the lo parts) plus the accumulation roundings of either path
__device__ __forceinline__ float sign_threshold(const Shared& sh, const float (&z)[kIn]) {
  float zmax = 0.f;
#pragma unroll
  for (int i = 0; i < kIn; ++i) zmax = fmaxf(zmax, fabsf(z[i]));
  return 4e-6f * fmaf(zmax, __uint_as_float(sh.w1max_bits), __uint_as_float(sh.bmax_bits));
}

// this thread's decoder input -> its TMEM lane, split into hi / lo (all 32 lanes of the warp call this)
__device__ __forceinline__ void store_inputs(uint32_t tlane, const float (&z)[kIn]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < kIn; ++i) split_tf32(z[i], hi[i], lo[i]);
  hi[kIn] = kOneBits;  // bias column
  lo[kIn] = 0u;
#pragma unroll
  for (int i = kIn + 1; i < 16; ++i) { hi[i] = 0u; lo[i] = 0u; }
  tmem_st16(tlane + kColZhi, hi);
  tmem_st16(tlane + kColZlo, lo);
  tmem_wait_st();
  fence_before_sync();
}

// one thread, after the CTA barrier that follows store_inputs
__device__ __forceinline__ void issue_layer1(Shared& sh, uint32_t tmem) {
  fence_after_sync();
  constexpr uint32_t id = idesc_tf32(kH);
  const uint64_t whi = smem_desc(sh.b1[0], 1024, 128), wlo = smem_desc(sh.b1[1], 1024, 128);
  constexpr uint64_t kStep = 2048 >> 4;  // two K chunks per K = 8 step
  mma_ts(tmem + kColD1, tmem + kColZhi, whi, id, false);
  mma_ts(tmem + kColD1, tmem + kColZhi + 8, whi + kStep, id, true);
  mma_ts(tmem + kColD1, tmem + kColZlo, whi, id, true);
  mma_ts(tmem + kColD1, tmem + kColZlo + 8, whi + kStep, id, true);
  mma_ts(tmem + kColD1, tmem + kColZhi, wlo, id, true);
  mma_ts(tmem + kColD1, tmem + kColZhi + 8, wlo + kStep, id, true);
  mma_commit(&sh.bar[0]);
}

// pre-activation of unit j exactly as mlp_l1_pairs computes it (two chains over even / odd inputs)
template <int H>
__device__ __forceinline__ float exact_pre(const float* __restrict__ sm_dec, const float (&z)[kIn], int j) {
  using Lay = MlpLayout<H, 1>;
  const float* w = sm_dec + Lay::kW0;
  float e = sm_dec[Lay::kB0 + j], o = 0.f;
  e = fmaf(w[Lay::w0_index(j, 0)], z[0], e);  o = fmaf(w[Lay::w0_index(j, 1)], z[1], o);
  e = fmaf(w[Lay::w0_index(j, 2)], z[2], e);  o = fmaf(w[Lay::w0_index(j, 3)], z[3], o);
  e = fmaf(w[Lay::w0_index(j, 4)], z[4], e);  o = fmaf(w[Lay::w0_index(j, 5)], z[5], o);
  e = fmaf(w[Lay::w0_index(j, 6)], z[6], e);  o = fmaf(w[Lay::w0_index(j, 7)], z[7], o);
  e = fmaf(w[Lay::w0_index(j, 8)], z[8], e);  o = fmaf(w[Lay::w0_index(j, 9)], z[9], o);
  e = fmaf(w[Lay::w0_index(j, 10)], z[10], e);
  return e + o;
}

// every thread, after waiting on bar[0]: pre-activations out of TMEM, logit, activation masks back into TMEM as the
// A operand of layer 2.  sm_dec: the fp32 weights in MlpLayout<64,1> (for the sign safeguard).
// mask0 / mask1 (kMask): unit j -> bit j % 32 of word j / 32 (the layout mlp_l1_pairs records).
// Per unit: max (activation), FFMA (logit), FSET (mask as 1.0 / 0.0) -- against ~39 instructions of the FMA decoder.
template <bool kMask, bool kLeaky>
__device__ __forceinline__ void hidden_epilogue_t(const Shared& sh, const float* __restrict__ sm_dec, uint32_t tlane,
                                                  const float (&z)[kIn], float& out, uint32_t& mask0, uint32_t& mask1) {
  fence_after_sync();
  const float thr = sign_threshold(sh, z);
  float o = sh.bout;
  uint32_t m0 = 0u, m1 = 0u;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[16];
    tmem_ld16(tlane + kColD1 + 16 * c, r);
    tmem_wait_ld();
    float mn = __int_as_float(0x7f800000);
#pragma unroll
    for (int u = 0; u < 16; ++u) mn = fminf(mn, fabsf(__uint_as_float(r[u])));
    if (mn < thr) {  // rare: a sign too close to call from the split product
#pragma unroll 1
      for (int u = 0; u < 16; ++u) {
        // r[] is indexed with a loop variable only on this cold path: select by predicate, no local memory
        const float e = exact_pre<kH>(sm_dec, z, 16 * c + u);
#pragma unroll
        for (int v = 0; v < 16; ++v)
          if (v == u && fabsf(__uint_as_float(r[v])) < thr) r[v] = __float_as_uint(e);
      }
    }
    uint32_t bits = 0u;
    const float4* wo4 = reinterpret_cast<const float4*>(sh.wout + 16 * c);
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) {
      const float4 wo = wo4[u4];
      const float wv[4] = {wo.x, wo.y, wo.z, wo.w};
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int u = 4 * u4 + v;
        const float pre = __uint_as_float(r[u]);
        const float act = kLeaky ? fmaxf(pre, kLeakySlope * pre) : fmaxf(pre, 0.f);
        o = fmaf(wv[v], act, o);
        r[u] = __float_as_uint(pre > 0.f ? 1.0f : 0.0f);
        if (kMask) bits |= pre > 0.f ? (1u << u) : 0u;
      }
    }
    tmem_st16(tlane + kColD1 + 16 * c, r);
    if (kMask) {
      if (c == 0) m0 = bits;
      else if (c == 1) m0 |= bits << 16;
      else if (c == 2) m1 = bits;
      else m1 |= bits << 16;
    }
  }
  tmem_wait_st();
  fence_before_sync();
  out = o;
  mask0 = m0;
  mask1 = m1;
}

template <bool kMask>
__device__ __forceinline__ void hidden_epilogue(const Shared& sh, const float* __restrict__ sm_dec, uint32_t tlane,
                                                const float (&z)[kIn], float slope, float& out, uint32_t& mask0,
                                                uint32_t& mask1) {
  if (slope == 0.f) hidden_epilogue_t<kMask, false>(sh, sm_dec, tlane, z, out, mask0, mask1);
  else hidden_epilogue_t<kMask, true>(sh, sm_dec, tlane, z, out, mask0, mask1);
}

// one thread, after the CTA barrier that follows hidden_epilogue
__device__ __forceinline__ void issue_layer2(Shared& sh, uint32_t tmem) {
  fence_after_sync();
  constexpr uint32_t id = idesc_tf32(kK1);
  const uint64_t bhi = smem_desc(sh.b2[0], 256, 128), blo = smem_desc(sh.b2[1], 256, 128);
  constexpr uint64_t kStep = 512 >> 4;
#pragma unroll
  for (int s = 0; s < kH / 8; ++s) mma_ts(tmem + kColD2, tmem + kColD1 + 8 * s, bhi + s * kStep, id, s > 0);
#pragma unroll
  for (int s = 0; s < kH / 8; ++s) mma_ts(tmem + kColD2, tmem + kColD1 + 8 * s, blo + s * kStep, id, true);
  mma_commit(&sh.bar[1]);
}

// every thread, after waiting on bar[1]: a = d out / d z
__device__ __forceinline__ void load_input_grad(const Shared& sh, uint32_t tlane, float slope, float (&a)[kIn]) {
  fence_after_sync();
  uint32_t r[16];
  tmem_ld16(tlane + kColD2, r);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < kIn; ++i) {
    const float on_part = __uint_as_float(r[i]);  // sum over active units
    a[i] = slope == 0.f ? on_part : fmaf(1.f - slope, on_part, slope * sh.a_all[i]);
  }
}

__device__ __forceinline__ uint32_t lane_base(uint32_t tmem) { return tmem + ((uint32_t)((threadIdx.x >> 5) & 3) << 21); }  // 32 lanes << 16

}  // namespace tc

// Stand-alone decoder evaluation on given inputs (model/decoder.py:58-82 Decoder.mlp + d / d z): the unit test of
// the tensor-core path, and what a caller with its own feature vectors uses.  One CTA = 128 samples per pass.
struct DecoderEvalParams {
  ClidDecoder dec;
  const float* z;   // [n,11]
  float* out;       // [n] un-scaled logit
  float* a;         // [n,11] or NULL
  uint32_t* mask;   // [n,2] or NULL
  int64_t n;
  uint32_t flags;
};

#ifdef CLID_PLAIN_KERNELS
__global__ void __launch_bounds__(128, 4) decoder_eval_tc_kernel(const __grid_constant__ DecoderEvalParams p) {
  extern __shared__ __align__(16) float smem[];
  tc::Shared& sh = *reinterpret_cast<tc::Shared*>(smem);
  float* sm_dec = smem + (sizeof(tc::Shared) + 3) / 4;
  stage_decoder<tc::kH, 1>(sm_dec, p.dec);
  const uint32_t tmem = tc::prologue(sh, sm_dec);
  const uint32_t tlane = tc::lane_base(tmem);
  const float slope = (p.flags & CLID_LEAKY_RELU) ? kLeakySlope : 0.f;
  const int64_t n_tiles = (p.n + 127) / 128;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t q = tile * 128 + threadIdx.x;
    float z[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) z[i] = q < p.n ? p.z[q * kIn + i] : 0.f;
    tc::store_inputs(tlane, z);
    __syncthreads();
    if (threadIdx.x == 0) tc::issue_layer1(sh, tmem);
    tc::mbar_wait_bounded(&sh.bar[0], parity);
    float out;
    uint32_t m0, m1;
    tc::hidden_epilogue<true>(sh, sm_dec, tlane, z, slope, out, m0, m1);
    __syncthreads();
    if (threadIdx.x == 0) tc::issue_layer2(sh, tmem);
    tc::mbar_wait_bounded(&sh.bar[1], parity);
    float a[kIn];
    tc::load_input_grad(sh, tlane, slope, a);
    if (q < p.n) {
      p.out[q] = out;
      if (p.a) {
#pragma unroll
        for (int i = 0; i < kIn; ++i) p.a[q * kIn + i] = a[i];
      }
      if (p.mask) { p.mask[2 * q] = m0; p.mask[2 * q + 1] = m1; }
    }
    parity ^= 1u;
  }
  tc::epilogue_free(sh, tmem);
}
#endif

}  // namespace clid
