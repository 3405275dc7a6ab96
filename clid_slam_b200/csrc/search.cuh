// Candidate enumeration of a query: top-K bookkeeping, the probe through the reference's hash table
// and the thread-per-query walk through the brick index (model/neural_points.py:971-1030, :595-612).
#pragma once
#include "common.cuh"

namespace clid {

#ifndef CLID_QUERY_MIN_BLOCKS
#define CLID_QUERY_MIN_BLOCKS 4  // resident CTAs per SM the forward kernel is register-budgeted for
#endif
#ifndef CLID_QUERY_THREADS
#define CLID_QUERY_THREADS 128
#endif
constexpr int kQueryThreads = CLID_QUERY_THREADS;
constexpr int kBrickSlots = 8;  // span 2: a neighbourhood touches at most 2x2x2 bricks


__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {  // LDG.E.256 (sm_100+)
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

#ifndef CLID_TOPK_PARALLEL
#define CLID_TOPK_PARALLEL 1
#endif
// Ascending top-K by squared distance with an integer payload; ties keep the earlier candidate.
template <int K>
struct TopK {
  float d[K];
  int id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = __int_as_float(0x7f800000); id[k] = -1; }
  }
  __device__ __forceinline__ void insert(float dc, int ic) {
    if (!(dc < d[K - 1])) return;
#if CLID_TOPK_PARALLEL
    // all K comparisons are independent, every slot then takes {itself, its left neighbour, the candidate}
    // from the OLD values: dependency depth 3 instead of a K-long compare-exchange chain (the list is
    // ascending, so lt[k-1] implies lt[k]); ties keep the earlier candidate as before
    bool lt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) lt[k] = dc < d[k];
#pragma unroll
    for (int k = K - 1; k > 0; --k) {
      d[k] = lt[k] ? (lt[k - 1] ? d[k - 1] : dc) : d[k];
      id[k] = lt[k] ? (lt[k - 1] ? id[k - 1] : ic) : id[k];
    }
    d[0] = lt[0] ? dc : d[0];
    id[0] = lt[0] ? ic : id[0];
#else
    // one pass of compare-exchange from the front: the carried element is always the larger one
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool lt = dc < d[k];
      const float dk = d[k];
      const int ik = id[k];
      d[k] = lt ? dc : dk;
      id[k] = lt ? ic : ik;
      dc = lt ? dk : dc;
      ic = lt ? ik : ic;
    }
#endif
  }
};

// ---- candidate enumeration through the reference's hash table ----------------------------
// Payload of the top-K: gather row (local row with CLID_QUERY_LOCALLY, else global id).
template <int K>
__device__ __forceinline__ int search_hashed(const ClidMap& m, const int64_t* __restrict__ cell_mod, float px,
                                             float py, float pz, const bool local, const bool time_filter,
                                             TopK<K>& top) {
  constexpr int U = 9;
  const int gx = cell_of(px, m.resolution), gy = cell_of(py, m.resolution), gz = cell_of(pz, m.resolution);
  const int64_t B = m.buffer_size;
  const int64_t m0 = floor_mod((int64_t)gx * m.primes[0] + (int64_t)gy * m.primes[1] + (int64_t)gz * m.primes[2], B);
  float td_cur = 0.f;
  if (time_filter) td_cur = m.travel_dist[m.cur_ts];
  int count = 0;
  for (int c0 = 0; c0 < m.kc; c0 += U) {
    int gi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int c = c0 + u;
      int64_t v = -1;
      if (c < m.kc) {
        int64_t slot = m0 + cell_mod[c];
        slot = slot >= B ? slot - B : slot;
        v = __ldg(m.buffer_pt_index + slot);
      }
      gi[u] = (int)v;
    }
    float cx[U], cy[U], cz[U];
    int li[U];
    int tsc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int g = gi[u] < 0 ? 0 : gi[u];  // invalid lanes read row 0 (always mapped); result discarded
      const float* p = m.neural_points + 3 * (int64_t)g;
      cx[u] = __ldg(p); cy[u] = __ldg(p + 1); cz[u] = __ldg(p + 2);
      li[u] = local ? (int)__ldg(m.global2local + g) : g;
      tsc[u] = time_filter ? __ldg(m.point_ts_create + g) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool ok = gi[u] >= 0;
      if (time_filter) {
        float gap = fabsf(td_cur - __ldg(m.travel_dist + tsc[u]));
        ok = ok && (gap < m.diff_travel_dist_local);
      }
      float d2 = dist2_torch(cx[u] - px, cy[u] - py, cz[u] - pz);  // neighbour - query, as the reference
      ok = ok && !(d2 > m.max_valid_dist2) && li[u] >= 0;
      if (ok) {
        ++count;
        top.insert(d2, li[u]);
      }
    }
  }
  return count;
}

// ---- candidate enumeration through the brick index ---------------------------------------
// Payload of the top-K: record index.  A 64-cell brick is walked as two 32-cell halves (z < 2,
// z >= 2) so every bit operation is a single 32-bit instruction.  The grid carries a one-brick
// empty apron (ClidBricks.apron) and a neighbourhood spans 2 x 2 x 2 bricks (span == 2): a query is
// range-tested once, the eight header addresses follow by constant strides.
// Phase 1: load the 8 brick headers, AND with the stencil, compact the non-empty (want, occupancy,
// first-record) half-brick triples into this lane's scratch column (word s of the column is
// col[s * kStride]: want in slots 0..11, occupancy in 12..23, first record in 24..35).
// Phase 2 (warp-converged): every lane pops up to kWalkBatch candidates, issues their record loads
// together, then ranks them; a warp iterates ceil(max-over-lanes(candidates) / kWalkBatch) times
// instead of diverging inside nested loops.
// A neighbourhood is at most 5 cells wide (reach <= 2): 5 consecutive z cells touch at most 3 of the
// 2-cell z halves, so at most 3 x 2 x 2 half-bricks can be non-empty.
constexpr int kHalfSlots = 12;
#ifndef CLID_WALK_BATCH
#define CLID_WALK_BATCH 7   // r1: 2 / 4 / 6 / 8: 61.5 / 58.4 / 57.4 / 57.4 us (forward, 131072 queries, cold L2); r2 (trimmed means of
                            // 200 launches, the event clock ticks every 1.9 us): 6 / 7 / 11: 51.4 / 50.7 / 50.9 us -- a fully occupied
                            // planar neighbourhood has 21 candidates = three batches of 7
#endif
constexpr int kWalkBatch = CLID_WALK_BATCH;
#ifndef CLID_WALK_PIPELINE
#define CLID_WALK_PIPELINE 0  // depth-2 software pipeline of the walk (next batch's loads in flight while this one is
                              // ranked): measured SLOWER, forward 53.3 -> 60.4 us cold (register pressure, code size)
#endif
#ifndef CLID_PF_RECORDS
#define CLID_PF_RECORDS 0   // L2 prefetch of the record lines of every non-empty half-brick: paid off with
                            // 16-B-per-iteration walks, no longer with 6 record loads in flight per lane
#endif
#ifndef CLID_FEAT_BATCH
#define CLID_FEAT_BATCH 3   // feature rows of the top-K requested together before the first is blended
#endif
constexpr int kFeatBatch = CLID_FEAT_BATCH;
#ifndef CLID_PF_NEXT_TILE
#define CLID_PF_NEXT_TILE 1  // L2 prefetch of the next tile's coordinates (and of this tile's labels in training)
#endif
#ifndef CLID_PF_FEATURES
#define CLID_PF_FEATURES 0  // L2 prefetch of the feature row of every candidate that enters the top-K: ~11 extra L1TEX requests per
                            // query; trimmed means of 200 launches, on / off: forward 50.8 / 50.5 us cold, 45.5 / 43.9 warm,
                            // training kernel 94.0 / 93.6 cold, 88.8 / 86.7 warm -- off (the same prefetch of the certainties: +3 us)
#endif

#ifndef CLID_STAGE_RECORDS
#define CLID_STAGE_RECORDS 0  // > 0: candidate records are staged through shared memory by cp.async (LDGSTS), this many
                              // slots per lane and chunk; the walk then ranks out of shared memory (see search_bricks)
#endif
constexpr int kStage = CLID_STAGE_RECORDS;

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct BrickScratch {  // [slot][thread] columns of a 128-thread CTA
  uint32_t want[kHalfSlots][kQueryThreads];
  uint32_t occ[kHalfSlots][kQueryThreads];
  int base[kHalfSlots][kQueryThreads];
};

template <int K, int kStride, bool kStaged = (kStage > 0)>
__device__ __forceinline__ int search_bricks(const ClidMap& m, const ClidBricks& b, const uint64_t* stencil,
                                           uint32_t* col, float4* stg, bool live, float px, float py, float pz,
                                           TopK<K>& top) {
  const int rx = cell_of(px, m.resolution) - b.origin[0] - b.reach;
  const int ry = cell_of(py, m.resolution) - b.origin[1] - b.reach;
  const int rz = cell_of(pz, m.resolution) - b.origin[2] - b.reach;
  const int bx0 = rx >> 2, by0 = ry >> 2, bz0 = rz >> 2;
  const int D0 = b.dims[0], D1 = b.dims[1];
  const bool in = live && (unsigned)bx0 < (unsigned)(D0 - 1) && (unsigned)by0 < (unsigned)(D1 - 1) &&
                  (unsigned)bz0 < (unsigned)(b.dims[2] - 1);
  const float4* records = reinterpret_cast<const float4*>(b.records);
  int nfill = 0;
  if (in) {
    const uint2* st = reinterpret_cast<const uint2*>(stencil) + (((rz & 3) * 4 + (ry & 3)) * 4 + (rx & 3)) * 8;
    uint4 h[8];
    if (b.hood != nullptr) {
      // one 128-byte line holds the masks and first records of all eight bricks (ClidBricks.hood):
      // three 256-bit loads from ONE line instead of eight 128-bit loads from eight lines
      const uint32_t* hl = b.hood + (((int64_t)bz0 * D1 + by0) * D0 + bx0) * 32;
      uint32_t ma[8], mb[8], bs[8];
      ldg256(hl, ma);
      ldg256(hl + 8, mb);
      ldg256(hl + 16, bs);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        h[s] = make_uint4(ma[2 * s], ma[2 * s + 1], bs[s], 0u);
        h[s + 4] = make_uint4(mb[2 * s], mb[2 * s + 1], bs[s + 4], 0u);
      }
    } else {
      const uint4* h0 = reinterpret_cast<const uint4*>(b.headers) + ((int64_t)bz0 * D1 + by0) * D0 + bx0;
      const int sy = D0, sz = D0 * D1;
#pragma unroll
      for (int s = 0; s < 8; ++s) h[s] = __ldg(h0 + (s & 1) + ((s >> 1) & 1) * sy + (s >> 2) * sz);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const uint2 sten = st[s];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t occ = half ? h[s].y : h[s].x;
        const uint32_t want = occ & (half ? sten.y : sten.x);
        if (want) {
          const int base = (int)h[s].z + (half ? __popc(h[s].x) : 0);
          col[nfill * kStride] = want;
          col[(kHalfSlots + nfill) * kStride] = occ;
          col[(2 * kHalfSlots + nfill) * kStride] = (uint32_t)base;
          ++nfill;
          // the records of a half-brick are contiguous: pull their first and last line towards L2
#if CLID_PF_RECORDS
          prefetch_l2(records + base);
          prefetch_l2(records + base + __popc(occ) - 1);
#endif
        }
      }
    }
  }
  // cursor over the filled slots: a pointer into the lane's column and the number of slots left
  int count = 0, left = nfill;
  const uint32_t* sp = col;
  uint32_t w = 0, occ = 0;
  int base = 0;
  if (nfill > 0) { w = sp[0]; occ = sp[kHalfSlots * kStride]; base = (int)sp[2 * kHalfSlots * kStride]; }
  // pop up to kWalkBatch candidates and issue their record loads
  auto fetch = [&](int (&rec)[kWalkBatch], float4 (&r)[kWalkBatch]) {
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) {
      rec[j] = -1;
      if (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        rec[j] = base + __popc(occ & ((1u << bit) - 1u));
        if (w == 0 && left > 1) {
          --left;
          sp += kStride;
          w = sp[0]; occ = sp[kHalfSlots * kStride]; base = (int)sp[2 * kHalfSlots * kStride];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) r[j] = __ldg(records + (rec[j] < 0 ? 0 : rec[j]));
  };
  auto rank = [&](const int (&rec)[kWalkBatch], const float4 (&r)[kWalkBatch]) {
#pragma unroll
    for (int j = 0; j < kWalkBatch; ++j) {
      const float d2 = dist2_torch(r[j].x - px, r[j].y - py, r[j].z - pz);
      if (rec[j] >= 0 && !(d2 > m.max_valid_dist2)) {
        ++count;
        if (d2 < top.d[K - 1]) {
#if CLID_PF_FEATURES
          prefetch_l2(m.gather_features + (int64_t)__float_as_int(r[j].w) * kFeat);  // likely neighbour
#endif
          top.insert(d2, rec[j]);
        }
      }
    }
  };
  if constexpr (kStaged) {
  // Staged walk: a lane pops up to kStage candidates and hands every record to the async copy unit
  // (cp.async 16 B, global -> this lane's column of shared memory: no destination registers, so all of them are
  // in flight at once), waits ONCE for the whole chunk and ranks it out of shared memory in batches of four.
  // A neighbourhood of ~19 candidates costs one or two global round trips instead of four or five.
  constexpr int kSt = kStaged ? kStage : 4;
  while (__any_sync(0xffffffffu, w != 0)) {
    int rec[kSt];
#pragma unroll
    for (int j = 0; j < kSt; ++j) {
      rec[j] = -1;
      if (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        rec[j] = base + __popc(occ & ((1u << bit) - 1u));
        cp_async16(reinterpret_cast<float*>(stg + j * kStride), reinterpret_cast<const float*>(records + rec[j]));
        if (w == 0 && left > 1) {
          --left;
          sp += kStride;
          w = sp[0]; occ = sp[kHalfSlots * kStride]; base = (int)sp[2 * kHalfSlots * kStride];
        }
      }
    }
    cp_async_wait_all();
#pragma unroll
    for (int j0 = 0; j0 < kSt; j0 += 4) {
      if (!__any_sync(0xffffffffu, rec[j0] >= 0)) break;  // slots fill in order: nobody has a candidate from j0 on
      float4 r[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) r[jj] = stg[(j0 + jj) * kStride];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj;
        const float d2 = dist2_torch(r[jj].x - px, r[jj].y - py, r[jj].z - pz);
        if (rec[j] >= 0 && !(d2 > m.max_valid_dist2)) {
          ++count;
          if (d2 < top.d[K - 1]) {
#if CLID_PF_FEATURES
            prefetch_l2(m.gather_features + (int64_t)__float_as_int(r[jj].w) * kFeat);
#endif
            top.insert(d2, rec[j]);
          }
        }
      }
    }
  }
  return count;
  }
#if CLID_WALK_PIPELINE
  // software pipeline of depth two: the loads of the next batch are in flight while this one is ranked
  // (the exposed wait on the record loads was 11 % of the kernel's stall samples, profiles/r2a_*)
  if (__any_sync(0xffffffffu, w != 0)) {
    int recA[kWalkBatch], recB[kWalkBatch];
    float4 rA[kWalkBatch], rB[kWalkBatch];
    fetch(recA, rA);
    while (true) {
      const bool moreB = __any_sync(0xffffffffu, w != 0);
      if (moreB) fetch(recB, rB);
      rank(recA, rA);
      if (!moreB) break;
      const bool moreA = __any_sync(0xffffffffu, w != 0);
      if (moreA) fetch(recA, rA);
      rank(recB, rB);
      if (!moreA) break;
    }
  }
#else
  while (__any_sync(0xffffffffu, w != 0)) {
    int rec[kWalkBatch];
    float4 r[kWalkBatch];
    fetch(rec, r);
    rank(rec, r);
  }
#endif
  return count;
}

}  // namespace clid
