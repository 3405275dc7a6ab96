// One-shot all-reduce of a small vector ([decoder gradients | loss], ~3 kB) over peer-mapped memory, and the
// device-side barrier that comes with it (include/clid_sdf.h ClidPeerArgs).  The reference is single-GPU; this is
// the collective of the sample-/space-sharded mapping step, written against NVLink peer memory
// instead of NCCL because the payload is latency-bound: one kernel pushes the vector into a slot on every rank and
// raises a flag there, one kernel waits for all flags and sums the slots in rank order.  Both run inside the step's
// CUDA graph; nothing goes through the host.
#pragma once
#include "common.cuh"

namespace clid {

#ifdef CLID_PLAIN_KERNELS
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_publish_kernel(const ClidPeerArgs a, const float* __restrict__ src0,
                                                           const float* __restrict__ src1) {
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) epoch = *a.epoch + 1u;
  const int n = a.n0 + a.n1;
  for (int r = 0; r < a.world; ++r) {
    float* dst = a.slots_of[r] + (int64_t)a.rank * a.stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = i < a.n0 ? src0[i] : src1[i - a.n0];
  }
  __threadfence_system();  // the slot contents (and every earlier write of this stream, e.g. the remote gradient adds
  __syncthreads();         // of clid_train_fused) are visible system-wide before any flag is
  if (threadIdx.x < a.world) st_release_sys(a.flags_of[threadIdx.x] + a.rank, epoch);
  if (threadIdx.x == 0) *a.epoch = epoch;
}

__global__ void __launch_bounds__(256) peer_reduce_kernel(const ClidPeerArgs a, float* __restrict__ dst0, float* __restrict__ dst1) {
  __shared__ int failed;
  if (threadIdx.x == 0) failed = 0;
  __syncthreads();
  const uint32_t epoch = *a.epoch;  // set by this rank's publish of the same step (stream order)
  if (threadIdx.x < a.world) {
    const uint32_t* flag = a.flags_of[a.rank] + threadIdx.x;
    const long long t0 = clock64();
    const long long budget = (long long)(a.timeout_ms > 0 ? a.timeout_ms : 2000) * 2000000ll;  // ~2 GHz
    // epochs only grow; a peer that is a step ahead already shows epoch + 1
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
      if (clock64() - t0 > budget) { failed = 1; break; }
    }
  }
  __syncthreads();
  if (failed) {
    if (threadIdx.x == 0 && a.error) *a.error = 1;
    return;  // leave dst untouched: the caller checks *error
  }
  const float* slots = a.slots_of[a.rank];
  const int n = a.n0 + a.n1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < a.world; ++r) s += __ldcv(slots + (int64_t)r * a.stride + i);  // rank order: identical sums everywhere
    if (i < a.n0) dst0[i] = s;
    else dst1[i - a.n0] = s;
  }
}
#endif

}  // namespace clid
