// Dev micro-benchmark: issue rates of FFMA, FFMA2 (fma.rn.f32x2) and legacy mma.sync TF32 on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/ub/ub_pipes.cu -o scripts/ub/ub_pipes
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float x = out[0], y = out[1];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  float2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
  float2 x = make_float2(out[0], out[1]), y = make_float2(out[1], out[0]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __ffma2_rn(a[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {__float_as_uint(out[0]), __float_as_uint(out[1]), __float_as_uint(out[2]), __float_as_uint(out[3])};
  unsigned b[2] = {__float_as_uint(out[4]), __float_as_uint(out[5])};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_tf32(c[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void k_mma_bf16(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {__float_as_uint(out[0]), __float_as_uint(out[1]), __float_as_uint(out[2]), __float_as_uint(out[3])};
  unsigned b[2] = {__float_as_uint(out[4]), __float_as_uint(out[5])};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_bf16(c[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float run(F f, int blocks, int threads, float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f<<<blocks, threads>>>(d, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); f<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4); cudaMemset(d, 0, 148 * 8 * 1024 * 4);
  int iters = 4096;
  for (int wps : {1, 2, 4, 8}) {  // warps per SMSP
    int threads = wps * 4 * 32, blocks = 148;
    double clk = 1.965e9;
    float t1 = run(k_ffma, blocks, threads, d, iters);
    float t2 = run(k_ffma2, blocks, threads, d, iters);
    float t3 = run(k_mma, blocks, threads, d, iters);
    float t4 = run(k_mma_bf16, blocks, threads, d, iters);
    // cycles per warp-instruction per SMSP
    printf("warps/SMSP %d: FFMA %.2f cyc/inst  FFMA2 %.2f cyc/inst  mma.tf32.m16n8k8 %.2f cyc/inst  mma.bf16.m16n8k16 %.2f cyc/inst\n", wps,
           t1 * 1e-3 * clk / (iters * 16.0 * wps), t2 * 1e-3 * clk / (iters * 16.0 * wps), t3 * 1e-3 * clk / (iters * 8.0 * wps),
           t4 * 1e-3 * clk / (iters * 8.0 * wps));
  }
  return 0;
}
