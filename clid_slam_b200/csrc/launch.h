// Host-side launch plumbing shared by the translation units of libclid_sdf.so.  The template
// kernels are instantiated in separate inst_*.cu files so that nvcc can compile them in parallel
// (clid_slam_b200/build.py); api.cu holds the C ABI, the argument validation and the plain kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "clid_sdf.h"

namespace clid {

struct QueryParams;
struct TrainBwdParams;
struct TrainFusedParams;
struct QueryBwdParams;
struct DecoderGradParams;
struct DecoderGradL2Params;

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

struct DeviceInfo {
  int sm_count = 0;
};
int device_info(DeviceInfo* info);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// inst_query_{bricks,hashed}.cu
int dispatch_query_bricks(const QueryParams& p, bool has_dec, cudaStream_t stream);
int dispatch_query_hashed(const QueryParams& p, bool has_dec, cudaStream_t stream);
// inst_query_tc.cu: 64 x 1 decoder on the tensor cores, brick index (CLID_TC_DECODER)
int dispatch_query_tc(const QueryParams& p, cudaStream_t stream);
// inst_query_bwd.cu
int launch_query_backward_first(const QueryBwdParams& p, int grid, cudaStream_t stream);
int launch_query_backward_second(const QueryBwdParams& p, int grid, cudaStream_t stream);
// inst_train_bwd.cu
int dispatch_train_backward(const TrainBwdParams& p, cudaStream_t stream);
// inst_fused_{bricks,hashed}.cu
int dispatch_train_fused_bricks(const TrainFusedParams& p, cudaStream_t stream);
int dispatch_train_fused_hashed(const TrainFusedParams& p, cudaStream_t stream);

// inst_decoder_grad.cu
int launch_decoder_grad(const DecoderGradParams& p, cudaStream_t stream);
int launch_decoder_grad_l2(const DecoderGradL2Params& p, cudaStream_t stream);

}  // namespace clid
