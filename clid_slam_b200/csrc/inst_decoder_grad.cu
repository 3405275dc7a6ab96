// decoder_grad_kernel instantiations and launch geometry.
#include "launch.h"
#include "decoder_grad.cuh"
#include "mlp_l2.cuh"

namespace clid {

template <int H>
static int launch_decoder_grad_t(const DecoderGradParams& p, int sm_count, cudaStream_t stream) {
  auto kern = decoder_grad_kernel<H>;
  const int64_t tiles = p.n_rows >> 5;
  const int64_t want = (tiles + DgSmem<H>::kWarps - 1) / DgSmem<H>::kWarps;
  const int grid = (int)(want < sm_count ? want : sm_count);
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DgSmem<H>::kBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decoder_grad_kernel)");
    configured = true;
  }
  kern<<<grid, DgSmem<H>::kWarps * 32, DgSmem<H>::kBytes, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "decoder_grad_kernel launch");
  return CLID_OK;
}

int launch_decoder_grad_l2(const DecoderGradL2Params& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  if (p.n_rows == 0) return CLID_OK;
  if (p.dec.hidden_dim != 32 || p.dec.levels != 2)
    return set_error(CLID_EUNSUPPORTED, "decoder_grad_l2_kernel is compiled for 32 x 2 decoders; got %d x %d", p.dec.hidden_dim, p.dec.levels);
  const int64_t want = (p.n_rows + 7) / 8;
  const int grid = (int)(want < info.sm_count ? want : info.sm_count);
  decoder_grad_l2_kernel<32><<<grid, 256, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "decoder_grad_l2_kernel launch");
  return CLID_OK;
}

int launch_decoder_grad(const DecoderGradParams& p, cudaStream_t stream) {
  DeviceInfo info;
  if (int rc = device_info(&info)) return rc;
  const int64_t tiles = p.n_rows >> 5;
  if (tiles == 0) return CLID_OK;
  const int H = p.dec.hidden_dim;
  if (H == 64) return launch_decoder_grad_t<64>(p, info.sm_count, stream);
  if (H == 32) return launch_decoder_grad_t<32>(p, info.sm_count, stream);
  if (H == 128) return launch_decoder_grad_t<128>(p, info.sm_count, stream);
  return set_error(CLID_EUNSUPPORTED, "decoder_grad_kernel is compiled for H in {32,64,128}; got %d", H);
}

}  // namespace clid
