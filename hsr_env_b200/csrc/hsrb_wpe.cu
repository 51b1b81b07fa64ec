// Warp-per-environment action kernel (hsrb_wpe.cuh): launch helpers.
#define HSR_COMPACT 1
#define HSRB_WPE_IMPL 1
#include "hsrb_wpe.cuh"

cudaError_t hsrb_wpe_prepare(size_t smem, int threads, int* bps) {
  cudaError_t e = cudaFuncSetAttribute(hsrb_wpe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_wpe_kernel, threads, smem);
}

cudaError_t hsrb_wpe_launch(const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s) {
  hsrb_wpe_kernel<<<grid, threads, smem, s>>>(a, f);
  return cudaGetLastError();
}
