// Phase-locked general action kernel (hsrb_kernels.cuh: hsrb_step_lock_kernel): launch helpers.
#define HSRB_STEP_LOCK_IMPL 1
#include "hsrb_kernels.cuh"

cudaError_t hsrb_prepare_step_lock(size_t smem, int threads, int* bps) {
  cudaError_t e = cudaFuncSetAttribute(hsrb_step_lock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_step_lock_kernel, threads, smem);
}

cudaError_t hsrb_launch_step_lock(const KArgs& a, int grid, int threads, size_t smem, cudaStream_t s) {
  hsrb_step_lock_kernel<<<grid, threads, smem, s>>>(a);
  return cudaGetLastError();
}
