// Warp-per-environment action kernel (hsrb_wpe.cuh): launch helpers.
#define HSR_COMPACT 1
#define HSRB_WPE_IMPL 1
#include "hsrb_wpe.cuh"

cudaError_t hsrb_wpe_prepare(size_t smem, int threads, int* bps) {
  cudaError_t e = cudaFuncSetAttribute(hsrb_wpe_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hsrb_wpe_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_wpe_kernel_t<false>, threads, smem);
}

cudaError_t hsrb_wpe_launch(const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s, bool lock) {
  if (lock) hsrb_wpe_kernel_t<true><<<grid, threads, smem, s>>>(a, f);
  else hsrb_wpe_kernel_t<false><<<grid, threads, smem, s>>>(a, f);
  return cudaGetLastError();
}

#if defined(WPE_CHAIN_CLOCKS)
// experiment hook (not part of include/hsrb.h): out[0..31] = cycles, out[32..63] = occurrences per chain section; resets
extern "C" int hsrb_debug_chain_clocks(unsigned long long* out) {
  unsigned long long z[32] = {0};
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, wpe::wpe_ck_cyc, sizeof(z)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out + 32, wpe::wpe_ck_cnt, sizeof(z)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(wpe::wpe_ck_cyc, z, sizeof(z)); cudaMemcpyToSymbol(wpe::wpe_ck_cnt, z, sizeof(z));
  return 0;
}
#endif
