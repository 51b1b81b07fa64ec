// Fast-path action kernel instantiations (hsrb_fast.cuh): 8 lanes per environment, NV = 8 (one block) and 2 (none).
#define HSR_COMPACT 1
#include "hsrb_fast.cuh"

template <int NV>
static cudaError_t prepare_t(size_t smem, int threads, int* bps) {
  cudaError_t e = cudaFuncSetAttribute(hsrb_fast_kernel<8, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_fast_kernel<8, NV>, threads, smem);
}

cudaError_t hsrb_fast_prepare(int nv, size_t smem, int threads, int* bps) {
  return nv == 8 ? prepare_t<8>(smem, threads, bps) : prepare_t<2>(smem, threads, bps);
}

cudaError_t hsrb_fast_launch(int nv, const KArgs& a, const FastInfo& f, int grid, int threads, size_t smem, cudaStream_t s) {
  if (nv == 8) hsrb_fast_kernel<8, 8><<<grid, threads, smem, s>>>(a, f);
  else hsrb_fast_kernel<8, 2><<<grid, threads, smem, s>>>(a, f);
  return cudaGetLastError();
}
