// Action-kernel instantiations of hsrb_push.cuh: G = 8 / 16 / 32 lanes per environment, NV = 8 (one block) and 2 (none).
#define HSR_COMPACT 1
#include "hsrb_push.cuh"

template <int G, int NV>
static cudaError_t prepare_t(size_t smem, int threads, int* bps) {
  cudaError_t e = cudaFuncSetAttribute(hsrb_push_kernel<G, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per function, not per handle: handles with different layouts share it
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_push_kernel<G, NV>, threads, smem);
}

template <int G>
static cudaError_t prepare_g(int nv, size_t smem, int threads, int* bps) {
  return nv == 8 ? prepare_t<G, 8>(smem, threads, bps) : prepare_t<G, 2>(smem, threads, bps);
}

cudaError_t hsrb_push_prepare(int G, int nv, size_t smem, int threads, int* bps) {
  switch (G) {
    case 8: return prepare_g<8>(nv, smem, threads, bps);
    case 16: return prepare_g<16>(nv, smem, threads, bps);
    default: return prepare_g<32>(nv, smem, threads, bps);
  }
}

template <int G>
static void launch_g(int nv, const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s) {
  if (nv == 8) hsrb_push_kernel<G, 8><<<grid, threads, smem, s>>>(a, f);
  else hsrb_push_kernel<G, 2><<<grid, threads, smem, s>>>(a, f);
}

cudaError_t hsrb_push_launch(int G, int nv, const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s) {
  switch (G) {
    case 8: launch_g<8>(nv, a, f, grid, threads, smem, s); break;
    case 16: launch_g<16>(nv, a, f, grid, threads, smem, s); break;
    default: launch_g<32>(nv, a, f, grid, threads, smem, s); break;
  }
  return cudaGetLastError();
}
