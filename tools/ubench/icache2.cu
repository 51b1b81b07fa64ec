// Instruction-fetch microbenchmark 2 (sm_100a): W warps per SM loop over the SAME straight-line body of KB kilobytes,
// but DESYNCHRONISED: warp w enters the loop after a delay of w/W of one pass, so at any time the warps fetch W different
// places of the body (what free-running warps of a big kernel do).  Reports cycles per instruction per warp and the SM's
// aggregate IPC.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache2 icache2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define F8 "fma.rn.f32 %0,%0,%8,%9;\n\tfma.rn.f32 %1,%1,%8,%9;\n\tfma.rn.f32 %2,%2,%8,%9;\n\tfma.rn.f32 %3,%3,%8,%9;\n\t" \
           "fma.rn.f32 %4,%4,%8,%9;\n\tfma.rn.f32 %5,%5,%8,%9;\n\tfma.rn.f32 %6,%6,%8,%9;\n\tfma.rn.f32 %7,%7,%8,%9;\n\t"
#define F32 F8 F8 F8 F8
#define ASM32 asm volatile(F32 : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));
#define R4(X) X X X X
#define R16(X) R4(R4(X))
#define R64(X) R4(R16(X))
#define R256(X) R4(R64(X))
template <int KB>
__global__ void body(float* out, long long* cyc, int iters, float m, float c, long long stagger) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const int w = threadIdx.x / 32;
  if (stagger > 0) { long long t = clock64(); while (clock64() - t < stagger * w) { } }
  long long t0 = 0;
  for (int it = 0; it < iters + 1; it++) {
    if (it == 1) t0 = clock64();
    if (KB >= 8) { R16(ASM32) }
    if (KB >= 16) { R16(ASM32) }
    if (KB >= 32) { R16(ASM32) R16(ASM32) }
    if (KB >= 64) { R64(ASM32) }
    if (KB >= 128) { R64(ASM32) R64(ASM32) }
    if (KB >= 256) { R256(ASM32) }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + w] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int KB>
void run(int warps, bool stag) {
  float* out; long long* cyc;
  int grid = 148, threads = 32 * warps, iters = 100;
  cudaMalloc(&out, sizeof(float) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid * warps);
  // one pass of the body takes about KB*64*cpi cycles; spread the warps over it (cpi guess 4)
  long long stagger = stag ? (long long)KB * 64 * 4 / warps + 97 : 0;
  body<KB><<<grid, threads>>>(out, cyc, iters, 1.0001f, 0.5f, stagger);
  cudaDeviceSynchronize();
  static long long h[148 * 32];
  cudaMemcpy(h, cyc, sizeof(long long) * grid * warps, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < grid * warps; i++) s += (double)h[i];
  s /= grid * warps;
  const double instr = (double)KB * 64 * iters;
  printf("{\"body_kb\": %d, \"staggered\": %d, \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.3f, \"sm_ipc\": %.3f}\n", KB, (int)stag, warps,
         s / instr, warps * instr / s);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  int ws[4] = {4, 8, 16, 28};
  for (int st = 0; st < 2; st++)
    for (int k = 0; k < 4; k++) {
      int w = ws[k];
      run<16>(w, st); run<32>(w, st); run<64>(w, st); run<128>(w, st); run<256>(w, st);
    }
  return 0;
}
