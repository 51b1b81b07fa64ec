// Instruction-fetch microbenchmark for sm_100a: cycles per instruction of a loop whose body is a straight line of
// BODY independent FFMAs (8 accumulators), as a function of the body size and of the number of warps per SM;
// a second variant takes a short forward branch every 32 instructions.  Used to size the action kernel's hot path
// (DESIGN.md "instruction footprint").   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

#define F8 "fma.rn.f32 %0,%0,%8,%9;\n\tfma.rn.f32 %1,%1,%8,%9;\n\tfma.rn.f32 %2,%2,%8,%9;\n\tfma.rn.f32 %3,%3,%8,%9;\n\t" \
           "fma.rn.f32 %4,%4,%8,%9;\n\tfma.rn.f32 %5,%5,%8,%9;\n\tfma.rn.f32 %6,%6,%8,%9;\n\tfma.rn.f32 %7,%7,%8,%9;\n\t"
#define F32 F8 F8 F8 F8
#define ASM32 asm volatile(F32 : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));
// 24 FFMAs, then a branch that skips 8 more (always taken at run time, not provable at compile time)
#define ASMBR asm volatile("{\n\t.reg .pred p;\n\t" F8 F8 F8 "setp.neu.f32 p, %8, %9;\n\t@p bra SKIP;\n\t" F8 "SKIP:\n\t}" \
                           : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));
#define R4(X) X X X X
#define R16(X) R4(R4(X))
#define R64(X) R4(R16(X))
#define R256(X) R4(R64(X))

template <int KB, bool BR>
__global__ void body(float* out, long long* cyc, int iters, float m, float c) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  long long t0 = 0;
  for (int it = 0; it < iters + 1; it++) {
    if (it == 1) t0 = clock64();   // first pass warms the caches
    // KB kilobytes of code = KB*64 instructions = KB*2 blocks of 32
    if (KB >= 8) { if (BR) { R16(ASMBR) } else { R16(ASM32) } }
    if (KB >= 16) { if (BR) { R16(ASMBR) } else { R16(ASM32) } }
    if (KB >= 32) { if (BR) { R16(ASMBR) R16(ASMBR) } else { R16(ASM32) R16(ASM32) } }
    if (KB >= 64) { if (BR) { R64(ASMBR) } else { R64(ASM32) } }
    if (KB >= 128) { if (BR) { R64(ASMBR) R64(ASMBR) } else { R64(ASM32) R64(ASM32) } }
    if (KB >= 256) { if (BR) { R256(ASMBR) } else { R256(ASM32) } }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int KB, bool BR>
void run(int warps) {
  float* out; long long* cyc;
  int grid = 148, threads = 32 * warps, iters = 200;
  cudaMalloc(&out, sizeof(float) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid * warps);
  body<KB, BR><<<grid, threads>>>(out, cyc, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  static long long h[148 * 32];
  cudaMemcpy(h, cyc, sizeof(long long) * grid * warps, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < grid * warps; i++) s += (double)h[i];
  s /= grid * warps;
  const double instr = (double)KB * 64 * iters * (BR ? 0.78 : 1.0);  // executed instructions per warp (branch variant skips 8 of 34)
  printf("{\"body_kb\": %d, \"branchy\": %d, \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.3f, \"err\": \"%s\"}\n", KB, (int)BR, warps,
         s / instr, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  int ws[4] = {1, 4, 8, 16};
  for (int k = 0; k < 4; k++) {
    int w = ws[k];
    run<8, false>(w); run<16, false>(w); run<32, false>(w); run<64, false>(w); run<128, false>(w); run<256, false>(w);
    run<8, true>(w); run<16, true>(w); run<32, true>(w); run<64, true>(w); run<128, true>(w); run<256, true>(w);
  }
  return 0;
}
