// TEST INFRASTRUCTURE: a minimal SIMT emulator so that the CUDA kernels of hsr_env_b200/csrc can be executed on a
// CPU-only box (this container has nvcc but no GPU).  The kernel source is compiled unchanged by g++ with
//     -D__CUDACC__ -DHSRB_SIMT_EMU -include tests/simt_emu/simt_emu.h
// Every CUDA thread of a block becomes one std::thread; warp intrinsics (__shfl*_sync, __ballot_sync, __syncwarp)
// are exchanges through a per-warp slot array bracketed by a barrier over the lanes named in the member mask, which
// is exactly their semantics for converged callers.  Blocks run one after the other.  It checks logic and
// arithmetic (modulo FMA contraction / approximate intrinsics), not performance.  Never linked into the product.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __align__(n) alignas(n)

struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };

namespace emu {

struct SpinBarrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit SpinBarrier(int n_) : n(n_) {}
  void wait() {
    int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_acq_rel);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 64) std::this_thread::yield();
    }
  }
};

struct Warp {
  uint64_t slot[32];
  std::mutex mu;
  std::map<unsigned, std::unique_ptr<SpinBarrier>> bars;
  SpinBarrier& bar(unsigned mask) {
    std::lock_guard<std::mutex> l(mu);
    auto it = bars.find(mask);
    if (it == bars.end()) it = bars.emplace(mask, std::make_unique<SpinBarrier>(__builtin_popcount(mask))).first;
    return *it->second;
  }
};

struct Block {
  std::vector<std::unique_ptr<Warp>> warps;
  std::unique_ptr<SpinBarrier> all;
  std::atomic<int> vote{0};
  // named barriers (bar.sync id, nthreads): one barrier object + vote counter per id
  std::mutex nmu;
  std::map<int, std::unique_ptr<SpinBarrier>> named;
  std::atomic<int> nvote[16];
  SpinBarrier& nbar(int id, int n) {
    std::lock_guard<std::mutex> l(nmu);
    auto it = named.find(id);
    if (it == named.end()) it = named.emplace(id, std::make_unique<SpinBarrier>(n)).first;
    return *it->second;
  }
};

inline thread_local Block* tl_block = nullptr;
inline thread_local Warp* tl_warp = nullptr;
inline thread_local int tl_lane = 0;
inline unsigned char* dyn_smem = nullptr;

inline void named_barrier(int id, int nthreads) { tl_block->nbar(id, nthreads).wait(); }
// bar.red.and / .or over the threads of a named barrier
inline bool named_vote(int id, int nthreads, bool p, bool is_and) {
  Block& b = *tl_block;
  SpinBarrier& bar = b.nbar(id, nthreads);
  bar.wait();
  b.nvote[id].store(0);
  bar.wait();
  if (is_and ? !p : p) b.nvote[id].fetch_add(1);
  bar.wait();
  return is_and ? b.nvote[id].load() == 0 : b.nvote[id].load() != 0;
}

}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::Warp& w = *emu::tl_warp;
  int lane = emu::tl_lane;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu::SpinBarrier& b = w.bar(mask);
  w.slot[lane] = bits;
  b.wait();
  int s = (lane & ~(width - 1)) + (src & (width - 1));
  T r;
  std::memcpy(&r, &w.slot[s], sizeof(T));
  b.wait();
  return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  return __shfl_sync(mask, v, (emu::tl_lane ^ lanemask) & (width - 1), width);
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  for (int o = 16; o > 0; o >>= 1) { unsigned t = __shfl_xor_sync(mask, v, o); v = t > v ? t : v; }
  return v;
}
inline int __reduce_min_sync(unsigned mask, int v) {
  for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(mask, v, o); v = t < v ? t : v; }
  return v;
}
inline unsigned __ballot_sync(unsigned mask, bool p) {
  emu::Warp& w = *emu::tl_warp;
  emu::SpinBarrier& b = w.bar(mask);
  w.slot[emu::tl_lane] = p ? 1 : 0;
  b.wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (((mask >> i) & 1u) && w.slot[i]) r |= 1u << i;
  b.wait();
  return r;
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) == mask; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::tl_warp->bar(mask).wait(); }
inline void __syncthreads() { emu::tl_block->all->wait(); }
inline int __syncthreads_or(int p) {
  emu::Block& b = *emu::tl_block;
  b.all->wait();
  if (threadIdx.x == 0) b.vote.store(0);
  b.all->wait();
  if (p) b.vote.fetch_add(1);
  b.all->wait();
  return b.vote.load() != 0;
}
inline int __syncthreads_and(int p) {
  emu::Block& b = *emu::tl_block;
  b.all->wait();
  if (threadIdx.x == 0) b.vote.store(0);
  b.all->wait();
  if (!p) b.vote.fetch_add(1);
  b.all->wait();
  return b.vote.load() == 0;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline long long clock64() { return 0; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }

namespace emu {

// Run `kernel(args...)` for a grid of `grid` blocks of `threads` threads with `smem_bytes` of dynamic shared memory.
template <typename F>
void launch(int grid, int threads, size_t smem_bytes, F&& body) {
  std::vector<unsigned char> smem(smem_bytes + 64);
  dyn_smem = smem.data() + (16 - (reinterpret_cast<uintptr_t>(smem.data()) & 15)) % 16;
  blockDim.x = (unsigned)threads;
  gridDim.x = (unsigned)grid;
  for (int b = 0; b < grid; b++) {
    Block blk;
    int nw = (threads + 31) / 32;
    for (int w = 0; w < nw; w++) blk.warps.emplace_back(std::make_unique<Warp>());
    blk.all = std::make_unique<SpinBarrier>(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        tl_block = &blk;
        tl_warp = blk.warps[t / 32].get();
        tl_lane = t % 32;
        threadIdx.x = (unsigned)t;
        blockIdx.x = (unsigned)b;
        body();
      });
    for (auto& x : th) x.join();
  }
}

}  // namespace emu
