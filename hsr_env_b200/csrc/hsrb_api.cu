// libhsrb.so - C ABI of the batched physics backend (include/hsrb.h).
//
// Host side of the action kernel: owns the device copy of the model blob, the resident per-environment
// state [N, S] (qpos | qvel | qacc_warmstart | mocap_pos), the episode counters that key the Philox reset
// streams, and the launch configuration.  Everything is asynchronous on the caller's stream; no entry point
// throws; errors come back as negative codes with a thread-local message.
//
// Replaces mujoco_py.load_model_from_path / MjSim / sim.step / sim.reset / sim.forward / get_state /
// set_state as used by /root/reference/hsr/mujoco_env.py:33-34,83-94 and /root/reference/hsr/env.py:115-177.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/hsrb.h"
#include "hsrb_wpe.cuh"

cudaError_t hsrb_push_prepare(int G, int nv, size_t smem, int threads, int* bps);
cudaError_t hsrb_push_launch(int G, int nv, const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s);
cudaError_t hsrb_wpe_prepare(size_t smem, int threads, int* bps);
cudaError_t hsrb_wpe_launch(const KArgs& a, const PushInfo& f, int grid, int threads, size_t smem, cudaStream_t s, bool lock);

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail(-2, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

}  // namespace

struct hsrb {
  HostModel<float> hm;      // host copy (pointers into hm.buf)
  ModelT<float> dm;         // same table with device pointers
  unsigned char* d_model = nullptr;
  EnvCfg<float> cfg;
  int n = 0, device = 0, S = 0;
  unsigned long long seed = 0, env_off = 0;
  float* d_state = nullptr;
  unsigned* d_episode = nullptr;
  unsigned long long* d_stats = nullptr;
  // launch configuration
  int lanes = 0, lanes_req = 0;
  unsigned ws_bytes = 0;
  int blocks_per_sm = 0, grid = 0, num_sm = 0;
  bool configured = false;
  // scratch for the host-buffer entry point
  float *d_ctrl = nullptr, *d_obs = nullptr, *d_reward = nullptr;
  unsigned char* d_done = nullptr;
  int* d_taken = nullptr;
  long long launches = 0;
  // fast path (hsrb_push.cuh): sliding base + at most one free box
  PushInfo fast;
  PushTables fast_tab;
  unsigned char* d_fast_tab = nullptr;
  bool fast_ok = false;
  int path = 0;             // 0 auto, 1 general kernel, 2 fast kernel
  bool fast_configured = false;
  int fast_lanes = 8, fast_threads = 0, fast_grid = 0, fast_bps = 0;
  unsigned fast_ws = 0;
  char fast_why[128] = "";
  // phase-locked general kernel (hsrb_step_lock_kernel): STEP launches of the general path
  bool lock_configured = false;
  int lock_threads = 0, lock_grid = 0, lock_bps = 0;
  // warp-per-environment kernel of the same family (hsrb_wpe.cuh)
  bool wpe_ok = false, wpe_configured = false;
  int wpe_threads = 0, wpe_grid = 0, wpe_bps = 0;
  unsigned wpe_ws = 0;
  // work-sorted launch order of the wpe kernel: the kernel writes a work estimate per environment, the next launch takes
  // the environments heaviest first (slot r -> block r % grid, warp r / grid: the heaviest warps of every block form team 0)
  int *d_work = nullptr, *d_work_sorted = nullptr, *d_iota = nullptr, *d_order = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  bool order_valid = false;
};

namespace {

__global__ void fill_iota(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

cudaError_t prepare(int G, size_t smem, int* bps) {
  switch (G) {
    case 4: return hsrb_prepare_step_4(smem, bps);
    case 8: return hsrb_prepare_step_8(smem, bps);
    case 16: return hsrb_prepare_step_16(smem, bps);
    default: return hsrb_prepare_step_32(smem, bps);
  }
}
cudaError_t launch(int G, const KArgs& a, int grid, size_t smem, cudaStream_t s) {
  switch (G) {
    case 4: return hsrb_launch_step_4(a, grid, smem, s);
    case 8: return hsrb_launch_step_8(a, grid, smem, s);
    case 16: return hsrb_launch_step_16(a, grid, smem, s);
    default: return hsrb_launch_step_32(a, grid, smem, s);
  }
}
cudaError_t launch_aux(int G, const KArgs& a, int grid, size_t smem, cudaStream_t s) {
  switch (G) {
    case 4: return hsrb_launch_aux_4(a, grid, smem, s);
    case 8: return hsrb_launch_aux_8(a, grid, smem, s);
    case 16: return hsrb_launch_aux_16(a, grid, smem, s);
    default: return hsrb_launch_aux_32(a, grid, smem, s);
  }
}

// Choose lanes-per-environment and the grid: as many lanes per env as keeps every environment resident in
// one wave (more lanes = shorter dependent chain per substep), otherwise fewer lanes / a grid-stride loop.
int configure(hsrb* h) {
  if (h->configured) return 0;
  h->ws_bytes = (unsigned)ws_carve<float>(h->dm, nullptr, nullptr);
  // lanes per environment: the widest layout that holds every environment in one wave; if none does, shared memory
  // decides how many environments an SM holds (one warp per block, 32/G workspaces per block), and among the layouts
  // within 10 % of the best residency the narrowest one wins: the single-lane stages (kinematics, bias forces, Euler)
  // then run for 32/G environments at once.  Measured on B200 (tools/sweep_lanes.py): C3 (nv = 13, 16384 envs)
  // G = 8: 3.5 M substeps/s against 2.6 M (G = 4, one block per SM) and 1.1 M (G = 32); C5 (nv = 26): G = 32 0.53 M
  // against 0.29 M (G = 16)
  int cands[4] = {32, 16, 8, 4};
  int best = 0;
  long long res[4] = {0, 0, 0, 0};
  int bpsv[4] = {0, 0, 0, 0};
  long long resmax = 0;
  for (int k = 0; k < 4; k++) {
    int G = cands[k];
    if (h->lanes_req && G != h->lanes_req) continue;
    size_t smem = (size_t)h->ws_bytes * (32 / G);
    if (smem > 227 * 1024) continue;
    int bps = 0;
    if (prepare(G, smem, &bps) != cudaSuccess || bps <= 0) { cudaGetLastError(); continue; }
    bpsv[k] = bps;
    res[k] = (long long)bps * h->num_sm * (32 / G);
    if (res[k] > resmax) resmax = res[k];
  }
  for (int k = 0; k < 4 && !best; k++)
    if (res[k] >= h->n) { best = cands[k]; h->blocks_per_sm = bpsv[k]; }
  for (int k = 3; k >= 0 && !best; k--)
    if (res[k] > 0 && res[k] * 10 >= resmax * 9) { best = cands[k]; h->blocks_per_sm = bpsv[k]; }
  if (!best) return fail(-3, "no launch configuration fits: workspace %u bytes per environment", h->ws_bytes);
  h->lanes = best;
  size_t smem = (size_t)h->ws_bytes * (32 / best);
  CU(prepare(best, smem, &h->blocks_per_sm));
  int gpb = 32 / best;
  int need = (h->n + gpb - 1) / gpb;
  int cap = h->blocks_per_sm * h->num_sm;
  h->grid = need < cap ? need : cap;
  if (h->grid < 1) h->grid = 1;
  h->configured = true;
  return 0;
}

int configure_fast(hsrb* h) {
  if (h->fast_configured) return 0;
  ModelT<float> mm = h->dm;
  h->fast_ws = (unsigned)push::carve(mm, nullptr, nullptr);
  if (const char* o = getenv("HSRB_PUSH_PAD")) h->fast_ws += 16u * (unsigned)atoi(o);   // experiments: stride of the environment slices
  // lanes per environment: the caller's choice (hsrb_config) when it is one this kernel has, else 8
  const int G = (h->lanes_req == 8 || h->lanes_req == 16 || h->lanes_req == 32) ? h->lanes_req : 8;
  h->fast_lanes = G;
  const int epw = 32 / G;
  // warps per block so that all environments are resident in one wave when they fit: n / (envs per warp) / SMs
  int warps_needed = (h->n + epw - 1) / epw;
  int wpb = (warps_needed + h->num_sm - 1) / h->num_sm;
  if (wpb < 1) wpb = 1;
  const int wpb_max = G == 16 ? 14 : 8;   // __launch_bounds__ of the kernel (hsrb_push.cuh)
  if (wpb == 7 && wpb_max >= 8) wpb = 8;   // two warps on each of the four schedulers: 128 blocks x 32 envs beat 147 x 28 (measured +2 %)
  if (wpb > wpb_max) wpb = wpb_max;
  if (const char* o = getenv("HSRB_PUSH_WPB")) { int v = atoi(o); if (v >= 1 && v <= wpb_max) wpb = v; }   // experiments
  // shared memory: at most 227 KB per block
  const size_t tail = push::shared_tail(mm);
  while (wpb > 1 && (size_t)h->fast_ws * (wpb * epw) + tail > 227 * 1024) wpb--;
  h->fast_threads = 32 * wpb;
  size_t smem = (size_t)h->fast_ws * (h->fast_threads / G) + tail;
  CU(hsrb_push_prepare(G, h->fast.nv, smem, h->fast_threads, &h->fast_bps));
  if (h->fast_bps < 1) return fail(-3, "fast kernel does not fit on an SM (%zu bytes of shared memory)", smem);
  int epb = h->fast_threads / G;
  int need = (h->n + epb - 1) / epb;
  int cap = h->fast_bps * h->num_sm;
  h->fast_grid = need < cap ? need : cap;
  h->fast_configured = true;
  return 0;
}

// the fast kernels test the env_wrapper-form goal only (one block against the goal point)
bool use_fast(const hsrb* h) { return h->fast_ok && h->path != 1 && h->cfg.ngoal == 0; }
// which of the two: path 3 = warp-per-environment kernel (hsrb_wpe.cuh), 2 = 8-lane lock-step kernel (hsrb_push.cuh);
// 0 (auto) = the warp-per-environment kernel, phase-locked, two teams per block - measured on B200, round 2, TimeLimit
// workload: 48.4 M substeps/s at 4096 envs (lock-step 8-lane kernel 34.6 M, free-running warps 32.9 M), 65.8 M at
// 131072 envs (52.3 M / 53.2 M); HSRB_FAST_KERNEL=push|wpe overrides (experiments)
bool use_wpe(const hsrb* h) {
  if (!use_fast(h) || !h->wpe_ok) return false;
  if (h->path == 3) return true;
  if (h->path == 2) return false;
  const char* o = getenv("HSRB_FAST_KERNEL");
  return !(o && o[0] == 'p');
}

int configure_wpe(hsrb* h) {
  if (h->wpe_configured) return 0;
  h->wpe_ws = (unsigned)wpe::slice_bytes();
  const size_t tail = wpe::shared_tail(h->dm);
  // one warp per environment; all environments resident in one wave when they fit (warps per block = n / SMs),
  // otherwise the most warps shared memory and the register file hold, and a grid-stride loop over the environments
  int wpb = (h->n + h->num_sm - 1) / h->num_sm;
  if (wpb < 1) wpb = 1;
  if (wpb > WPE_MAXWARPS) wpb = WPE_MAXWARPS;
  if (const char* o = getenv("HSRB_WPE_WPB")) { int v = atoi(o); if (v >= 1 && v <= WPE_MAXWARPS) wpb = v; }   // experiments
  while (wpb > 1 && (size_t)h->wpe_ws * wpb + tail > 227 * 1024) wpb--;
  h->wpe_threads = 32 * wpb;
  const size_t smem = (size_t)h->wpe_ws * wpb + tail;
  CU(hsrb_wpe_prepare(smem, h->wpe_threads, &h->wpe_bps));
  if (h->wpe_bps < 1) return fail(-3, "warp-per-environment kernel does not fit on an SM (%zu bytes of shared memory)", smem);
  const int need = (h->n + wpb - 1) / wpb;
  const int cap = h->wpe_bps * h->num_sm;
  h->wpe_grid = need < cap ? need : cap;
  if (wpb > 1 && h->wpe_grid < cap && 2 * h->wpe_grid > cap) h->wpe_grid = cap;   // 4096 envs: 148 blocks of 27-28 instead of 147 of 28 + an idle SM
  if (const char* o = getenv("HSRB_WPE_GRID")) { int v = atoi(o); if (v >= 1 && v < h->wpe_grid) h->wpe_grid = v; }   // experiments
  h->wpe_configured = true;
  return 0;
}

// STEP launches of the general path: one warp per environment, as many warps per block as shared memory holds
int configure_lock(hsrb* h) {
  if (h->lock_configured) return 0;
  h->ws_bytes = (unsigned)ws_carve<float>(h->dm, nullptr, nullptr);
  const size_t tail = lock_tail_bytes();
  int wpb = (int)((227 * 1024 - tail) / h->ws_bytes);
  if (wpb < 1) return fail(-3, "no launch configuration fits: workspace %u bytes per environment", h->ws_bytes);
  if (wpb > HSRB_LOCK_MAXWARPS) wpb = HSRB_LOCK_MAXWARPS;
  const int need_w = (h->n + h->num_sm - 1) / h->num_sm;
  if (wpb > need_w) wpb = need_w;
  if (const char* o = getenv("HSRB_LOCK_WPB")) { int v = atoi(o); if (v >= 1 && v <= wpb) wpb = v; }   // experiments
  h->lock_threads = 32 * wpb;
  const size_t smem = (size_t)h->ws_bytes * wpb + tail;
  CU(hsrb_prepare_step_lock(smem, h->lock_threads, &h->lock_bps));
  if (h->lock_bps < 1) return fail(-3, "phase-locked general kernel does not fit on an SM (%zu bytes of shared memory)", smem);
  const int need = (h->n + wpb - 1) / wpb;
  const int cap = h->lock_bps * h->num_sm;
  h->lock_grid = need < cap ? need : cap;
  h->lock_configured = true;
  return 0;
}
bool use_lock(const hsrb* h) {
  if (h->hm.m.npair > 256) return false;
  const char* o = getenv("HSRB_GENERAL_LOCK");
  return !(o && o[0] == '0');
}

// Work-sorted launch order of the warp-per-environment kernel (hsrb_wpe_kernel_t): the action
// kernel writes a work estimate per environment, a radix sort (cub, descending, stable) turns it into the order of the next
// launch.  HSRB_WPE_SORT=0 switches it off (experiments; results do not depend on it).
bool use_sorted_order() {
  const char* so = getenv("HSRB_WPE_SORT");
  return !(so && so[0] == '0');
}
int sort_prepare(hsrb* h, KArgs& a, void* stream) {
  if (!h->d_work) {
    const size_t nb = sizeof(int) * (size_t)h->n;
    CU(cudaMalloc(&h->d_work, nb)); CU(cudaMalloc(&h->d_work_sorted, nb)); CU(cudaMalloc(&h->d_iota, nb)); CU(cudaMalloc(&h->d_order, nb));
    CU(cudaMemsetAsync(h->d_work, 0, nb, (cudaStream_t)stream));
    fill_iota<<<(h->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->d_iota, h->n);
    CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, h->sort_tmp_bytes, h->d_work, h->d_work_sorted, h->d_iota, h->d_order, h->n, 0, 24,
                                                 (cudaStream_t)stream));
    CU(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes));
  }
  a.work = h->d_work; a.order = h->order_valid ? h->d_order : nullptr;
  return 0;
}
int sort_update(hsrb* h, void* stream) {   // the next launch's order (ties keep the environment order)
  CU(cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp, h->sort_tmp_bytes, h->d_work, h->d_work_sorted, h->d_iota, h->d_order, h->n, 0, 24,
                                               (cudaStream_t)stream));
  h->order_valid = true;
  return 0;
}

KArgs base_args(hsrb* h) {
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.m = h->dm; a.cfg = h->cfg; a.n = h->n; a.S = h->S; a.ws_bytes = h->ws_bytes;
  a.seed = h->seed; a.env_off = h->env_off; a.state = h->d_state; a.episode = h->d_episode; a.stats = h->d_stats;
  if (const char* o = getenv("HSRB_OPTS")) a.opts = (unsigned)strtoul(o, nullptr, 0);   // experiment switches
  return a;
}

int run(hsrb* h, KArgs& a, void* stream) {
  if (a.mode == MODE_STEP && use_wpe(h)) {
    int rc = configure_wpe(h);
    if (rc) return rc;
    a.ws_bytes = h->wpe_ws;
    a.m.ncon_max = WPE_MAXCON; a.m.nefc_max = WPE_MAXROW;
    const size_t smem = (size_t)h->wpe_ws * (h->wpe_threads / 32) + wpe::shared_tail(h->dm);
    // phase-locked variant with two teams per block by default; HSRB_WPE_LOCK=0: free-running warps, HSRB_WPE_TEAMS=1|2|4
    const char* lk = getenv("HSRB_WPE_LOCK");
    const char* tm = getenv("HSRB_WPE_TEAMS");
    const bool lock = !(lk && lk[0] == '0');
    const unsigned teams = tm ? (unsigned)atoi(tm) : 2u;
    if (!(a.opts & 0xf0u)) a.opts |= (teams & 15u) << 4;
    const bool sorted = use_sorted_order();
    if (sorted) { int rc2 = sort_prepare(h, a, stream); if (rc2) return rc2; }
    CU(hsrb_wpe_launch(a, h->fast, h->wpe_grid, h->wpe_threads, smem, (cudaStream_t)stream, lock));
    if (sorted) { int rc2 = sort_update(h, stream); if (rc2) return rc2; }
    h->launches++;
    return 0;
  }
  if (a.mode == MODE_STEP && use_fast(h)) {
    int rc = configure_fast(h);
    if (rc) return rc;
    a.ws_bytes = h->fast_ws;
    a.m.ncon_max = PUSH_MAXCON; a.m.nefc_max = 2 + PUSH_ROWS;
    size_t smem = (size_t)h->fast_ws * (h->fast_threads / h->fast_lanes) + push::shared_tail(h->dm);
    CU(hsrb_push_launch(h->fast_lanes, h->fast.nv, a, h->fast, h->fast_grid, h->fast_threads, smem, (cudaStream_t)stream));
    h->launches++;
    return 0;
  }
  if (a.mode == MODE_STEP && h->path == 2) return fail(-3, "fast path not available for this model: %s", h->fast_why);
  if (a.mode == MODE_STEP && use_lock(h) && !h->lanes_req) {
    int rc = configure_lock(h);
    if (rc) return rc;
    a.ws_bytes = h->ws_bytes;
    a.m.ncon_max = h->dm.ncon_max; a.m.nefc_max = h->dm.nefc_max;
    const size_t smem = (size_t)h->ws_bytes * (h->lock_threads / 32) + lock_tail_bytes();
    // (the work-sorted launch order of the wpe kernel was tried here too: no gain on configs[2] / [4], 2.80 vs 2.88 and 1.28 vs
    // 1.35 M substeps/s - configs[2] resets every environment before every action, configs[4] runs four warps per block)
    CU(hsrb_launch_step_lock(a, h->lock_grid, h->lock_threads, smem, (cudaStream_t)stream));
    h->launches++;
    return 0;
  }
  int rc = configure(h);
  if (rc) return rc;
  a.ws_bytes = h->ws_bytes;
  a.m.ncon_max = h->dm.ncon_max; a.m.nefc_max = h->dm.nefc_max;
  size_t smem = (size_t)h->ws_bytes * (32 / h->lanes);
  if (a.mode == MODE_RESET || a.mode == MODE_FORWARD) CU(launch_aux(h->lanes, a, h->grid, smem, (cudaStream_t)stream));   // the lean instance
  else CU(launch(h->lanes, a, h->grid, smem, (cudaStream_t)stream));
  h->launches++;
  return 0;
}

// FP32 FMA-loop peak of the device (the denominator of bench.py's FP32 roofline): 8 independent chains per thread
__global__ void __launch_bounds__(1024) fma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------------- 'openai' observation
// The 25-d Fetch-style observation of HSREnv._get_observation (/root/reference/hsr/env.py:72-110; the branch is dead code in
// the snapshot, SURVEY.md App. C #8: this is its intent, stated in hsr_env_b200/kin.py, which stays as the test-side
// oracle) from the resident state, one thread per environment, fp64: forward kinematics with 6-D body velocities down
// the tree (bodies are in tree order), then
//   grip_pos | object_pos | object_pos - grip_pos | finger qpos (2) | mat2euler(object xmat) | (object_velp - grip_velp) dt |
//   object_velr dt | grip_velp dt | finger qvel dt (2)
__device__ __forceinline__ void oq_mul(const double* a, const double* b, double* r) {
  const double r0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], r1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const double r2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], r3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3;
}
__device__ __forceinline__ void oq_mat(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ void oq_norm(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  n = n < 1e-15 ? 1e-15 : n;
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
#define OMV(R, v, o) { o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2]; o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2]; o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2]; }
#define OCROSS(a, b, o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
__global__ void __launch_bounds__(128) openai_obs_kernel(const ModelT<float> m, const float* __restrict__ state, int n, int S, int block_body,
                                                        int adr_l, int adr_r, int dof_l, int dof_r, float* __restrict__ obs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* qpos = state + (size_t)e * S;
  const float* qvel = qpos + m.nq;
  double xpos[HSRB_MAXBODY][3], xquat[HSRB_MAXBODY][4], velp[HSRB_MAXBODY][3], velr[HSRB_MAXBODY][3];
  for (int k = 0; k < 3; k++) { xpos[0][k] = 0; velp[0][k] = 0; velr[0][k] = 0; }
  xquat[0][0] = 1; xquat[0][1] = xquat[0][2] = xquat[0][3] = 0;
  for (int b = 1; b < m.nbody; b++) {
    const int p = m.body_parent[b], j0 = m.body_jntadr[b], nj = m.body_jntnum[b];
    if (nj == 1 && m.jnt_type[j0] == JNT_FREE) {
      const int a = m.jnt_qposadr[j0], v = m.jnt_dofadr[j0];
      double q[4] = {(double)qpos[a + 3], (double)qpos[a + 4], (double)qpos[a + 5], (double)qpos[a + 6]}, R[9];
      oq_norm(q); oq_mat(q, R);
      const double w[3] = {(double)qvel[v + 3], (double)qvel[v + 4], (double)qvel[v + 5]};
      double wr[3];
      OMV(R, w, wr);
      for (int k = 0; k < 3; k++) { xpos[b][k] = (double)qpos[a + k]; velp[b][k] = (double)qvel[v + k]; velr[b][k] = wr[k]; }
      for (int k = 0; k < 4; k++) xquat[b][k] = q[k];
      continue;
    }
    double Rp[9], pos[3], quat[4], w[3], vel[3], t3[3], d3[3];
    oq_mat(xquat[p], Rp);
    const double bp[3] = {(double)m.body_pos[3 * b], (double)m.body_pos[3 * b + 1], (double)m.body_pos[3 * b + 2]};
    const double bq[4] = {(double)m.body_quat[4 * b], (double)m.body_quat[4 * b + 1], (double)m.body_quat[4 * b + 2], (double)m.body_quat[4 * b + 3]};
    OMV(Rp, bp, t3);
    for (int k = 0; k < 3; k++) { pos[k] = xpos[p][k] + t3[k]; w[k] = velr[p][k]; d3[k] = pos[k] - xpos[p][k]; }
    oq_mul(xquat[p], bq, quat);
    OCROSS(velr[p], d3, t3);
    for (int k = 0; k < 3; k++) vel[k] = velp[p][k] + t3[k];   // origin of b carried rigidly by its parent
    for (int j = j0; j < j0 + nj; j++) {
      double R[9], anchor[3], axis[3];
      oq_mat(quat, R);
      const double jp[3] = {(double)m.jnt_pos[3 * j], (double)m.jnt_pos[3 * j + 1], (double)m.jnt_pos[3 * j + 2]};
      const double ja[3] = {(double)m.jnt_axis[3 * j], (double)m.jnt_axis[3 * j + 1], (double)m.jnt_axis[3 * j + 2]};
      OMV(R, jp, t3);
      for (int k = 0; k < 3; k++) anchor[k] = pos[k] + t3[k];
      OMV(R, ja, axis);
      const int a = m.jnt_qposadr[j], v = m.jnt_dofadr[j];
      const double q = (double)qpos[a] - (double)m.qpos0[a], qd = (double)qvel[v];
      if (m.jnt_type[j] == JNT_SLIDE) {
        for (int k = 0; k < 3; k++) { pos[k] += axis[k] * q; vel[k] += axis[k] * qd; }
      } else {
        const double half = 0.5 * q, sh = sin(half);
        const double dq[4] = {cos(half), sh * ja[0], sh * ja[1], sh * ja[2]};
        oq_mul(quat, dq, quat);
        oq_mat(quat, R);
        OMV(R, jp, t3);
        for (int k = 0; k < 3; k++) { pos[k] = anchor[k] - t3[k]; w[k] += axis[k] * qd; d3[k] = pos[k] - anchor[k]; }
        const double aw[3] = {axis[0] * qd, axis[1] * qd, axis[2] * qd};
        OCROSS(aw, d3, t3);
        for (int k = 0; k < 3; k++) vel[k] += t3[k];
      }
    }
    oq_norm(quat);
    for (int k = 0; k < 3; k++) { xpos[b][k] = pos[k]; velp[b][k] = vel[k]; velr[b][k] = w[k]; }
    for (int k = 0; k < 4; k++) xquat[b][k] = quat[k];
  }
  double gp[3] = {0, 0, 0}, gv[3] = {0, 0, 0};
  for (int f = 0; f < 2; f++) {
    const int b = m.finger_body[f];
    double R[9], r[3], t3[3];
    oq_mat(xquat[b], R);
    const double fp[3] = {(double)m.finger_pos[3 * f], (double)m.finger_pos[3 * f + 1], (double)m.finger_pos[3 * f + 2]};
    OMV(R, fp, r);
    OCROSS(velr[b], r, t3);
    for (int k = 0; k < 3; k++) { gp[k] += 0.5 * (xpos[b][k] + r[k]); gv[k] += 0.5 * (velp[b][k] + t3[k]); }
  }
  const double dt = (double)m.timestep;
  double Rb[9];
  oq_mat(xquat[block_body], Rb);
  // mat2euler, /root/reference/hsr/env.py:256-272
  const double cy = sqrt(Rb[8] * Rb[8] + Rb[5] * Rb[5]);
  const bool cond = cy > 2.220446049250313e-16 * 4;
  const double e2 = cond ? -atan2(Rb[1], Rb[0]) : -atan2(-Rb[3], Rb[4]);
  const double e1 = -atan2(-Rb[2], cy);
  const double e0 = cond ? -atan2(Rb[5], Rb[8]) : 0.0;
  float* o = obs + (size_t)e * 25;
  const double* op = xpos[block_body];
  for (int k = 0; k < 3; k++) {
    o[k] = (float)gp[k]; o[3 + k] = (float)op[k]; o[6 + k] = (float)(op[k] - gp[k]);
    o[14 + k] = (float)((velp[block_body][k] - gv[k]) * dt); o[17 + k] = (float)(velr[block_body][k] * dt); o[20 + k] = (float)(gv[k] * dt);
  }
  o[9] = adr_l >= 0 ? qpos[adr_l] : 0.f; o[10] = adr_r >= 0 ? qpos[adr_r] : 0.f;
  o[11] = (float)e0; o[12] = (float)e1; o[13] = (float)e2;
  o[23] = dof_l >= 0 ? (float)((double)qvel[dof_l] * dt) : 0.f; o[24] = dof_r >= 0 ? (float)((double)qvel[dof_r] * dt) : 0.f;
}
#undef OMV
#undef OCROSS

__global__ void gather_state(const float* __restrict__ st, int n, int S, int off, int w, float* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * w) return;
  int e = (int)(i / w), k = (int)(i % w);
  out[i] = st[(size_t)e * S + off + k];
}
__global__ void scatter_state(float* __restrict__ st, int n, int S, int off, int w, const float* __restrict__ in) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * w) return;
  int e = (int)(i / w), k = (int)(i % w);
  st[(size_t)e * S + off + k] = in[i];
}

}  // namespace

extern "C" {

const char* hsrb_last_error(void) { return g_err; }

int hsrb_create(const void* model_blob, size_t bytes, int n_envs, int device, uint64_t seed, uint64_t env_id_offset,
                hsrb_t** out) {
  if (!model_blob || !out || n_envs <= 0) return fail(-1, "hsrb_create: bad arguments");
  *out = nullptr;
  hsrb* h = new (std::nothrow) hsrb();
  if (!h) return fail(-1, "out of host memory");
  std::string err;
  if (!h->hm.parse(model_blob, bytes, err)) { delete h; return fail(-1, "model blob: %s", err.c_str()); }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    delete h;
    return fail(-2, "no CUDA device (%s): this backend has no CPU fallback", cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) { delete h; return fail(-1, "device %d out of range (%d visible)", device, ndev); }
  h->n = n_envs; h->device = device; h->seed = seed; h->env_off = env_id_offset;
  h->S = h->hm.m.nq + 2 * h->hm.m.nv + 3;
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.qidx0 = 0; h->cfg.qidx1 = 2;
#define CUH(call)                                                                                    \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) { int rc_ = fail(-2, "%s: %s", #call, cudaGetErrorString(e_)); hsrb_destroy(h); return rc_; } \
  } while (0)
  CUH(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUH(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    int rc = fail(-2, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    hsrb_destroy(h);
    return rc;
  }
  h->num_sm = prop.multiProcessorCount;
  CUH(cudaMalloc(&h->d_model, h->hm.buf.size()));
  CUH(cudaMemcpy(h->d_model, h->hm.buf.data(), h->hm.buf.size(), cudaMemcpyHostToDevice));
  HostModel<float> tmp = h->hm;  // copy the table, then point it at the device buffer
  tmp.rebase(h->d_model);
  h->dm = tmp.m;
  h->hm.rebase(h->hm.buf.data());
  CUH(cudaMalloc(&h->d_state, sizeof(float) * (size_t)h->n * h->S));
  CUH(cudaMalloc(&h->d_episode, sizeof(unsigned) * (size_t)h->n));
  CUH(cudaMalloc(&h->d_stats, sizeof(unsigned long long) * ST_COUNT));
  CUH(cudaMemset(h->d_episode, 0, sizeof(unsigned) * (size_t)h->n));
  CUH(cudaMemset(h->d_stats, 0, sizeof(unsigned long long) * ST_COUNT));
  // state after load = mj_resetData: qpos0, zero velocities (MjSim(model), mujoco_env.py:34)
  {
    std::vector<float> row(h->S, 0.f);
    for (int i = 0; i < h->hm.m.nq; i++) row[i] = h->hm.m.qpos0[i];
    for (int k = 0; k < 3; k++) row[h->hm.m.nq + 2 * h->hm.m.nv + k] = h->hm.m.mocap_pos0[k];
    std::vector<float> all((size_t)h->n * h->S);
    for (int e2 = 0; e2 < h->n; e2++) memcpy(all.data() + (size_t)e2 * h->S, row.data(), sizeof(float) * h->S);
    CUH(cudaMemcpy(h->d_state, all.data(), sizeof(float) * all.size(), cudaMemcpyHostToDevice));
  }
#undef CUH
  h->fast_ok = push_fill_info(h->hm.m, h->fast, h->fast_tab, h->fast_why, sizeof(h->fast_why));
  if (h->fast_ok) {
    cudaError_t e2 = cudaMalloc(&h->d_fast_tab, h->fast_tab.bytes());
    if (e2 == cudaSuccess) e2 = cudaMemcpy(h->d_fast_tab, h->fast_tab.tab.data(), h->fast_tab.bytes(), cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) { int rc_ = fail(-2, "fast-path tables: %s", cudaGetErrorString(e2)); hsrb_destroy(h); return rc_; }
    h->fast_tab.point(h->fast, h->d_fast_tab);
    int nbg = 0;
    for (int gi = 0; gi < h->hm.m.ngeom; gi++) nbg += h->hm.m.geom_body[gi] == h->fast.block_body ? 1 : 0;
    h->wpe_ok = nbg <= WPE_MAXBG && h->hm.m.npair <= WPE_MAXPAIR && h->hm.m.ngeom <= WPE_MAXGEOM;
  }
  *out = h;
  return 0;
}

int hsrb_set_path(hsrb_t* h, int path) {
  if (!h) return fail(-1, "null handle");
  if (path < 0 || path > 3) return fail(-1, "path must be 0 (auto), 1 (general kernel), 2 (lock-step fast kernel) or 3 (warp-per-environment fast kernel)");
  if (path >= 2 && !h->fast_ok) return fail(-3, "fast path not available for this model: %s", h->fast_why);
  if (path == 3 && !h->wpe_ok) return fail(-3, "warp-per-environment kernel not available for this model");
  h->path = path;
  return use_wpe(h) ? 3 : (use_fast(h) ? 2 : 1);
}

int hsrb_destroy(hsrb_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->d_fast_tab); cudaFree(h->d_model); cudaFree(h->d_state); cudaFree(h->d_episode); cudaFree(h->d_stats);
  cudaFree(h->d_work); cudaFree(h->d_work_sorted); cudaFree(h->d_iota); cudaFree(h->d_order); cudaFree(h->d_sort_tmp);
  cudaFree(h->d_ctrl); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_done); cudaFree(h->d_taken);
  delete h;
  return 0;
}

int hsrb_dims(hsrb_t* h, int* nq, int* nv, int* nu, int* nbody, int* nblock) {
  if (!h) return fail(-1, "null handle");
  if (nq) *nq = h->hm.m.nq;
  if (nv) *nv = h->hm.m.nv;
  if (nu) *nu = h->hm.m.nu;
  if (nbody) *nbody = h->hm.m.nbody;
  if (nblock) *nblock = h->hm.m.nblock;
  return 0;
}

int hsrb_config(hsrb_t* h, int lanes_per_env, int ncon_max, int nefc_max) {
  if (!h) return fail(-1, "null handle");
  if (lanes_per_env != 0 && lanes_per_env != 4 && lanes_per_env != 8 && lanes_per_env != 16 && lanes_per_env != 32)
    return fail(-1, "lanes_per_env must be 0, 4, 8, 16 or 32");
  h->lanes_req = lanes_per_env;
  if (ncon_max > 0) h->dm.ncon_max = h->hm.m.ncon_max = ncon_max;
  if (nefc_max > 0) h->dm.nefc_max = h->hm.m.nefc_max = nefc_max;
  h->configured = false;
  h->fast_configured = false;
  h->wpe_configured = false;
  h->lock_configured = false;
  CU(cudaSetDevice(h->device));
  return configure(h);
}

int hsrb_set_goals(hsrb_t* h, const float* goal_lohi, const float* block_lohi, float geofence, float min_sep, int qidx0,
                   int qidx1) {
  if (!h) return fail(-1, "null handle");
  if (qidx0 < 0 || qidx0 > 3 || qidx1 < 0 || qidx1 > 3) return fail(-1, "quaternion indices must be in 0..3");
  EnvCfg<float>& c = h->cfg;
  c.has_goal = goal_lohi ? 1 : 0;
  c.has_block = block_lohi ? 1 : 0;
  c.ngoal = 0;
  c.qidx0 = qidx0; c.qidx1 = qidx1; c.geofence = geofence; c.min_sep = min_sep;
  for (int k = 0; k < 3; k++) { c.goal_lo[k] = goal_lohi ? goal_lohi[k] : 0.f; c.goal_hi[k] = goal_lohi ? goal_lohi[3 + k] : 0.f; }
  for (int k = 0; k < 4; k++) { c.block_lo[k] = block_lohi ? block_lohi[k] : 0.f; c.block_hi[k] = block_lohi ? block_lohi[4 + k] : 0.f; }
  return 0;
}

int hsrb_set_goal_list(hsrb_t* h, int ngoal, const int32_t* a_codes, const int32_t* b_codes, const float* distance,
                       const float* point_lohi, const float* fixed_pts, int nfixed) {
  if (!h) return fail(-1, "null handle");
  if (ngoal < 0 || ngoal > HSRB_MAXGOAL) return fail(-1, "at most %d GoalSpecs", HSRB_MAXGOAL);
  if (nfixed < 0 || nfixed > HSRB_MAXFIXED) return fail(-1, "at most %d fixed goal points", HSRB_MAXFIXED);
  if (ngoal > 0 && (!a_codes || !b_codes || !distance)) return fail(-1, "goal list arrays are NULL");
  EnvCfg<float>& c = h->cfg;
  for (int k = 0; k < ngoal; k++) {
    const int codes[2] = {a_codes[k], b_codes[k]};
    for (int e = 0; e < 2; e++) {
      if (codes[e] >= h->hm.m.nbody) return fail(-1, "goal %d: body id %d out of range", k, codes[e]);
      if (codes[e] < -1 - nfixed) return fail(-1, "goal %d: fixed point %d not given", k, -2 - codes[e]);
    }
  }
  c.ngoal = ngoal;
  c.has_goal = ngoal > 0 ? 1 : 0;
  c.has_block = 0;
  for (int k = 0; k < ngoal; k++) { c.goal_a[k] = a_codes[k]; c.goal_b[k] = b_codes[k]; c.goal_dist[k] = distance[k]; }
  for (int k = 0; k < 3; k++) {
    c.goal_lo[k] = point_lohi ? point_lohi[k] : h->hm.m.mocap_pos0[k];
    c.goal_hi[k] = point_lohi ? point_lohi[3 + k] : h->hm.m.mocap_pos0[k];
  }
  for (int k = 0; k < nfixed; k++) for (int i = 0; i < 3; i++) c.fixed_pt[k][i] = fixed_pts[3 * k + i];
  return 0;
}

int hsrb_set_starts(hsrb_t* h, int nstart, const int32_t* qpos_adr, const int32_t* width, const float* lo, const float* hi) {
  if (!h) return fail(-1, "null handle");
  if (nstart < 0 || nstart > HSRB_MAXSTART) return fail(-1, "at most %d joints with a start space", HSRB_MAXSTART);
  if (nstart > 0 && (!qpos_adr || !width || !lo || !hi)) return fail(-1, "start arrays are NULL");
  EnvCfg<float>& c = h->cfg;
  for (int s = 0; s < nstart; s++) {
    if (width[s] != 1 && width[s] != 7) return fail(-1, "start %d: width must be 1 (slide / hinge) or 7 (free joint)", s);
    if (qpos_adr[s] < 0 || qpos_adr[s] + width[s] > h->hm.m.nq) return fail(-1, "start %d: qpos slice out of range", s);
  }
  c.nstart = nstart;
  for (int s = 0; s < nstart; s++) {
    c.start_adr[s] = qpos_adr[s]; c.start_width[s] = width[s];
    for (int k = 0; k < 7; k++) { c.start_lo[s][k] = k < width[s] ? lo[7 * s + k] : 0.f; c.start_hi[s][k] = k < width[s] ? hi[7 * s + k] : 0.f; }
  }
  return 0;
}

#if defined(WPE_CHAIN_CLOCKS)
// experiment hook (not part of include/hsrb.h): the per-environment work array of the last wpe launch -> host
extern "C" int hsrb_debug_work(hsrb_t* h, int* out_host) {
  if (!h || !h->d_work) return -1;
  cudaDeviceSynchronize();
  return cudaMemcpy(out_host, h->d_work, sizeof(int) * (size_t)h->n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif

int hsrb_reset(hsrb_t* h, const uint8_t* mask, float* obs, void* stream) {
  if (!h) return fail(-1, "null handle");
  CU(cudaSetDevice(h->device));
  KArgs a = base_args(h);
  a.mode = MODE_RESET; a.mask = mask; a.obs = obs; a.nsub = 0;
  return run(h, a, stream);
}

int hsrb_step(hsrb_t* h, const float* ctrl, int nsubsteps, float* obs, float* reward, uint8_t* done, uint8_t* success,
              int32_t* substeps_taken, uint8_t* bad_state, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!ctrl && h->hm.m.nu > 0) return fail(-1, "ctrl is NULL");
  if (nsubsteps < 0) return fail(-1, "nsubsteps < 0");
  CU(cudaSetDevice(h->device));
  KArgs a = base_args(h);
  a.mode = MODE_STEP; a.ctrl = ctrl; a.nsub = nsubsteps; a.obs = obs; a.reward = reward; a.done = done;
  a.success = success; a.taken = substeps_taken; a.bad = bad_state;
  return run(h, a, stream);
}

int hsrb_step_host(hsrb_t* h, const float* ctrl_host, int nsubsteps, float* obs_host, float* reward_host,
                   uint8_t* done_host, int32_t* taken_host, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!ctrl_host && h->hm.m.nu > 0) return fail(-1, "ctrl_host is NULL");
  CU(cudaSetDevice(h->device));
  const int n = h->n, nu = h->hm.m.nu, nobs = h->hm.m.nq + h->hm.m.nv;
  if (!h->d_obs) {
    CU(cudaMalloc(&h->d_ctrl, sizeof(float) * (size_t)n * (nu > 0 ? nu : 1)));
    CU(cudaMalloc(&h->d_obs, sizeof(float) * (size_t)n * nobs));
    CU(cudaMalloc(&h->d_reward, sizeof(float) * (size_t)n));
    CU(cudaMalloc(&h->d_done, (size_t)n));
    CU(cudaMalloc(&h->d_taken, sizeof(int) * (size_t)n));
  }
  // Everything is enqueued on the CALLER's stream, behind whatever the caller launched there before (hsrb_reset,
  // hsrb_set_state, a previous step), so no cross-stream ordering is left to the caller; the call returns after the
  // device->host copies have completed.
  cudaStream_t s = (cudaStream_t)stream;
  if (nu > 0) CU(cudaMemcpyAsync(h->d_ctrl, ctrl_host, sizeof(float) * (size_t)n * nu, cudaMemcpyHostToDevice, s));
  int rc = hsrb_step(h, h->d_ctrl, nsubsteps, h->d_obs, h->d_reward, h->d_done, nullptr, h->d_taken, nullptr, s);
  if (rc) return rc;
  if (obs_host) CU(cudaMemcpyAsync(obs_host, h->d_obs, sizeof(float) * (size_t)n * nobs, cudaMemcpyDeviceToHost, s));
  if (reward_host) CU(cudaMemcpyAsync(reward_host, h->d_reward, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s));
  if (done_host) CU(cudaMemcpyAsync(done_host, h->d_done, (size_t)n, cudaMemcpyDeviceToHost, s));
  if (taken_host) CU(cudaMemcpyAsync(taken_host, h->d_taken, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int hsrb_get_state(hsrb_t* h, float* qpos, float* qvel, float* qacc_warm, float* mocap_pos, void* stream) {
  if (!h) return fail(-1, "null handle");
  CU(cudaSetDevice(h->device));
  const int nq = h->hm.m.nq, nv = h->hm.m.nv;
  float* outs[4] = {qpos, qvel, qacc_warm, mocap_pos};
  int offs[4] = {0, nq, nq + nv, nq + 2 * nv}, ws[4] = {nq, nv, nv, 3};
  for (int k = 0; k < 4; k++) {
    if (!outs[k] || ws[k] == 0) continue;
    long long tot = (long long)h->n * ws[k];
    gather_state<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d_state, h->n, h->S, offs[k], ws[k], outs[k]);
  }
  CU(cudaGetLastError());
  return 0;
}

int hsrb_set_state(hsrb_t* h, const float* qpos, const float* qvel, const float* qacc_warm, const float* mocap_pos,
                   void* stream) {
  if (!h) return fail(-1, "null handle");
  CU(cudaSetDevice(h->device));
  const int nq = h->hm.m.nq, nv = h->hm.m.nv;
  const float* ins[4] = {qpos, qvel, qacc_warm, mocap_pos};
  int offs[4] = {0, nq, nq + nv, nq + 2 * nv}, ws[4] = {nq, nv, nv, 3};
  for (int k = 0; k < 4; k++) {
    if (!ins[k] || ws[k] == 0) continue;
    long long tot = (long long)h->n * ws[k];
    scatter_state<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d_state, h->n, h->S, offs[k], ws[k], ins[k]);
  }
  CU(cudaGetLastError());
  return 0;
}

int hsrb_forward(hsrb_t* h, float* body_xpos, float* gripper_pos, void* stream) {
  if (!h) return fail(-1, "null handle");
  CU(cudaSetDevice(h->device));
  KArgs a = base_args(h);
  a.mode = MODE_FORWARD; a.body_xpos = body_xpos; a.gripper = gripper_pos;
  return run(h, a, stream);
}

int hsrb_openai_obs(hsrb_t* h, int finger_qposadr_l, int finger_qposadr_r, int finger_dofadr_l, int finger_dofadr_r, float* obs25,
                    void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!obs25) return fail(-1, "obs25 is NULL");
  const auto& m = h->hm.m;
  if (m.nblock < 1) return fail(-3, "the 'openai' observation needs a block (hsr/env.py:58)");
  if (m.nbody > HSRB_MAXBODY) return fail(-3, "more than %d bodies", HSRB_MAXBODY);
  if (finger_qposadr_l >= m.nq || finger_qposadr_r >= m.nq || finger_dofadr_l >= m.nv || finger_dofadr_r >= m.nv)
    return fail(-1, "finger joint address out of range");
  CU(cudaSetDevice(h->device));
  openai_obs_kernel<<<(h->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->dm, h->d_state, h->n, h->S, m.block_body[0], finger_qposadr_l,
                                                                        finger_qposadr_r, finger_dofadr_l, finger_dofadr_r, obs25);
  CU(cudaGetLastError());
  return 0;
}

int hsrb_compute_reward(hsrb_t* h, float* reward, uint8_t* success, void* stream) {
  if (!h) return fail(-1, "null handle");
  CU(cudaSetDevice(h->device));
  KArgs a = base_args(h);
  a.mode = MODE_FORWARD; a.reward = reward; a.success = success;
  return run(h, a, stream);
}

int hsrb_debug_size(hsrb_t* h) {
  if (!h) return fail(-1, "null handle");
  return (int)debug_size(h->dm);
}

int hsrb_debug_substep(hsrb_t* h, const float* ctrl, double* dump, void* stream) {
  if (!h) return fail(-1, "null handle");
  if (!dump) return fail(-1, "dump is NULL");
  CU(cudaSetDevice(h->device));
  KArgs a = base_args(h);
  a.mode = MODE_DEBUG; a.ctrl = ctrl; a.nsub = 1; a.dump = dump; a.dump_stride = (unsigned)debug_size(h->dm);
  return run(h, a, stream);
}

int hsrb_stats(hsrb_t* h, int64_t* out9, void* stream) {
  static_assert(ST_COUNT == 16, "hsrb.h documents 16 counters");
  if (!h || !out9) return fail(-1, "bad arguments");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  unsigned long long v[ST_COUNT];
  CU(cudaMemcpy(v, h->d_stats, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < ST_COUNT; i++) out9[i] = (int64_t)v[i];
  out9[ST_LAUNCHES] = h->launches;
  return 0;
}

int hsrb_measure_fp32_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail(-1, "bad arguments");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 4096;
  float* out = nullptr;
  CU(cudaMalloc(&out, sizeof(float) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, 0);
    fma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  CU(cudaGetLastError());
  *tflops_out = best;
  return 0;
}

int hsrb_launch_info(hsrb_t* h, int* out4) {  // out4: 6 ints
  if (!h || !out4) return fail(-1, "bad arguments");
  CU(cudaSetDevice(h->device));
  int rc = configure(h);
  if (rc) return rc;
  if (use_wpe(h)) {
    rc = configure_wpe(h);
    if (rc) return rc;
    out4[0] = 32; out4[1] = (int)h->wpe_ws; out4[2] = h->wpe_bps * (h->wpe_threads / 32); out4[3] = h->wpe_grid;
    out4[4] = 3; out4[5] = h->wpe_threads;
    return 0;
  }
  if (use_fast(h)) {
    rc = configure_fast(h);
    if (rc) return rc;
    out4[0] = h->fast_lanes; out4[1] = (int)h->fast_ws; out4[2] = h->fast_bps * (h->fast_threads / h->fast_lanes); out4[3] = h->fast_grid;
    out4[4] = 2; out4[5] = h->fast_threads;
    return 0;
  }
  if (use_lock(h) && !h->lanes_req) {
    rc = configure_lock(h);
    if (rc) return rc;
    out4[0] = 32; out4[1] = (int)h->ws_bytes; out4[2] = h->lock_bps * (h->lock_threads / 32); out4[3] = h->lock_grid;
    out4[4] = 1; out4[5] = h->lock_threads;
    return 0;
  }
  out4[0] = h->lanes; out4[1] = (int)h->ws_bytes; out4[2] = h->blocks_per_sm * (32 / h->lanes); out4[3] = h->grid;
  out4[4] = 1; out4[5] = 32;
  return 0;
}

}  // extern "C"
