// Kernel-side definitions shared by the per-G translation units (hsrb_step_g*.cu) and the API TU (hsrb_api.cu).
//
// One persistent kernel per action: a group of G lanes of a warp owns one environment, loads its state
// (qpos, qvel, qacc_warmstart, mocap, ctrl) from HBM once, keeps state, kinematics, contacts, constraint
// Jacobian and solver vectors in its slice of shared memory for all <=300 substeps, and writes state, obs,
// reward, done and substep counts back once.  No tensor cores: the per-environment systems are 8x8..26x26.
//
// Replaces HSREnv.step's loop over sim.step()  (/root/reference/hsr/env.py:115-135).
#pragma once
#include <cuda_runtime.h>

#include "hsr_core.h"

using namespace hsr;

// dynamic shared memory of the block (the SIMT emulator of tests/simt_emu substitutes a host buffer)
#if defined(HSRB_SIMT_EMU)
#define HSRB_DYN_SMEM(name) unsigned char* name = emu::dyn_smem
#else
#define HSRB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

enum { ST_SUBSTEPS = 0, ST_ITERS, ST_NARROW, ST_LSEVAL, ST_CONTACTS, ST_ROWS, ST_LAUNCHES, ST_BAD, ST_FLOPS, ST_PHASE0, ST_COUNT = ST_PHASE0 + PH_COUNT };

struct KArgs {
  ModelT<float> m;
  EnvCfg<float> cfg;
  int n, S, nsub, mode;
  unsigned opts;           // experiment switches (HSRB_OPTS): bit 0 = no separating-direction cache in the fast kernel
  unsigned ws_bytes;
  unsigned long long seed, env_off;
  float* state;            // [N,S]: qpos nq | qvel nv | qacc_warmstart nv | mocap 3
  unsigned* episode;       // [N]
  const float* ctrl;       // [N,nu]
  const unsigned char* mask;
  float* obs; float* reward; unsigned char* done; unsigned char* success; int* taken; unsigned char* bad;
  float* body_xpos; float* gripper;
  double* dump; unsigned dump_stride;
  unsigned long long* stats;
};

enum { MODE_STEP = 0, MODE_DEBUG = 1, MODE_FORWARD = 2, MODE_RESET = 3 };

template <int G>
__device__ __forceinline__ void load_state(const KArgs& a, WS<float>& w, const DevGrp<G>& g, int env) {
  const float* st = a.state + (size_t)env * a.S;
  int n0 = a.m.nq + 2 * a.m.nv;
  for (int i = g.lane; i < n0; i += G) w.qpos[i] = st[i];  // qpos|qvel|warm are contiguous in the workspace
  for (int i = g.lane; i < 3; i += G) w.mocap[i] = st[n0 + i];
  for (int i = g.lane; i < WI_COUNT; i += G) w.wi[i] = 0;
  for (int i = g.lane; i < 4 * a.m.npair; i += G) w.sep[i] = 0.f;   // no cached separating directions
}
template <int G>
__device__ __forceinline__ void store_state(const KArgs& a, WS<float>& w, const DevGrp<G>& g, int env) {
  float* st = a.state + (size_t)env * a.S;
  int n0 = a.m.nq + 2 * a.m.nv;
  for (int i = g.lane; i < n0; i += G) st[i] = w.qpos[i];
  for (int i = g.lane; i < 3; i += G) st[n0 + i] = w.mocap[i];
}

// The action kernel.  blockDim.x = 32 (one warp), 32/G environments per block, grid-stride over environments.
template <int G>
__global__ void __launch_bounds__(32) hsrb_step_kernel(const __grid_constant__ KArgs a) {
  HSRB_DYN_SMEM(smem);
  DevGrp<G> g;
  const int gpb = 32 / G;
  const int gi = threadIdx.x / G;
  WS<float> w;
  ws_carve<float>(a.m, &w, smem + (size_t)gi * a.ws_bytes);
  const int nobs = a.m.nq + a.m.nv;
  for (int env = blockIdx.x * gpb + gi; env < a.n; env += gridDim.x * gpb) {
    load_state<G>(a, w, g, env);
    for (int i = g.lane; i < a.m.nu; i += G) w.ctrl[i] = a.ctrl ? a.ctrl[(size_t)env * a.m.nu + i] : 0.f;
    g.sync();
    bool success = false;
    int taken = 0;
    if (a.mode == MODE_STEP) {
      taken = env_action(a.m, a.cfg, w, g, a.nsub, success);
    } else if (a.mode == MODE_DEBUG) {
      forward(a.m, w, g);
      euler_solve(a.m, w, g);
      if (g.lane == 0) euler_lane0(a.m, w);
      g.sync();
      if (g.lane == 0) debug_dump(a.m, w, a.dump + (size_t)env * a.dump_stride);
      success = goal_reached(a.m, a.cfg, w);
      taken = 1;
    } else if (a.mode == MODE_RESET) {
      // MujocoEnv.reset + HSREnv.reset_model (mujoco_env.py:83-85, env.py:158-177) for the masked environments
      // environments outside the mask keep their state bit for bit (only their observation is emitted)
      if (!a.mask || a.mask[env]) {
        if (g.lane == 0) {
          unsigned ep = a.episode[env];
          a.episode[env] = ep + 1;
          reset_lane0(a.m, a.cfg, w, a.seed, (uint32_t)(a.env_off + (unsigned long long)env), ep);
          kinematics_lane0(a.m, w);  // sim.forward(): normalises free-joint quaternions in qpos
        }
        g.sync();
      }
    } else {  // MODE_FORWARD: body positions of the current state (data.get_body_xpos, env.py:144,180,184)
      if (g.lane == 0) kinematics_lane0(a.m, w);
      g.sync();
      if (a.body_xpos)
        for (int i = g.lane; i < a.m.nbody * 3; i += G) a.body_xpos[(size_t)env * a.m.nbody * 3 + i] = (float)w.xpos[i];
      if (a.gripper && g.lane < 3) {
        GT s = 0;
        for (int k = 0; k < 2; k++) {
          int b = a.m.finger_body[k];
          const GT* R = w.xmat + 9 * b;
          const float* p = a.m.finger_pos + 3 * k;
          s += w.xpos[3 * b + g.lane] + R[3 * g.lane] * (GT)p[0] + R[3 * g.lane + 1] * (GT)p[1] + R[3 * g.lane + 2] * (GT)p[2];
        }
        a.gripper[(size_t)env * 3 + g.lane] = (float)(0.5 * s);
      }
      success = goal_reached(a.m, a.cfg, w);
    }
    store_state<G>(a, w, g, env);
    if (a.obs) for (int i = g.lane; i < nobs; i += G) a.obs[(size_t)env * nobs + i] = w.qpos[i];
    if (g.lane == 0) {
      int flags = w.wi[WI_FLAGS];
      if (a.reward) a.reward[env] = success ? 1.0f : 0.0f;
      if (a.done) a.done[env] = success ? 1 : 0;
      if (a.success) a.success[env] = success ? 1 : 0;
      if (a.taken) a.taken[env] = taken;
      if (a.bad) a.bad[env] = (unsigned char)flags;
      if (a.mode <= MODE_DEBUG) {
      atomicAdd(a.stats + ST_SUBSTEPS, (unsigned long long)taken);
      atomicAdd(a.stats + ST_ITERS, (unsigned long long)w.wi[WI_ITER]);
      atomicAdd(a.stats + ST_NARROW, (unsigned long long)w.wi[WI_NARROW]);
      atomicAdd(a.stats + ST_LSEVAL, (unsigned long long)w.wi[WI_LSEVAL]);
      atomicAdd(a.stats + ST_CONTACTS, (unsigned long long)w.wi[WI_SUMCON]);
      atomicAdd(a.stats + ST_ROWS, (unsigned long long)w.wi[WI_SUMEFC]);
      atomicAdd(a.stats + ST_FLOPS, (unsigned long long)w.wi[WI_KFLOP]);
#ifdef HSRB_PHASE_CLOCKS
      for (int k = 0; k < PH_COUNT; k++) atomicAdd(a.stats + ST_PHASE0 + k, (unsigned long long)w.wi[WI_PHASE0 + k]);
#endif
      if (flags) atomicAdd(a.stats + ST_BAD, 1ull);
      }
    }
    g.sync();
  }
}

// per-G entry points, one translation unit each (parallel compilation)
#define HSRB_DECL_G(G)                                                \
  cudaError_t hsrb_prepare_step_##G(size_t smem, int* blocks_per_sm); \
  cudaError_t hsrb_launch_step_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s);
HSRB_DECL_G(4) HSRB_DECL_G(8) HSRB_DECL_G(16) HSRB_DECL_G(32)

#define HSRB_DEFINE_G(G)                                                                                               \
  cudaError_t hsrb_prepare_step_##G(size_t smem, int* bps) {                                                           \
    cudaError_t e = cudaFuncSetAttribute(hsrb_step_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); /* per function, shared by all handles */ \
    if (e != cudaSuccess) return e;                                                                                    \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_step_kernel<G>, 32, smem);                          \
  }                                                                                                                    \
  cudaError_t hsrb_launch_step_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s) {                            \
    hsrb_step_kernel<G><<<grid, 32, smem, s>>>(a);                                                                     \
    return cudaGetLastError();                                                                                         \
  }
