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
  const int* order;        // wpe kernel: environment of launch slot r (heaviest first), or null = identity
  int* work;               // wpe kernel: [N] work estimate of this action per environment (the next launch's sort key)
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
// FULL = false: the same kernel without the STEP / DEBUG modes (reset and forward launches): a few KB of code instead of
// several hundred, which matters when it runs between two action kernels on a cold L2 (bench.py flushes L2 between timed
// steps: the masked reset through the full kernel's binary cost 1.4-2.6 ms of instruction fetch from DRAM, 0.03 ms warm).
template <int G, bool FULL = true>
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
    if (FULL && a.mode == MODE_STEP) {
      taken = env_action(a.m, a.cfg, w, g, a.nsub, success);
    } else if (FULL && a.mode == MODE_DEBUG) {
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
        }
        g.sync();
        kinematics_trig(a.m, w, g);
        if (g.lane == 0) kinematics_lane0(a.m, w);  // sim.forward(): normalises free-joint quaternions in qpos
        g.sync();
      }
    } else {  // MODE_FORWARD: body positions of the current state (data.get_body_xpos, env.py:144,180,184)
      kinematics_trig(a.m, w, g);
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

// ------------------------------------------------------------------------------------------------ phase-locked general kernel
// The action kernel of the general path for STEP launches: one WARP per environment (32 lanes), a block of several warps
// that run the phases of a substep together (team barriers: the warps share the instruction cache instead of thrashing it)
// and share ONE queue of convex-convex narrowphase jobs: any warp refines any environment's candidate pair (the owner's
// poses, the serving warp's registers), so an environment with ten hull pairs near contact (the gripper on the pan,
// configs[2]) no longer refines them one after the other.  Same arithmetic as hsrb_step_kernel<32>: the stages are the
// functions of hsr_core.h, bitwise.
#define HSRB_LOCK_MAXWARPS 12
struct LockQueue {
  int cnt[4];                                    // [0] long jobs, [1] quick jobs, [2] next to serve
  int jobs[HSRB_LOCK_MAXWARPS * HSR_MAXJOBS];    // (owner warp << 16) | (slot << 8) | pair
};
__host__ __device__ inline size_t lock_tail_bytes() { return sizeof(LockQueue) + 16; }

#if defined(HSRB_STEP_LOCK_IMPL)   // defined by the translation unit that owns the kernel (hsrb_step_lock.cu, the emulated build)
__global__ void __launch_bounds__(32 * HSRB_LOCK_MAXWARPS) hsrb_step_lock_kernel(const __grid_constant__ KArgs a) {
  HSRB_DYN_SMEM(smem);
  typedef DevGrp<32> Grp;
  const Grp g;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const ModelT<float>& m = a.m;
  WS<float> w;
  ws_carve<float>(m, &w, smem + (size_t)wib * a.ws_bytes);
  LockQueue& Q = *reinterpret_cast<LockQueue*>(smem + (size_t)wpb * a.ws_bytes);
  if (threadIdx.x < 4) Q.cnt[threadIdx.x] = 0;
  __syncthreads();
  const int nobs = m.nq + m.nv;
  const int qcap = wpb * HSR_MAXJOBS;
  for (int env0 = blockIdx.x * wpb; env0 < a.n; env0 += gridDim.x * wpb) {
    const int env = env0 + wib;
    const bool valid = env < a.n;
    if (valid) {
      load_state<32>(a, w, g, env);
      for (int i = g.lane; i < m.nu; i += 32) w.ctrl[i] = a.ctrl ? a.ctrl[(size_t)env * m.nu + i] : 0.f;
    }
    g.sync();
    bool success = false, finished = !valid || a.nsub <= 0;
    int taken = 0;
    for (int sb = 0; sb < a.nsub; sb++) {
      if (__syncthreads_and(finished)) break;
      int nlimit = 0, nrow = 0, ncon = 0, it0 = 0, ls0 = 0;
      // PH: block barriers between ALL phases (not only around the job service), PL: one block vote per Newton pass - the
      // warps of the block then run the same code region at the same time (the instruction stream of a substep of this
      // kernel is several hundred KB; free-running warps spent half of their stall samples waiting for instructions).
      // a.opts bit 1 / bit 2 switch them off (experiments)
      const bool PH = !(a.opts & 2u), PL = !(a.opts & 4u);
      if (!finished) {
        HSR_PHASE_START(w, g);
        kinematics_trig(m, w, g);
        if (g.lane == 0) kinematics_lane0(m, w);
        g.sync();
        HSR_PHASE(w, g, PH_KIN);
      }
      if (PH) __syncthreads();
      if (!finished) {
        cdof_geoms(m, w, g);
        g.sync();
        mass_matrix(m, w, g);
        g.sync();
        HSR_PHASE(w, g, PH_CRB);
      }
      if (PH) __syncthreads();
      if (!finished) {
        if (g.lane == 0) smooth_lane0(m, w);
        g.sync();
        smooth_solve(m, w, g);
        HSR_PHASE(w, g, PH_SMOOTH);
      }
      if (PH) __syncthreads();
      if (!finished) {
        for (int j = 0; j < m.njnt; j++) {
          if (!m.jnt_limited[j] || m.jnt_type[j] == JNT_FREE) continue;
          const float q = w.qpos[m.jnt_qposadr[j]];
          if (q - m.jnt_range[2 * j] < 0) nlimit++;
          if (m.jnt_range[2 * j + 1] - q < 0) nlimit++;
        }
        nrow = nlimit;
        collision_cull(m, w, g);
        // queue the convex-convex candidates: long jobs (no cached separating direction) from the front, quick ones from the back
        if (g.lane == 0) {
          int kc = 0;
          for (int base = 0; base < m.npair; base += 32) {
            unsigned bb = w.cand[base >> 5];
            while (bb) {
              const int pk = base + __ffs((int)bb) - 1;
              bb &= bb - 1;
              if (m.pair_func[pk] != NP_CONVEX_CONVEX) continue;
              if (kc < HSR_MAXJOBS) {
                const int code = (wib << 16) | (kc << 8) | pk;
                if (w.sep[4 * pk + 3] == 1.f) Q.jobs[qcap - 1 - atomicAdd(&Q.cnt[1], 1)] = code;
                else Q.jobs[atomicAdd(&Q.cnt[0], 1)] = code;
              }
              kc++;
            }
          }
        }
      }
      __syncthreads();
      {
        // every warp serves jobs: the owner's poses and separating-direction cache, this warp's registers
        const int nlong = Q.cnt[0], njobs = nlong + Q.cnt[1];
        while (true) {
          int j = 0;
          if (g.lane == 0) j = atomicAdd(&Q.cnt[2], 1);
          j = __shfl_sync(0xffffffffu, j, 0);
          if (j >= njobs) break;
          const int code = j < nlong ? Q.jobs[j] : Q.jobs[qcap - 1 - (j - nlong)];
          const int pk = code & 255;
          WS<float> wo;
          ws_carve<float>(m, &wo, smem + (size_t)(code >> 16) * a.ws_bytes);
          Geom<float> A, B;
          load_geom(m, wo, m.pair_geom1[pk], A);
          load_geom(m, wo, m.pair_geom2[pk], B);
          GT depth = 0; V3<GT> dir = mk<GT>(0, 0, 1), pos = mk<GT>(0, 0, 0);
          const bool hit = mpr_penetration_inl(A, B, (GT)m.mpr_tolerance, m.mpr_iterations, g, depth, dir, pos, wo.sep + 4 * pk);
          if (g.lane == 0) {
            GT* r = wo.jres + 8 * ((code >> 8) & 255);
            r[0] = hit ? 1.0 : 0.0; r[1] = depth; r[2] = dir.x; r[3] = dir.y; r[4] = dir.z; r[5] = pos.x; r[6] = pos.y; r[7] = pos.z;
          }
          g.sync();
        }
      }
      __syncthreads();
      if (threadIdx.x < 3) Q.cnt[threadIdx.x] = 0;
      if (!finished) {
        ncon = collision_assemble(m, w, g, nrow);
        g.sync();
        HSR_PHASE(w, g, PH_COLLIDE);
      }
      if (PH) __syncthreads();
      if (!finished) {
        make_constraint(m, w, g, nlimit, ncon);
        it0 = w.wi[WI_ITER]; ls0 = w.wi[WI_LSEVAL];
        g.sync();
        if (g.lane == 0) { w.wi[WI_NCON] = ncon; w.wi[WI_NEFC] = nrow; w.wi[WI_NLIMIT] = nlimit; }
        g.sync();
        HSR_PHASE(w, g, PH_ROWS);
      }
      if (PH) __syncthreads();
      if (PL) {
        NewtonState<float> ns;
        bool act = !finished && newton_begin(m, w, g, nlimit, ncon, nrow, ns);
        const bool solved = act;
        while (__syncthreads_or(act)) {
          if (act) act = newton_pass(m, w, g, nlimit, ncon, nrow, ns);
        }
        if (solved) newton_end(m, w, g, nrow, ns);
      } else if (!finished) {
        solve_newton(m, w, g, nlimit, ncon, nrow);
      }
      if (!finished) {
        HSR_PHASE(w, g, PH_SOLVE);
        if (g.lane == 0) {
          w.wi[WI_SUMCON] += ncon; w.wi[WI_SUMEFC] += nrow;
          w.wi[WI_KFLOP] += algorithmic_flops(m, ncon, nrow, w.wi[WI_ITER] - it0, w.wi[WI_LSEVAL] - ls0, w.wi[WI_NPFLOP]);
        }
        g.sync();
      }
      __syncthreads();
      if (!finished) {
        euler_solve(m, w, g);
        if (g.lane == 0) euler_lane0(m, w);
        g.sync();
        HSR_PHASE(w, g, PH_EULER);
        taken++;
        if (goal_reached(m, a.cfg, w)) success = true;
        if (success || sb == a.nsub - 1) finished = true;
      }
    }
    if (valid) {
      store_state<32>(a, w, g, env);
      if (a.obs) for (int i = g.lane; i < nobs; i += 32) a.obs[(size_t)env * nobs + i] = w.qpos[i];
      if (g.lane == 0) {
        const int flags = w.wi[WI_FLAGS];
        if (a.reward) a.reward[env] = success ? 1.0f : 0.0f;
        if (a.done) a.done[env] = success ? 1 : 0;
        if (a.success) a.success[env] = success ? 1 : 0;
        if (a.taken) a.taken[env] = taken;
        if (a.bad) a.bad[env] = (unsigned char)flags;
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
    __syncthreads();
  }
}
#endif  // HSRB_STEP_LOCK_IMPL
cudaError_t hsrb_prepare_step_lock(size_t smem, int threads, int* blocks_per_sm);
cudaError_t hsrb_launch_step_lock(const KArgs& a, int grid, int threads, size_t smem, cudaStream_t s);

// per-G entry points, one translation unit each (parallel compilation)
#define HSRB_DECL_G(G)                                                \
  cudaError_t hsrb_prepare_step_##G(size_t smem, int* blocks_per_sm); \
  cudaError_t hsrb_launch_step_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s); \
  cudaError_t hsrb_launch_aux_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s);
HSRB_DECL_G(4) HSRB_DECL_G(8) HSRB_DECL_G(16) HSRB_DECL_G(32)

#define HSRB_DEFINE_G(G)                                                                                               \
  cudaError_t hsrb_prepare_step_##G(size_t smem, int* bps) {                                                           \
    cudaError_t e = cudaFuncSetAttribute(hsrb_step_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); /* per function, shared by all handles */ \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hsrb_step_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    if (e != cudaSuccess) return e;                                                                                    \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, hsrb_step_kernel<G>, 32, smem);                          \
  }                                                                                                                    \
  cudaError_t hsrb_launch_step_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s) {                            \
    hsrb_step_kernel<G><<<grid, 32, smem, s>>>(a);                                                                     \
    return cudaGetLastError();                                                                                         \
  }                                                                                                                    \
  cudaError_t hsrb_launch_aux_##G(const KArgs& a, int grid, size_t smem, cudaStream_t s) { /* reset / forward launches */ \
    hsrb_step_kernel<G, false><<<grid, 32, smem, s>>>(a);                                                              \
    return cudaGetLastError();                                                                                         \
  }
