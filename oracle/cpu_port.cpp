// ORACLE-SIDE CPU PORT (test infrastructure + CPU baseline; never on the product path).
//
// Host build (g++, G=1 "group") of the same templated substep the CUDA kernels instantiate
// (hsr_env_b200/csrc/hsr_core.h), in fp64 and fp32.  It serves three purposes:
//   1. a second, compiled checker beside the independent numpy restatement (oracle/mjstep.py):
//      tests assert  numpy-fp64 == port-fp64 (tight)  and  CUDA-fp32 ~= port-fp64 (1e-4);
//   2. the fp32 instantiation predicts the GPU's rounding behaviour on a box without a GPU;
//   3. the multi-threaded fp64 instantiation is the "port" CPU baseline of bench.py
//      (one environment per thread at a time, all host cores) - a stand-in for the reference's
//      mujoco-py loop (/root/reference/hsr/env.py:115-135), which cannot run here.
//      It is NOT MuJoCo and is never labelled as such.
// PARITY UNPINNED (see oracle/mjstep.py header).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define HSR_SCAN_NOISE 1
#include "../hsr_env_b200/csrc/hsr_core.h"

using namespace hsr;

namespace {

struct Handle {
  HostModel<double> md;
  HostModel<float> mf;
  EnvCfg<double> cfgd;
  EnvCfg<float> cfgf;
};

template <typename T> HostModel<T>& model_of(Handle* h);
template <> HostModel<double>& model_of<double>(Handle* h) { return h->md; }
template <> HostModel<float>& model_of<float>(Handle* h) { return h->mf; }
template <typename T> EnvCfg<T>& cfg_of(Handle* h);
template <> EnvCfg<double>& cfg_of<double>(Handle* h) { return h->cfgd; }
template <> EnvCfg<float>& cfg_of<float>(Handle* h) { return h->cfgf; }

template <typename T>
void run_range(Handle* h, int lo, int hi, int nsub, const double* qpos, const double* qvel, const double* warm,
               const double* ctrl, const double* mocap, double* qpos_o, double* qvel_o, double* warm_o, int* taken,
               unsigned char* success, int* flags, double* dbg, long long* counters) {
  const ModelT<T>& m = model_of<T>(h).m;
  const EnvCfg<T>& cfg = cfg_of<T>(h);
  std::vector<unsigned char> buf(ws_carve<T>(m, nullptr, nullptr) + 64);
  WS<T> w;
  ws_carve<T>(m, &w, buf.data());
  HostGrp g;
  size_t dsz = debug_size(m);
  for (int e = lo; e < hi; e++) {
    for (int i = 0; i < m.nq; i++) w.qpos[i] = (T)qpos[(size_t)e * m.nq + i];
    for (int i = 0; i < m.nv; i++) { w.qvel[i] = (T)qvel[(size_t)e * m.nv + i]; w.warm[i] = (T)warm[(size_t)e * m.nv + i]; }
    for (int i = 0; i < m.nu; i++) w.ctrl[i] = (T)ctrl[(size_t)e * m.nu + i];
    for (int i = 0; i < 3; i++) w.mocap[i] = (T)mocap[(size_t)e * 3 + i];
    for (int i = 0; i < WI_COUNT; i++) w.wi[i] = 0;
    bool ok = false;
    int t;
    if (dbg) {
      // one forward pass, dump every stage, then integrate
      forward(m, w, g);
      T q0[64], v0[32];
      for (int i = 0; i < m.nq; i++) q0[i] = w.qpos[i];
      for (int i = 0; i < m.nv; i++) v0[i] = w.qvel[i];
      euler_solve(m, w, g);
      euler_lane0(m, w);
      debug_dump(m, w, dbg + (size_t)e * dsz);
      ok = goal_reached(m, cfg, w);
      t = 1;
    } else {
      t = env_action(m, cfg, w, g, nsub, ok);
    }
    for (int i = 0; i < m.nq; i++) qpos_o[(size_t)e * m.nq + i] = (double)w.qpos[i];
    for (int i = 0; i < m.nv; i++) { qvel_o[(size_t)e * m.nv + i] = (double)w.qvel[i]; warm_o[(size_t)e * m.nv + i] = (double)w.warm[i]; }
    if (taken) taken[e] = t;
    if (success) success[e] = ok ? 1 : 0;
    if (flags) flags[e] = w.wi[WI_FLAGS];
    if (counters) {
      counters[(size_t)e * 4 + 0] = w.wi[WI_ITER]; counters[(size_t)e * 4 + 1] = w.wi[WI_NARROW];
      counters[(size_t)e * 4 + 2] = w.wi[WI_LSEVAL]; counters[(size_t)e * 4 + 3] = w.wi[WI_KFLOP];
    }
  }
}

template <typename T>
int run(Handle* h, int n, int nsub, int nthreads, const double* qpos, const double* qvel, const double* warm,
        const double* ctrl, const double* mocap, double* qpos_o, double* qvel_o, double* warm_o, int* taken,
        unsigned char* success, int* flags, double* dbg, long long* counters) {
  if (nthreads <= 1) {
    run_range<T>(h, 0, n, nsub, qpos, qvel, warm, ctrl, mocap, qpos_o, qvel_o, warm_o, taken, success, flags, dbg, counters);
    return 0;
  }
  std::vector<std::thread> th;
  std::atomic<int> next(0);
  const int chunk = 1;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&]() {
      while (true) {
        int lo = next.fetch_add(chunk);
        if (lo >= n) break;
        int hi = lo + chunk < n ? lo + chunk : n;
        run_range<T>(h, lo, hi, nsub, qpos, qvel, warm, ctrl, mocap, qpos_o, qvel_o, warm_o, taken, success, flags, dbg, counters);
      }
    });
  for (auto& t : th) t.join();
  return 0;
}

template <typename T> void set_cfg(EnvCfg<T>& c, int has_goal, int has_block, const double* goal_lohi, const double* block_lohi,
                                   double geofence, double min_sep, int qidx0, int qidx1) {
  c.ngoal = 0;
  c.has_goal = has_goal; c.has_block = has_block; c.qidx0 = qidx0; c.qidx1 = qidx1; c.geofence = (T)geofence; c.min_sep = (T)min_sep;
  for (int k = 0; k < 3; k++) { c.goal_lo[k] = (T)goal_lohi[k]; c.goal_hi[k] = (T)goal_lohi[3 + k]; }
  for (int k = 0; k < 4; k++) { c.block_lo[k] = (T)block_lohi[k]; c.block_hi[k] = (T)block_lohi[4 + k]; }
}

template <typename T> void set_goal_list(EnvCfg<T>& c, const ModelT<T>& m, int ngoal, const int* a, const int* b, const double* dist,
                                         const double* point_lohi, const double* fixed, int nfixed) {
  c.ngoal = ngoal; c.has_goal = ngoal > 0; c.has_block = 0;
  for (int k = 0; k < ngoal; k++) { c.goal_a[k] = a[k]; c.goal_b[k] = b[k]; c.goal_dist[k] = (T)dist[k]; }
  for (int k = 0; k < 3; k++) {
    c.goal_lo[k] = point_lohi ? (T)point_lohi[k] : m.mocap_pos0[k];
    c.goal_hi[k] = point_lohi ? (T)point_lohi[3 + k] : m.mocap_pos0[k];
  }
  for (int k = 0; k < nfixed; k++) for (int i = 0; i < 3; i++) c.fixed_pt[k][i] = (T)fixed[3 * k + i];
}
template <typename T> void set_starts(EnvCfg<T>& c, int n, const int* adr, const int* width, const double* lo, const double* hi) {
  c.nstart = n;
  for (int s = 0; s < n; s++) {
    c.start_adr[s] = adr[s]; c.start_width[s] = width[s];
    for (int k = 0; k < 7; k++) { c.start_lo[s][k] = k < width[s] ? (T)lo[7 * s + k] : T(0); c.start_hi[s][k] = k < width[s] ? (T)hi[7 * s + k] : T(0); }
  }
}

}  // namespace

extern "C" {

void* hsrp_create(const void* blob, size_t bytes) {
  Handle* h = new Handle();
  std::string err;
  if (!h->md.parse(blob, bytes, err) || !h->mf.parse(blob, bytes, err)) {
    fprintf(stderr, "hsrp_create: %s\n", err.c_str());
    delete h;
    return nullptr;
  }
  memset(&h->cfgd, 0, sizeof(h->cfgd)); memset(&h->cfgf, 0, sizeof(h->cfgf));
  double z3[6] = {0, 0, 0, 0, 0, 0}, z4[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  set_cfg(h->cfgd, 0, 0, z3, z4, 0.0, 0.0, 0, 2);
  set_cfg(h->cfgf, 0, 0, z3, z4, 0.0, 0.0, 0, 2);
  return h;
}

void hsrp_destroy(void* h) { delete (Handle*)h; }

void hsrp_set_caps(void* hv, int ncon_max, int nefc_max) {
  Handle* h = (Handle*)hv;
  h->md.m.ncon_max = h->mf.m.ncon_max = ncon_max;
  h->md.m.nefc_max = h->mf.m.nefc_max = nefc_max;
}

void hsrp_dims(void* hv, int* out) {
  const ModelT<double>& m = ((Handle*)hv)->md.m;
  out[0] = m.nq; out[1] = m.nv; out[2] = m.nu; out[3] = m.nbody; out[4] = m.ncon_max; out[5] = m.nefc_max;
  out[6] = (int)debug_size(m); out[7] = (int)ws_carve<float>(((Handle*)hv)->mf.m, nullptr, nullptr);
}

void hsrp_set_goals(void* hv, int has_goal, int has_block, const double* goal_lohi, const double* block_lohi, double geofence,
                    double min_sep, int qidx0, int qidx1) {
  Handle* h = (Handle*)hv;
  set_cfg(h->cfgd, has_goal, has_block, goal_lohi, block_lohi, geofence, min_sep, qidx0, qidx1);
  set_cfg(h->cfgf, has_goal, has_block, goal_lohi, block_lohi, geofence, min_sep, qidx0, qidx1);
}

// mirrors of hsrb_set_goal_list / hsrb_set_starts (include/hsrb.h)
void hsrp_set_goal_list(void* hv, int ngoal, const int* a, const int* b, const double* dist, const double* point_lohi,
                        const double* fixed, int nfixed) {
  Handle* h = (Handle*)hv;
  set_goal_list(h->cfgd, h->md.m, ngoal, a, b, dist, point_lohi, fixed, nfixed);
  set_goal_list(h->cfgf, h->mf.m, ngoal, a, b, dist, point_lohi, fixed, nfixed);
}
void hsrp_set_starts(void* hv, int n, const int* adr, const int* width, const double* lo, const double* hi) {
  Handle* h = (Handle*)hv;
  set_starts(h->cfgd, n, adr, width, lo, hi);
  set_starts(h->cfgf, n, adr, width, lo, hi);
}

// fp32-rounding-sized noise on the hull-vertex scans of the calling thread (0 = off): see HSR_SCAN_NOISE in hsr_core.h
void hsrp_set_scan_noise(unsigned long long seed) { hsr_scan_noise_seed = seed; }

// Step n environments by up to nsub substeps each (teacher-forced from the given states).
// dbg != NULL: exactly one substep, per-stage dump (debug_size doubles per env).
int hsrp_step(void* hv, int use_float, int n, int nsub, int nthreads, const double* qpos, const double* qvel,
              const double* warm, const double* ctrl, const double* mocap, double* qpos_o, double* qvel_o,
              double* warm_o, int* taken, unsigned char* success, int* flags, double* dbg, long long* counters) {
  Handle* h = (Handle*)hv;
  if (use_float)
    return run<float>(h, n, nsub, nthreads, qpos, qvel, warm, ctrl, mocap, qpos_o, qvel_o, warm_o, taken, success, flags, dbg, counters);
  return run<double>(h, n, nsub, nthreads, qpos, qvel, warm, ctrl, mocap, qpos_o, qvel_o, warm_o, taken, success, flags, dbg, counters);
}

// Philox reset of env `env_id` at `episode` (same stream as the CUDA reset kernel): writes qpos[nq], mocap[3].
void hsrp_reset(void* hv, unsigned long long seed, unsigned env_id, unsigned episode, double* qpos, double* mocap) {
  Handle* h = (Handle*)hv;
  const ModelT<float>& m = h->mf.m;
  std::vector<unsigned char> buf(ws_carve<float>(m, nullptr, nullptr) + 64);
  WS<float> w;
  ws_carve<float>(m, &w, buf.data());
  reset_lane0(m, h->cfgf, w, seed, env_id, episode);
  for (int i = 0; i < m.nq; i++) qpos[i] = (double)w.qpos[i];
  for (int i = 0; i < 3; i++) mocap[i] = (double)w.mocap[i];
}

}  // extern "C"
