// TEST INFRASTRUCTURE: runs hsr_env_b200/csrc/hsrb_push.cuh (the CUDA action kernel, unchanged source) on the CPU
// through the SIMT emulator (simt_emu.h) so that its logic can be checked against the oracle on a box without a GPU.
// Build: see tests/simt_emu/build.py.  Never linked into the product.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define HSRB_WPE_IMPL 1
#include "../../hsr_env_b200/csrc/hsrb_wpe.cuh"

namespace {

template <int G, int NV>
void run(const KArgs& a, const PushInfo& fi, int grid, int threads, size_t smem) {
  emu::launch(grid, threads, smem, [&]() { hsrb_push_kernel<G, NV>(a, fi); });
}

}  // namespace

extern "C" int emu_push_step(const void* blob, size_t bytes, int G, int threads, int n, int nsub, const double* qpos,
                             const double* qvel, const double* warm, const double* ctrl, const double* mocap, int has_goal,
                             double geofence, double* qpos_o, double* qvel_o, double* warm_o, int* taken,
                             unsigned char* success, unsigned char* flags, long long* stats_out) {
  HostModel<float> hm;
  std::string err;
  if (!hm.parse(blob, bytes, err)) { fprintf(stderr, "emu_push_step: %s\n", err.c_str()); return -1; }
  PushInfo fi;
  PushTables tb;
  char why[128];
  if (!push_fill_info(hm.m, fi, tb, why, sizeof(why))) { fprintf(stderr, "emu_push_step: %s\n", why); return -2; }
  tb.point(fi, (const unsigned char*)tb.tab.data());
  const ModelT<float>& m = hm.m;
  const int S = m.nq + 2 * m.nv + 3, nobs = m.nq + m.nv;
  std::vector<float> state((size_t)n * S), c32((size_t)n * (m.nu > 0 ? m.nu : 1)), obs((size_t)n * nobs), reward(n);
  std::vector<unsigned char> done(n), succ(n), bad(n);
  std::vector<int> tk(n);
  std::vector<unsigned long long> stats(ST_COUNT, 0);
  for (int e = 0; e < n; e++) {
    float* st = state.data() + (size_t)e * S;
    for (int i = 0; i < m.nq; i++) st[i] = (float)qpos[(size_t)e * m.nq + i];
    for (int i = 0; i < m.nv; i++) { st[m.nq + i] = (float)qvel[(size_t)e * m.nv + i]; st[m.nq + m.nv + i] = (float)warm[(size_t)e * m.nv + i]; }
    for (int i = 0; i < 3; i++) st[m.nq + 2 * m.nv + i] = (float)mocap[(size_t)e * 3 + i];
    for (int i = 0; i < m.nu; i++) c32[(size_t)e * m.nu + i] = (float)ctrl[(size_t)e * m.nu + i];
  }
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.m = m;
  a.m.ncon_max = PUSH_MAXCON; a.m.nefc_max = 2 + PUSH_ROWS;
  a.cfg.has_goal = has_goal; a.cfg.geofence = (float)geofence; a.cfg.qidx0 = 0; a.cfg.qidx1 = 2;
  a.n = n; a.S = S; a.nsub = nsub; a.mode = MODE_STEP;
  if (const char* o = getenv("HSRB_OPTS")) a.opts = (unsigned)strtoul(o, nullptr, 0);
  a.ws_bytes = (unsigned)(G <= 2 ? wpe::slice_bytes() : push::carve(a.m, nullptr, nullptr));
  a.state = state.data(); a.ctrl = c32.data(); a.obs = obs.data(); a.reward = reward.data(); a.done = done.data();
  a.success = succ.data(); a.taken = tk.data(); a.bad = bad.data(); a.stats = stats.data();
  if (G == 1 || G == 2) {   // warp-per-environment kernel (hsrb_wpe.cuh): `threads` / 32 environments per block; G = 2: phase-locked variant
    const int wpb = threads / 32;
    const int grid1 = (n + wpb - 1) / wpb;
    a.m.nefc_max = WPE_MAXROW;
    if (G == 1) emu::launch(grid1, threads, (size_t)a.ws_bytes * wpb + wpe::shared_tail(a.m), [&]() { hsrb_wpe_kernel_t<false>(a, fi); });
    else emu::launch(grid1, threads, (size_t)a.ws_bytes * wpb + wpe::shared_tail(a.m), [&]() { hsrb_wpe_kernel_t<true>(a, fi); });
  }
  const int epb = G <= 2 ? 1 : threads / G;
  const int grid = (n + epb - 1) / epb;
  const size_t smem = (size_t)a.ws_bytes * epb + push::shared_tail(a.m);
  const int NV = m.nv;
#define RUN(G_) { if (NV == 8) run<G_, 8>(a, fi, grid, threads, smem); else run<G_, 2>(a, fi, grid, threads, smem); }
  if (G == 8) RUN(8) else if (G == 16) RUN(16) else if (G == 32) RUN(32) else if (G > 2) return -3;
#undef RUN
  for (int e = 0; e < n; e++) {
    const float* st = state.data() + (size_t)e * S;
    for (int i = 0; i < m.nq; i++) qpos_o[(size_t)e * m.nq + i] = st[i];
    for (int i = 0; i < m.nv; i++) { qvel_o[(size_t)e * m.nv + i] = st[m.nq + i]; warm_o[(size_t)e * m.nv + i] = st[m.nq + m.nv + i]; }
    taken[e] = tk[e]; success[e] = succ[e]; flags[e] = bad[e];
  }
  if (stats_out) for (int i = 0; i < ST_COUNT; i++) stats_out[i] = (long long)stats[i];
  return 0;
}
