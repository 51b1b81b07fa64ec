// Register-resident action kernel for the "sliding base + at most one free box" models (README block-push:
// --use-dof slide_x slide_y, --n-blocks 0/1; BASELINE.json configs[0], [1] and [3]).
//
// Same substep as the general kernel (hsrb_kernels.cuh / hsr_core.h), restructured for the
// hardware instead of for generality:
//   * G = 8 lanes of a warp own one environment; four environments share a warp.
//   * State (qpos, qvel, qacc_warmstart, ctrl, goal) is replicated in the registers of the 8 lanes for all 300
//     substeps; it touches HBM once per action.
//   * The mass matrix of this model family is constant and diagonal (slides on one world-attached body, free
//     box with principal-axis frame), so CRB / RNE / factorisation collapse to closed forms.
//   * One contact per lane: the lane keeps its contact's 6 x NV Jacobian block, regularisers, reference
//     accelerations and cone state in registers.  Cross-contact sums (gradient, Hessian, line-search
//     derivatives) are xor-shuffle butterflies over the 8 lanes, after which every lane holds the full NV x NV
//     Hessian and redundantly runs the (fully unrolled) Cholesky factorisation and solve: no shared-memory
//     traffic and no divergence inside the Newton iteration.
//   * Poses and the narrowphase (double precision, see hsr_core.h "Geometry precision") reuse the general
//     device functions on a compact per-environment shared-memory workspace.
//
// Replaces the loop over sim.step() in HSREnv.step (/root/reference/hsr/env.py:115-135).
#pragma once
#include "hsrb_kernels.cuh"

#define HSRB_FAST_MAXCON 8

// Model constants of the fast path, filled on the host (hsrb_api.cu: fill_fast_info) and passed by value as a
// kernel parameter: every read is a uniform constant-bank load.
struct FastInfo {
  int nv;                 // 2 (no block) or 8
  int robot_body, block_body;
  float Mdiag[8];         // [m_r, m_r, m_b, m_b, m_b, I1, I2, I3]
  float damp[8];
  float axis[2][3];       // slide axes in the world frame
  float gq[2];            // generalised gravity force on the slides: m_r * (g . axis)
  float q0[2];
  int act_n; int act_dof[2]; int act_q[2];
  float kp[2], gear[2], cr_lo[2], cr_hi[2], fr_lo[2], fr_hi[2];
  int ctrllimited[2], forcelimited[2];
  int limited[2];
  float range[2][2], lsolref[2][2], lsolimp[2][5], linvw[2];
  float gravity[3];
};

// compact workspace: what kinematics_lane0 / cdof_geoms / collision touch
__host__ __device__ inline size_t ws_carve_fast(const ModelT<float>& m, WS<float>* w, unsigned char* base) {
  size_t off = 0;
  WS<float> dummy;
  if (!w) w = &dummy;
#define CARVE(field, type, n) { w->field = (type*)(base + off); off += sizeof(type) * (size_t)(n); }
  int nv = m.nv, nb = m.nbody, nc = HSRB_FAST_MAXCON;
  CARVE(xpos, GT, nb * 3) CARVE(xquat, GT, nb * 4) CARVE(xmat, GT, nb * 9) CARVE(xipos, GT, nb * 3)
  CARVE(anchor, GT, m.njnt * 3) CARVE(axis, GT, m.njnt * 3) CARVE(gpos, GT, m.ngeom * 3) CARVE(com, GT, nb * 3)
  CARVE(qpos, float, m.nq) CARVE(cdof, float, nv * 6) CARVE(cinert, float, nb * 10) CARVE(binert, float, nb * 10)
  CARVE(gaabb, float, m.ngeom * 3)
  CARVE(con_dist, float, nc) CARVE(con_pos, float, nc * 3) CARVE(con_frame, float, nc * 9)
  CARVE(con_pair, int, nc) CARVE(con_adr, int, nc) CARVE(wi, int, WI_COUNT)
#undef CARVE
  off += (16 - off % 16) % 16;
  return off;
}

namespace fast {

__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

}  // namespace fast

template <int G, int NV>
__global__ void __launch_bounds__(256) hsrb_fast_kernel(const __grid_constant__ KArgs a, const __grid_constant__ FastInfo fi) {
  extern __shared__ __align__(16) unsigned char smem[];
  DevGrp<G> g;
  constexpr int NT = NV * (NV + 1) / 2;
  constexpr bool HASB = NV == 8;
  const int epb = blockDim.x / G;
  const int gi = threadIdx.x / G;
  WS<float> w;
  ws_carve_fast(a.m, &w, smem + (size_t)gi * a.ws_bytes);
  const ModelT<float>& m = a.m;
  const int nq = m.nq;
  const float dt = m.timestep;
  const float scale = 1.0f / (m.meaninertia * (float)(NV > 1 ? NV : 1));

  for (int env = blockIdx.x * epb + gi; env < a.n; env += gridDim.x * epb) {
    // ------------------------------------------------------------------ state -> registers (replicated)
    float qpos[9], qvel[8], warm[8], ctrl[2], mocap[3];
#pragma unroll
    for (int i = 0; i < 9; i++) qpos[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { qvel[i] = 0.f; warm[i] = 0.f; }
    {
      const float* st = a.state + (size_t)env * a.S;
#pragma unroll
      for (int i = 0; i < (HASB ? 9 : 2); i++) qpos[i] = st[i];
#pragma unroll
      for (int i = 0; i < NV; i++) { qvel[i] = st[nq + i]; warm[i] = st[nq + NV + i]; }
#pragma unroll
      for (int i = 0; i < 3; i++) mocap[i] = st[nq + 2 * NV + i];
      ctrl[0] = ctrl[1] = 0.f;
      if (a.ctrl)
        for (int i = 0; i < fi.act_n; i++) ctrl[i] = a.ctrl[(size_t)env * m.nu + i];
      for (int i = g.lane; i < WI_COUNT; i += G) w.wi[i] = 0;
    }
    int n_iter = 0, n_ls = 0, sumcon = 0, sumefc = 0, kflop = 0, flags = 0;
    bool success = false;
    int taken = 0;
    g.sync();

    for (int sub = 0; sub < a.nsub; sub++) {
      // ---------------------------------------------------------------- poses + collision (general code, double)
      for (int i = g.lane; i < nq; i += G) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < (HASB ? 9 : 2); k++) if (k == i) v = qpos[k];
        w.qpos[i] = v;
      }
      g.sync();
      HSR_PHASE_START(w, g);
      if (g.lane == 0) kinematics_lane0(m, w);
      g.sync();
      HSR_PHASE(w, g, PH_KIN);
      if (HASB) {
#pragma unroll
        for (int k = 5; k < 9; k++) qpos[k] = w.qpos[k];  // quaternion normalised in place
      }
      cdof_geoms(m, w, g);
      g.sync();
      HSR_PHASE(w, g, PH_CRB);
      // active joint limits (redundant in every lane)
      float lim_sg[2], lim_D[2], lim_aref[2];
      int nlimit = 0;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        lim_sg[j] = 0.f; lim_D[j] = 0.f; lim_aref[j] = 0.f;
        if (fi.limited[j]) {
          float q = qpos[j];
          float dlo = q - fi.range[j][0], dhi = fi.range[j][1] - q;
          float dist = 0.f, sg = 0.f;
          if (dlo < 0) { dist = dlo; sg = 1.f; }
          else if (dhi < 0) { dist = dhi; sg = -1.f; }
          if (sg != 0.f) {
            GT R; float ar;
            row_params(m, fi.lsolref[j], fi.lsolimp[j], dist, sg * qvel[j], fi.linvw[j], false, R, ar);
            lim_sg[j] = sg; lim_D[j] = (float)(GT(1) / R); lim_aref[j] = ar;
            nlimit++;
          }
        }
      }
      int nrow = nlimit;
      int ncon = collision(m, w, g, nrow);
      g.sync();
      if (ncon > HSRB_FAST_MAXCON) { ncon = HSRB_FAST_MAXCON; }
      flags |= w.wi[WI_FLAGS];
      HSR_PHASE(w, g, PH_COLLIDE);

      // ---------------------------------------------------------------- smooth forces (closed form, B.6)
      float qs[8], as[8];  // qfrc_smooth, qacc_smooth
#pragma unroll
      for (int j = 0; j < 2; j++) qs[j] = -fi.damp[j] * qvel[j] + fi.gq[j];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        if (k < fi.act_n) {
          float c = ctrl[k];
          if (fi.ctrllimited[k]) c = fminf(fmaxf(c, fi.cr_lo[k]), fi.cr_hi[k]);
          float qa = fi.act_q[k] == 0 ? qpos[0] : qpos[1];
          float fo = fi.kp[k] * c - fi.kp[k] * fi.gear[k] * qa;
          if (fi.forcelimited[k]) fo = fminf(fmaxf(fo, fi.fr_lo[k]), fi.fr_hi[k]);
          float ga = fi.gear[k] * fo;
          if (fi.act_dof[k] == 0) qs[0] += ga; else qs[1] += ga;
        }
      }
      if (HASB) {
        // free box: gravity on the translational dofs, -w x (I w) on the body-frame rotational dofs
#pragma unroll
        for (int k = 0; k < 3; k++) qs[2 + k] = fi.Mdiag[2] * fi.gravity[k] - fi.damp[2 + k] * qvel[2 + k];
        float Iw0 = fi.Mdiag[5] * qvel[5], Iw1 = fi.Mdiag[6] * qvel[6], Iw2 = fi.Mdiag[7] * qvel[7];
        qs[5] = -(qvel[6] * Iw2 - qvel[7] * Iw1) - fi.damp[5] * qvel[5];
        qs[6] = -(qvel[7] * Iw0 - qvel[5] * Iw2) - fi.damp[6] * qvel[6];
        qs[7] = -(qvel[5] * Iw1 - qvel[6] * Iw0) - fi.damp[7] * qvel[7];
      }
#pragma unroll
      for (int i = 0; i < NV; i++) as[i] = qs[i] / fi.Mdiag[i];

      // ---------------------------------------------------------------- constraint rows: one contact per lane (B.4/B.5)
      float J[6][NV], D[6], aref[6], jar[6], jv[6], f[6], fri[5], mu = 0.f;
      int dim = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        D[r] = 0.f; aref[r] = 0.f; jar[r] = 0.f; jv[r] = 0.f; f[r] = 0.f;
#pragma unroll
        for (int d = 0; d < NV; d++) J[r][d] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 5; k++) fri[k] = 1.f;
      if (g.lane < ncon) {
        const int c = g.lane;
        const int pk = w.con_pair[c];
        dim = m.pair_condim[pk];
        const int g1 = m.pair_geom1[pk], g2 = m.pair_geom2[pk];
        const int b1 = m.geom_body[g1], b2 = m.geom_body[g2];
        const float sr = (float)(b2 == fi.robot_body) - (float)(b1 == fi.robot_body);
        const float sb = (float)(b2 == fi.block_body) - (float)(b1 == fi.block_body);
        float fr[9];
#pragma unroll
        for (int k = 0; k < 9; k++) fr[k] = w.con_frame[9 * c + k];
#pragma unroll
        for (int k = 0; k < 5; k++) fri[k] = m.pair_friction[5 * pk + k];
        // robot slides: translational rows only
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int j = 0; j < 2; j++)
            J[r][j] = sr * (fr[3 * r] * fi.axis[j][0] + fr[3 * r + 1] * fi.axis[j][1] + fr[3 * r + 2] * fi.axis[j][2]);
        if (HASB) {
          const GT* Rb = w.xmat + 9 * fi.block_body;
          const GT* xb = w.xpos + 3 * fi.block_body;
          float rel[3];
#pragma unroll
          for (int k = 0; k < 3; k++) rel[k] = (float)((GT)w.con_pos[3 * c + k] - xb[k]);
#pragma unroll
          for (int k = 0; k < 3; k++) {
            float ax0 = (float)Rb[k], ax1 = (float)Rb[3 + k], ax2 = (float)Rb[6 + k];  // body axis k in the world frame
            float jp0 = ax1 * rel[2] - ax2 * rel[1], jp1 = ax2 * rel[0] - ax0 * rel[2], jp2 = ax0 * rel[1] - ax1 * rel[0];
#pragma unroll
            for (int r = 0; r < 3; r++) {
              J[r][2 + k] = sb * fr[3 * r + k];
              J[r][5 + k] = sb * (fr[3 * r] * jp0 + fr[3 * r + 1] * jp1 + fr[3 * r + 2] * jp2);
              J[3 + r][5 + k] = sb * (fr[3 * r] * ax0 + fr[3 * r + 1] * ax1 + fr[3 * r + 2] * ax2);
            }
          }
        }
        const float diag = m.geom_invweight[g1] + m.geom_invweight[g2];
        GT R0 = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
          if (r < dim) {
            float vel = 0.f;
#pragma unroll
            for (int d = 0; d < NV; d++) vel += J[r][d] * qvel[d];
            GT R; float ar;
            row_params(m, m.pair_solref + 2 * pk, m.pair_solimp + 5 * pk, r == 0 ? w.con_dist[c] : 0.f, vel, diag, r > 0, R, ar);
            if (r == 0) R0 = R;
            else if (r == 1) R = R0 / (GT)m.impratio;
            else R = (R0 / (GT)m.impratio) * (GT)fri[0] * (GT)fri[0] / ((GT)fri[r - 1] * (GT)fri[r - 1]);
            D[r] = (float)(GT(1) / R); aref[r] = ar;
          } else {
#pragma unroll
            for (int d = 0; d < NV; d++) J[r][d] = 0.f;
          }
        }
        mu = dim > 1 ? (float)((GT)fri[0] * sqrt((R0 / (GT)m.impratio) / R0)) : fri[0];
      }
      const int nefc = nrow;
      sumcon += ncon; sumefc += nefc;
      HSR_PHASE(w, g, PH_ROWS);

      // ---------------------------------------------------------------- Newton solver (B.7), registers only
      float x[8], Ma[8], qfc[8];  // qacc, M qacc, J^T f
#pragma unroll
      for (int i = 0; i < NV; i++) qfc[i] = 0.f;
      int it = 0, ls_used = 0;

      // cone state of this lane's contact for the rows in jar[]: cost, and (full) forces + zone
      auto update = [&](bool full, int& zone, float& cN, float& cT) -> float {
        float cost = 0.f;
        zone = 0; cN = 0.f; cT = 0.f;
        if (dim == 0) return 0.f;
        if (dim == 1) {
          zone = jar[0] < 0 ? 1 : 0;
        } else {
          cN = jar[0] * mu;
          float tt = 0.f;
#pragma unroll
          for (int j = 1; j < 6; j++) if (j < dim) { float u = jar[j] * fri[j - 1]; tt += u * u; }
          cT = sqrtf(tt);
          if (cN >= mu * cT || (cT <= 0 && cN >= 0)) zone = 0;
          else if (mu * cN + cT <= 0 || (cT <= 0 && cN < 0)) zone = 1;
          else zone = 2;
        }
        if (zone == 0) {
          if (full) {
#pragma unroll
            for (int r = 0; r < 6; r++) f[r] = 0.f;
          }
        } else if (zone == 1) {
#pragma unroll
          for (int r = 0; r < 6; r++) if (r < dim) {
            cost += 0.5f * D[r] * jar[r] * jar[r];
            if (full) f[r] = -D[r] * jar[r];
          }
        } else {
          float Dm = D[0] / (mu * mu * (1 + mu * mu));
          float NTv = cN - mu * cT;
          cost += 0.5f * Dm * NTv * NTv;
          if (full) {
            float f0 = -Dm * NTv * mu;
            f[0] = f0;
#pragma unroll
            for (int j = 1; j < 6; j++) if (j < dim) { float U = jar[j] * fri[j - 1]; f[j] = -f0 / cT * U * fri[j - 1]; }
          }
        }
        return cost;
      };
      // limit rows (redundant): cost and, optionally, force added to qfc / Hessian diagonal
      auto limit_cost = [&](const float* xx) -> float {
        float cst = 0.f;
#pragma unroll
        for (int j = 0; j < 2; j++) if (lim_sg[j] != 0.f) {
          float jr = lim_sg[j] * xx[j] - lim_aref[j];
          if (jr < 0) cst += 0.5f * lim_D[j] * jr * jr;
        }
        return cst;
      };
      auto total_cost = [&](const float* xx) -> float {
        // jar = J x - aref for this lane's contact, then the cone cost; Gauss term; limits (lane 0 only counts them)
#pragma unroll
        for (int r = 0; r < 6; r++) {
          float s = -aref[r];
#pragma unroll
          for (int d = 0; d < NV; d++) s += J[r][d] * xx[d];
          jar[r] = s;
        }
        int z; float cn, ct;
        float c = update(false, z, cn, ct);
        c = g.sum(c);
        float gauss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++) { float mx = fi.Mdiag[i] * xx[i]; gauss += 0.5f * (mx - qs[i]) * (xx[i] - as[i]); }
        return c + gauss + limit_cost(xx);
      };

      if (nefc == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) x[i] = as[i];
      } else {
        float cw = total_cost(warm);
        float cs = total_cost(as);
        const bool use_warm = cw <= cs;
#pragma unroll
        for (int i = 0; i < NV; i++) { x[i] = use_warm ? warm[i] : as[i]; Ma[i] = fi.Mdiag[i] * x[i]; }
#pragma unroll
        for (int r = 0; r < 6; r++) {
          float s = -aref[r];
#pragma unroll
          for (int d = 0; d < NV; d++) s += J[r][d] * x[d];
          jar[r] = s;
        }
        int zone; float cN, cT;
        float cost;
        {
          float c = update(true, zone, cN, cT);
          float gauss = 0.f;
#pragma unroll
          for (int i = 0; i < NV; i++) gauss += 0.5f * (Ma[i] - qs[i]) * (x[i] - as[i]);
          cost = g.sum(c) + gauss + limit_cost(x);
        }
        while (true) {
          // ---- gradient: Ma - qfrc_smooth - J^T f  (contacts: butterfly sum; limits: redundant)
          float grad[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 6; r++) s += J[r][i] * f[r];
            qfc[i] = g.sum(s);
          }
          float lim_f[2];
#pragma unroll
          for (int j = 0; j < 2; j++) {
            lim_f[j] = 0.f;
            if (lim_sg[j] != 0.f) {
              float jr = lim_sg[j] * x[j] - lim_aref[j];
              if (jr < 0) lim_f[j] = -lim_D[j] * jr;
              qfc[j] += lim_sg[j] * lim_f[j];
            }
          }
          float gn = 0.f;
#pragma unroll
          for (int i = 0; i < NV; i++) { grad[i] = Ma[i] - qs[i] - qfc[i]; gn += grad[i] * grad[i]; }
          gn = sqrtf(gn);
          if (it > 0 && scale * gn < m.tolerance) break;
          if (it >= m.iterations) break;
          // ---- Hessian: M + J^T W J, W = cone Hessian of this lane's contact
          float H[NT];
#pragma unroll
          for (int e = 0; e < NT; e++) H[e] = 0.f;
          if (zone != 0) {
            float U[6], scl[6];
            float Dm = 0.f, NTv = 0.f, invT = 0.f;
            if (zone == 2) {
              Dm = D[0] / (mu * mu * (1 + mu * mu)); NTv = cN - mu * cT; invT = 1.0f / cT;
              scl[0] = mu; U[0] = cN;
#pragma unroll
              for (int j = 1; j < 6; j++) { scl[j] = j < dim ? fri[j - 1] : 0.f; U[j] = j < dim ? jar[j] * fri[j - 1] : 0.f; }
            }
#pragma unroll
            for (int aa = 0; aa < 6; aa++) {
              if (aa < dim) {
                float Wa[NV];
                if (zone == 1) {
#pragma unroll
                  for (int d = 0; d < NV; d++) Wa[d] = D[aa] * J[aa][d];
                } else {
#pragma unroll
                  for (int d = 0; d < NV; d++) Wa[d] = 0.f;
#pragma unroll
                  for (int b = 0; b < 6; b++) {
                    if (b < dim) {
                      float h;
                      if (aa == 0 && b == 0) h = 1.f;
                      else if (aa == 0) h = -mu * U[b] * invT;
                      else if (b == 0) h = -mu * U[aa] * invT;
                      else {
                        float uu = U[aa] * U[b] * invT * invT;
                        h = mu * mu * uu - mu * NTv * ((aa == b ? invT : 0.f) - uu * invT);
                      }
                      float hc = Dm * scl[aa] * h * scl[b];
#pragma unroll
                      for (int d = 0; d < NV; d++) Wa[d] += hc * J[b][d];
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < NV; i++)
#pragma unroll
                  for (int j = 0; j <= i; j++) H[fast::tri(i, j)] += J[aa][i] * Wa[j];
              }
            }
          }
#pragma unroll
          for (int e = 0; e < NT; e++) H[e] = g.sum(H[e]);
#pragma unroll
          for (int i = 0; i < NV; i++) H[fast::tri(i, i)] += fi.Mdiag[i];
#pragma unroll
          for (int j = 0; j < 2; j++) if (lim_sg[j] != 0.f && lim_sg[j] * x[j] - lim_aref[j] < 0) H[fast::tri(j, j)] += lim_D[j];
          // ---- Cholesky (in registers, every lane redundantly) and search = -H^-1 grad
#pragma unroll
          for (int k = 0; k < NV; k++) {
            float dkk = H[fast::tri(k, k)];
#pragma unroll
            for (int j = 0; j < k; j++) dkk -= H[fast::tri(k, j)] * H[fast::tri(k, j)];
            if (!(dkk > 1e-15f)) { dkk = 1e-15f; flags |= FLAG_CHOL; }
            dkk = sqrtf(dkk);
            H[fast::tri(k, k)] = dkk;
            float inv = 1.0f / dkk;
#pragma unroll
            for (int i = k + 1; i < NV; i++) {
              float s = H[fast::tri(i, k)];
#pragma unroll
              for (int j = 0; j < k; j++) s -= H[fast::tri(i, j)] * H[fast::tri(k, j)];
              H[fast::tri(i, k)] = s * inv;
            }
          }
          float srch[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) {
            float s = -grad[i];
#pragma unroll
            for (int j = 0; j < i; j++) s -= H[fast::tri(i, j)] * srch[j];
            srch[i] = s / H[fast::tri(i, i)];
          }
#pragma unroll
          for (int i = NV - 1; i >= 0; i--) {
            float s = srch[i];
#pragma unroll
            for (int j = i + 1; j < NV; j++) s -= H[fast::tri(j, i)] * srch[j];
            srch[i] = s / H[fast::tri(i, i)];
          }
          float sn = 0.f, dec = 0.f;
#pragma unroll
          for (int i = 0; i < NV; i++) { sn += srch[i] * srch[i]; dec -= grad[i] * srch[i]; }
          sn = sqrtf(sn);
          if (sn < 1e-15f) break;
          float Mv[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) Mv[i] = fi.Mdiag[i] * srch[i];
#pragma unroll
          for (int r = 0; r < 6; r++) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < NV; d++) s += J[r][d] * srch[d];
            jv[r] = s;
          }
          const float gtol = m.tolerance * m.ls_tolerance * sn / scale;
          // ---- exact line search: root of the 1-D derivative (generic: linesearch / ls_eval)
          float q1 = 0.f, q2 = 0.f;
#pragma unroll
          for (int i = 0; i < NV; i++) { q1 += srch[i] * (Ma[i] - qs[i]); q2 += 0.5f * srch[i] * Mv[i]; }
          float lq[10];
          {
            float uu = 0.f, uv = 0.f, vv = 0.f, Q1 = 0.f, Q2 = 0.f;
#pragma unroll
            for (int r = 0; r < 6; r++) if (r < dim) {
              float xx = jar[r], v = jv[r], Dr = D[r];
              Q1 += Dr * xx * v; Q2 += 0.5f * Dr * v * v;
              if (r > 0) { float u = xx * fri[r - 1], s = v * fri[r - 1]; uu += u * u; uv += u * s; vv += s * s; }
            }
            if (dim == 1) { lq[0] = jar[0]; lq[1] = jv[0]; lq[8] = 1.f; lq[9] = -1.f; }
            else { lq[0] = jar[0] * mu; lq[1] = jv[0] * mu; lq[8] = mu; lq[9] = dim > 0 ? D[0] / (mu * mu * (1 + mu * mu)) : 0.f; }
            lq[2] = uu; lq[3] = uv; lq[4] = vv; lq[6] = Q1; lq[7] = Q2;
          }
          auto ls_eval = [&](float alpha, float& d1, float& d2) {
            float l1 = 0.f, l2 = 0.f;
            if (dim > 0) {
              float mu_ = lq[8];
              float N = lq[0] + alpha * lq[1];
              float tsq = lq[2] + alpha * (2 * lq[3] + alpha * lq[4]);
              float Tn = tsq > 0 ? sqrtf(tsq) : 0.f;
              bool top = (N >= mu_ * Tn) || (Tn <= 0 && N >= 0);
              bool bottom = (mu_ * N + Tn <= 0) || (Tn <= 0 && N < 0);
              if (lq[9] < 0) { top = !(N < 0); bottom = N < 0; }
              if (!top) {
                if (bottom) { l1 = lq[6] + 2 * alpha * lq[7]; l2 = 2 * lq[7]; }
                else {
                  float Dm = lq[9];
                  float NTv = N - mu_ * Tn;
                  float N1 = lq[1];
                  float T1 = (lq[3] + alpha * lq[4]) / Tn;
                  float T2 = lq[4] / Tn - T1 * T1 / Tn;
                  l1 = Dm * NTv * (N1 - mu_ * T1);
                  l2 = Dm * ((N1 - mu_ * T1) * (N1 - mu_ * T1) - NTv * mu_ * T2);
                }
              }
            }
            l1 = g.sum(l1); l2 = g.sum(l2);
#pragma unroll
            for (int j = 0; j < 2; j++) if (lim_sg[j] != 0.f) {
              float xv = lim_sg[j] * srch[j];
              float xx = lim_sg[j] * x[j] - lim_aref[j] + alpha * xv;
              if (xx < 0) { l1 += lim_D[j] * xx * xv; l2 += lim_D[j] * xv * xv; }
            }
            d1 = l1 + q1 + 2 * alpha * q2;
            d2 = l2 + 2 * q2;
          };
          float alpha = 0.f;
          {
            float d1, d2;
            ls_eval(0.f, d1, d2);
            int nev = 1;
            float lo = 0.f, hi = -1.f, dlo = d1, dhi = 0.f;
            const float rel = 3.4526698e-4f;  // sqrt(FLT_EPSILON)
            bool conv = fabsf(d1) < gtol;
            for (int li = 0; li < m.ls_iterations && !conv; li++) {
              float step = d2 > 1e-15f ? -d1 / d2 : (d1 < 0 ? 1.f : -1.f);
              float nxt = alpha + step;
              if (hi >= 0 && !(lo < nxt && nxt < hi)) nxt = 0.5f * (lo + hi);
              if (nxt <= 0 && hi < 0) nxt = alpha * 0.5f;
              if (nxt == alpha) break;
              bool tiny = fabsf(nxt - alpha) <= rel * fabsf(nxt);
              alpha = nxt;
              ls_eval(alpha, d1, d2);
              nev++;
              if (d1 < 0) { if (alpha > lo) { lo = alpha; dlo = d1; } }
              else if (hi < 0 || alpha < hi) { hi = alpha; dhi = d1; }
              conv = fabsf(d1) < gtol || tiny;
            }
            ls_used += nev;
            if (!conv) {
              if (hi >= 0 && (lo <= 0 || fabsf(dhi) < fabsf(dlo))) alpha = (lo > 0 || fabsf(dhi) < fabsf(dlo)) ? hi : 0.f;
              else alpha = lo;
            }
          }
          if (alpha == 0.f) break;
          float gsum = 0.f;
#pragma unroll
          for (int i = 0; i < NV; i++) {
            x[i] += alpha * srch[i]; Ma[i] += alpha * Mv[i];
            gsum += 0.5f * (Ma[i] - qs[i]) * (x[i] - as[i]);
          }
#pragma unroll
          for (int r = 0; r < 6; r++) jar[r] += alpha * jv[r];
          const float old = cost;
          cost = g.sum(update(true, zone, cN, cT)) + gsum + limit_cost(x);
          it++;
          const float improvement = alpha < 2.f ? alpha * (1.f - 0.5f * alpha) * dec : old - cost;
          if (scale * improvement < m.tolerance) {
            // forces of the final point for qfrc_constraint
#pragma unroll
            for (int i = 0; i < NV; i++) {
              float s = 0.f;
#pragma unroll
              for (int r = 0; r < 6; r++) s += J[r][i] * f[r];
              qfc[i] = g.sum(s);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) if (lim_sg[j] != 0.f) {
              float jr = lim_sg[j] * x[j] - lim_aref[j];
              if (jr < 0) qfc[j] += lim_sg[j] * (-lim_D[j] * jr);
            }
            break;
          }
        }
      }
      HSR_PHASE(w, g, PH_SOLVE);
      n_iter += it; n_ls += ls_used;
      kflop += algorithmic_flops(m, ncon, nefc, it, ls_used, w.wi[WI_NPFLOP]);

      // ---------------------------------------------------------------- goal test on the poses of this forward pass
      bool reached = false;
      if (HASB && a.cfg.has_goal) {
        const GT* p = w.xpos + 3 * fi.block_body;
        GT dx = p[0] - (GT)mocap[0], dy = p[1] - (GT)mocap[1], dz = p[2] - (GT)mocap[2];
        reached = sqrt(dx * dx + dy * dy + dz * dz) < (GT)a.cfg.geofence;
      }
      // ---------------------------------------------------------------- Euler with implicit joint damping (B.8)
#pragma unroll
      for (int i = 0; i < NV; i++) {
        float ai = m.any_damping ? (qs[i] + qfc[i]) / (fi.Mdiag[i] + dt * fi.damp[i]) : x[i];
        warm[i] = x[i];
        qvel[i] += dt * ai;
        if (!(fabsf(qvel[i]) < 1e6f)) flags |= FLAG_BAD_NUM;
      }
      qpos[0] += dt * qvel[0]; qpos[1] += dt * qvel[1];
      if (HASB) {
#pragma unroll
        for (int k = 0; k < 3; k++) qpos[2 + k] += dt * qvel[2 + k];
        float om0 = qvel[5], om1 = qvel[6], om2 = qvel[7];
        float ang = sqrtf(om0 * om0 + om1 * om1 + om2 * om2);
        quatnormalize(qpos + 5);
        if (ang * dt > 1e-15f) {
          float hh = 0.5f * ang * dt, sn_ = sinf(hh) / ang;
          float dq[4] = {cosf(hh), sn_ * om0, sn_ * om1, sn_ * om2};
          quatmul(qpos + 5, dq, qpos + 5);
        }
        quatnormalize(qpos + 5);
      }
      taken++;
      g.sync();
      HSR_PHASE(w, g, PH_EULER);
      if (reached) { success = true; break; }
    }

    // ------------------------------------------------------------------ results: HBM once per action
    {
      float* st = a.state + (size_t)env * a.S;
      const int nobs = nq + NV;
      for (int i = g.lane; i < nq + 2 * NV; i += G) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < (HASB ? 9 : 2); k++) if (i == k) v = qpos[k];
#pragma unroll
        for (int k = 0; k < NV; k++) { if (i == nq + k) v = qvel[k]; if (i == nq + NV + k) v = warm[k]; }
        st[i] = v;
        if (a.obs && i < nobs) a.obs[(size_t)env * nobs + i] = v;
      }
      if (g.lane == 0) {
        if (a.reward) a.reward[env] = success ? 1.0f : 0.0f;
        if (a.done) a.done[env] = success ? 1 : 0;
        if (a.success) a.success[env] = success ? 1 : 0;
        if (a.taken) a.taken[env] = taken;
        if (a.bad) a.bad[env] = (unsigned char)flags;
        atomicAdd(a.stats + ST_SUBSTEPS, (unsigned long long)taken);
        atomicAdd(a.stats + ST_ITERS, (unsigned long long)n_iter);
        atomicAdd(a.stats + ST_NARROW, (unsigned long long)w.wi[WI_NARROW]);
        atomicAdd(a.stats + ST_LSEVAL, (unsigned long long)n_ls);
        atomicAdd(a.stats + ST_CONTACTS, (unsigned long long)sumcon);
        atomicAdd(a.stats + ST_ROWS, (unsigned long long)sumefc);
        atomicAdd(a.stats + ST_FLOPS, (unsigned long long)kflop);
        if (flags) atomicAdd(a.stats + ST_BAD, 1ull);
#ifdef HSRB_PHASE_CLOCKS
        for (int k = 0; k < PH_COUNT; k++) atomicAdd(a.stats + ST_PHASE0 + k, (unsigned long long)w.wi[WI_PHASE0 + k]);
#endif
      }
    }
    g.sync();
  }
}
