// Flat model description shared by the CUDA kernels and the host (CPU port) build.
//
// Mirrors hsr_env_b200/model.py::_FIELDS one to one (same order, same shapes); the blob written by
// Model.to_blob() is walked with this table.  Replaces what the reference keeps inside mjModel after
// mujoco_py.load_model_from_path (/root/reference/hsr/mujoco_env.py:33).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#define HSRB_MAGIC 0x42525348
#define HSRB_VERSION 3
#define HSRB_MAXBODY 12

enum { JNT_FREE = 0, JNT_SLIDE = 1, JNT_HINGE = 2 };
enum { GEOM_PLANE = 0, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { NP_PLANE_BOX = 0, NP_PLANE_CONVEX = 1, NP_BOX_BOX = 2, NP_CONVEX_CONVEX = 3 };

// X(name, is_float, count expression over the dims)
#define HSRB_FIELDS(X)                    \
  X(opt, 1, 16)                           \
  X(qpos0, 1, nq)                         \
  X(body_parent, 0, nbody)                \
  X(body_pos, 1, nbody * 3)               \
  X(body_quat, 1, nbody * 4)              \
  X(body_mass, 1, nbody)                  \
  X(body_ipos, 1, nbody * 3)              \
  X(body_inertia, 1, nbody * 6)           \
  X(body_jntadr, 0, nbody)                \
  X(body_jntnum, 0, nbody)                \
  X(body_dofadr, 0, nbody)                \
  X(body_dofnum, 0, nbody)                \
  X(jnt_type, 0, njnt)                    \
  X(jnt_body, 0, njnt)                    \
  X(jnt_qposadr, 0, njnt)                 \
  X(jnt_dofadr, 0, njnt)                  \
  X(jnt_axis, 1, njnt * 3)                \
  X(jnt_pos, 1, njnt * 3)                 \
  X(jnt_limited, 0, njnt)                 \
  X(jnt_range, 1, njnt * 2)               \
  X(jnt_solref, 1, njnt * 2)              \
  X(jnt_solimp, 1, njnt * 5)              \
  X(dof_body, 0, nv)                      \
  X(dof_parent, 0, nv)                    \
  X(dof_jnt, 0, nv)                       \
  X(dof_damping, 1, nv)                   \
  X(dof_invweight0, 1, nv)                \
  X(act_dof, 0, nu)                       \
  X(act_qposadr, 0, nu)                   \
  X(act_gear, 1, nu)                      \
  X(act_kp, 1, nu)                        \
  X(act_ctrllimited, 0, nu)               \
  X(act_ctrlrange, 1, nu * 2)             \
  X(act_forcelimited, 0, nu)              \
  X(act_forcerange, 1, nu * 2)            \
  X(geom_type, 0, ngeom)                  \
  X(geom_body, 0, ngeom)                  \
  X(geom_pos, 1, ngeom * 3)               \
  X(geom_mat, 1, ngeom * 9)               \
  X(geom_size, 1, ngeom * 3)              \
  X(geom_rbound, 1, ngeom)                \
  X(geom_aabb, 1, ngeom * 3)              \
  X(geom_vertadr, 0, ngeom)               \
  X(geom_vertnum, 0, ngeom)               \
  X(geom_invweight, 1, ngeom)             \
  X(hull_vert, 1, nvert * 3)              \
  X(pair_geom1, 0, npair)                 \
  X(pair_geom2, 0, npair)                 \
  X(pair_func, 0, npair)                  \
  X(pair_condim, 0, npair)                \
  X(pair_friction, 1, npair * 5)          \
  X(pair_solref, 1, npair * 2)            \
  X(pair_solimp, 1, npair * 5)            \
  X(block_body, 0, nblock)                \
  X(finger_body, 0, 2)                    \
  X(finger_pos, 1, 6)                     \
  X(mocap_pos0, 1, 3)

// Environment-level configuration: the goal / block spaces and geofence of the reference's
// env_wrapper (/root/reference/hsr/util.py:53-74) in the form App. C #2 of SURVEY.md defines, the general
// list-of-GoalSpec form of HSREnv (/root/reference/hsr/env.py:126,137-147,161-172) and the per-joint start spaces
// of HSREnv.new_state (/root/reference/hsr/env.py:149-156).
#define HSRB_MAXGOAL 4     // GoalSpecs in the general form
#define HSRB_MAXFIXED 4    // fixed (ndarray) goal endpoints beside the sampled point
#define HSRB_MAXSTART 8    // joints with a start space
enum { GOAL_EP_POINT = -1 };   // endpoint codes: >= 0 body id, -1 the per-environment goal point (mocap_pos), -2-k fixed point k
template <typename T>
struct EnvCfg {
  int has_goal;        // 0: goals=None (README run before the first reset): never done
  int has_block;       // 0: no block-space: blocks reset to their qpos0 pose
  int qidx0, qidx1;    // which quaternion components block-space dims 2,3 drive
  T goal_lo[3], goal_hi[3];   // the sampled goal point (a 3-d Space endpoint; lo == hi for an ndarray) -> mocap_pos
  T block_lo[4], block_hi[4];
  T geofence;
  T min_sep;           // >0: rejection-sample block (x,y) so that blocks start at least this far apart
  // general form: success = all_k |pos(a_k) - pos(b_k)| < dist_k.  ngoal == 0: the env_wrapper form above
  // (every block within `geofence` of the goal point).
  int ngoal;
  int goal_a[HSRB_MAXGOAL], goal_b[HSRB_MAXGOAL];
  T goal_dist[HSRB_MAXGOAL];
  T fixed_pt[HSRB_MAXFIXED][3];
  // start spaces: qpos[adr .. adr+width) ~ U[lo, hi] at reset (width 1, or 7 for a free joint)
  int nstart;
  int start_adr[HSRB_MAXSTART], start_width[HSRB_MAXSTART];
  T start_lo[HSRB_MAXSTART][7], start_hi[HSRB_MAXSTART][7];
};

template <typename T>
struct ModelT {
  int nq, nv, nu, nbody, njnt, ngeom, nvert, npair, nblock;
  int ncon_max, nefc_max;
  // options
  T timestep, gravity[3], impratio, tolerance, ls_tolerance, mpr_tolerance, meaninertia;
  int iterations, ls_iterations, mpr_iterations;
  int any_damping;
  uint32_t body_dofmask[HSRB_MAXBODY];  // dofs that move each body (ancestors' dofs included)
  int body_root[HSRB_MAXBODY];          // the ancestor attached to the world (root of the body's kinematic tree)
#define X(name, isf, cnt) const typename std::conditional<isf, T, int>::type* name;
  HSRB_FIELDS(X)
#undef X
};

// Host-side owner of a parsed blob: one contiguous byte buffer + a ModelT whose pointers point into it.
// `base` may be re-targeted (device copy) with rebase().
template <typename T>
struct HostModel {
  ModelT<T> m;
  std::vector<unsigned char> buf;  // all arrays, converted to T / int32
  std::vector<size_t> offsets;     // byte offset of every field inside buf

  bool parse(const void* blob, size_t bytes, std::string& err) {
    const unsigned char* p = (const unsigned char*)blob;
    if (bytes < 64) { err = "blob too small"; return false; }
    int32_t head[11];
    memcpy(head, p, sizeof(head));
    if (head[0] != HSRB_MAGIC || head[1] != HSRB_VERSION) { err = "bad magic/version"; return false; }
    memset(&m, 0, sizeof(m));
    m.nq = head[2]; m.nv = head[3]; m.nu = head[4]; m.nbody = head[5]; m.njnt = head[6];
    m.ngeom = head[7]; m.nvert = head[8]; m.npair = head[9]; m.nblock = head[10];
    if (m.nbody > HSRB_MAXBODY || m.nv > 32) { err = "model too large (nbody<=12, nv<=32)"; return false; }
    int nq = m.nq, nv = m.nv, nu = m.nu, nbody = m.nbody, njnt = m.njnt, ngeom = m.ngeom, nvert = m.nvert,
        npair = m.npair, nblock = m.nblock;
    (void)nq; (void)nv; (void)nu; (void)nbody; (void)njnt; (void)ngeom; (void)nvert; (void)npair; (void)nblock;
    size_t off = 64, out = 0;
    offsets.clear();
    // pass 1: sizes
#define X(name, isf, cnt) { size_t n = (size_t)(cnt); size_t b = n * (isf ? 8 : 4); b += (8 - b % 8) % 8; off += b; \
      offsets.push_back(out); size_t ob = n * (isf ? sizeof(T) : 4); ob += (16 - ob % 16) % 16; out += ob; }
    HSRB_FIELDS(X)
#undef X
    if (off != bytes) { err = "blob size mismatch"; return false; }
    buf.assign(out + 16, 0);
    off = 64;
    int fi = 0;
#define X(name, isf, cnt) { size_t n = (size_t)(cnt); unsigned char* dst = buf.data() + offsets[fi++]; \
      if (isf) { for (size_t i = 0; i < n; i++) { double v; memcpy(&v, p + off + 8 * i, 8); ((T*)dst)[i] = (T)v; } off += n * 8; } \
      else { memcpy(dst, p + off, n * 4); size_t b = n * 4; off += b + (8 - b % 8) % 8; } }
    HSRB_FIELDS(X)
#undef X
    rebase(buf.data());
    const T* o = m.opt;
    m.timestep = o[0]; m.gravity[0] = o[1]; m.gravity[1] = o[2]; m.gravity[2] = o[3]; m.impratio = o[4];
    m.tolerance = o[5]; m.ls_tolerance = o[6]; m.iterations = (int)o[7]; m.ls_iterations = (int)o[8];
    m.mpr_tolerance = o[9]; m.mpr_iterations = (int)o[10]; m.meaninertia = o[11];
    m.any_damping = 0;
    for (int i = 0; i < m.nv; i++) if (m.dof_damping[i] > 0) m.any_damping = 1;
    for (int b = 0; b < m.nbody; b++) {
      uint32_t mask = 0;
      int c = b;
      while (c > 0) {
        for (int k = 0; k < m.body_dofnum[c]; k++) mask |= 1u << (m.body_dofadr[c] + k);
        c = m.body_parent[c];
      }
      m.body_dofmask[b] = mask;
      int r = b;
      while (r > 0 && m.body_parent[r] > 0) r = m.body_parent[r];
      m.body_root[b] = r;
    }
    // default capacities: 4 plane contacts per block, up to 8 box-box per block pair / block-pan, a few hull contacts
    m.ncon_max = 8 + 4 * m.nblock + (m.nblock > 1 ? 4 * m.nblock : 0);
    // an articulated arm adds finger / palm contacts with the block and the pan (configs[2]: 4 pan-block + up to 4 hand-pan +
    // up to 4 hand-block were seen along gripper roll-outs)
    for (int b = 1; b < m.nbody; b++) if (m.body_parent[b] > 0) { m.ncon_max += 4; break; }
    if (m.ncon_max > 32) m.ncon_max = 32;
    int nlim = 0;
    for (int j = 0; j < m.njnt; j++) nlim += m.jnt_limited[j] ? 1 : 0;
    m.nefc_max = nlim + 6 * m.ncon_max;
    return true;
  }

  void rebase(const unsigned char* base) {
    int fi = 0;
#define X(name, isf, cnt) m.name = (decltype(m.name))(base + offsets[fi++]);
    HSRB_FIELDS(X)
#undef X
  }
};
