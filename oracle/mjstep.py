"""ORACLE (test infrastructure, not product code) - fp64 numpy restatement of ``mj_step`` for the HSR models.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in MuJoCo (via mujoco-py), which is neither
vendored in /root/reference, nor pinned (setup.py:26), nor installable here, and the reference has no tests or
golden vectors.  This file therefore restates MuJoCo's *published* algorithm
(MuJoCo "Computation" chapter) for the call sites

    /root/reference/hsr/env.py:123          self.sim.step()              -> step()
    /root/reference/hsr/env.py:176          self.sim.forward()           -> forward()
    /root/reference/hsr/mujoco_env.py:84    self.sim.reset()             -> Data.reset()

and is pinned only by the analytic known-answer tests of SURVEY.md B.10 (tests/test_oracle_kat.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Stage map:  B.1 kinematics() | B.2 com_crb() | B.3 collision() | B.4/B.5 make_constraint()
| B.6 smooth_dynamics() | B.7 solve_newton() | B.8 euler().
"""
from __future__ import annotations

import numpy as np

JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 1, 2
GEOM_PLANE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 0, 5, 6, 7
NP_PLANE_BOX, NP_PLANE_CONVEX, NP_BOX_BOX, NP_CONVEX_CONVEX = 0, 1, 2, 3
MINVAL = 1e-15
MINIMP, MAXIMP = 1e-4, 0.9999
EPS = np.finfo(np.float64).eps
MAXCON_PAIR_BOX = 8


# ----------------------------------------------------------------------------------------------- math
def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def q2mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def qnormalize(q):
    n = np.linalg.norm(q)
    if n < MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])  # zero-norm quaternion -> identity (SURVEY B.10 case 6)
    return q / n


def sym6(v):
    return np.array([[v[0], v[3], v[4]], [v[3], v[1], v[5]], [v[4], v[5], v[2]]])


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def cross_motion(v, s):
    return np.concatenate([np.cross(v[:3], s[:3]), np.cross(v[:3], s[3:]) + np.cross(v[3:], s[:3])])


def cross_force(v, f):
    return np.concatenate([np.cross(v[:3], f[:3]) + np.cross(v[3:], f[3:]), np.cross(v[:3], f[3:])])


# ----------------------------------------------------------------------------------------------- data
class Data:
    """Per-environment state (the part of mjData the path touches)."""

    def __init__(self, m):
        self.m = m
        self.reset()

    def reset(self):
        """mj_resetData: qpos=qpos0, qvel=0, ctrl=0, time=0, mocap_pos=body_pos, warmstart=0."""
        m = self.m
        self.qpos = m.qpos0.copy()
        self.qvel = np.zeros(m.nv)
        self.ctrl = np.zeros(m.nu)
        self.qacc_warmstart = np.zeros(m.nv)
        self.mocap_pos = m.mocap_pos0.copy()
        self.time = 0.0
        self.stats = {}

    def copy_state(self):
        return dict(qpos=self.qpos.copy(), qvel=self.qvel.copy(), qacc_warmstart=self.qacc_warmstart.copy(),
                    ctrl=self.ctrl.copy(), time=self.time)


# ----------------------------------------------------------------------------------------------- B.1
def kinematics(m, d):
    nb = m.nbody
    d.xpos = np.zeros((nb, 3)); d.xquat = np.tile([1.0, 0, 0, 0], (nb, 1)); d.xmat = np.tile(np.eye(3), (nb, 1, 1))
    d.xipos = np.zeros((nb, 3)); d.anchor = np.zeros((m.njnt, 3)); d.axis = np.zeros((m.njnt, 3))
    for b in range(1, nb):
        p = m.body_parent[b]
        j0, nj = m.body_jntadr[b], m.body_jntnum[b]
        if nj == 1 and m.jnt_type[j0] == JNT_FREE:
            a = m.jnt_qposadr[j0]
            d.xpos[b] = d.qpos[a:a + 3]
            d.qpos[a + 3:a + 7] = qnormalize(d.qpos[a + 3:a + 7])  # MuJoCo normalises in place
            d.xquat[b] = d.qpos[a + 3:a + 7]
            d.anchor[j0] = d.xpos[b]
        else:
            d.xpos[b] = d.xpos[p] + d.xmat[p] @ m.body_pos[b]
            d.xquat[b] = qmul(d.xquat[p], m.body_quat[b])
            for j in range(j0, j0 + nj):
                R = q2mat(d.xquat[b])
                d.anchor[j] = d.xpos[b] + R @ m.jnt_pos[j]
                d.axis[j] = R @ m.jnt_axis[j]
                q = d.qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
                if m.jnt_type[j] == JNT_SLIDE:
                    d.xpos[b] = d.xpos[b] + d.axis[j] * q
                else:
                    d.xquat[b] = qmul(d.xquat[b], np.concatenate([[np.cos(q / 2)], np.sin(q / 2) * m.jnt_axis[j]]))
                    d.xpos[b] = d.anchor[j] - q2mat(d.xquat[b]) @ m.jnt_pos[j]
        d.xquat[b] = qnormalize(d.xquat[b])
        d.xmat[b] = q2mat(d.xquat[b])
        d.xipos[b] = d.xpos[b] + d.xmat[b] @ m.body_ipos[b]
    # geom frames
    d.gpos = np.zeros((m.ngeom, 3)); d.gmat = np.zeros((m.ngeom, 3, 3))
    for g in range(m.ngeom):
        b = m.geom_body[g]
        d.gpos[g] = d.xpos[b] + d.xmat[b] @ m.geom_pos[g]
        d.gmat[g] = d.xmat[b] @ m.geom_mat[g].reshape(3, 3)


# ----------------------------------------------------------------------------------------------- B.2
def com_crb(m, d):
    """cdof (spatial motion axes about the world origin, [ang; lin]), composite inertias, dense M."""
    nv = m.nv
    d.cdof = np.zeros((nv, 6))
    for j in range(m.njnt):
        a = m.jnt_dofadr[j]; b = m.jnt_body[j]; t = m.jnt_type[j]
        if t == JNT_FREE:
            for k in range(3):
                d.cdof[a + k, 3 + k] = 1.0
                ax = d.xmat[b][:, k]
                d.cdof[a + 3 + k, :3] = ax
                d.cdof[a + 3 + k, 3:] = -np.cross(ax, d.xpos[b])
        elif t == JNT_SLIDE:
            d.cdof[a, 3:] = d.axis[j]
        else:
            d.cdof[a, :3] = d.axis[j]
            d.cdof[a, 3:] = -np.cross(d.axis[j], d.anchor[j])
    # spatial inertia of each body about the world origin
    d.cinert = np.zeros((m.nbody, 6, 6))
    for b in range(1, m.nbody):
        mass = m.body_mass[b]; c = d.xipos[b]
        Iw = d.xmat[b] @ sym6(m.body_inertia[b]) @ d.xmat[b].T
        C = skew(c)
        d.cinert[b, :3, :3] = Iw - mass * C @ C
        d.cinert[b, :3, 3:] = mass * C
        d.cinert[b, 3:, :3] = -mass * C
        d.cinert[b, 3:, 3:] = mass * np.eye(3)
    crb = d.cinert.copy()
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parent[b]
        if p > 0:
            crb[p] += crb[b]
    d.crb = crb
    M = np.zeros((nv, nv))
    for i in range(nv):
        f = crb[m.dof_body[i]] @ d.cdof[i]
        j = i
        while j >= 0:
            M[i, j] = M[j, i] = d.cdof[j] @ f
            j = m.dof_parent[j]
    d.M = M


def jac(m, d, body, point):
    """translational / rotational Jacobians (3 x nv) of a world point fixed to `body` (mj_jac)."""
    Jp = np.zeros((3, m.nv)); Jr = np.zeros((3, m.nv))
    if body <= 0:
        return Jp, Jr
    i = m.body_dofadr[body] + m.body_dofnum[body] - 1
    while i >= 0:
        Jr[:, i] = d.cdof[i, :3]
        Jp[:, i] = d.cdof[i, 3:] + np.cross(d.cdof[i, :3], point)
        i = m.dof_parent[i]
    return Jp, Jr


# ----------------------------------------------------------------------------------------------- B.3
def make_frame(n):
    """contact frame rows: normal, then two tangents (mju_makeFrame)."""
    n = n / np.linalg.norm(n)
    t = np.array([0.0, 1.0, 0.0]) if -0.5 < n[1] < 0.5 else np.array([0.0, 0.0, 1.0])
    t = t - n * np.dot(n, t)
    t = t / np.linalg.norm(t)
    return np.stack([n, t, np.cross(n, t)])


def plane_box(ppos, pmat, bpos, bmat, size):
    """mjc_PlaneBox: corners at or below the plane, at most 4 contacts. normal = plane normal."""
    n = pmat[:, 2]
    dist0 = np.dot(bpos - ppos, n)
    out = []
    for i in range(8):
        s = np.array([size[0] if i & 1 else -size[0], size[1] if i & 2 else -size[1], size[2] if i & 4 else -size[2]])
        vec = bmat @ s
        ldist = np.dot(n, vec)
        if dist0 + ldist > 0 or ldist > 0:
            continue
        dist = dist0 + ldist
        out.append((dist, bpos + vec - n * dist * 0.5, n))
        if len(out) == 4:
            break
    return out


SUPPORT_TIE = 1e-12   # support ties (see support())


def support(gtype, size, verts, pos, mat, dirw):
    """Support point of a convex geom in world direction dirw (mjccd_support, margin 0).

    Ties.  A box face / edge perpendicular to the direction, or a hull face whose vertices share the maximal support,
    is a TIE in exact arithmetic, and such directions are not accidents: the portal refinement converges to face
    normals of the Minkowski difference, and resting contacts line directions up with box axes.  With plain comparisons
    (libccd / MuJoCo: `dir[i] > 0`, first maximum of the vertex scan) the winner is decided by rounding noise of order
    1e-17, differently in every implementation, and the contact point jumps across the face.  This restatement (and the
    kernels, hsr_core.h support_d) resolves ties DETERMINISTICALLY: a local direction component within SUPPORT_TIE of
    zero counts as positive, and among the hull vertices whose support is within SUPPORT_TIE of the maximum the lowest
    index wins.  It only changes results where MuJoCo's own answer is rounding noise."""
    dl = mat.T @ dirw
    if gtype == GEOM_BOX:
        res = np.where(dl >= -SUPPORT_TIE, 1.0, -1.0) * size
    elif gtype == GEOM_CYLINDER:
        n = np.hypot(dl[0], dl[1])
        res = np.array([dl[0] / n * size[0], dl[1] / n * size[0], 0.0]) if n > MINVAL else np.zeros(3)
        res[2] = size[1] if dl[2] >= -SUPPORT_TIE else -size[1]
    else:
        vals = verts[:, 0] * dl[0] + verts[:, 1] * dl[1] + verts[:, 2] * dl[2]
        res = verts[int(np.argmax(vals >= vals.max() - SUPPORT_TIE))]
    return pos + mat @ res


def plane_convex(ppos, pmat, g2):
    """Deepest point of a convex geom against a plane (one contact; see DESIGN.md 'plane-convex')."""
    n = pmat[:, 2]
    p = support(*g2, -n)
    dist = np.dot(p - ppos, n)
    if dist > 0:
        return []
    return [(dist, p - n * dist * 0.5, n)]


def _iszero(x):
    return abs(x) < EPS


def _eq(a, b):
    ab = abs(a - b)
    if ab < EPS:
        return True
    a, b = abs(a), abs(b)
    return ab < EPS * (b if b > a else a)


def _tri_dist2(P, a, b, c):
    """squared distance from P to triangle (a,b,c) and the closest point (ccdVec3PointTriDist2)."""
    ab, ac, ap = b - a, c - a, P - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        q = a
    else:
        bp = P - b
        d3, d4 = ab @ bp, ac @ bp
        if d3 >= 0 and d4 <= d3:
            q = b
        else:
            vc = d1 * d4 - d3 * d2
            cp = P - c
            d5, d6 = ab @ cp, ac @ cp
            if vc <= 0 and d1 >= 0 and d3 <= 0:
                q = a + ab * (d1 / (d1 - d3))
            elif d6 >= 0 and d5 <= d6:
                q = c
            else:
                vb = d5 * d2 - d1 * d6
                va = d3 * d6 - d5 * d4
                if vb <= 0 and d2 >= 0 and d6 <= 0:
                    q = a + ac * (d2 / (d2 - d6))
                elif va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
                    q = b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))
                else:
                    den = 1.0 / (va + vb + vc)
                    q = a + ab * (vb * den) + ac * (vc * den)
    return (q - P) @ (q - P), q


def mpr_penetration(g1, g2, tol, max_iter, stats=None):
    """Minkowski Portal Refinement penetration query (libccd ccdMPRPenetration, as called by mjc_Convex).

    g = (type, size, verts, pos, mat).  Returns None or (depth, dir (geom1->geom2), pos).
    """
    nsup = [0]

    def sup(dirv):
        nsup[0] += 1
        a = support(*g1, dirv)
        b = support(*g2, -dirv)
        return (a - b, a, b)

    def portal_dir(v1, v2, v3):
        n = np.cross(v2[0] - v1[0], v3[0] - v1[0])
        return n / np.linalg.norm(n)

    def reach_tol(v1, v2, v3, v4, dirv):
        dv4 = v4[0] @ dirv
        dot = min(dv4 - v1[0] @ dirv, dv4 - v2[0] @ dirv, dv4 - v3[0] @ dirv)
        return _eq(dot, tol) or dot < tol

    def expand(v0, v1, v2, v3, v4):
        v4v0 = np.cross(v4[0], v0[0])
        if v1[0] @ v4v0 > 0:
            if v2[0] @ v4v0 > 0:
                v1 = v4
            else:
                v3 = v4
        else:
            if v3[0] @ v4v0 > 0:
                v2 = v4
            else:
                v1 = v4
        return v1, v2, v3

    c1, c2 = g1[3], g2[3]
    v0 = (c1 - c2, c1, c2)
    if np.all(np.abs(v0[0]) < EPS):
        v0 = (v0[0] + np.array([EPS * 10, 0, 0]), c1, c2)
    # ---- discover portal
    dirv = -v0[0] / np.linalg.norm(v0[0])
    v1 = sup(dirv)
    dot = v1[0] @ dirv
    if _iszero(dot) or dot < 0:
        return None
    dirv = np.cross(v0[0], v1[0])
    if _iszero(dirv @ dirv):
        if np.all(np.abs(v1[0]) < EPS):
            return None  # touching contact: depth 0, undefined normal -> MuJoCo rejects it
        depth = np.linalg.norm(v1[0])
        return depth, v1[0] / depth, 0.5 * (v1[1] + v1[2])
    dirv = dirv / np.linalg.norm(dirv)
    v2 = sup(dirv)
    dot = v2[0] @ dirv
    if _iszero(dot) or dot < 0:
        return None
    dirv = np.cross(v1[0] - v0[0], v2[0] - v0[0])
    dirv = dirv / np.linalg.norm(dirv)
    if dirv @ v0[0] > 0:
        v1, v2 = v2, v1
        dirv = -dirv
    while True:
        v3 = sup(dirv)
        dot = v3[0] @ dirv
        if _iszero(dot) or dot < 0:
            return None
        cont = False
        dot = np.cross(v1[0], v3[0]) @ v0[0]
        if dot < 0 and not _iszero(dot):
            v2 = v3; cont = True
        if not cont:
            dot = np.cross(v3[0], v2[0]) @ v0[0]
            if dot < 0 and not _iszero(dot):
                v1 = v3; cont = True
        if cont:
            dirv = np.cross(v1[0] - v0[0], v2[0] - v0[0])
            dirv = dirv / np.linalg.norm(dirv)
        else:
            break
    # ---- refine portal until it encapsulates the origin
    while True:
        dirv = portal_dir(v1, v2, v3)
        dot = dirv @ v1[0]
        if _iszero(dot) or dot > 0:
            break
        v4 = sup(dirv)
        dot = v4[0] @ dirv
        if not (_iszero(dot) or dot > 0) or reach_tol(v1, v2, v3, v4, dirv):
            return None
        v1, v2, v3 = expand(v0, v1, v2, v3, v4)
    # ---- find penetration
    it = 0
    while True:
        dirv = portal_dir(v1, v2, v3)
        v4 = sup(dirv)
        if reach_tol(v1, v2, v3, v4, dirv) or it > max_iter:
            d2, q = _tri_dist2(np.zeros(3), v1[0], v2[0], v3[0])
            depth = np.sqrt(d2)
            if _iszero(depth):
                return None
            pdir = q / np.linalg.norm(q)
            # position: barycentric coordinates of the origin in the portal tetrahedron
            b = np.array([np.cross(v1[0], v2[0]) @ v3[0], np.cross(v3[0], v2[0]) @ v0[0],
                          np.cross(v0[0], v1[0]) @ v3[0], np.cross(v2[0], v1[0]) @ v0[0]])
            s = b.sum()
            if _iszero(s) or s < 0:
                b = np.array([0.0, np.cross(v2[0], v3[0]) @ dirv, np.cross(v3[0], v1[0]) @ dirv,
                              np.cross(v1[0], v2[0]) @ dirv])
                s = b.sum()
            vs = (v0, v1, v2, v3)
            p1 = sum(b[k] * vs[k][1] for k in range(4)) / s
            p2 = sum(b[k] * vs[k][2] for k in range(4)) / s
            if stats is not None:
                stats["mpr_support"] = stats.get("mpr_support", 0) + nsup[0]
            return depth, pdir, 0.5 * (p1 + p2)
        v1, v2, v3 = expand(v0, v1, v2, v3, v4)
        it += 1


def box_box(p1, R1, s1, p2, R2, s2):
    """Box-box manifold: SAT over 15 axes, face clipping (<= 8 points) or one edge-edge contact.

    MuJoCo's mjc_BoxBox is implementation-defined and cannot be restated from documentation; this routine is
    the definition both the oracle and the CUDA kernel follow (DESIGN.md 'box-box').  normal: box1 -> box2.
    """
    d = p2 - p1
    C = R1.T @ R2
    Q = np.abs(C) + 1e-10
    dl1 = R1.T @ d
    dl2 = R2.T @ d
    best = (-np.inf, -1, None)  # (separation (<0 = penetration), code, axis)
    for i in range(3):
        sep = abs(dl1[i]) - (s1[i] + Q[i] @ s2)
        if sep > 0:
            return []
        if sep > best[0]:
            best = (sep, i, R1[:, i] * (1.0 if dl1[i] >= 0 else -1.0))
    for i in range(3):
        sep = abs(dl2[i]) - (s2[i] + Q[:, i] @ s1)
        if sep > 0:
            return []
        if sep > best[0]:
            best = (sep, 3 + i, R2[:, i] * (1.0 if dl2[i] >= 0 else -1.0))
    ebest = (-np.inf, -1, None)
    for i in range(3):
        for j in range(3):
            ax = np.cross(R1[:, i], R2[:, j])
            ln = np.linalg.norm(ax)
            if ln < 1e-4:
                continue
            ax = ax / ln
            i1, i2 = (i + 1) % 3, (i + 2) % 3
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            ra = s1[i1] * abs(R1[:, i1] @ ax) + s1[i2] * abs(R1[:, i2] @ ax)
            rb = s2[j1] * abs(R2[:, j1] @ ax) + s2[j2] * abs(R2[:, j2] @ ax)
            dd = d @ ax
            sep = abs(dd) - (ra + rb)
            if sep > 0:
                return []
            if sep > ebest[0]:
                ebest = (sep, 6 + 3 * i + j, ax * (1.0 if dd >= 0 else -1.0))
    # prefer face axes unless an edge axis is clearly less penetrating
    if ebest[1] >= 0 and 1.05 * ebest[0] > best[0]:
        best = ebest
    sep, code, n = best
    if code >= 6:
        i, j = divmod(code - 6, 3)
        # edge of box 1: the one furthest along +n ; edge of box 2: furthest along -n
        pa = p1.copy()
        for k in range(3):
            if k != i:
                pa = pa + R1[:, k] * s1[k] * (1.0 if R1[:, k] @ n > 0 else -1.0)
        pb = p2.copy()
        for k in range(3):
            if k != j:
                pb = pb - R2[:, k] * s2[k] * (1.0 if R2[:, k] @ n > 0 else -1.0)
        ua, ub = R1[:, i], R2[:, j]
        w = pa - pb
        a_, b_, c_ = ua @ ua, ua @ ub, ub @ ub
        d_, e_ = ua @ w, ub @ w
        den = a_ * c_ - b_ * b_
        ta = (b_ * e_ - c_ * d_) / den
        tb = (a_ * e_ - b_ * d_) / den
        ta = np.clip(ta, -s1[i], s1[i]); tb = np.clip(tb, -s2[j], s2[j])
        ca, cb = pa + ua * ta, pb + ub * tb
        return [(sep, 0.5 * (ca + cb), n)]
    # face contact: reference box owns the axis
    if code < 3:
        pr, Rr, sr, pi_, Ri, si, nr, ax = p1, R1, s1, p2, R2, s2, n, code
    else:
        pr, Rr, sr, pi_, Ri, si, nr, ax = p2, R2, s2, p1, R1, s1, -n, code - 3
    # nr: outward normal of the reference face (pointing to the incident box)
    # incident face: the face of the incident box most anti-parallel to nr
    dots = Ri.T @ nr
    k = int(np.argmax(np.abs(dots)))
    sgn = -1.0 if dots[k] > 0 else 1.0
    fc = pi_ + Ri[:, k] * si[k] * sgn
    k1, k2 = (k + 1) % 3, (k + 2) % 3
    u, v = Ri[:, k1] * si[k1], Ri[:, k2] * si[k2]
    poly = [fc + u + v, fc - u + v, fc - u - v, fc + u - v]
    # clip against the four side planes of the reference face
    a1, a2 = (ax + 1) % 3, (ax + 2) % 3
    for axis_id, half in ((a1, sr[a1]), (a2, sr[a2])):
        for sg in (1.0, -1.0):
            pn = Rr[:, axis_id] * sg
            off = pn @ pr + half
            new = []
            for q in range(len(poly)):
                A, B = poly[q], poly[(q + 1) % len(poly)]
                da, db = pn @ A - off, pn @ B - off
                if da <= 0:
                    new.append(A)
                if (da < 0 < db) or (db < 0 < da):
                    new.append(A + (B - A) * (da / (da - db)))
            poly = new
            if not poly:
                return []
    face_off = nr @ pr + sr[ax]
    out = []
    for P in poly:
        depth = face_off - nr @ P
        if depth < 0:
            continue
        out.append((-depth, P + nr * depth * 0.5, n))
        if len(out) == MAXCON_PAIR_BOX:
            break
    return out


def collision(m, d, midphase=True):
    """Static candidate pairs -> bounding-sphere cull (+ optional conservative AABB cull) -> narrowphase."""
    d.contacts = []
    tol, mit = m.opt[9], int(m.opt[10])
    aabb = np.zeros((m.ngeom, 3))
    for g in range(m.ngeom):
        aabb[g] = np.abs(d.gmat[g]) @ m.geom_aabb[g]
    for k in range(m.npair):
        a, b = m.pair_geom1[k], m.pair_geom2[k]
        func = m.pair_func[k]
        if m.geom_type[a] == GEOM_PLANE:
            n = d.gmat[a][:, 2]
            if np.dot(d.gpos[b] - d.gpos[a], n) > m.geom_rbound[b]:
                continue
        else:
            if np.linalg.norm(d.gpos[b] - d.gpos[a]) > m.geom_rbound[a] + m.geom_rbound[b]:
                continue
            if midphase and np.any(np.abs(d.gpos[b] - d.gpos[a]) > aabb[a] + aabb[b]):
                continue
        ga = (m.geom_type[a], m.geom_size[a], m.hull_vert[m.geom_vertadr[a]:m.geom_vertadr[a] + m.geom_vertnum[a]],
              d.gpos[a], d.gmat[a])
        gb = (m.geom_type[b], m.geom_size[b], m.hull_vert[m.geom_vertadr[b]:m.geom_vertadr[b] + m.geom_vertnum[b]],
              d.gpos[b], d.gmat[b])
        d.stats["narrowphase"] = d.stats.get("narrowphase", 0) + 1
        if func == NP_PLANE_BOX:
            cons = plane_box(d.gpos[a], d.gmat[a], d.gpos[b], d.gmat[b], m.geom_size[b])
        elif func == NP_PLANE_CONVEX:
            cons = plane_convex(d.gpos[a], d.gmat[a], gb)
        elif func == NP_BOX_BOX:
            cons = box_box(d.gpos[a], d.gmat[a], m.geom_size[a], d.gpos[b], d.gmat[b], m.geom_size[b])
        else:
            r = mpr_penetration(ga, gb, tol, mit, d.stats)
            cons = [(-r[0], r[2], r[1])] if r is not None else []
        for dist, pos, n in cons:
            d.contacts.append(dict(pair=k, dist=dist, pos=pos, frame=make_frame(n), dim=int(m.pair_condim[k]),
                                   friction=m.pair_friction[k], solref=m.pair_solref[k], solimp=m.pair_solimp[k],
                                   body1=int(m.geom_body[a]), body2=int(m.geom_body[b]),
                                   invweight=m.geom_invweight[a] + m.geom_invweight[b]))


# ----------------------------------------------------------------------------------------------- B.4 / B.5
def impedance(solimp, pos):
    d0, dmax, width, mid, power = solimp
    d0 = min(max(d0, MINIMP), MAXIMP); dmax = min(max(dmax, MINIMP), MAXIMP)
    width = max(MINVAL, width); mid = min(max(mid, MINIMP), MAXIMP); power = max(1.0, power)
    if d0 == dmax or width <= MINVAL:
        return 0.5 * (d0 + dmax)
    x = abs(pos) / width
    if x >= 1:
        return dmax
    if x == 0:
        return d0
    if power == 1:
        y = x
    elif x <= mid:
        y = (1.0 / mid ** (power - 1)) * x ** power
    else:
        y = 1.0 - (1.0 / (1 - mid) ** (power - 1)) * (1 - x) ** power
    return d0 + y * (dmax - d0)


def make_constraint(m, d):
    """Rows: active joint limits (joint order), then contacts (elliptic cones, dim rows each)."""
    dt = m.opt[0]
    rows_J, pos, diag, sref, simp = [], [], [], [], []
    d.efc_contact = []  # per contact: (first row, dim, mu, friction)
    d.efc_type = []
    for j in range(m.njnt):
        if not m.jnt_limited[j] or m.jnt_type[j] == JNT_FREE:
            continue
        q = d.qpos[m.jnt_qposadr[j]]
        for side, dist in ((1.0, q - m.jnt_range[j, 0]), (-1.0, m.jnt_range[j, 1] - q)):
            if dist < 0:
                r = np.zeros(m.nv); r[m.jnt_dofadr[j]] = side
                rows_J.append(r); pos.append(dist); diag.append(m.dof_invweight0[m.jnt_dofadr[j]])
                sref.append(m.jnt_solref[j]); simp.append(m.jnt_solimp[j]); d.efc_type.append(0)
    d.nlimit = len(rows_J)
    for c in d.contacts:
        Jp1, Jr1 = jac(m, d, c["body1"], c["pos"])
        Jp2, Jr2 = jac(m, d, c["body2"], c["pos"])
        dJp, dJr = Jp2 - Jp1, Jr2 - Jr1
        first = len(rows_J)
        for r in range(c["dim"]):
            rows_J.append(c["frame"][r] @ dJp if r < 3 else c["frame"][r - 3] @ dJr)
            pos.append(c["dist"] if r == 0 else 0.0)
            diag.append(c["invweight"])
            sref.append(c["solref"]); simp.append(c["solimp"]); d.efc_type.append(1 if r == 0 else 2)
        d.efc_contact.append([first, c["dim"], 0.0, c["friction"]])
    ne = len(rows_J)
    d.nefc = ne
    d.efc_J = np.array(rows_J).reshape(ne, m.nv)
    d.efc_pos = np.array(pos)
    d.efc_R = np.zeros(ne); d.efc_aref = np.zeros(ne); d.efc_imp = np.zeros(ne)
    vel = d.efc_J @ d.qvel
    d.efc_vel = vel
    for i in range(ne):
        tc = max(sref[i][0], 2 * dt); dr = sref[i][1]
        dmax = min(max(simp[i][1], MINIMP), MAXIMP)
        imp = impedance(simp[i], d.efc_pos[i])
        d.efc_imp[i] = imp
        d.efc_R[i] = max(MINVAL, (1 - imp) / imp * diag[i])
        k = 1.0 / (dmax * dmax * tc * tc * dr * dr)
        bb = 2.0 / (dmax * tc)
        if d.efc_type[i] == 2:
            k = 0.0  # friction rows: no position term (their pos is 0 anyway)
        d.efc_aref[i] = -bb * vel[i] - k * imp * d.efc_pos[i]
    # elliptic cones: friction-row regularisation tied to the normal row (mj_makeImpedance)
    for c in d.efc_contact:
        i, dim, _, fr = c
        if dim > 1:
            d.efc_R[i + 1] = d.efc_R[i] / m.opt[4]
            c[2] = fr[0] * np.sqrt(d.efc_R[i + 1] / d.efc_R[i])
            for j in range(2, dim):
                d.efc_R[i + j] = d.efc_R[i + 1] * fr[0] * fr[0] / (fr[j - 1] * fr[j - 1])
        else:
            c[2] = fr[0]
    d.efc_D = 1.0 / d.efc_R


# ----------------------------------------------------------------------------------------------- B.6
def smooth_dynamics(m, d):
    nv = m.nv
    d.qfrc_passive = -m.dof_damping * d.qvel
    # RNE with qacc = 0: bias = Coriolis/centrifugal + gravity
    cvel = np.zeros((m.nbody, 6)); cacc = np.zeros((m.nbody, 6)); cfrc = np.zeros((m.nbody, 6))
    cacc[0, 3:] = -m.opt[1:4]
    for b in range(1, m.nbody):
        p = m.body_parent[b]
        v = cvel[p].copy(); a = cacc[p].copy()
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            d0 = m.jnt_dofadr[j]
            if m.jnt_type[j] == JNT_FREE:
                for k in range(3):
                    v = v + d.cdof[d0 + k] * d.qvel[d0 + k]
                vt = v.copy()
                for k in range(3, 6):
                    a = a + cross_motion(vt, d.cdof[d0 + k]) * d.qvel[d0 + k]
                    v = v + d.cdof[d0 + k] * d.qvel[d0 + k]
            else:
                a = a + cross_motion(v, d.cdof[d0]) * d.qvel[d0]
                v = v + d.cdof[d0] * d.qvel[d0]
        cvel[b] = v; cacc[b] = a
        cfrc[b] = d.cinert[b] @ a + cross_force(v, d.cinert[b] @ v)
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parent[b]
        if p > 0:
            cfrc[p] += cfrc[b]
    d.cvel = cvel
    d.qfrc_bias = np.array([d.cdof[i] @ cfrc[m.dof_body[i]] for i in range(nv)])
    # <position> actuators: force = clamp(kp*(clamp(ctrl) - gear*q)); qfrc = gear*force   
    d.qfrc_actuator = np.zeros(nv)
    d.actuator_force = np.zeros(m.nu)
    for a in range(m.nu):
        c = d.ctrl[a]
        if m.act_ctrllimited[a]:
            c = min(max(c, m.act_ctrlrange[a, 0]), m.act_ctrlrange[a, 1])
        f = m.act_kp[a] * c - m.act_kp[a] * m.act_gear[a] * d.qpos[m.act_qposadr[a]]
        if m.act_forcelimited[a]:
            f = min(max(f, m.act_forcerange[a, 0]), m.act_forcerange[a, 1])
        d.actuator_force[a] = f
        d.qfrc_actuator[m.act_dof[a]] += m.act_gear[a] * f
    d.qfrc_smooth = d.qfrc_passive - d.qfrc_bias + d.qfrc_actuator
    d.qacc_smooth = np.linalg.solve(d.M, d.qfrc_smooth) if nv else np.zeros(0)


# ----------------------------------------------------------------------------------------------- B.7
def constraint_update(m, d, jar, want_hessian=False):
    """efc_force, constraint cost and (optionally) the per-row / per-cone second derivatives."""
    ne = d.nefc
    force = np.zeros(ne)
    cost = 0.0
    Hrows = np.zeros(ne)  # diagonal (quadratic) second derivatives
    cones = []  # (first,dim,Hc) for contacts in the middle zone
    for i in range(d.nlimit):
        if jar[i] < 0:
            force[i] = -d.efc_D[i] * jar[i]
            cost += 0.5 * d.efc_D[i] * jar[i] * jar[i]
            Hrows[i] = d.efc_D[i]
    for (i, dim, mu, fr) in d.efc_contact:
        if dim == 1:
            if jar[i] < 0:
                force[i] = -d.efc_D[i] * jar[i]; cost += 0.5 * d.efc_D[i] * jar[i] ** 2; Hrows[i] = d.efc_D[i]
            continue
        scl = np.concatenate([[mu], fr[:dim - 1]])
        U = jar[i:i + dim] * scl
        N = U[0]; T = np.linalg.norm(U[1:])
        if N >= mu * T or (T <= 0 and N >= 0):
            continue  # top zone: separated / inside the dual cone
        if mu * N + T <= 0 or (T <= 0 and N < 0):
            sl = slice(i, i + dim)  # bottom zone: fully quadratic
            force[sl] = -d.efc_D[sl] * jar[sl]
            cost += 0.5 * np.sum(d.efc_D[sl] * jar[sl] ** 2)
            Hrows[sl] = d.efc_D[sl]
            continue
        Dm = d.efc_D[i] / (mu * mu * (1 + mu * mu))
        NT = N - mu * T
        cost += 0.5 * Dm * NT * NT
        force[i] = -Dm * NT * mu
        force[i + 1:i + dim] = -force[i] / T * U[1:] * fr[:dim - 1]
        if want_hessian:
            Hu = np.zeros((dim, dim))
            Hu[0, 0] = 1.0
            Hu[0, 1:] = Hu[1:, 0] = -mu * U[1:] / T
            Hu[1:, 1:] = mu * mu * np.outer(U[1:], U[1:]) / (T * T) \
                - mu * NT * (np.eye(dim - 1) / T - np.outer(U[1:], U[1:]) / T ** 3)
            cones.append((i, dim, Dm * (scl[:, None] * Hu * scl[None, :])))
    return force, cost, Hrows, cones


def _total_cost(m, d, qacc):
    jar = d.efc_J @ qacc - d.efc_aref
    _, cost, _, _ = constraint_update(m, d, jar)
    Ma = d.M @ qacc
    return cost + 0.5 * np.dot(Ma - d.qfrc_smooth, qacc - d.qacc_smooth)


def _linesearch(m, d, qacc, jar, Ma, search, jv, Mv, gtol, max_ls):
    """Exact 1-D minimisation of cost(qacc + alpha*search): safeguarded Newton with bracketing."""
    quad_gauss1 = np.dot(search, Ma) - np.dot(d.qfrc_smooth, search)
    quad_gauss2 = 0.5 * np.dot(search, Mv)

    def ev(alpha):
        c = alpha * quad_gauss1 + alpha * alpha * quad_gauss2
        d1 = quad_gauss1 + 2 * alpha * quad_gauss2
        d2 = 2 * quad_gauss2
        x = jar + alpha * jv
        for i in range(d.nlimit):
            if x[i] < 0:
                D = d.efc_D[i]
                c += 0.5 * D * x[i] * x[i]; d1 += D * x[i] * jv[i]; d2 += D * jv[i] * jv[i]
        for (i, dim, mu, fr) in d.efc_contact:
            scl = np.concatenate([[mu], fr[:dim - 1]])
            U = x[i:i + dim] * scl; V = jv[i:i + dim] * scl
            N = U[0]; T = np.linalg.norm(U[1:])
            if N >= mu * T or (T <= 0 and N >= 0):
                continue
            if mu * N + T <= 0 or (T <= 0 and N < 0):
                sl = slice(i, i + dim)
                c += 0.5 * np.sum(d.efc_D[sl] * x[sl] ** 2)
                d1 += np.sum(d.efc_D[sl] * x[sl] * jv[sl]); d2 += np.sum(d.efc_D[sl] * jv[sl] ** 2)
                continue
            Dm = d.efc_D[i] / (mu * mu * (1 + mu * mu))
            NT = N - mu * T
            N1 = V[0]
            T1 = np.dot(U[1:], V[1:]) / T
            T2 = np.dot(V[1:], V[1:]) / T - T1 * T1 / T
            c += 0.5 * Dm * NT * NT
            d1 += Dm * NT * (N1 - mu * T1)
            d2 += Dm * ((N1 - mu * T1) ** 2 - NT * mu * T2)
        return c, d1, d2

    c0, d10, d20 = ev(0.0)
    nev = 1
    # Newton iterations; maintain a bracket [lo, hi] with d1(lo) < 0 < d1(hi) once both signs were seen
    lo, hi = 0.0, None
    alpha = 0.0; d1, d2 = d10, d20
    best = (c0, 0.0)
    for _ in range(max_ls):
        if abs(d1) < gtol:
            break
        step = -d1 / d2 if d2 > MINVAL else (1.0 if d1 < 0 else -1.0)
        nxt = alpha + step
        if hi is not None and not (lo < nxt < hi):
            nxt = 0.5 * (lo + hi)
        if nxt <= 0 and hi is None:
            nxt = alpha * 0.5
        alpha = nxt
        c, d1, d2 = ev(alpha); nev += 1
        if c < best[0]:
            best = (c, alpha)
        if d1 < 0:
            lo = max(lo, alpha)
        else:
            hi = alpha if hi is None else min(hi, alpha)
    d.stats["ls_evals"] = d.stats.get("ls_evals", 0) + nev
    return best[1] if best[0] < c0 or abs(d1) < gtol else 0.0


def solve_newton(m, d):
    """Primal Newton solver on the convex cost (mj_solNewton semantics; elliptic cones)."""
    nv = m.nv
    if d.nefc == 0:
        d.qacc = d.qacc_smooth.copy()
        d.efc_force = np.zeros(0); d.qfrc_constraint = np.zeros(nv); d.solver_iter = 0
        return
    tol, ls_tol = m.opt[5], m.opt[6]
    max_it, max_ls = int(m.opt[7]), int(m.opt[8])
    # warm start: keep whichever of qacc_warmstart / qacc_smooth has the lower cost
    cw = _total_cost(m, d, d.qacc_warmstart)
    cs = _total_cost(m, d, d.qacc_smooth)
    qacc = d.qacc_warmstart.copy() if cw <= cs else d.qacc_smooth.copy()
    scale = 1.0 / (m.opt[11] * max(1, nv))
    it = 0
    Ma = d.M @ qacc
    jar = d.efc_J @ qacc - d.efc_aref
    force, ccost, Hrows, cones = constraint_update(m, d, jar, True)
    cost = ccost + 0.5 * np.dot(Ma - d.qfrc_smooth, qacc - d.qacc_smooth)
    while True:
        grad = Ma - d.qfrc_smooth - d.efc_J.T @ force
        H = d.M + d.efc_J.T @ (Hrows[:, None] * d.efc_J)
        for (i, dim, Hc) in cones:
            Jc = d.efc_J[i:i + dim]
            H += Jc.T @ Hc @ Jc
        L = np.linalg.cholesky(H)
        search = -np.linalg.solve(L.T, np.linalg.solve(L, grad))
        if it >= max_it:
            break
        snorm = np.linalg.norm(search)
        if snorm < MINVAL:
            break
        Mv = d.M @ search; jv = d.efc_J @ search
        gtol = tol * ls_tol * snorm / scale
        alpha = _linesearch(m, d, qacc, jar, Ma, search, jv, Mv, gtol, max_ls)
        if alpha == 0.0:
            break
        qacc = qacc + alpha * search; Ma = Ma + alpha * Mv; jar = jar + alpha * jv
        old = cost
        force, ccost, Hrows, cones = constraint_update(m, d, jar, True)
        cost = ccost + 0.5 * np.dot(Ma - d.qfrc_smooth, qacc - d.qacc_smooth)
        it += 1
        grad = Ma - d.qfrc_smooth - d.efc_J.T @ force
        if scale * (old - cost) < tol or scale * np.linalg.norm(grad) < tol:
            break
    d.qacc = qacc
    d.efc_force = force
    d.qfrc_constraint = d.efc_J.T @ force
    d.solver_iter = it
    d.stats["newton_iter"] = d.stats.get("newton_iter", 0) + it


# ----------------------------------------------------------------------------------------------- B.8
def euler(m, d):
    dt = m.opt[0]
    nv = m.nv
    if np.any(m.dof_damping > 0):
        qacc_int = np.linalg.solve(d.M + dt * np.diag(m.dof_damping), d.qfrc_smooth + d.qfrc_constraint)
    else:
        qacc_int = d.qacc
    d.qacc_warmstart = d.qacc.copy()
    d.qvel = d.qvel + dt * qacc_int
    for j in range(m.njnt):
        a, v = m.jnt_qposadr[j], m.jnt_dofadr[j]
        if m.jnt_type[j] == JNT_FREE:
            d.qpos[a:a + 3] += dt * d.qvel[v:v + 3]
            w = d.qvel[v + 3:v + 6]
            ang = np.linalg.norm(w)
            q = qnormalize(d.qpos[a + 3:a + 7])
            if ang * dt > MINVAL:
                half = 0.5 * ang * dt
                dq = np.concatenate([[np.cos(half)], np.sin(half) * w / ang])
                q = qmul(q, dq)
            d.qpos[a + 3:a + 7] = qnormalize(q)
        else:
            d.qpos[a] += dt * d.qvel[v]
    d.time += dt


# ----------------------------------------------------------------------------------------------- drivers
def forward(m, d, midphase=True):
    """mj_forward: position -> velocity -> actuation -> acceleration -> constraint."""
    d.stats = {}
    kinematics(m, d)
    com_crb(m, d)
    collision(m, d, midphase)
    make_constraint(m, d)
    smooth_dynamics(m, d)
    solve_newton(m, d)


def step(m, d, midphase=True):
    """mj_step = mj_forward + mj_Euler (sim.step(), /root/reference/hsr/env.py:123)."""
    forward(m, d, midphase)
    euler(m, d)


def body_xpos(m, d, body):
    return d.xpos[body].copy()


def flops_estimate(m, d):
    """Algorithmic flop count of the substep just executed, per the stage formulas of SURVEY.md §8(d)."""
    nb = m.nbody - 1
    nc = len(d.contacts); ne = d.nefc; nv = m.nv
    nnz = nv
    f = 100 * nb + 60 * m.ngeom + 9 * m.npair
    f += 80 * d.stats.get("narrowphase", 0) + 2 * 246 * d.stats.get("mpr_support", 0)
    f += 150 * nc + 2 * ne * nnz + 60 * nc + 10 * nv + 150 * nb
    f += 2 * (2 * ne * nnz + 50 * nc + 4 * nv)
    it = getattr(d, "solver_iter", 0)
    f += max(it, 1 if ne else 0) * (50 * nc + ne * nnz * nnz + nv ** 3 // 3 + 2 * ne * nnz + 2 * nv * nv)
    f += d.stats.get("ls_evals", 0) * (2 * ne + 70 * nc)
    f += 6 * nv + 60 * m.nblock + 10
    return f
