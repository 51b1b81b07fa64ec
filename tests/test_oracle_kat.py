"""Analytic known-answer tests of the fp64 numpy oracle: the oracle is pinned by these
(no MuJoCo binary, golden vector or reference test exists for this path: parity unpinned, see oracle/mjstep.py)."""
import numpy as np
import pytest

from oracle import mjstep


def fresh(model):
    d = mjstep.Data(model)
    return d


def test_kat1_pd_actuator_with_implicit_damping(models):
    """C1, no contact / limit: v' = v + dt (gear clamp(kp(u - gear q), F) - D v)/(m + dt D), q' = q + dt v'."""
    m = models["c1_readme"]
    d = fresh(m)
    d.qpos[:] = [0.05, -0.03]; d.qvel[:] = [0.02, -0.01]; d.ctrl[:] = [0.4, -0.2]
    dt = m.opt[0]
    mass = m.body_mass[1]
    for _ in range(5):
        q, v = d.qpos.copy(), d.qvel.copy()
        mjstep.step(m, d)
        assert d.nefc == 0
        for i in range(2):
            u = np.clip(d.ctrl[i], *m.act_ctrlrange[i])
            f = np.clip(m.act_kp[i] * (u - m.act_gear[i] * q[i]), *m.act_forcerange[i])
            vn = v[i] + dt * (m.act_gear[i] * f - m.dof_damping[i] * v[i]) / (mass + dt * m.dof_damping[i])
            assert d.qvel[i] == pytest.approx(vn, rel=1e-12, abs=1e-15)
            assert d.qpos[i] == pytest.approx(q[i] + dt * vn, rel=1e-12, abs=1e-15)


def test_kat2_free_fall_and_gyroscopic(models):
    """Block in the air: v_z' = v_z - g dt; |quat| = 1; I w' = I w - dt w x I w (body frame, explicit)."""
    m = models["c2_push"]
    d = fresh(m)
    d.qpos[2:5] = [0.3, 0.5, 1.0]
    d.qpos[5:9] = [np.cos(.4), np.sin(.4) * .6, 0, np.sin(.4) * .8]
    d.qvel[2:8] = [0.1, -0.2, 0.3, 2.0, -1.0, 0.5]
    dt, g = m.opt[0], -m.opt[3]
    I = m.body_inertia[2][:3]
    for _ in range(10):
        v = d.qvel.copy()
        mjstep.step(m, d)
        assert len(d.contacts) == 0
        assert d.qvel[4] == pytest.approx(v[4] - g * dt, rel=1e-12)
        assert d.qvel[2] == pytest.approx(v[2], abs=1e-14) and d.qvel[3] == pytest.approx(v[3], abs=1e-14)
        w = v[5:8]
        wn = w - dt * np.cross(w, I * w) / I
        assert np.allclose(d.qvel[5:8], wn, rtol=1e-10, atol=1e-13)
        assert np.linalg.norm(d.qpos[5:9]) == pytest.approx(1.0, abs=1e-14)


def test_kat3_resting_block_carries_its_weight(models):
    """Block settled on the floor: sum of normal contact forces = m g; 4 contacts, symmetric."""
    m = models["c2_push"]
    d = fresh(m)
    d.qpos[2:5] = [0.3, 0.4, 0.017]; d.qpos[5:9] = [1, 0, 0, 0]
    for _ in range(600):
        mjstep.step(m, d)
    assert len(d.contacts) == 4
    normal = sum(d.efc_force[c[0]] for c in d.efc_contact)
    assert normal == pytest.approx(m.body_mass[2] * 9.81, rel=1e-5)
    f = [d.efc_force[c[0]] for c in d.efc_contact]
    assert max(f) - min(f) < 1e-5
    assert abs(d.qvel[2:8]).max() < 1e-7
    # equilibrium penetration consistent with the reference acceleration: aref = a_normal = force/D per corner
    c0 = d.efc_contact[0][0]
    assert d.efc_aref[c0] > 0 and d.efc_pos[c0] < 0


def test_kat4_joint_limit_holds(models):
    """Base driven into its lower slide_x limit (the upper one is out of reach: the pan edge stops the base first,
    SURVEY.md App. A.6): the limit row activates and stops it within a millimetre."""
    m = models["c1_readme"]
    d = fresh(m)
    d.qpos[:] = [-0.115, 0.0]; d.ctrl[:] = [-1.0, 0.0]
    active = 0
    for _ in range(1500):
        mjstep.step(m, d)
        active += d.nlimit
    lo = m.jnt_range[0, 0]
    assert active > 0
    assert len(d.contacts) == 0
    assert lo - 2e-3 < d.qpos[0] < lo
    assert abs(d.qvel[0]) < 1e-6
    # steady state: constraint force balances the saturated actuator (gear * forcemax)
    assert d.qfrc_constraint[0] == pytest.approx(m.act_gear[0] * m.act_forcerange[0, 1], rel=1e-6)


def test_kat5_mirror_symmetry(models):
    """The floor-block-pan subsystem is symmetric under y -> -y: mirrored initial conditions, mirrored trajectory."""
    m = models["c2_push"]

    def run(sign):
        d = fresh(m)
        d.qpos[0:2] = [-0.1, 0.0]
        d.qpos[2:5] = [0.35, 0.3 * sign, 0.05]
        a = 0.3 * sign
        d.qpos[5:9] = [np.cos(a), 0, 0, np.sin(a)]
        d.qvel[2:8] = [0.2, 0.1 * sign, 0.0, 0.5 * sign, 0.0, 1.0 * sign]   # (wx, wz flip; wy keeps)
        for _ in range(300):
            mjstep.step(m, d)
        return d

    a, b = run(+1), run(-1)
    assert len(a.contacts) > 0
    S = np.array([1, -1, 1])
    assert np.allclose(a.qpos[2:5], b.qpos[2:5] * S, atol=1e-9)
    assert np.allclose(a.qvel[2:5], b.qvel[2:5] * S, atol=1e-8)


def test_kat6_zero_norm_quaternion_becomes_identity(models):
    """README block-space (0,0)x4 writes a zero quaternion: MuJoCo's normalisation turns it into identity."""
    m = models["c2_push"]
    d = fresh(m)
    d.qpos[2:5] = [0.3, 0.3, 0.5]; d.qpos[5:9] = 0
    mjstep.forward(m, d)
    assert np.array_equal(d.qpos[5:9], [1, 0, 0, 0])
    assert np.allclose(d.xmat[2], np.eye(3))


def test_solver_satisfies_optimality(models):
    """At the solver's output the gradient M a - qfrc_smooth - J^T f vanishes and forces obey the friction cone."""
    m = models["c5_clutter"]
    d = fresh(m)
    rng = np.random.default_rng(0)
    d.qvel[:] = rng.normal(0, .2, m.nv)
    d.ctrl[:] = [1, -1]
    for _ in range(30):
        mjstep.step(m, d)
        if d.nefc:
            grad = d.M @ d.qacc - d.qfrc_smooth - d.qfrc_constraint
            assert np.abs(grad).max() < 1e-6 * max(1.0, np.abs(d.qfrc_smooth).max())
            for (i, dim, mu, fr) in d.efc_contact:
                f = d.efc_force[i:i + dim]
                assert f[0] >= -1e-9
                assert np.linalg.norm(f[1:] / fr[:dim - 1]) <= f[0] * (1 + 1e-6) + 1e-9
