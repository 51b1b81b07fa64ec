"""Seeded state generators shared by the CPU and GPU parity tests: teacher-forcing states are taken along
trajectories of the fp64 oracle port so that they cover free motion, joint limits, resting and pushing contacts."""
import numpy as np

BLOCK_LO, BLOCK_HI = np.array([-.25, -.2, 0., -1.]), np.array([-.05, .1, 1., 1.])
GOAL_LO, GOAL_HI = np.array([-.15, -.2, .017]), np.array([0., .1, .017])


def block_adr(model):
    """qpos addresses of the free joints of the blocks."""
    return [int(model.jnt_qposadr[model.body_jntadr[b]]) for b in model.block_body]


BASE_EDGE_X = -0.234  # x of the front edge of the base hull at slide_x = 0


def initial_states(model, n, seed, pan=False):
    """n start states: robot dofs inside their ranges; blocks resting (slightly sunk, as at equilibrium) on the
    floor just in front of the base so that the base reaches them within a few dozen substeps (or on the pan),
    random yaw, not overlapping each other or the robot."""
    rng = np.random.default_rng(seed)
    q = np.tile(model.qpos0, (n, 1))
    first_slide = min(j for j in range(model.njnt) if model.jnt_type[j] != 0)
    cupboard = "blockjoint" in list(model.names.get("joint", []))
    for j in range(model.njnt):
        if model.jnt_type[j] == 0:
            continue
        a = model.jnt_qposadr[j]
        lo, hi = (model.jnt_range[j] if model.jnt_limited[j] else (-0.5, 0.0))
        if j == first_slide:
            # slide_x: keeps the base clear of the pan edge and of the joint limits; in the cupboard scene (its own
            # `blockjoint`) also keeps the arm out of the cupboard doors (16 cm deep at the upper joint limit)
            lo, hi = (-0.05, 0.02) if cupboard else (-0.05, 0.12)
        q[:, a] = rng.uniform(lo, hi, n)
    for e in range(n):
        placed = []
        edge = BASE_EDGE_X + q[e, 0]
        for a in block_adr(model):
            for _ in range(200):
                if pan:
                    xy = rng.uniform([-.1, -.2], [.1, .2])
                else:
                    xy = np.array([edge + 0.0565 + rng.uniform(0.0, 0.012) + 0.13 * (len(placed) // 2),
                                   q[e, 1] - 0.08 + rng.uniform(-.12, .12)])
                if all(np.linalg.norm(xy - p) > 0.125 for p in placed):
                    break
            placed.append(xy)
            yaw = rng.uniform(-np.pi, np.pi)
            q[e, a:a + 3] = [xy[0], xy[1], (.405 + .017) if pan else .017]
            q[e, a + 3:a + 7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    return q


def rollout_states(port, model, n, seed, n_actions=3, substeps=(1, 120), pan=False, float32=False):
    """States (qpos, qvel, warm, ctrl) reached after a few random actions of random length with the fp64 port."""
    rng = np.random.default_rng(seed + 1)
    qpos = initial_states(model, n, seed, pan)
    qvel = np.zeros((n, model.nv)); warm = np.zeros((n, model.nv))
    lo, hi = model.act_ctrlrange[:, 0], model.act_ctrlrange[:, 1]
    ctrl = rng.uniform(lo, hi, (n, model.nu))
    for a in range(n_actions):
        ctrl = rng.uniform(lo, hi, (n, model.nu))
        if not pan:
            ctrl[:, 0] = np.abs(ctrl[:, 0])  # drive towards the blocks
        for e in range(n):
            k = int(rng.integers(substeps[0], substeps[1] + 1))
            out = port.step(qpos[e], qvel[e], warm[e], ctrl[e], nsub=k)
            qpos[e], qvel[e], warm[e] = out["qpos"][0], out["qvel"][0], out["warm"][0]
    if float32:  # states exactly representable in fp32 so that GPU and oracle start from identical numbers
        qpos, qvel, warm, ctrl = [x.astype(np.float32).astype(np.float64) for x in (qpos, qvel, warm, ctrl)]
    return qpos, qvel, warm, ctrl


def rel_err(got, want):
    """one-step error of a state vector: max |diff| / max(1, max |want|)  (per environment)."""
    got = np.atleast_2d(got); want = np.atleast_2d(want)
    return np.abs(got - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
