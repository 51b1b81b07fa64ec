"""Seeded state generators shared by the CPU and GPU parity tests: teacher-forcing states are taken along
trajectories of the fp64 oracle port so that they cover free motion, joint limits, resting and pushing contacts,
robot-hull / pan contacts, gripper / arm contacts with the block on the pan (configs[2]) and block-block box-box
contacts (configs[4]).  `regimes()` classifies a state by the constraint rows its forward pass produces, so that the
tests can assert that each regime is actually present."""
import numpy as np

BLOCK_LO, BLOCK_HI = np.array([-.25, -.2, 0., -1.]), np.array([-.05, .1, 1., 1.])
GOAL_LO, GOAL_HI = np.array([-.15, -.2, .017]), np.array([0., .1, .017])


def block_adr(model):
    """qpos addresses of the free joints of the blocks."""
    return [int(model.jnt_qposadr[model.body_jntadr[b]]) for b in model.block_body]


BASE_EDGE_X = -0.234  # x of the front edge of the base hull at slide_x = 0
ARM_LO = np.array([-.05, -.1, 0., -1.9, -1.57, 0., 0.])      # slide_x, slide_y, arm_lift, arm_flex, wrist_roll, hand_l, hand_r:
ARM_HI = np.array([.12, .1, .12, -1.25, 1.57, .349, .349])   # the hand hovers at the height of a block lying on the pan


def _joint_names(model):
    return list(model.names.get("joint", []))


def _has_arm(model):
    return "arm_flex_joint" in _joint_names(model)


def initial_states(model, n, seed, pan=False, port=None):
    """n start states: robot dofs inside their ranges; blocks resting (slightly sunk, as at equilibrium) on the
    floor just in front of the base so that the base reaches them within a few dozen substeps (or on the pan),
    random yaw, not overlapping each other or the robot.  On top of that, by environment index:
      * e % 6 == 5: a slide joint starts within 4 mm of (or just beyond) one of its limits (limit rows);
      * no-block model, e % 4 == 3: the base starts against the pan edge (robot-hull / pan contacts);
      * several blocks, e % 3 == 0: blocks 0 and 1 start face to face, 2 mm into each other (box-box contacts);
      * arm model on the pan (needs `port` for the kinematics), e % 5 < 4: the hand starts just above the pan and the
        block is put under / between the fingers (gripper-block, gripper-pan contacts)."""
    rng = np.random.default_rng(seed)
    q = np.tile(model.qpos0, (n, 1))
    slides = [j for j in range(model.njnt) if model.jnt_type[j] == 1]
    first_slide = min(j for j in range(model.njnt) if model.jnt_type[j] != 0)
    cupboard = "blockjoint" in _joint_names(model)
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
    if not cupboard:
        for e in range(5, n, 6):       # near a slide limit
            j = slides[(e // 6) % len(slides)]
            if j == first_slide and (e // 12) % 2 == 0 and not model.nblock == 0:
                j = slides[-1]         # mostly slide_y: slide_x's upper limit is behind the pan edge
            lo, hi = model.jnt_range[j]
            side = (e // 6 // len(slides)) % 2
            if j == first_slide:
                side = 0               # lower limit of slide_x: the base backs away from the pan
            q[e, model.jnt_qposadr[j]] = (lo + rng.uniform(-.002, .004)) if side == 0 else (hi - rng.uniform(-.002, .004))
        if model.nblock == 0:
            for e in range(3, n, 4):   # base against the pan edge
                q[e, model.jnt_qposadr[first_slide]] = rng.uniform(.155, .165)
    arm = _has_arm(model) and pan and port is not None
    hand = None
    if arm:
        # hand position for candidate arm configurations (block far away), keep those hovering over the pan
        cand = np.tile(model.qpos0, (4 * n, 1))
        cand[:, :7] = rng.uniform(ARM_LO, ARM_HI, (4 * n, 7))
        a0 = block_adr(model)[0]
        cand[:, a0:a0 + 3] = [0, 0, 5.0]
        dbg = port.step(cand, np.zeros((4 * n, model.nv)), np.zeros((4 * n, model.nv)), np.zeros((4 * n, model.nu)), nsub=1, debug=True)["debug"]
        names = list(model.names["body"])
        bl, br = names.index("hand_l_proximal_link"), names.index("hand_r_proximal_link")
        mid = np.stack([0.5 * (r["xpos"][bl] + r["xpos"][br]) for r in dbg])
        ok = np.nonzero((mid[:, 2] > .43) & (mid[:, 2] < .53) & (np.abs(mid[:, 0]) < .12) & (np.abs(mid[:, 1]) < .2))[0]
        hand = (cand, mid, ok)
    for e in range(n):
        placed = []
        edge = BASE_EDGE_X + q[e, 0]
        grip = arm and e % 5 < 4 and len(hand[2]) > 0
        if grip:
            cand, mid, ok = hand
            c = ok[e % len(ok)]
            q[e, :7] = cand[c, :7]
        for ib, a in enumerate(block_adr(model)):
            for _ in range(200):
                if grip:
                    xy = np.clip(mid[c, :2] + rng.uniform(-.03, .03, 2), [-.12, -.2], [.12, .2])
                elif pan:
                    xy = rng.uniform([-.1, -.2], [.1, .2])
                else:
                    xy = np.array([edge + 0.0565 + rng.uniform(0.0, 0.012) + 0.13 * (len(placed) // 2),
                                   q[e, 1] - 0.08 + rng.uniform(-.12, .12)])
                if all(np.linalg.norm(xy - p) > 0.125 for p in placed):
                    break
            yaw = rng.uniform(-np.pi, np.pi)
            if ib == 1 and e % 3 == 0 and not pan:
                # face to face with block 0: same yaw, centres 2 * 0.025 - 0.002 apart along the blocks' y axis
                yaw = yaw0
                xy = placed[0] + (2 * .025 - .002) * np.array([-np.sin(yaw), np.cos(yaw)])
            if ib == 0:
                yaw0 = yaw
            placed.append(xy)
            q[e, a:a + 3] = [xy[0], xy[1], (.405 + .017) if pan else .017]
            q[e, a + 3:a + 7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    return q


def rollout_states(port, model, n, seed, n_actions=3, substeps=(1, 120), pan=False, float32=False):
    """States (qpos, qvel, warm, ctrl) reached after a few random actions of random length with the fp64 port."""
    rng = np.random.default_rng(seed + 1)
    qpos = initial_states(model, n, seed, pan, port)
    qvel = np.zeros((n, model.nv)); warm = np.zeros((n, model.nv))
    lo, hi = model.act_ctrlrange[:, 0], model.act_ctrlrange[:, 1]
    ctrl = rng.uniform(lo, hi, (n, model.nu))
    slides = [j for j in range(model.njnt) if model.jnt_type[j] == 1]
    cupboard = "blockjoint" in _joint_names(model)
    arm = _has_arm(model) and pan
    for a in range(n_actions):
        ctrl = rng.uniform(lo, hi, (n, model.nu))
        if not pan:
            ctrl[:, 0] = np.abs(ctrl[:, 0])  # drive towards the blocks
        for e in range(n):
            k = int(rng.integers(substeps[0], substeps[1] + 1))
            if not cupboard and e % 6 == 5:
                # keep pushing into the limit the environment started at (short roll-outs: the row stays active)
                for iu in range(model.nu):
                    j = int(model.dof_jnt[model.act_dof[iu]])
                    if j in slides:
                        ql, qh = model.jnt_range[j]
                        qj = qpos[e, model.jnt_qposadr[j]]
                        if qj < ql + .01:
                            ctrl[e, iu] = lo[iu]
                        elif qj > qh - .01:
                            ctrl[e, iu] = hi[iu]
                k = min(k, 25)
            if model.nblock == 0 and e % 4 == 3:
                ctrl[e, 0] = hi[0]; k = min(k, 40)      # keep the base against the pan edge
            if model.nblock > 1 and e % 3 == 0 and not pan:
                k = min(k, 8)                            # the face-to-face blocks are still pressed together
            if arm and e % 5 < 4:
                k = min(k, 30)                           # the position servo lifts the arm off the pan within ~50 substeps
            out = port.step(qpos[e], qvel[e], warm[e], ctrl[e], nsub=k)
            qpos[e], qvel[e], warm[e] = out["qpos"][0], out["qvel"][0], out["warm"][0]
    if float32:  # states exactly representable in fp32 so that GPU and oracle start from identical numbers
        qpos, qvel, warm, ctrl = [x.astype(np.float32).astype(np.float64) for x in (qpos, qvel, warm, ctrl)]
    return qpos, qvel, warm, ctrl


def regimes(model, dbg):
    """Classify the forward pass of one state (an entry of port.step(..., debug=True)['debug'], or any dict with
    nlimit / con_pair) by the constraint rows it holds: returns a set of
    'limit', 'world-block' (plane-box), 'base-block' / 'base-world' (robot base hulls), 'arm-block' / 'arm-world'
    (arm, wrist, hand hulls), 'block-block', 'pan-block' (box-box)."""
    names = list(model.names["body"])
    base = names.index("base_link") if "base_link" in names else -1
    blocks = set(int(b) for b in model.block_body)
    out = set()
    if dbg["nlimit"] > 0:
        out.add("limit")
    for pk in dbg["con_pair"]:
        g1, g2 = int(model.pair_geom1[pk]), int(model.pair_geom2[pk])
        b1, b2 = int(model.geom_body[g1]), int(model.geom_body[g2])
        bs = {b1, b2}
        if bs <= blocks and len(bs) == 2:
            out.add("block-block")
        elif bs & blocks:
            other = (bs - blocks).pop()
            if other == 0:
                out.add("world-block" if int(model.geom_type[g1 if b1 == 0 else g2]) == 0 else "pan-block")
            else:
                out.add("base-block" if other == base else "arm-block")
        else:
            other = (bs - {0}).pop() if bs - {0} else 0
            out.add("base-world" if other == base else "arm-world")
    return out


def regime_counts(port, model, qpos, qvel, warm, ctrl):
    """Number of states in each regime (forward pass of the fp64 port)."""
    dbg = port.step(qpos, qvel, warm, ctrl, nsub=1, debug=True)["debug"]
    counts = {}
    for d in dbg:
        for r in regimes(model, d):
            counts[r] = counts.get(r, 0) + 1
    return counts


def rel_err(got, want):
    """one-step error of a state vector: max |diff| / max(1, max |want|)  (per environment)."""
    got = np.atleast_2d(got); want = np.atleast_2d(want)
    return np.abs(got - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
