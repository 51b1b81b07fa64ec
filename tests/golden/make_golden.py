"""Generates the committed golden fixtures of the hot path: teacher-forcing states along oracle trajectories and the
one-substep result of the fp64 numpy oracle (oracle/mjstep.py) for each of them.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

PARITY UNPINNED: the reference holds no golden vectors and MuJoCo cannot run here, so these vectors pin
the *oracle* (and through it the CUDA path) against regressions; they are not MuJoCo outputs.  When MuJoCo is available,
tools/dump_mujoco_golden.py writes files of the same layout from the real engine.

Layout of <model>.npz: qpos, qvel, warm, ctrl (inputs, exactly representable in fp32); qpos1, qvel1, qacc (outputs of
one mj_step), ncon, nefc (contact / constraint-row counts), success (goal test with geofence .05 and the goal stored in
`mocap`).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from hsr_env_b200.model import Model  # noqa: E402
from oracle import mjstep, port  # noqa: E402
from scenarios import rollout_states  # noqa: E402

CASES = {"c1_readme": (12, False), "c1b_readme_block": (12, False), "c2_push": (24, False), "c3_arm": (8, True), "c5_clutter": (8, False),
         "f2_cupboard": (8, True)}


def main():
    out = Path(__file__).resolve().parent
    only = sys.argv[1:]
    for name, (n, pan) in CASES.items():
        if only and name not in only:
            continue
        model = Model.load(ROOT / "hsr_env_b200" / "blobs" / f"{name}.hsrb")
        cp = port.CpuPort(model)
        qpos, qvel, warm, ctrl = rollout_states(cp, model, n, seed=1234 + len(name), pan=pan, float32=True)
        rng = np.random.default_rng(len(name))
        res = dict(qpos=qpos, qvel=qvel, warm=warm, ctrl=ctrl, qpos1=np.zeros_like(qpos), qvel1=np.zeros_like(qvel),
                   qacc=np.zeros_like(qvel), ncon=np.zeros(n, np.int32), nefc=np.zeros(n, np.int32),
                   mocap=np.zeros((n, 3)), success=np.zeros(n, np.uint8))
        for e in range(n):
            d = mjstep.Data(model)
            d.qpos[:] = qpos[e]; d.qvel[:] = qvel[e]; d.qacc_warmstart[:] = warm[e]; d.ctrl[:] = ctrl[e]
            mjstep.step(model, d)
            res["qpos1"][e], res["qvel1"][e], res["qacc"][e] = d.qpos, d.qvel, d.qacc
            res["ncon"][e], res["nefc"][e] = len(d.contacts), d.nefc
            if model.nblock:
                # a goal 4.5 .. 5.5 cm from the first block: straddles the geofence
                b = int(model.block_body[0])
                p = mjstep.body_xpos(model, d, b)
                r, a = rng.uniform(.045, .055), rng.uniform(0, 2 * np.pi)
                g = (p + np.array([r * np.cos(a), r * np.sin(a), 0.0])).astype(np.float32).astype(np.float64)
                res["mocap"][e] = g
                # HSREnv.step: all(in_range(block_i, goal, geofence))  (hsr/env.py:126)
                res["success"][e] = all(np.linalg.norm(mjstep.body_xpos(model, d, int(bb)) - g) < np.float32(.05)
                                        for bb in model.block_body)
        np.savez_compressed(out / f"{name}.npz", **res)
        print(name, "contacts", res["ncon"].tolist())


if __name__ == "__main__":
    main()
