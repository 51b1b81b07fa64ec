"""Generates the committed golden fixtures of the hot path: teacher-forcing states along oracle trajectories and the
one-substep result of the INDEPENDENT fp64 numpy oracle (oracle/mjstep.py) for each of them, 256 states per model.

    python tests/golden/make_golden.py [model ...]          # rewrites tests/golden/*.npz

PARITY UNPINNED: the reference holds no golden vectors and MuJoCo cannot run here, so these vectors pin
the *oracle* (and through it the CUDA path) against regressions; they are not MuJoCo outputs.  When MuJoCo is available,
tools/dump_mujoco_golden.py writes files of the same layout from the real engine.

Layout of <model>.npz: qpos, qvel, warm, ctrl (inputs, exactly representable in fp32); qpos1, qvel1, qacc (outputs of
one mj_step), ncon, nefc, nlimit (contact / constraint-row / limit-row counts), mocap + success (goal test with geofence
.05 and the goal stored in `mocap`, placed 4.5 .. 5.5 cm from block 0 with a margin of at least 1e-5 to the geofence),
regime (bit mask over REGIMES: which kinds of constraint rows the state holds), overflow (1 where the state holds more
contacts than the kernels' default capacity for the model: the CUDA path flags it in bad_state and drops the excess),
sensitive (1 where the one-step map is ill-defined at the precision of the comparison: the oracle's own result moves by
more than 2e-5 when the inputs are perturbed by one fp32 ulp, or the two fp64 implementations of the same algorithm -
numpy and the g++ port - disagree by more than 1e-6, or the fp32 port's result moves by more than 2e-5 when its
hull-vertex scans get noise of the size of an fp32 rounding error.  Cause: the portal refinement (libccd MPR) ends on a face of the
Minkowski difference, where several hull vertices tie for the support point in exact arithmetic and rounding picks
the final portal; closest point and normal then differ by O(0.1).  MuJoCo itself is subject to it; no fp32
implementation can be held to 1e-4 there).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from hsr_env_b200.model import Model  # noqa: E402
from oracle import mjstep, port  # noqa: E402
from scenarios import regimes, rel_err, rollout_states  # noqa: E402

N = 256
CASES = {"c1_readme": False, "c1b_readme_block": False, "c2_push": False, "c3_arm": True, "c5_clutter": False,
         "f2_cupboard": True}
REGIMES = ["limit", "world-block", "base-block", "base-world", "arm-block", "arm-world", "block-block", "pan-block"]


def oracle_step(model, qpos, qvel, warm, ctrl):
    d = mjstep.Data(model)
    d.qpos[:] = qpos; d.qvel[:] = qvel; d.qacc_warmstart[:] = warm; d.ctrl[:] = ctrl
    mjstep.step(model, d)
    return d


def main():
    out = Path(__file__).resolve().parent
    only = sys.argv[1:]
    for name, pan in CASES.items():
        if only and name not in only:
            continue
        model = Model.load(ROOT / "hsr_env_b200" / "blobs" / f"{name}.hsrb")
        cp = port.CpuPort(model)
        n = N
        qpos, qvel, warm, ctrl = rollout_states(cp, model, n, seed=1234 + len(name), pan=pan, float32=True)
        rng = np.random.default_rng(len(name))
        res = dict(qpos=qpos, qvel=qvel, warm=warm, ctrl=ctrl, qpos1=np.zeros_like(qpos), qvel1=np.zeros_like(qvel),
                   qacc=np.zeros_like(qvel), ncon=np.zeros(n, np.int32), nefc=np.zeros(n, np.int32), nlimit=np.zeros(n, np.int32),
                   mocap=np.zeros((n, 3)), success=np.zeros(n, np.uint8), regime=np.zeros(n, np.int32),
                   sensitive=np.zeros(n, np.uint8), overflow=np.zeros(n, np.uint8))
        cp.set_caps(64, 64 * 6 + 8)     # regime labels and the port cross-check: no capacity limit, like the numpy oracle
        pref = cp.step(qpos, qvel, warm, ctrl, nsub=1)
        dbg = cp.step(qpos, qvel, warm, ctrl, nsub=1, debug=True)["debug"]   # contact pairs of the forward pass (regime labels)
        cap = port.CpuPort(model).ncon_max                                  # the kernels' default capacity
        for e in range(n):
            d = oracle_step(model, qpos[e], qvel[e], warm[e], ctrl[e])
            res["qpos1"][e], res["qvel1"][e], res["qacc"][e] = d.qpos, d.qvel, d.qacc
            res["ncon"][e], res["nefc"][e] = len(d.contacts), d.nefc
            res["nlimit"][e] = dbg[e]["nlimit"]
            assert dbg[e]["ncon"] == len(d.contacts) and dbg[e]["nefc"] == d.nefc, (name, e)
            for r in regimes(model, dbg[e]):
                res["regime"][e] |= 1 << REGIMES.index(r)
            # sensitivity of the oracle's own one-step map to one-ulp (fp32) input perturbations
            worst = 0.0
            res["overflow"][e] = len(d.contacts) > cap
            worst = max(worst, float(rel_err(pref["qvel"][e], d.qvel)[0]) * 20, float(rel_err(pref["qpos"][e], d.qpos)[0]) * 20)   # port vs numpy: 1e-6
            for trial in range(4):
                sg = rng.choice([-1.0, 1.0], size=model.nq)
                qp = np.nextafter(qpos[e].astype(np.float32), (sg * np.inf).astype(np.float32)).astype(np.float64)
                dp = oracle_step(model, qp, qvel[e], warm[e], ctrl[e])
                worst = max(worst, float(rel_err(dp.qvel, d.qvel)[0]), float(rel_err(dp.qpos, d.qpos)[0]))
            res["sensitive"][e] = worst > 2e-5
            if model.nblock:
                # a goal 4.5 .. 5.5 cm from the first block: straddles the geofence, clear of it by >= 1e-5
                b = int(model.block_body[0])
                p = mjstep.body_xpos(model, d, b)
                while True:
                    r, a = rng.uniform(.045, .055), rng.uniform(0, 2 * np.pi)
                    g = (p + np.array([r * np.cos(a), r * np.sin(a), 0.0])).astype(np.float32).astype(np.float64)
                    if abs(np.linalg.norm(p - g) - np.float32(.05)) > 1e-5:
                        break
                res["mocap"][e] = g
                # HSREnv.step: all(in_range(block_i, goal, geofence))  (hsr/env.py:126)
                res["success"][e] = all(np.linalg.norm(mjstep.body_xpos(model, d, int(bb)) - g) < np.float32(.05)
                                        for bb in model.block_body)
        # ... or an fp32 implementation's result depends on how near-ties of the hull-vertex scan (support values within
        # an fp32 rounding error, ~3e-8 m) are rounded: the fp32 port with rounding-sized noise on the scanned values
        base = cp.step(qpos, qvel, warm, ctrl, nsub=1, use_float=True)
        for trial in range(6):
            cp.set_scan_noise(1000 + trial)
            try:
                pert = cp.step(qpos, qvel, warm, ctrl, nsub=1, use_float=True)
            finally:
                cp.set_scan_noise(0)
            res["sensitive"] |= (rel_err(pert["qvel"], base["qvel"]) > 2e-5) | (rel_err(pert["qpos"], base["qpos"]) > 2e-5)
        np.savez_compressed(out / f"{name}.npz", **res)
        counts = {r: int(((res["regime"] >> i) & 1).sum()) for i, r in enumerate(REGIMES) if ((res["regime"] >> i) & 1).any()}
        print(name, "states", n, "regimes", counts, "sensitive", int(res["sensitive"].sum()), "overflow", int(res["overflow"].sum()),
              "max ncon", int(res["ncon"].max()))


if __name__ == "__main__":
    main()
