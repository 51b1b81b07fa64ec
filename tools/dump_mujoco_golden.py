"""Pinning hook: on a machine where MuJoCo imports (`mujoco` official bindings, or `mujoco_py`), dump
(1) the compiled mjModel fields the model blob needs and (2) teacher-forcing vectors
(qpos, qvel, ctrl, qacc_warmstart) -> (qpos', qvel') of the real engine, in the layout of tests/golden/*.npz, so that
the oracle (and the CUDA path) can be pinned against MuJoCo itself.  Not runnable in this container (no MuJoCo).

    python tools/dump_mujoco_golden.py /path/to/hsr/models/world.xml out_dir [--use-dof slide_x slide_y] [--n-blocks 1]
"""
import argparse
import sys
from pathlib import Path

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("xml")
    ap.add_argument("out")
    ap.add_argument("--states", help="npz with qpos/qvel/warm/ctrl to teacher-force (default: tests/golden/<name>.npz)")
    args = ap.parse_args()
    try:
        import mujoco
    except ImportError:
        sys.exit("MuJoCo is not importable here: this script is the hook for a machine that has it (SURVEY.md §8c)")
    model = mujoco.MjModel.from_xml_path(args.xml)
    data = mujoco.MjData(model)
    out = Path(args.out); out.mkdir(parents=True, exist_ok=True)
    fields = ["body_mass", "body_inertia", "body_ipos", "body_iquat", "body_invweight0", "dof_invweight0", "dof_damping",
              "geom_size", "geom_rbound", "geom_friction", "geom_solref", "geom_solimp", "jnt_range", "qpos0"]
    np.savez_compressed(out / "mjmodel_fields.npz", **{f: np.array(getattr(model, f)) for f in fields},
                        opt=np.array([model.opt.timestep, *model.opt.gravity, model.opt.impratio, model.opt.tolerance]))
    if args.states:
        s = dict(np.load(args.states))
        n = len(s["qpos"])
        q1 = np.zeros_like(s["qpos"]); v1 = np.zeros_like(s["qvel"])
        for e in range(n):
            mujoco.mj_resetData(model, data)
            data.qpos[:] = s["qpos"][e]; data.qvel[:] = s["qvel"][e]; data.ctrl[:] = s["ctrl"][e]
            data.qacc_warmstart[:] = s["warm"][e]
            mujoco.mj_step(model, data)
            q1[e], v1[e] = data.qpos, data.qvel
        np.savez_compressed(out / "mujoco_onestep.npz", qpos=s["qpos"], qvel=s["qvel"], warm=s["warm"], ctrl=s["ctrl"], qpos1=q1, qvel1=v1)


if __name__ == "__main__":
    main()
