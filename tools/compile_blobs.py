"""Compile the benchmark / test configurations of SURVEY.md §8(d) from the reference's MJCF + STL assets into
model blobs under hsr_env_b200/blobs/ (committed: the GPU box has no /root/reference).

    python tools/compile_blobs.py [--assets /root/reference/hsr] [--mesh-inertia legacy|exact|convex]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hsr_env_b200 import mjcf  # noqa: E402

ALL_DOFS = ["slide_x", "slide_y", "arm_lift_joint", "arm_flex_joint", "wrist_roll_joint", "hand_l_proximal_joint",
            "hand_r_proximal_joint"]
SLIDE = ["slide_x", "slide_y"]

CONFIGS = {
    # name: (use_dof, n_blocks, block positions at mutation time (= goal_space.sample(), hsr/util.py:108))
    "c1_readme": (SLIDE, 0, []),                                   # README.md:5 verbatim (--n-blocks defaults to 0)
    "c1b_readme_block": (SLIDE, 1, [(0.0, 0.0, 0.0)]),              # + --n-blocks 1, goal-space (0,0)x3
    "c2_push": (SLIDE, 1, [(-0.15, 0.0, 0.017)]),                  # block on the floor in front of the base
    "c3_arm": (ALL_DOFS, 1, [(0.0, 0.0, 0.422)]),                  # block on the pan, 7 robot dofs
    "c5_clutter": (SLIDE, 4, [(-0.20, -0.12, 0.017), (-0.08, -0.10, 0.017), (-0.20, 0.05, 0.017), (-0.07, 0.06, 0.017)]),
    # the scene of the reference's own smoke loop (hsr/__init__.py:9-28): cupboard + its `block` body
    "f2_cupboard": (SLIDE, 0, [], "cupboard-world.xml"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assets", type=Path, default=mjcf.default_assets_root())
    ap.add_argument("--mesh-inertia", default="legacy")
    ap.add_argument("--out", type=Path, default=ROOT / "hsr_env_b200" / "blobs")
    args = ap.parse_args()
    args.out.mkdir(exist_ok=True, parents=True)
    opts = mjcf.CompileOptions(mesh_inertia=args.mesh_inertia)
    for name, cfg in CONFIGS.items():
        dofs, nb, pos = cfg[:3]
        xml = cfg[3] if len(cfg) > 3 else "world.xml"
        m = mjcf.compile_model(args.assets / "models" / xml, dofs, n_blocks=nb, block_pos=pos, opts=opts)
        m.save_with_names(args.out / f"{name}.hsrb")
        print(f"{name}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom} npair={m.npair} "
              f"nvert={m.nvert} bytes={len(m.to_blob())}")


if __name__ == "__main__":
    main()
