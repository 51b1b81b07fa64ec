"""Driver of tools/emu_sanitize.sh: a few substeps of every fast-family kernel layout on the SIMT emulator (library given by
EMU_LIB, built with a sanitizer), on states with floor, hull and limit contacts."""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "simt_emu"))
import build as emu  # noqa: E402
from hsr_env_b200.model import Model  # noqa: E402

emu.LIB = Path(os.environ["EMU_LIB"])
emu.build = lambda force=False: emu.LIB
g = np.load(ROOT / "tests" / "golden" / "c2_push.npz")
g1 = np.load(ROOT / "tests" / "golden" / "c1_readme.npz")
m2 = Model.load(ROOT / "hsr_env_b200" / "blobs" / "c2_push.hsrb")
m1 = Model.load(ROOT / "hsr_env_b200" / "blobs" / "c1_readme.hsrb")
sel = np.r_[np.nonzero(g["regime"] & 4)[0][:6], np.nonzero(g["regime"] & 1)[0][:4], 0, 1][:12]   # hull contacts, limits, plain
for name, model, fx, idx in (("c2_push", m2, g, sel), ("c1_readme", m1, g1, np.nonzero(g1["regime"])[0][:8])):
    for G, threads in ((1, 64), (8, 32), (16, 32), (32, 32)):
        out = emu.step(model, fx["qpos"][idx], fx["qvel"][idx], fx["warm"][idx], fx["ctrl"][idx], nsub=3, G=G, threads=threads)
        print(name, "G", G, "threads", threads, "envs", len(idx), "flags", np.unique(out["flags"]), "stats", out["stats"][:6], flush=True)
print("done")
