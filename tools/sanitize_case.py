"""One small invocation of each kernel for compute-sanitizer (tools/gpu_sanitize.sh): few environments, few substeps,
states taken along oracle roll-outs so that contacts (plane-box, hull-box, box-box) and the Newton solver are exercised."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from hsr_env_b200.env import BatchedHSREnv  # noqa: E402
from hsr_env_b200.model import Model  # noqa: E402
from hsr_env_b200.spaces import Box  # noqa: E402
from hsr_env_b200.util import GoalSpec  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "fast"
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if case == "fast":
    goals = [GoalSpec(a=Box([-.25, -.2, 0, -1], [-.05, .1, 1, 1]), b=Box([-.15, -.2, .017], [0, .1, .017]), distance=.05)]
    env = BatchedHSREnv("c2_push.hsrb", goals, n_envs=64, device="cuda:0", seed=0, kernel="fast")
    env.reset()
    act = torch.ones(64, 2)
    act[:, 1] = torch.linspace(-1, 1, 64)
    for k in range(3):
        obs, r, d, info = env.step(act, steps=nsub)
        env.reset(mask=d)
else:
    name = {"c3": "c3_arm", "c5": "c5_clutter", "general": "c2_push"}[case]
    from oracle import port
    from scenarios import rollout_states
    model = Model.load(ROOT / "hsr_env_b200" / "blobs" / f"{name}.hsrb")
    cp = port.CpuPort(model)
    qpos, qvel, warm, ctrl = rollout_states(cp, model, 16, seed=3, pan=(case == "c3"), float32=True)
    env = BatchedHSREnv(f"{name}.hsrb", None, n_envs=16, device="cuda:0", kernel="general")
    env.set_state(qpos, qvel, warm)
    obs, r, d, info = env.step(torch.tensor(ctrl, dtype=torch.float32), steps=nsub)
torch.cuda.synchronize()
print(case, "ok", env.launch_info(), env.stats())
env.close()
