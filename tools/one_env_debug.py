"""Re-run one stored environment state for one substep under several kernel layouts (debugging aid)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hsr_env_b200.env import BatchedHSREnv
d = np.load(ROOT / "tools" / "smoke_env38.npz")
for kernel, lanes, n in (("fast", 8, 1), ("fast", 16, 1), ("fast", 32, 1), ("general", 0, 1), ("fast", 8, 64)):
    env = BatchedHSREnv("c2_push.hsrb", None, n_envs=n, device="cuda:0", seed=0, kernel=kernel, lanes_per_env=lanes)
    env.reset()
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda:0").reshape(1, -1).repeat(n, 1)
    env.set_state(qpos=t(d["qpos"]), qvel=t(d["qvel"]), qacc_warmstart=t(d["warm"]), mocap_pos=t(d["mocap"]))
    qp, qv, wm, mo = env.get_state()
    s0 = env.stats()
    obs, *_ = env.step(t(d["act"]), steps=1)
    s1 = env.stats()
    print(kernel, lanes, n, "quat", ["%.9g" % x for x in obs[0, 5:9].cpu().numpy()], "qvel[5:]", obs[0, 14:].cpu().numpy(), "warm ok", bool(torch.allclose(wm[0].cpu(), torch.tensor(d["warm"], dtype=torch.float32))),
          {k: (s1[k] - s0[k]) // n for k in ("newton_iters", "ls_evals", "contacts", "efc_rows", "narrowphase")})
    env.close()
