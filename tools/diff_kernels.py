"""Differential check on the GPU: one substep of the fast-path kernel vs the general kernel (same fp32 algorithm,
different code) from states reached along bench-style roll-outs, and the outliers against the fp64 oracle port."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hsr_env_b200.env import BatchedHSREnv
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = "cuda:0"
goals = [GoalSpec(a=Box([-.25, -.2, 0, -1], [-.05, .1, 1, 1]), b=Box([-.15, -.2, .017], [0, .1, .017]), distance=.05)]
fast = BatchedHSREnv("c2_push.hsrb", goals, n_envs=n, device=dev, seed=3, kernel="fast")
gen = BatchedHSREnv("c2_push.hsrb", None, n_envs=n, device=dev, seed=3, kernel="general")
fast.reset(); gen.reset()
g = torch.Generator(device=dev).manual_seed(1)
bad_total = tot = 0
worst = []
for rnd in range(6):
    act = torch.rand(n, fast.nu, generator=g, device=dev) * 2 - 1
    fast.step(act, steps=int(torch.randint(5, 120, (1,)).item()))
    qpos, qvel, warm, mocap = fast.get_state()
    gen.set_state(qpos=qpos, qvel=qvel, qacc_warmstart=warm, mocap_pos=mocap)
    of, *_ = fast.step(act, steps=1)
    og, *_ = gen.step(act, steps=1)
    a, b = of.double().cpu().numpy(), og.double().cpu().numpy()
    err = np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))
    out = np.nonzero(err > 1e-4)[0]
    bad_total += len(out); tot += n
    if len(out):
        cp = port.CpuPort(fast.model)
        q, v, w, m_ = [t.double().cpu().numpy() for t in (qpos, qvel, warm, mocap)]
        ref = cp.step(q[out], v[out], w[out], act.double().cpu().numpy()[out], m_[out], nsub=1)
        want = np.concatenate([ref["qpos"], ref["qvel"]], axis=1)
        ef = np.abs(a[out] - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
        eg = np.abs(b[out] - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
        worst.append((float(ef.max()), float(eg.max()), int((ef > 1e-4).sum()), int((eg > 1e-4).sum())))
    print(f"round {rnd}: fast vs general > 1e-4: {len(out)} / {n}; max {err.max():.2e}; median {np.median(err):.2e}", flush=True)
print(f"TOTAL outliers {bad_total} / {tot} = {bad_total / tot:.4%}; vs fp64 oracle on the outliers (fast max, general max, #fast bad, #general bad): {worst}")
