"""Dump the teacher-forced comparison of __graft_entry__.smoke() per environment (debugging aid)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hsr_env_b200.env import BatchedHSREnv
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec
from oracle import port
n = 64
goals = [GoalSpec(a=Box([-.25, -.2, 0, -1], [-.05, .1, 1, 1]), b=Box([-.15, -.2, .017], [0, .1, .017]), distance=.05)]
env = BatchedHSREnv("c2_push.hsrb", goals, n_envs=n, device="cuda:0", seed=0)
env.reset()
cp = port.CpuPort(env.model)
cp.set_goals(np.r_[goals[0].b.low, goals[0].b.high], np.r_[goals[0].a.low, goals[0].a.high], .05)
gen = torch.Generator().manual_seed(0)
for k in range(3):
    act = (torch.rand(n, env.nu, generator=gen) * 2 - 1)
    env.step(act.cuda(), steps=20)
qpos, qvel, warm, mocap = [t.double().cpu().numpy() for t in env.get_state()]
ref = cp.step(qpos, qvel, warm, act.double().numpy(), mocap, nsub=1, debug=True)
ref32 = cp.step(qpos, qvel, warm, act.double().numpy(), mocap, nsub=1, use_float=True)
obs, reward, done, info = env.step(act.cuda(), steps=1)
got = obs.double().cpu().numpy()
want = np.concatenate([ref["qpos"], ref["qvel"]], axis=1)
want32 = np.concatenate([ref32["qpos"], ref32["qvel"]], axis=1)
err = np.abs(got - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
err32 = np.abs(want32 - want).max(axis=1) / np.maximum(1.0, np.abs(want).max(axis=1))
o = np.argsort(err)[::-1][:5]
print("worst envs", o, err[o], "fp32 port vs fp64 port", err32[o])
for e in o[:2]:
    d = ref["debug"][e]
    print("env", e, "ncon", d["ncon"], "nefc", d["nefc"], "iters", d["iters"], "con_dist", d["con_dist"], "pairs", d["con_pair"])
    print(" got ", got[e]); print(" want", want[e]); print(" w32 ", want32[e])
np.savez(ROOT / "gpurun_out" / "smoke_debug.npz", qpos=qpos, qvel=qvel, warm=warm, act=act.numpy(), mocap=mocap, got=got, want=want)
