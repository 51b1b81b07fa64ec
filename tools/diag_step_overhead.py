"""Host-side enqueue time against GPU time of reset + step (is anything in the step path synchronising?).  Measured: 0.12 ms of
CPU per step, GPU brackets back to back - the path is fully asynchronous."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench as B
from hsr_env_b200 import dist as D
from hsr_env_b200.env import BatchedHSREnv
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec
dev = torch.device("cuda:0"); D.init_from_env()
goals = [GoalSpec(a=Box(B.BLOCK_LO, B.BLOCK_HI), b=Box(B.GOAL_LO, B.GOAL_HI), distance=B.GEOFENCE)]
n = 4096
env = BatchedHSREnv(B.BLOB, goals, steps_per_action=300, n_envs=n, device=dev, seed=0)
acts = torch.from_numpy(B.host_actions(0, 30, n, env.model.act_ctrlrange[:, 0], env.model.act_ctrlrange[:, 1])).to(dev)
env.reset()
done = torch.zeros(n, dtype=torch.bool, device=dev)
for k in range(5):
    env.reset(mask=done); obs, r, done, info = env.step(acts[k])
torch.cuda.synchronize()
t_reset = t_step = 0.0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * 20)]
t0 = time.perf_counter()
for k in range(20):
    ev[4 * k].record()
    a = time.perf_counter(); env.reset(mask=done); b = time.perf_counter()
    ev[4 * k + 1].record()
    obs, r, done, info = env.step(acts[5 + k]); c = time.perf_counter()
    ev[4 * k + 2].record()
    t_reset += b - a; t_step += c - b
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
g_reset = sum(ev[4 * k].elapsed_time(ev[4 * k + 1]) for k in range(20)) / 20
g_step = sum(ev[4 * k + 1].elapsed_time(ev[4 * k + 2]) for k in range(20)) / 20
print(f"CPU: loop {1e3 * t_cpu / 20:.3f} ms/step enqueue (reset {1e3 * t_reset / 20:.3f}, step {1e3 * t_step / 20:.3f}); wall incl. final sync {1e3 * t_all / 20:.3f} ms/step")
print(f"GPU: reset bracket {g_reset:.3f} ms, step bracket {g_step:.3f} ms")
