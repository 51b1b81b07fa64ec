"""Throughput of the block-push action kernel on a uniformly LIGHT workload (robot parked: every environment is the block
resting on the floor, 4 contacts, 1 Newton iteration) and on the bench workload, for a given batch size.

    python tools/light_env_rate.py [n_envs=4096] [actions=4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
from hsr_env_b200 import dist as D
from hsr_env_b200.env import BatchedHSREnv
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec


def timed(env, acts, dev):
    torch.cuda.synchronize(dev)
    st0 = env.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a in acts:
        env.step(a)
    e1.record()
    torch.cuda.synchronize(dev)
    st1 = env.stats()
    sub = st1["substeps"] - st0["substeps"]
    ms = e0.elapsed_time(e1)
    return sub / ms / 1e3, ms / len(acts), (st1["contacts"] - st0["contacts"]) / sub, (st1["newton_iters"] - st0["newton_iters"]) / sub


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda:0")
    D.init_from_env()
    goals = [GoalSpec(a=Box(B.BLOCK_LO, B.BLOCK_HI), b=Box(B.GOAL_LO, B.GOAL_HI), distance=B.GEOFENCE)]
    env = BatchedHSREnv(B.BLOB, goals, steps_per_action=int(os.environ.get('NSUB', B.NSUB)), n_envs=n, device=dev, seed=0, kernel="wpe")
    obs0 = env.reset()
    lo, hi = env.model.act_ctrlrange[:, 0], env.model.act_ctrlrange[:, 1]
    parked = obs0[:, :2].clone().contiguous()   # servo targets = the base position after the reset: nothing moves
    for _ in range(3):
        env.step(parked)
    r = timed(env, [parked] * k, dev)
    print(f"parked robot   n={n}: {r[0]:7.2f} M substeps/s, {r[1]:7.2f} ms/action, contacts {r[2]:.2f}, newton {r[3]:.2f}")
    acts = torch.from_numpy(B.host_actions(0, 6 + k, n, lo, hi)).to(dev)
    env.reset()
    for i in range(6):
        env.step(acts[i])
    r = timed(env, [acts[6 + i] for i in range(k)], dev)
    print(f"bench actions  n={n}: {r[0]:7.2f} M substeps/s, {r[1]:7.2f} ms/action, contacts {r[2]:.2f}, newton {r[3]:.2f} (no resets)")


if __name__ == "__main__":
    main()
