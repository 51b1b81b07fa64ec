"""Latency of the sections of one environment's dependent chain in the warp-per-environment kernel.

    python -m hsr_env_b200.build --define=WPE_CHAIN_CLOCKS --tag=ck
    HSRB_LIB=hsr_env_b200/csrc/libhsrb_ck.so HSRB_WPE_LOCK=0 python tools/chain_clocks.py [n_envs=148] [actions=10]

n = 148 -> one free-running warp per SM: the cycles of a section are its latency without contention."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
from hsr_env_b200 import dist as D
from hsr_env_b200 import lib as L
from hsr_env_b200.spaces import Box
from hsr_env_b200.util import GoalSpec

NAMES = {0: "poses", 1: "limits+cull", 2: "plane-box", 3: "mpr hit", 4: "mpr miss (cached dir)", 5: "mpr miss (full)",
         6: "smooth", 7: "rows", 8: "jar (cost phases)", 9: "cost/zones", 10: "gradient", 11: "hessian", 12: "cholesky+solves",
         13: "jv + ls coeffs", 14: "ls evaluation", 16: "update", 17: "goal+euler", 18: "box-box", 20: "support pair", 21: "portal logic", 22: "support: setup (hull)", 23: "support: hull_argmax", 24: "support: tail (hull)", 25: "support: box total"}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    acts = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    D.init_from_env()
    goals = [GoalSpec(a=Box(B.BLOCK_LO, B.BLOCK_HI), b=Box(B.GOAL_LO, B.GOAL_HI), distance=B.GEOFENCE)]
    r = B.run_workload(torch, D, dev, blob=B.BLOB, goals=goals, starts=None, n_local=n, env_offset=0, steps=1, warmup=5, seed=0,
                       kernel="wpe")
    env = r["env"]
    lib = L.load()
    out = (ctypes.c_ulonglong * 64)()
    lib.hsrb_debug_chain_clocks(out)   # reset
    st0 = env.stats()
    a = torch.from_numpy(B.host_actions(1, acts, n, env.model.act_ctrlrange[:, 0], env.model.act_ctrlrange[:, 1])).to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(acts):
        env.step(a[k])
    ev1.record()
    torch.cuda.synchronize()
    lib.hsrb_debug_chain_clocks(out)
    st1 = env.stats()
    sub = st1["substeps"] - st0["substeps"]
    print(f"{n} envs x {acts} actions: {sub} substeps, {ev0.elapsed_time(ev1):.1f} ms, contacts/substep {(st1['contacts'] - st0['contacts']) / sub:.2f}, "
          f"newton/substep {(st1['newton_iters'] - st0['newton_iters']) / sub:.2f}, ls evals/substep {(st1['ls_evals'] - st0['ls_evals']) / sub:.2f}")
    tot = sum(out[i] for i in range(20))
    print(f"{'section':26s} {'count/substep':>13s} {'cycles each':>12s} {'cycles/substep':>14s} share")
    for i in range(32):
        if out[32 + i]:
            print(f"{NAMES.get(i, str(i)):26s} {out[32 + i] / sub:13.3f} {out[i] / out[32 + i]:12.0f} {out[i] / sub:14.0f} {out[i] / tot if i < 20 else float('nan'):5.3f}")
    print(f"sum of kernel sections per substep: {tot / sub:.0f} cycles")


if __name__ == "__main__":
    main()
