"""Distribution over environments of the cycles one environment's 300-substep action takes on an uncontended warp.

    python -m hsr_env_b200.build --define=WPE_CHAIN_CLOCKS --tag=ck
    HSRB_LIB=hsr_env_b200/csrc/libhsrb_ck.so HSRB_WPE_LOCK=0 HSRB_WPE_WPB=1 HSRB_WPE_SORT=1 python tools/env_work_dist.py

HSRB_WPE_WPB=1: one warp per block, 148 blocks, grid-stride over the 4096 environments -> every warp runs alone on its SM."""
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
    D.init_from_env()
    goals = [GoalSpec(a=Box(B.BLOCK_LO, B.BLOCK_HI), b=Box(B.GOAL_LO, B.GOAL_HI), distance=B.GEOFENCE)]
    r = B.run_workload(torch, D, dev, blob=B.BLOB, goals=goals, starts=None, n_local=n, env_offset=0, steps=3, warmup=5, seed=0, kernel="wpe")
    env = r["env"]
    lib = L.load()
    w = np.zeros(n, np.int32)
    lib.hsrb_debug_work(env._h, w.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    kc = w.astype(np.float64)   # k-cycles (1024 cycles) of the last action per environment
    q = np.percentile(kc, [0, 10, 25, 50, 75, 90, 95, 99, 99.9, 100])
    print("k-cycles per 300-substep action, percentiles 0/10/25/50/75/90/95/99/99.9/100:", np.round(q).astype(int).tolist())
    print(f"mean {kc.mean():.0f} k-cycles = {kc.mean() * 1024 / 1.965e6:.2f} ms; max {kc.max() * 1024 / 1.965e6:.2f} ms; "
          f"sum / (148 SMs x 28 warps) = {kc.sum() * 1024 / 1.965e6 / (148 * 28):.2f} ms")
    # persistence: how well does the work of one action predict the next one's?
    a = torch.from_numpy(B.host_actions(5, 2, n, env.model.act_ctrlrange[:, 0], env.model.act_ctrlrange[:, 1])).to(dev)
    ws = []
    for k in range(2):
        env.step(a[k])
        lib.hsrb_debug_work(env._h, w.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        ws.append(w.astype(np.float64).copy())
    w0, w1 = ws
    r0 = np.argsort(np.argsort(-w0)); r1 = np.argsort(np.argsort(-w1))   # ranks, 0 = heaviest
    for top in (0.01, 0.05, 0.10):
        heavy1 = r1 < top * n
        print(f"of the heaviest {top:.0%} of action k+1: in the top 5% / 10% / 25% / 50% of action k:",
              [round(float((r0[heavy1] < f * n).mean()), 3) for f in (.05, .10, .25, .5)], " (no reset between the two actions)")
    print("rank correlation:", round(float(np.corrcoef(r0, r1)[0, 1]), 3))
    srt = np.sort(kc)[::-1]
    print("heaviest 16:", srt[:16].astype(int).tolist())
    print("share of total work in the heaviest 5% / 10% / 25%:", [round(float(srt[: int(n * f)].sum() / srt.sum()), 3) for f in (.05, .1, .25)])


if __name__ == "__main__":
    main()
