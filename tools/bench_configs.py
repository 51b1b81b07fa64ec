"""Throughput of the other BASELINE.json configurations (parity-test cases, not bench.py lines) on one GPU:

    C3  full arm + gripper dofs, 1 block on the pan, 16384 environments   (general kernel)
    C5  slide_x/slide_y, 4 blocks, 4096 environments                      (general kernel)
    C4  the per-GPU share of the 1 M-environment sweep: 131072 environments of the C2 model (fast kernel)

States come from tests/scenarios.initial_states (blocks in front of the base / on the pan), actions ~ U(ctrlrange),
300 substeps per action, no goal (no early exit).  One JSON line per configuration.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from hsr_env_b200.env import BatchedHSREnv  # noqa: E402
from scenarios import initial_states  # noqa: E402


def run(name, blob, n, pan, steps=3, warmup=2, nsub=300, lanes=0):
    dev = torch.device("cuda", 0)
    env = BatchedHSREnv(blob, None, steps_per_action=nsub, n_envs=n, device=dev, lanes_per_env=lanes)
    base = initial_states(env.model, min(n, 512), seed=1, pan=pan)
    q = torch.tensor(np.tile(base, ((n + len(base) - 1) // len(base), 1))[:n], dtype=torch.float32, device=dev)
    env.reset()
    env.set_state(qpos=q, qvel=torch.zeros(n, env.nv, device=dev))
    lo = torch.tensor(env.model.act_ctrlrange[:, 0], dtype=torch.float32, device=dev)
    hi = torch.tensor(env.model.act_ctrlrange[:, 1], dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    acts = [lo + (hi - lo) * torch.rand(n, env.nu, generator=gen, device=dev) for _ in range(steps + warmup)]
    for k in range(warmup):
        env.step(acts[k])
    torch.cuda.synchronize()
    st0 = env.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    taken = 0
    for k in range(steps):
        _, _, _, inf = env.step(acts[warmup + k])
        taken = taken + inf["substeps_taken"].sum()
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    st1 = env.stats()
    info = env.launch_info()
    sub = float(taken.item())
    line = {"config": name, "model": blob, "n_envs": n, "nq": env.nq, "nv": env.nv, "nu": env.nu, "kernel": info["kernel"],
            "lanes_per_env": info["lanes_per_env"], "threads_per_block": info["threads_per_block"], "grid": info["grid"],
            "env_actions_per_s": n * steps / secs, "substeps_per_s": sub / secs, "ms_per_action_batch": 1e3 * secs / steps,
            "mean_contacts_per_substep": (st1["contacts"] - st0["contacts"]) / max(1.0, sub),
            "mean_newton_iters_per_substep": (st1["newton_iters"] - st0["newton_iters"]) / max(1.0, sub),
            "bad_envs": st1["bad_envs"] - st0["bad_envs"]}
    ph = {k: st1["phase_cycles"][k] - st0["phase_cycles"][k] for k in st1["phase_cycles"]}
    if sum(ph.values()) > 0:
        line["phase_share"] = {k: round(v / sum(ph.values()), 3) for k, v in ph.items()}
    env.close()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c5", "c4"]
    if "c3" in which:
        run("C3 arm+gripper pick-and-place, 16384 envs", "c3_arm.hsrb", 16384, pan=True)
    if "c5" in which:
        run("C5 4-block clutter, 4096 envs", "c5_clutter.hsrb", 4096, pan=False)
    if "c4" in which:
        run("C4 per-GPU share of the 1M-env sweep, 131072 envs", "c2_push.hsrb", 131072, pan=False, lanes=16)
        run("C4 per-GPU share of the 1M-env sweep, 131072 envs", "c2_push.hsrb", 131072, pan=False, lanes=8)
