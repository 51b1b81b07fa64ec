"""Benchmark of the hot path: env actions/sec (300 physics substeps each) of the slide_x/slide_y block-push
configuration (BASELINE.json configs[1]: 4096 batched environments per B200, randomized block-space /
goal-space resets), one process per GPU, weak scaling (4096 environments per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu 4096] [--impl reference]

A "step" is one env action for every environment of the batch: a masked reset of the environments that
finished on the previous action (the caller's ``if done: env.reset()``, /root/reference/hsr/control.py:73-75)
followed by ``HSREnv.step`` (/root/reference/hsr/env.py:115-135) = up to 300 substeps with the per-substep
goal test and early break.  Executed substeps are counted (``substeps_taken``), never assumed to be 300.

Prints ONE JSON line on rank 0.  ``--impl reference`` times the CPU side instead: the reference's own
implementation (mujoco-py) cannot run here, so it is the oracle's C++ port on all host cores,
labelled ``kind: "port"``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BLOCK_LO, BLOCK_HI = [-.25, -.2, 0., -1.], [-.05, .1, 1., 1.]   # (x, y, qw, qz)   SURVEY.md §8(d) C2
GOAL_LO, GOAL_HI = [-.15, -.2, .017], [0., .1, .017]
GEOFENCE = .05
NSUB = 300
BLOB = "c2_push.hsrb"
WORKLOAD = ("c2_push: --use-dof slide_x slide_y --n-blocks 1 --steps-per-action=300 --geofence=.05, "
            "block-space (-.25,-.05)(-.2,.1)(0,1)(-1,1), goal-space (-.15,0)(-.2,.1)(.017,.017), "
            "actions ~ U(ctrlrange), done envs reset every action")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs (NVML)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_port_run(n_envs: int, n_actions: int, threads: int, seed: int = 0, budget_s: float = 1e9):
    """The same workload on the oracle's fp64 C++ port (one environment per thread at a time): returns
    (env_actions, substeps, seconds, algorithmic flops).  Test infrastructure used as a reported baseline only."""
    from hsr_env_b200.model import Model
    from oracle import port

    model = Model.load(ROOT / "hsr_env_b200" / "blobs" / BLOB)
    cp = port.CpuPort(model)
    cp.set_goals(np.r_[GOAL_LO, GOAL_HI], np.r_[BLOCK_LO, BLOCK_HI], GEOFENCE)
    rng = np.random.default_rng(seed)
    qpos = np.zeros((n_envs, model.nq)); mocap = np.zeros((n_envs, 3))
    episode = np.zeros(n_envs, np.int64)
    for e in range(n_envs):
        qpos[e], mocap[e] = cp.reset(seed, e, 0)
    qvel = np.zeros((n_envs, model.nv)); warm = np.zeros((n_envs, model.nv))
    lo, hi = model.act_ctrlrange[:, 0], model.act_ctrlrange[:, 1]
    actions = substeps = flops = 0
    t_total = 0.0
    for a in range(n_actions):
        ctrl = rng.uniform(lo, hi, size=(n_envs, model.nu))
        t0 = time.perf_counter()
        out = cp.step(qpos, qvel, warm, ctrl, mocap, nsub=NSUB, nthreads=threads)
        t_total += time.perf_counter() - t0
        qpos, qvel, warm = out["qpos"], out["qvel"], out["warm"]
        actions += n_envs
        substeps += int(out["taken"].sum())
        flops += int(out["counters"][:, 3].sum())
        for e in np.nonzero(out["success"])[0]:
            episode[e] += 1
            qpos[e], mocap[e] = cp.reset(seed, int(e), int(episode[e]))
            qvel[e] = 0; warm[e] = 0
        if t_total > budget_s:
            break
    return actions, substeps, t_total, flops


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = max(cores * 32, 256)
    # warm-up + timed "steps": each step = one action for a bounded sample of n_envs environments (of the 4096)
    cpu_port_run(n_envs, max(1, min(args.warmup, 2)), cores)
    acts, subs, secs, flops = cpu_port_run(n_envs, args.steps, cores, budget_s=120.0)
    value = acts / secs
    line = {
        "impl": "reference", "metric": "env_actions_per_sec", "value": value, "unit": "env-actions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, acts / n_envs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_step": n_envs},
        "substeps_per_s": subs / secs, "mean_substeps_per_action": subs / acts,
        "cpu_baseline": {"value": value, "unit": "env-actions/s", "cores": cores, "kind": "port",
                         "sample": f"{acts} env-actions ({subs} substeps) of the workload on {cores} threads; the "
                                   "reference's mujoco-py cannot be installed here, this is the oracle's fp64 C++ port, NOT MuJoCo"},
        "e2e": {"value": value, "unit": "env-actions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU side
def run_gpu(args):
    import torch

    from hsr_env_b200 import dist as D
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    rank, local, world = D.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n = args.envs_per_gpu
    goals = [GoalSpec(a=Box(BLOCK_LO, BLOCK_HI), b=Box(GOAL_LO, GOAL_HI), distance=GEOFENCE)]
    env = BatchedHSREnv(BLOB, goals, steps_per_action=NSUB, n_envs=n, device=dev, seed=args.seed,
                        env_id_offset=rank * n, lanes_per_env=args.lanes, kernel=args.kernel)
    info = env.launch_info()
    lo = torch.tensor(env.model.act_ctrlrange[:, 0], dtype=torch.float32, device=dev)
    hi = torch.tensor(env.model.act_ctrlrange[:, 1], dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(args.seed + 1000 * rank)
    total = args.warmup + args.steps
    actions = [lo + (hi - lo) * torch.rand(n, env.nu, generator=gen, device=dev) for _ in range(total)]
    if args.action_scale != 1.0:   # experiments only (e.g. 0 = the robot stands still: floor contacts only)
        actions = [a * args.action_scale for a in actions]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    env.reset()
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    taken_sum = torch.zeros((), dtype=torch.int64, device=dev)
    succ_sum = torch.zeros((), dtype=torch.int64, device=dev)

    kstart = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kend = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def one_step(k, timed=None):
        nonlocal done
        env.reset(mask=done)
        if timed is not None:
            kstart[timed].record()   # the action kernel alone (torch's current stream = the launching stream)
        obs, reward, done, inf = env.step(actions[k])
        if timed is not None:
            kend[timed].record()
        return inf["substeps_taken"]

    for k in range(args.warmup):
        one_step(k)
    torch.cuda.synchronize(dev)
    st0 = env.stats()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    D.barrier()
    torch.cuda.synchronize(dev)
    # the workload drifts along the episodes (blocks get pushed away from the base: 36 -> 23 ms per action over 60
    # actions, tools/step_time_trend.py), so the end-to-end leg below restarts from this snapshot: both legs time the
    # same actions on the same states
    snap = [t.clone() for t in env.get_state()] + [done.clone()]
    with ClockSampler(local) as clk:
        for k in range(args.steps):
            flush.fill_(k & 0xff)           # evict L2 between timed iterations (outside the event bracket)
            starts[k].record()
            taken = one_step(args.warmup + k, timed=k)
            ends[k].record()
            taken_sum += taken.sum()
            succ_sum += done.sum()
        torch.cuda.synchronize(dev)
    D.barrier()
    secs = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) * 1e-3
    kernel_s = sum(s.elapsed_time(e) for s, e in zip(kstart, kend)) * 1e-3 / args.steps   # mean action-kernel launch
    st1 = env.stats()
    secs_max = D.max_over_ranks(secs, dev)
    sub_local = float(taken_sum.item())
    sub_total, succ_total, bad_total = D.sum_over_ranks([sub_local, float(succ_sum.item()), float(st1["bad_envs"] - st0["bad_envs"])], dev)
    flops_local = st1["flops"] - st0["flops"]
    flops_total = D.sum_over_ranks([float(flops_local)], dev)[0]
    launches = st1["launches"] - st0["launches"]
    gathered = D.gather_episode_stats(dict(episodes=float(succ_sum.item()), successes=float(succ_sum.item()),
                                           substeps=sub_local, bad_states=float(st1["bad_envs"] - st0["bad_envs"])), dev)

    # ---- end to end through the host-buffer API: pinned host actions in, obs/reward/done/substeps out, every step
    e2e_steps = args.steps
    act_host = [a.cpu().pin_memory() for a in actions[args.warmup:args.warmup + e2e_steps]]
    out = dict(obs=torch.empty(n, env.obs_dim).pin_memory(), reward=torch.empty(n).pin_memory(),
               done=torch.zeros(n, dtype=torch.uint8).pin_memory(), taken=torch.empty(n, dtype=torch.int32).pin_memory())
    mask_dev = torch.zeros(n, dtype=torch.uint8, device=dev)
    for k in range(min(3, e2e_steps)):  # warm-up of the host path
        env.step_host(act_host[k], out=out)
    env.set_state(qpos=snap[0], qvel=snap[1], qacc_warmstart=snap[2], mocap_pos=snap[3])
    out["done"].copy_(snap[4].to(torch.uint8))
    D.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        mask_dev.copy_(out["done"], non_blocking=True)            # H2D: which environments the caller resets
        env.reset(mask=mask_dev)
        env.step_host(act_host[k], out=out)                       # H2D ctrl, kernel, D2H obs/reward/done/taken, sync
    torch.cuda.synchronize(dev)
    e2e_secs = D.max_over_ranks(time.perf_counter() - t0, dev)
    h2d = n * env.nu * 4 + n
    d2h = n * (env.obs_dim * 4 + 4 + 1 + 4)

    hbm_peak, sm_max, which = peaks()
    nq, nv, nu = env.nq, env.nv, env.nu
    # algorithmic HBM bytes per env-action: state in/out once per action
    b_alg = 4 * ((nq + 2 * nv + nu + 3 + 2) + (nq + 2 * nv + (nq + nv) + 4))
    achieved_gbs = b_alg * n / kernel_s / 1e9
    traffic = None
    tp = ROOT / "profiles" / "r01_traffic.json"
    if tp.exists() and n == 4096 and info["kernel"] == "fast":
        traffic = json.loads(tp.read_text())["dram_bytes_per_launch"]
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    clocks = clk.summary()
    fp32_peak_at_clock = fp32_peak * (clocks["sm_mhz"] / sm_max) if clocks.get("sm_mhz") else None
    achieved_tf = flops_total / secs_max / 1e12 / world
    value = world * n * args.steps / secs_max
    line = {
        "metric": "env_actions_per_sec", "value": value, "unit": "env-actions/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n, "total_envs": n * world, "substeps_per_action": NSUB,
                   "l2": "flushed (256 MiB write) between timed steps", "kernel": info["kernel"], "threads_per_block": info["threads_per_block"], "lanes_per_env": info["lanes_per_env"],
                   "smem_per_env": info["smem_per_env"], "resident_envs_per_sm": info["envs_per_sm"], "grid": info["grid"]},
        "substeps_per_s": sub_total / secs_max, "mean_substeps_per_action": sub_total / (world * n * args.steps),
        "success_per_action": succ_total / (world * n * args.steps), "bad_states": bad_total,
        "clocks": clocks,
        "e2e": {"value": world * n * e2e_steps / e2e_secs, "unit": "env-actions/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "substeps_per_s": None},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": which + " (burst copy figure)",
                     "kernel": "hsrb_push_kernel" if info["kernel"] == "fast" else "hsrb_step_kernel",
                     "kernel_ms_per_launch": 1e3 * kernel_s, "algorithmic_bytes_per_launch": b_alg * n,
                     "algorithmic_bytes_per_env_action": b_alg,
                     "note": "state crosses HBM once per action (312 B per env-action): this path is bound by the instruction "
                             "stream of a warp (issue / fetch latency), not by HBM or the FP32 pipe; see fp32 and DESIGN.md 4.1"},
        "fp32": {"achieved_tflops": achieved_tf, "peak_tflops_at_max_clock": fp32_peak,
                 "peak_tflops_at_observed_clock": fp32_peak_at_clock,
                 "frac_of_max_clock_peak": achieved_tf / fp32_peak,
                 "mean_algorithmic_flops_per_substep": flops_total / max(1.0, sub_total),
                 "note": "algorithmic flops = SURVEY.md 8(d) stage formulas with the kernel's actual per-substep counts"},
        "episode_stats_per_rank": gathered,
    }
    ph = {k: st1["phase_cycles"][k] - st0["phase_cycles"][k] for k in st1["phase_cycles"]}
    if sum(ph.values()) > 0:
        tot = float(sum(ph.values()))
        line["phase_share"] = {k: round(v / tot, 4) for k, v in ph.items()}
        line["phase_cycles_per_substep_lane0"] = tot / max(1.0, sub_local)
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        acts, subs, csecs, _ = cpu_port_run(max(cores * 4, 32), 1000, cores, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": acts / csecs, "unit": "env-actions/s", "cores": cores, "kind": "port",
            "substeps_per_s": subs / csecs,
            "sample": f"{acts} env-actions ({subs} substeps, {csecs:.1f} s) of the same workload on {cores} host threads, "
                      "oracle fp64 C++ port (NOT mujoco-py: it cannot be installed here)"}
    env.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--lanes", type=int, default=0, help="lanes of a warp per environment (0 = auto)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "general", "fast", "wpe"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--action-scale", type=float, default=1.0, help="experiments: scale the sampled actions")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
