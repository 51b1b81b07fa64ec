"""Multi-GPU plumbing: one process per GPU, one contiguous slice of the global environment ids per rank.

Environments are independent (one MjSim per env in the reference, /root/reference/hsr/mujoco_env.py:34), so the
hot path has no collective.  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is
used only for the barrier / max-over-ranks timing and the optional all-gather of episode statistics.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of global env ids owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the process group if world>1."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device="cpu"):
    t = torch.as_tensor(values, dtype=torch.float64, device=device).clone()
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


EPISODE_STAT_KEYS = ("episodes", "successes", "substeps", "bad_states")


def gather_episode_stats(local: Dict[str, float], device="cpu") -> Dict[str, list]:
    """All-gather of the per-rank episode statistics: {key: [value of rank 0, rank 1, ...]}."""
    t = torch.tensor([float(local.get(k, 0.0)) for k in EPISODE_STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_initialized():
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
    else:
        out = [t]
    rows = torch.stack(out).cpu()
    return {k: rows[:, i].tolist() for i, k in enumerate(EPISODE_STAT_KEYS)}
