"""Multi-rank host logic on CPU: world_size 2 over gloo (the GPU box uses NCCL for the same calls).  The data path has
no collective; what is tested is the sharding of global env ids, the max-over-ranks timing reduction and the all-gather
of episode statistics (hsr_env_b200/dist.py)."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from hsr_env_b200 import dist as D

    r, l, w = D.init_from_env(backend="gloo")
    lo, hi = D.shard_range(1001, r, w)
    D.barrier()
    mx = D.max_over_ranks(1.0 + r)
    tot = D.sum_over_ranks([hi - lo, 1.0])
    g = D.gather_episode_stats(dict(episodes=10 + r, successes=r, substeps=300.0 * (r + 1), bad_states=0))
    q.put((r, lo, hi, mx, tot, g))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, mx0, tot0, g0), (r1, lo1, hi1, mx1, tot1, g1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)          # contiguous, sizes differ by at most one
    assert mx0 == mx1 == 2.0                                     # max over ranks
    assert tot0 == tot1 == [1001.0, 2.0]
    assert g0 == g1 and g0["episodes"] == [10.0, 11.0] and g0["substeps"] == [300.0, 600.0]


def test_shard_range_covers_everything():
    from hsr_env_b200.dist import shard_range

    for n in (1, 7, 4096, 1 << 20):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
