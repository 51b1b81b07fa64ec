"""Goal lists (HSREnv.step's all(in_range(*g)), /root/reference/hsr/env.py:126,137-147) and per-joint start spaces
(HSREnv.new_state, /root/reference/hsr/env.py:149-156): host logic on the oracle port (CPU) and the same configuration
through the C ABI on the GPU, bit-exact against the port."""
import numpy as np
import pytest

from scenarios import rollout_states


def _starts_tables(model, starts):
    names = list(model.names["joint"])
    adr, width, lo, hi = [], [], [], []
    for joint, (l, h) in starts.items():
        j = names.index(joint)
        w = 7 if int(model.jnt_type[j]) == 0 else 1
        L, H = np.zeros(7), np.zeros(7)
        L[:w], H[:w] = l, h
        adr.append(int(model.jnt_qposadr[j])); width.append(w); lo.append(L); hi.append(H)
    return adr, width, lo, hi


CUPBOARD_STARTS = {"blockjoint": ([-.1, -.2, .418, 0, 0, -1, 0], [.1, .2, .418, 1, 0, 1, 0]),      # hsr/__init__.py:13-18
                   "slide_x": ([-.05], [.02])}


def test_port_start_spaces_are_philox_streams(models, ports):
    """Start draws: inside [lo, hi), a function of (seed, global env id, episode) only, independent of the other draws."""
    model, port = models["f2_cupboard"], ports["f2_cupboard"]
    adr, width, lo, hi = _starts_tables(model, CUPBOARD_STARTS)
    port.set_starts(adr, width, lo, hi)
    try:
        seen = []
        for env_id in (0, 1, 77):
            for ep in (0, 1):
                q, _ = port.reset(5, env_id, ep)
                q2, _ = port.reset(5, env_id, ep)
                assert np.array_equal(q, q2)
                for a, w, L, H in zip(adr, width, lo, hi):
                    v = q[a:a + w]
                    assert np.all(v >= np.float32(L[:w]) - 1e-7) and np.all(v <= np.float32(H[:w]) + 1e-7)
                    assert np.all(v[np.asarray(L[:w]) == np.asarray(H[:w])] == np.float32(np.asarray(L[:w])[np.asarray(L[:w]) == np.asarray(H[:w])]))
                seen.append(q[adr[0]:adr[0] + 2].copy())
        assert len({tuple(s) for s in seen}) == len(seen)          # distinct streams per (env, episode)
        # joints without a start space keep qpos0
        q, _ = port.reset(5, 3, 0)
        touched = set(i for a, w in zip(adr, width) for i in range(a, a + w))
        rest = [i for i in range(model.nq) if i not in touched]
        assert np.array_equal(q[rest], np.float32(model.qpos0[rest]))
    finally:
        port.set_starts([], [], np.zeros((0, 7)), np.zeros((0, 7)))


def test_port_goal_list_is_and_over_goals(models, ports):
    """all(in_range(*g)) over body / point / fixed-point endpoints, strict <, against plain numpy on the port's poses."""
    model, port = models["c5_clutter"], ports["c5_clutter"]
    n = 48
    qpos, qvel, warm, ctrl = rollout_states(port, model, n, seed=2, float32=True)
    blocks = [int(b) for b in model.block_body]
    ref = port.step(qpos, qvel, warm, ctrl, nsub=1, debug=True)
    xpos = np.stack([r["xpos"] for r in ref["debug"]])          # poses of the forward pass the goal test uses
    rng = np.random.default_rng(0)
    # goal 0: block0 within d0 of the per-env point; goal 1: block1 within d1 of block2; goal 2: robot within d2 of a fixed point
    point = xpos[:, blocks[0]] + rng.normal(0, .03, (n, 3))
    fixed = np.array([[0.0, 0.0, 0.0]])
    d = np.array([.05, .25, .6])
    want = ((np.linalg.norm(xpos[:, blocks[0]] - point, axis=1) < d[0])
            & (np.linalg.norm(xpos[:, blocks[1]] - xpos[:, blocks[2]], axis=1) < d[1])
            & (np.linalg.norm(xpos[:, 1] - fixed[0], axis=1) < d[2]))
    assert 0 < want.sum() < n
    port.set_goal_list([blocks[0], blocks[1], 1], [-1, blocks[2], -2], d, None, fixed)
    try:
        out = port.step(qpos, qvel, warm, ctrl, point, nsub=1)
    finally:
        port.set_goals(None)
    assert np.array_equal(out["success"].astype(bool), want)


@pytest.mark.gpu
def test_gpu_start_spaces_bit_exact_and_shard_invariant(models, ports):
    torch = pytest.importorskip("torch")
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    model, port = models["f2_cupboard"], ports["f2_cupboard"]
    starts = {k: Box(np.array(l), np.array(h)) for k, (l, h) in CUPBOARD_STARTS.items()}
    adr, width, lo, hi = _starts_tables(model, CUPBOARD_STARTS)
    port.set_starts(adr, width, lo, hi)
    goals = [GoalSpec(a="block", b=np.array([0, 0, .498]), distance=.05)]
    try:
        env = BatchedHSREnv("f2_cupboard.hsrb", goals, starts=starts, n_envs=96, device="cuda:0", seed=9, env_id_offset=32)
        shard = BatchedHSREnv("f2_cupboard.hsrb", goals, starts=starts, n_envs=32, device="cuda:0", seed=9, env_id_offset=64)
        for episode in range(2):
            obs = env.reset(); obs_s = shard.reset()
            assert torch.equal(obs[32:64], obs_s)                    # rank-count invariant
            q = obs[:, :env.nq].cpu().numpy()
            for e in (0, 5, 95):
                want, _ = port.reset(9, 32 + e, episode)
                a = adr[0]
                assert np.array_equal(q[e, :a], np.float32(want[:a])) or a == 0
                assert np.array_equal(q[e, a:a + 3], np.float32(want[a:a + 3]))
                qn = want[a + 3:a + 7] / np.linalg.norm(want[a + 3:a + 7])   # sim.forward() normalises the quaternion
                np.testing.assert_allclose(q[e, a + 3:a + 7], qn, atol=1e-6)
                assert q[e, adr[1]] == np.float32(want[adr[1]])
        # a masked reset redraws only the masked environments, from their next episode
        mask = torch.zeros(96, dtype=torch.bool); mask[7] = True
        ob2 = env.reset(mask=mask)
        assert torch.equal(ob2[~mask.cuda()], obs[~mask.cuda()])
        want, _ = port.reset(9, 32 + 7, 2)
        assert np.array_equal(ob2[7, adr[0]:adr[0] + 3].cpu().numpy(), np.float32(want[adr[0]:adr[0] + 3]))
        env.close(); shard.close()
    finally:
        port.set_starts([], [], np.zeros((0, 7)), np.zeros((0, 7)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c5_clutter", "c2_push"])
def test_gpu_goal_list_flags_match_port(name, models, ports):
    """Several GoalSpecs with body-name / Space / ndarray endpoints: done, reward and substeps_taken equal the fp32 port's
    on matched states (c2_push: the sliding-base family switches to the general kernel for this form)."""
    torch = pytest.importorskip("torch")
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    model, port = models[name], ports[name]
    n = 128
    qpos, qvel, warm, ctrl = rollout_states(port, model, n, seed=4, float32=True)
    blocks = [int(b) for b in model.block_body]
    bnames = list(model.names["body"])
    rng = np.random.default_rng(1)
    a0 = 2 if name == "c2_push" else int(model.jnt_qposadr[model.body_jntadr[blocks[0]]])
    r, ang = rng.uniform(.04, .06, n), rng.uniform(0, 2 * np.pi, n)
    point = (qpos[:, a0:a0 + 3] + np.stack([r * np.cos(ang), r * np.sin(ang), np.zeros(n)], 1)).astype(np.float32).astype(np.float64)
    robot = bnames[1]
    goals = [GoalSpec(bnames[blocks[0]], Box([-1, -1, 0], [1, 1, 1]), .05),       # block0 near the sampled point
             GoalSpec(robot, np.array([0., 0., 0.]), 5.0)]                        # robot within 5 m of the origin (always)
    codes_a, codes_b, dist = [blocks[0], 1], [-1, -2], [.05, 5.0]
    if len(blocks) > 1:
        goals.append(GoalSpec(bnames[blocks[1]], bnames[blocks[2]], .3))
        codes_a.append(blocks[1]); codes_b.append(blocks[2]); dist.append(.3)
    env = BatchedHSREnv(f"{name}.hsrb", goals, n_envs=n, device="cuda:0")
    env.reset()
    assert env.launch_info()["kernel"] == "general"
    env.set_state(qpos, qvel, warm, point)
    port.set_goal_list(codes_a, codes_b, dist, np.r_[-1, -1, 0, 1, 1, 1.], np.zeros((1, 3)))
    try:
        ref = port.step(qpos, qvel, warm, ctrl, point, nsub=25, use_float=True)
    finally:
        port.set_goals(None)
    obs, reward, done, info = env.step(torch.tensor(ctrl, dtype=torch.float32), steps=25)
    got = done.cpu().numpy().astype(np.uint8)
    assert 0.05 < got.mean() < 0.95
    # blocks within 1e-6 of a geofence may round differently between the CUDA build and g++
    agree = got == ref["success"]
    assert agree.mean() >= 0.98, agree.mean()
    taken = info["substeps_taken"].cpu().numpy()
    assert np.mean(taken[agree] == ref["taken"][agree]) >= 0.98
    assert np.array_equal(reward.cpu().numpy(), got.astype(np.float32))
    env.close()


@pytest.mark.gpu
def test_step_host_is_ordered_after_reset_without_sync():
    """hsrb_step_host runs on the caller's stream: reset -> step_host with no synchronisation in between steps the
    freshly reset state (it used to run on a private non-blocking stream)."""
    torch = pytest.importorskip("torch")
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    n = 2048
    goals = [GoalSpec(Box([-.25, -.2, 0, -1], [-.05, .1, 1, 1]), Box([-.15, -.2, .017], [0, .1, .017]), .05)]
    act = np.random.default_rng(0).uniform(-1, 1, (n, 2))          # float64 on purpose: converted, not reinterpreted
    a = BatchedHSREnv("c2_push.hsrb", goals, n_envs=n, device="cuda:0", seed=1)
    b = BatchedHSREnv("c2_push.hsrb", goals, n_envs=n, device="cuda:0", seed=1)
    for rep in range(3):
        a.reset()
        out = a.step_host(act, steps=10)                            # no sync between reset and the host-buffer step
        b.reset(); torch.cuda.synchronize()
        obs, reward, done, info = b.step(torch.tensor(act, dtype=torch.float32), steps=10)
        assert np.array_equal(out["obs"], obs.cpu().numpy())
        assert np.array_equal(out["taken"], info["substeps_taken"].cpu().numpy())
    with pytest.raises(ValueError):
        a.step_host(act, steps=1, out=dict(obs=np.empty((n, 3), np.float32), reward=np.empty(n, np.float32),
                                           done=np.empty(n, np.uint8), taken=np.empty(n, np.int32)))
    a.close(); b.close()
