"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the fp64 numpy oracle):
the oracle reproduces them exactly, the compiled fp64 port agrees to 1e-9, and the CUDA path (GPU tests) agrees to the
1e-4 relative bound of BASELINE.json.  PARITY UNPINNED: these are oracle outputs, not MuJoCo outputs."""
import numpy as np
import pytest

from conftest import GOLDEN
from scenarios import rel_err

NAMES = ["c1_readme", "c1b_readme_block", "c2_push", "c3_arm", "c5_clutter", "f2_cupboard"]


def load(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_reproduces_golden(name, models):
    from oracle import mjstep

    g, m = load(name), models[name]
    for e in range(0, len(g["qpos"]), 3):
        d = mjstep.Data(m)
        d.qpos[:] = g["qpos"][e]; d.qvel[:] = g["qvel"][e]; d.qacc_warmstart[:] = g["warm"][e]; d.ctrl[:] = g["ctrl"][e]
        mjstep.step(m, d)
        assert np.array_equal(d.qpos, g["qpos1"][e]) and np.array_equal(d.qvel, g["qvel1"][e])
        assert len(d.contacts) == g["ncon"][e] and d.nefc == g["nefc"][e]


@pytest.mark.parametrize("name", NAMES)
def test_cpp_port_matches_golden(name, models, ports):
    g, port = load(name), ports[name]
    out = port.step(g["qpos"], g["qvel"], g["warm"], g["ctrl"], nsub=1)
    assert rel_err(out["qpos"], g["qpos1"]).max() < 1e-9
    assert rel_err(out["qvel"], g["qvel1"]).max() < 1e-7
    if models[name].nblock:
        port.set_goals(np.zeros(6), None, .05)
        try:
            out = port.step(g["qpos"], g["qvel"], g["warm"], g["ctrl"], g["mocap"], nsub=1)
        finally:
            port.set_goals(None)
        assert np.array_equal(out["success"], g["success"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(name, models):
    torch = pytest.importorskip("torch")
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    g, m = load(name), models[name]
    n = len(g["qpos"])
    goals = [GoalSpec(None, Box([0, 0, 0], [0, 0, 0]), .05)] if m.nblock else None
    env = BatchedHSREnv(f"{name}.hsrb", goals, n_envs=n, device="cuda:0")
    if goals:
        env.reset()
    env.set_state(g["qpos"], g["qvel"], g["warm"], g["mocap"] if goals else None)
    obs, reward, done, info = env.step(torch.tensor(g["ctrl"], dtype=torch.float32), steps=1)
    got = obs.double().cpu().numpy()
    assert rel_err(got[:, :m.nq], g["qpos1"]).max() <= 1e-4
    ev = rel_err(got[:, m.nq:], g["qvel1"])
    assert np.mean(ev > 1e-4) <= 0.05, np.sort(ev)[-3:]
    if goals and m.nblock == 1:
        # bit-exact flags wherever the block is not within fp32 rounding of the geofence
        a = 2 if name != "f2_cupboard" else 0   # qpos address of the block's free joint
        d = np.linalg.norm(g["qpos"][:, a:a + 3] - g["mocap"], axis=1)
        clear = np.abs(d - np.float32(.05)) > 1e-6
        assert np.array_equal(done.cpu().numpy().astype(np.uint8)[clear], g["success"][clear])
    env.close()
