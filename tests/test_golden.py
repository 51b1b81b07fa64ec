"""Committed golden vectors (tests/golden/*.npz, 256 states per model, written by tests/golden/make_golden.py from the
INDEPENDENT fp64 numpy oracle oracle/mjstep.py): the oracle reproduces them exactly, the compiled fp64 port agrees to
1e-6, and the CUDA path (GPU tests, every kernel that serves the model) agrees to the 1e-4 relative bound of
BASELINE.json with ZERO outliers on the states whose one-step map is well defined at that precision (`sensitive` = 0, see
make_golden.py: the rest sit on ties of the portal refinement, where MuJoCo itself depends on rounding) - and bit-exact
success flags.  PARITY UNPINNED: these are oracle outputs, not MuJoCo outputs."""
import numpy as np
import pytest

from conftest import GOLDEN
from scenarios import rel_err

NAMES = ["c1_readme", "c1b_readme_block", "c2_push", "c3_arm", "c5_clutter", "f2_cupboard"]
REGIMES = ["limit", "world-block", "base-block", "base-world", "arm-block", "arm-world", "block-block", "pan-block"]
TOL = 1e-4


def load(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def regime_count(g, r):
    return int(((g["regime"] >> REGIMES.index(r)) & 1).sum())


def test_fixtures_cover_the_regimes():
    """The regimes VERDICT r1 found missing are in the fixtures: joint-limit rows and robot-hull / pan contacts for the
    sliding-base kernels, gripper / arm contacts with the block and the pan for configs[2] (>= 25 % of the states),
    >= 32 block-block box-box states for configs[4]."""
    g = {n: load(n) for n in NAMES}
    for n in NAMES:
        assert len(g[n]["qpos"]) >= 256
        assert g[n]["sensitive"].mean() <= 0.08 and g[n]["overflow"].mean() <= 0.02
    assert regime_count(g["c1_readme"], "limit") >= 32 and regime_count(g["c1_readme"], "base-world") >= 32
    assert regime_count(g["c2_push"], "limit") >= 32 and regime_count(g["c2_push"], "base-block") >= 64
    assert regime_count(g["c2_push"], "world-block") >= 200
    assert regime_count(g["c3_arm"], "arm-block") >= 64 and regime_count(g["c3_arm"], "arm-world") >= 64
    assert regime_count(g["c3_arm"], "limit") >= 64
    assert regime_count(g["c5_clutter"], "block-block") >= 32
    assert regime_count(g["f2_cupboard"], "pan-block") >= 200
    assert 0.1 < g["c2_push"]["success"].mean() < 0.9


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_reproduces_golden(name, models):
    from oracle import mjstep

    g, m = load(name), models[name]
    for e in range(0, len(g["qpos"]), 16):
        d = mjstep.Data(m)
        d.qpos[:] = g["qpos"][e]; d.qvel[:] = g["qvel"][e]; d.qacc_warmstart[:] = g["warm"][e]; d.ctrl[:] = g["ctrl"][e]
        mjstep.step(m, d)
        assert np.array_equal(d.qpos, g["qpos1"][e]) and np.array_equal(d.qvel, g["qvel1"][e])
        assert len(d.contacts) == g["ncon"][e] and d.nefc == g["nefc"][e]


@pytest.mark.parametrize("name", NAMES)
def test_cpp_port_matches_golden(name, models, ports):
    """The g++ port shares its substep source with the product (hsr_core.h, fp64, one lane): this is a cross-check of
    two implementations of the same algorithm, not an independent derivation."""
    g, port = load(name), ports[name]
    ok = (g["sensitive"] == 0) & (g["overflow"] == 0)
    out = port.step(g["qpos"], g["qvel"], g["warm"], g["ctrl"], nsub=1)
    assert rel_err(out["qpos"], g["qpos1"])[ok].max() < 1e-8
    assert rel_err(out["qvel"], g["qvel1"])[ok].max() <= 1e-6
    if models[name].nblock:
        port.set_goals(np.zeros(6), None, .05)
        try:
            out = port.step(g["qpos"], g["qvel"], g["warm"], g["ctrl"], g["mocap"], nsub=1)
        finally:
            port.set_goals(None)
        assert np.array_equal(out["success"][ok], g["success"][ok])


def _kernels(name):
    return ["general", "fast", "wpe"] if name in ("c1_readme", "c1b_readme_block", "c2_push") else ["general"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,kernel", [(n, k) for n in NAMES for k in _kernels(n)])
def test_cuda_matches_golden(name, kernel, models):
    torch = pytest.importorskip("torch")
    from hsr_env_b200.env import BatchedHSREnv
    from hsr_env_b200.spaces import Box
    from hsr_env_b200.util import GoalSpec

    g, m = load(name), models[name]
    n = len(g["qpos"])
    goals = [GoalSpec(None, Box([0, 0, 0], [0, 0, 0]), .05)] if m.nblock else None
    env = BatchedHSREnv(f"{name}.hsrb", goals, n_envs=n, device="cuda:0", kernel=kernel)
    assert env.launch_info()["kernel"] == kernel
    if goals:
        env.reset()
    env.set_state(g["qpos"], g["qvel"], g["warm"], g["mocap"] if goals else None)
    obs, reward, done, info = env.step(torch.tensor(g["ctrl"], dtype=torch.float32), steps=1)
    got = obs.double().cpu().numpy()
    flags = info["bad_state"].cpu().numpy()
    ok = (g["sensitive"] == 0) & (g["overflow"] == 0)
    # the fast kernels hold 8 contacts per environment; the fixtures of their family stay below (max 5)
    assert int(((flags & 1) != 0)[ok].sum()) == 0, "contact overflow on a state within the default capacity"
    eq, ev = rel_err(got[:, :m.nq], g["qpos1"]), rel_err(got[:, m.nq:], g["qvel1"])
    print(f"{name}/{kernel}: {int(ok.sum())} well-defined states: qpos err max {eq[ok].max():.2e}, qvel err max {ev[ok].max():.2e} "
          f"median {np.median(ev[ok]):.2e}; {int((~ok).sum())} sensitive / overflowing states: qvel err max {ev[~ok].max() if (~ok).any() else 0:.2e}")
    assert eq[ok].max() <= TOL
    assert ev[ok].max() <= TOL, (np.nonzero(ok & (ev > TOL))[0], np.sort(ev[ok])[-3:])     # zero outliers
    if goals and m.nblock:
        # success flags bit-exact (the fixture keeps every goal >= 1e-5 clear of the geofence)
        assert np.array_equal(done.cpu().numpy().astype(np.uint8)[ok], g["success"][ok])
        assert np.array_equal(reward.cpu().numpy()[ok], g["success"][ok].astype(np.float32))
    env.close()
