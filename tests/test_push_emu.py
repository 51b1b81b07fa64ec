"""The CUDA fast-path kernel (hsr_env_b200/csrc/hsrb_push.cuh), compiled unchanged by g++ on top of the SIMT
emulator in tests/simt_emu, against the fp64 oracle port: teacher-forced single substep within 1e-4 relative
(BASELINE.json north_star), for every lanes-per-environment layout.  CPU only: this is how the kernel's logic is
checked on a box without a GPU; the GPU run of the same comparison is tests/test_gpu_parity.py."""
import sys
from pathlib import Path

import numpy as np
import pytest

from scenarios import rel_err, rollout_states

sys.path.insert(0, str(Path(__file__).resolve().parent / "simt_emu"))

TOL = 1e-4


@pytest.fixture(scope="module")
def emu():
    import build

    build.build()
    return build


@pytest.mark.parametrize("name,n,G", [("c2_push", 96, 8), ("c2_push", 48, 16), ("c2_push", 32, 32),
                                       ("c1b_readme_block", 32, 8), ("c1_readme", 32, 8),
                                       # G = 1: the warp-per-environment kernel (hsrb_wpe.cuh)
                                       ("c2_push", 96, 1), ("c1b_readme_block", 32, 1), ("c1_readme", 32, 1)])
def test_emulated_kernel_matches_oracle(name, n, G, models, ports, emu):
    model, port = models[name], ports[name]
    qpos, qvel, warm, ctrl = rollout_states(port, model, n, seed=len(name) * 7, float32=True)
    out = emu.step(model, qpos, qvel, warm, ctrl, nsub=1, G=G, threads=32)
    ref = port.step(qpos, qvel, warm, ctrl, nsub=1)
    eq, ev = rel_err(out["qpos"], ref["qpos"]), rel_err(out["qvel"], ref["qvel"])
    assert np.all(out["taken"] == 1) and np.all(out["flags"] == 0)
    assert eq.max() <= TOL
    assert np.mean(ev > TOL) <= 0.02, np.sort(ev)[-5:]


@pytest.mark.parametrize("G", [8, 1])
def test_emulated_kernel_limits_and_hull_contacts(G, models, ports, emu):
    """No-block model: joint-limit rows and robot-hull / pan contacts (states the roll-outs above do not reach)."""
    model, port = models["c1_readme"], ports["c1_readme"]
    rng = np.random.default_rng(0)
    n = 32
    qpos = np.zeros((n, 2)); qvel = rng.normal(0, .05, (n, 2)); warm = np.zeros((n, 2)); ctrl = rng.uniform(-1, 1, (n, 2))
    qpos[:16, 0] = rng.uniform(-.125, -.118, 16); ctrl[:16, 0] = -1     # lower slide_x limit
    qpos[16:, 0] = rng.uniform(.155, .165, 16); ctrl[16:, 0] = 1        # base against the pan edge
    qpos[:, 1] = rng.uniform(-.225, .245, n)
    qpos, qvel, warm, ctrl = [x.astype(np.float32).astype(np.float64) for x in (qpos, qvel, warm, ctrl)]
    seen_rows = 0
    for _ in range(3):
        out = emu.step(model, qpos, qvel, warm, ctrl, nsub=1, G=G)
        ref = port.step(qpos, qvel, warm, ctrl, nsub=1)
        assert rel_err(out["qpos"], ref["qpos"]).max() <= TOL
        assert rel_err(out["qvel"], ref["qvel"]).max() <= TOL
        seen_rows += int(out["stats"][5])
        qpos, qvel, warm = [ref[k].astype(np.float32).astype(np.float64) for k in ("qpos", "qvel", "warm")]
    assert seen_rows > 100


@pytest.mark.parametrize("G", [8, 1])
def test_emulated_kernel_goal_flags_and_early_exit(G, models, ports, emu):
    """Per-substep goal test with early break: flags and executed-substep counts equal the fp32 port's."""
    model, port = models["c2_push"], ports["c2_push"]
    n = 32
    rng = np.random.default_rng(3)
    qpos, qvel, warm, ctrl = rollout_states(port, model, n, seed=11, float32=True)
    d = rng.uniform(0.045, 0.055, n); ang = rng.uniform(0, 2 * np.pi, n)
    mocap = (qpos[:, 2:5] + np.stack([d * np.cos(ang), d * np.sin(ang), np.zeros(n)], 1)).astype(np.float32).astype(np.float64)
    port.set_goals(np.zeros(6), None, .05)
    try:
        ref = port.step(qpos, qvel, warm, ctrl, mocap, nsub=20, use_float=True)
    finally:
        port.set_goals(None)
    out = emu.step(model, qpos, qvel, warm, ctrl, mocap, nsub=20, G=G, geofence=.05)
    assert 0.05 < out["success"].mean() < 0.95
    same = out["success"] == ref["success"]
    assert same.mean() >= 0.9
    assert np.mean(out["taken"][same] == ref["taken"][same]) >= 0.9
    assert np.all(out["taken"][out["success"] == 0] == 20)


@pytest.mark.parametrize("G", [8, 1])
def test_emulated_kernel_separating_direction_cache(G, models, ports, emu, monkeypatch):
    """The cached separating direction of a candidate pair (mpr_penetration's `sep`) only shortens queries that end
    without a contact: a 40-substep action gives the same trajectory with the cache switched off (HSRB_OPTS bit 0)."""
    model, port = models["c2_push"], ports["c2_push"]
    qpos, qvel, warm, ctrl = rollout_states(port, model, 32, seed=5, float32=True)
    out = emu.step(model, qpos, qvel, warm, ctrl, nsub=40, G=G, threads=64)
    monkeypatch.setenv("HSRB_OPTS", "1")
    ref = emu.step(model, qpos, qvel, warm, ctrl, nsub=40, G=G, threads=64)
    assert np.all(out["flags"] == 0) and np.all(ref["flags"] == 0)
    assert int(ref["stats"][2]) == int(out["stats"][2]) > 40 * 32          # same narrowphase calls, some convex-convex
    same = np.all(out["qpos"] == ref["qpos"], axis=1) & np.all(out["qvel"] == ref["qvel"], axis=1)
    assert same.mean() >= 0.95, same.mean()                                 # bitwise except borderline touching pairs
    assert rel_err(out["qpos"], ref["qpos"]).max() <= 1e-4
