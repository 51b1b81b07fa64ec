"""Row f1: the 'openai' observation.  Host-side torch kinematics (hsr_env_b200/kin.py) against the numpy
oracle's kinematics and Jacobians on CPU; on the GPU, the observation kernel (hsrb_openai_obs) against that torch statement
and against the action kernel's own forward pass."""
import numpy as np
import pytest

from scenarios import rollout_states

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name,pan", [("c3_arm", True), ("c2_push", False), ("f2_cupboard", True)])
def test_body_kinematics_and_velocities_match_oracle(name, pan, models, ports):
    from hsr_env_b200 import kin
    from oracle import mjstep

    m = models[name]
    qpos, qvel, warm, ctrl = rollout_states(ports[name], m, 6, seed=21, pan=pan)
    rng = np.random.default_rng(0)
    qvel = qvel + rng.normal(0, .3, qvel.shape)           # every dof moving
    xpos, xmat, velp, velr = [t.numpy() for t in kin.body_kinematics(m, torch.tensor(qpos), torch.tensor(qvel))]
    for e in range(len(qpos)):
        d = mjstep.Data(m)
        d.qpos[:] = qpos[e]; d.qvel[:] = qvel[e]
        mjstep.kinematics(m, d); mjstep.com_crb(m, d)
        assert np.allclose(xpos[e], d.xpos, atol=1e-12) and np.allclose(xmat[e], d.xmat, atol=1e-12)
        for b in range(1, m.nbody):
            Jp, Jr = mjstep.jac(m, d, b, d.xpos[b])
            assert np.allclose(velp[e, b], Jp @ qvel[e], atol=1e-10), (name, b)
            assert np.allclose(velr[e, b], Jr @ qvel[e], atol=1e-10), (name, b)


def test_openai_observation_layout(models, ports):
    from hsr_env_b200 import kin
    from oracle import mjstep

    m = models["c3_arm"]
    qpos, qvel, warm, ctrl = rollout_states(ports["c3_arm"], m, 4, seed=3, pan=True)
    dt = float(m.timestep)
    bb = int(m.block_body[0])
    obs = kin.openai_observation(m, torch.tensor(qpos), torch.tensor(qvel), dt, bb).numpy()
    assert obs.shape == (4, 25)
    names = list(m.names["joint"])
    for e in range(4):
        d = mjstep.Data(m); d.qpos[:] = qpos[e]; d.qvel[:] = qvel[e]
        mjstep.kinematics(m, d); mjstep.com_crb(m, d)
        grip = np.mean([d.xpos[int(b)] + d.xmat[int(b)] @ m.finger_pos[k] for k, b in enumerate(m.finger_body)], axis=0)
        gv = np.mean([mjstep.jac(m, d, int(b), d.xpos[int(b)] + d.xmat[int(b)] @ m.finger_pos[k])[0] @ qvel[e]
                      for k, b in enumerate(m.finger_body)], axis=0)
        o = obs[e]
        assert np.allclose(o[0:3], grip, atol=1e-12) and np.allclose(o[3:6], d.xpos[bb], atol=1e-12)
        assert np.allclose(o[6:9], d.xpos[bb] - grip, atol=1e-12)
        gj = [names.index(f"hand_{x}_proximal_joint") for x in "lr"]
        assert np.allclose(o[9:11], [qpos[e][m.jnt_qposadr[j]] for j in gj])
        Jp, Jr = mjstep.jac(m, d, bb, d.xpos[bb])
        assert np.allclose(o[14:17], (Jp @ qvel[e] - gv) * dt, atol=1e-12)
        assert np.allclose(o[17:20], (Jr @ qvel[e]) * dt, atol=1e-12)
        assert np.allclose(o[20:23], gv * dt, atol=1e-12)
        assert np.allclose(o[23:25], [dt * qvel[e][m.jnt_dofadr[j]] for j in gj])
    # mat2euler: rotation about z by +0.3 rad -> (0, 0, 0.3) in the OpenAI robotics convention
    c, s = np.cos(.3), np.sin(.3)
    R = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1.]], dtype=torch.float64)
    assert np.allclose(kin.mat2euler(R).numpy(), [0, 0, .3])


@pytest.mark.gpu
def test_openai_observation_against_the_kernel_forward_pass():
    from hsr_env_b200.env import BatchedHSREnv

    n = 64
    env = BatchedHSREnv("c3_arm.hsrb", None, obs_type="openai", n_envs=n, device="cuda:0", steps_per_action=25)
    assert env.observation_space.shape == (25,)
    obs = env.reset()
    assert obs.shape == (n, 25)
    gen = torch.Generator().manual_seed(0)
    for _ in range(3):
        lo = torch.tensor(env.model.act_ctrlrange[:, 0], dtype=torch.float32); hi = torch.tensor(env.model.act_ctrlrange[:, 1], dtype=torch.float32)
        obs, reward, done, info = env.step(lo + (hi - lo) * torch.rand(n, env.nu, generator=gen))
        assert obs.shape == (n, 25) and torch.isfinite(obs).all()
        # the observation kernel (hsrb_openai_obs) = the torch / fp64 statement of the observation (kin.py, checked against
        # the numpy oracle above) on the same state: all 25 components, fp32 output rounding only
        from hsr_env_b200 import kin

        qpos, qvel, _, _ = env.get_state()
        ref = kin.openai_observation(env.model, qpos, qvel, float(env.model.timestep), int(env.model.block_body[0]))
        err = (obs.double() - ref).abs()
        assert float(err.max()) <= 2e-6, float(err.max())
        # and its positions = the action kernel's own forward pass (hsrb_forward)
        assert torch.allclose(obs[:, 0:3], env.gripper_pos(), atol=2e-6)
        assert torch.allclose(obs[:, 3:6], env.block_pos()[:, 0], atol=2e-6)
    env.close()


@pytest.mark.gpu
def test_openai_observation_without_finger_joints():
    """--use-dof slide_x slide_y: the finger joints are gone, their observation slots are zero (kin.py does the same)."""
    from hsr_env_b200 import kin
    from hsr_env_b200.env import BatchedHSREnv

    env = BatchedHSREnv("c2_push.hsrb", None, obs_type="openai", n_envs=32, device="cuda:0", steps_per_action=10)
    obs = env.reset()
    obs, _, _, _ = env.step(torch.zeros(32, env.nu))
    qpos, qvel, _, _ = env.get_state()
    ref = kin.openai_observation(env.model, qpos, qvel, float(env.model.timestep), int(env.model.block_body[0]))
    assert float((obs.double() - ref).abs().max()) <= 2e-6
    assert float(obs[:, 9:11].abs().max()) == 0.0 and float(obs[:, 23:25].abs().max()) == 0.0
    env.close()
