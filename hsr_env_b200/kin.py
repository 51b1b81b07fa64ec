"""Body poses and 6-D body velocities of a batch of states, as torch ops on the caller's device (fp64).

Used by the ``'openai'`` observation (``/root/reference/hsr/env.py:72-110``), which needs what
mujoco-py exposes as ``data.get_body_xpos / get_body_xmat / get_body_xvelp / get_body_xvelr``.  This is host-side
glue on the observation path, a few hundred flops per environment once per action -- not the physics hot path, which is
the CUDA kernel.  Same kinematic conventions as the kernel and the oracle: bodies in tree order,
free joints world-attached with their angular velocity in the body frame, slides / hinges about ``jnt_axis`` through
``jnt_pos`` in the body frame, reference configuration ``qpos0``.
"""
from __future__ import annotations

import torch

JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 1, 2


def _qmul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def _q2mat(q):
    w, x, y, z = q.unbind(-1)
    r = torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1)
    return r.reshape(*q.shape[:-1], 3, 3)


def _normalize(q):
    return q / q.norm(dim=-1, keepdim=True).clamp_min(1e-15)


def body_kinematics(model, qpos: torch.Tensor, qvel: torch.Tensor):
    """(xpos [N,nb,3], xmat [N,nb,3,3], velp [N,nb,3], velr [N,nb,3]): world pose of every (fused) body frame and the
    linear velocity of its origin / its angular velocity, both in the world frame."""
    dev = qpos.device
    f64 = dict(dtype=torch.float64, device=dev)
    qpos = qpos.to(torch.float64); qvel = qvel.to(torch.float64)
    n, nb = qpos.shape[0], model.nbody
    T = lambda a: torch.as_tensor(a, **f64)  # noqa: E731
    body_pos, body_quat, jnt_pos, jnt_axis, qpos0 = T(model.body_pos), T(model.body_quat), T(model.jnt_pos), T(model.jnt_axis), T(model.qpos0)
    xpos = [torch.zeros(n, 3, **f64)]; xmat = [torch.eye(3, **f64).expand(n, 3, 3)]
    xquat = [torch.tensor([1.0, 0, 0, 0], **f64).expand(n, 4)]
    velp = [torch.zeros(n, 3, **f64)]; velr = [torch.zeros(n, 3, **f64)]
    for b in range(1, nb):
        p = int(model.body_parent[b])
        j0, nj = int(model.body_jntadr[b]), int(model.body_jntnum[b])
        if nj == 1 and int(model.jnt_type[j0]) == JNT_FREE:
            a, v = int(model.jnt_qposadr[j0]), int(model.jnt_dofadr[j0])
            pos = qpos[:, a:a + 3]
            quat = _normalize(qpos[:, a + 3:a + 7])
            R = _q2mat(quat)
            xpos.append(pos); xquat.append(quat); xmat.append(R)
            velp.append(qvel[:, v:v + 3]); velr.append(torch.einsum("nij,nj->ni", R, qvel[:, v + 3:v + 6]))
            continue
        pos = xpos[p] + torch.einsum("nij,j->ni", xmat[p], body_pos[b])
        quat = _qmul(xquat[p], body_quat[b].expand(n, 4))
        w = velr[p]
        vel = velp[p] + torch.cross(velr[p], pos - xpos[p], dim=-1)   # origin of b carried rigidly by its parent
        for j in range(j0, j0 + nj):
            R = _q2mat(quat)
            anchor = pos + torch.einsum("nij,j->ni", R, jnt_pos[j])
            axis = torch.einsum("nij,j->ni", R, jnt_axis[j])
            a, v = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
            q = qpos[:, a] - qpos0[a]
            qd = qvel[:, v:v + 1]
            if int(model.jnt_type[j]) == JNT_SLIDE:
                pos = pos + axis * q[:, None]
                vel = vel + axis * qd
            else:
                half = 0.5 * q
                dq = torch.cat([torch.cos(half)[:, None], torch.sin(half)[:, None] * jnt_axis[j]], -1)
                quat = _qmul(quat, dq)
                pos = anchor - torch.einsum("nij,j->ni", _q2mat(quat), jnt_pos[j])
                w = w + axis * qd
                vel = vel + torch.cross(axis * qd, pos - anchor, dim=-1)
        quat = _normalize(quat)
        xpos.append(pos); xquat.append(quat); xmat.append(_q2mat(quat)); velp.append(vel); velr.append(w)
    return torch.stack(xpos, 1), torch.stack(xmat, 1), torch.stack(velp, 1), torch.stack(velr, 1)


def mat2euler(mat: torch.Tensor) -> torch.Tensor:
    """/root/reference/hsr/env.py:256-272 (the rotation.py convention of the OpenAI robotics environments)."""
    cy = torch.sqrt(mat[..., 2, 2] ** 2 + mat[..., 1, 2] ** 2)
    cond = cy > torch.finfo(torch.float64).eps * 4
    e2 = torch.where(cond, -torch.atan2(mat[..., 0, 1], mat[..., 0, 0]), -torch.atan2(-mat[..., 1, 0], mat[..., 1, 1]))
    e1 = -torch.atan2(-mat[..., 0, 2], cy)
    e0 = torch.where(cond, -torch.atan2(mat[..., 1, 2], mat[..., 2, 2]), torch.zeros_like(cy))
    return torch.stack([e0, e1, e2], -1)


def openai_observation(model, qpos, qvel, dt: float, block_body: int, gripper_joints=("hand_l_proximal_joint", "hand_r_proximal_joint")):
    """The 25-d ``'openai'`` observation per the *intent* of /root/reference/hsr/env.py:72-110.  That branch is dead code in
    the snapshot (it calls ``sim.get_body_xvelp`` / ``sim.timestep``, which do not exist, and fills
    ``gripper_state`` / ``qvels`` with qpos *addresses*); it is the Fetch observation of the OpenAI robotics
    environments, so the intent is taken from there:

        grip_pos(3) | object_pos(3) | object_pos - grip_pos(3) | gripper_state(2) = qpos of the two proximal finger joints |
        object_rot(3) = mat2euler(block xmat) | (object_velp - grip_velp) dt (3) | object_velr dt (3) | grip_velp dt (3) |
        gripper_vel(2) = dt * qvel of those joints

    with grip_pos / grip_velp the mean over the two distal finger links, dt = nsubsteps * timestep = timestep.  Finger
    joints removed by ``--use-dof`` contribute zeros."""
    xpos, xmat, velp, velr = body_kinematics(model, qpos, qvel)
    fb = [int(b) for b in model.finger_body]
    fp = torch.as_tensor(model.finger_pos, dtype=torch.float64, device=xpos.device)
    gp, gv = 0, 0
    for k, b in enumerate(fb):
        r = torch.einsum("nij,j->ni", xmat[:, b], fp[k])
        gp = gp + 0.5 * (xpos[:, b] + r)
        gv = gv + 0.5 * (velp[:, b] + torch.cross(velr[:, b], r, dim=-1))
    names = list(model.names.get("joint", []))
    gs, gvel = [], []
    for jn in gripper_joints:
        if jn in names:
            j = names.index(jn)
            gs.append(qpos[:, int(model.jnt_qposadr[j])].to(torch.float64)); gvel.append(qvel[:, int(model.jnt_dofadr[j])].to(torch.float64))
        else:
            gs.append(torch.zeros_like(gp[:, 0])); gvel.append(torch.zeros_like(gp[:, 0]))
    op = xpos[:, block_body]
    obs = torch.cat([gp, op, op - gp, torch.stack(gs, -1), mat2euler(xmat[:, block_body]), (velp[:, block_body] - gv) * dt,
                     velr[:, block_body] * dt, gv * dt, torch.stack(gvel, -1) * dt], -1)
    return obs
