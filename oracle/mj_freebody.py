"""ORACLE (test infrastructure only -- never imported by the product package).

CPU restatement of the MuJoCo semantics the SRBTrack hot path relies on, specialised to the
one-free-body model `gym_trackSRB/Characters/SRB5.xml` of the reference.

MuJoCo itself is a third-party dependency that is NOT vendored in /root/reference and is not
version-pinned (`requirements_SRBTrack2025.txt` lists the bare name `mujoco`); it is not installed
in this container.  This file therefore restates MuJoCo's *published* algorithm (engine_forward.c:
mj_step1 / mj_step2 / mj_Euler / mj_integratePos, engine_util_spatial.c: mju_quatIntegrate /
mju_quat2Mat, engine_support.c: mj_jac / mj_applyFT) for the call sites of the reference:

  env/SRBEnv5_mocap_list_window.py:382-450   (mj_step1 ... mj_step2 sub-step loop)
  env/SRBEnv5_mocap_list_window.py:892,1051  (mj_forward after reset / restore)
  env/testSRB_mujoco_original_viewer.py:471-511 (mj_jac, mj_applyFT, mj_fullM)

PARITY UNPINNED: the reference ships no golden vectors for this path and MuJoCo cannot be run
here, so this restatement is checked only against analytic properties (tests/test_oracle_*.py).

Model constants (gym_trackSRB/Characters/SRB5.xml:1-37):
  box geom mass 60, half sizes (0.19557607215607947, 0.25406692031825, 0.667757440991862)
  <default><joint armature="1" damping="1"/></default>  applies to the <joint type="free"/> dofs
  timestep 0.0083332, integrator "RK4" (split mj_step1/mj_step2 falls back to semi-implicit Euler)
  gravity default (0, 0, -9.81); floor friction "1 .1 .1", contype=conaffinity=0 (no collisions)
"""
import numpy as np

MJ_MINVAL = 1e-15


class FreeBodyModel:
    """Compiled constants of SRB5.xml (what mujoco.MjModel would hold)."""

    def __init__(self):
        self.nv = 6
        self.mass = 60.0
        hx, hy, hz = 0.19557607215607947, 0.25406692031825, 0.667757440991862
        self.half_size = np.array([hx, hy, hz])
        # box inertia about its centre, body frame = geom frame (geom at body origin, no rotation)
        self.inertia = self.mass / 3.0 * np.array([hy * hy + hz * hz, hx * hx + hz * hz, hx * hx + hy * hy])
        self.armature = np.ones(6)
        self.damping = np.ones(6)
        self.timestep = 0.0083332
        self.gravity = np.array([0.0, 0.0, -9.81])
        self.floor_friction = np.array([1.0, 0.1, 0.1])
        zf = -1.067757440991862
        yl = 0.14778803607803974
        # site local positions, SRB5.xml:21-33 (ids 0..9 in document order)
        self.site_pos = np.array([
            [0.0, yl, zf], [0.13, yl + 0.03, zf], [0.13, yl - 0.03, zf], [-0.075, yl + 0.03, zf], [-0.075, yl - 0.03, zf],
            [0.0, -yl, zf], [0.13, -yl + 0.03, zf], [0.13, -yl - 0.03, zf], [-0.075, -yl + 0.03, zf], [-0.075, -yl - 0.03, zf],
        ])
        self.body_pos0 = np.array([0.0, 0.0, 1.067757440991862])

    # diag of mj_fullM(qM): CRB inertia + armature (armature is part of qM in MuJoCo)
    def fullM_diag(self):
        return np.concatenate([np.full(3, self.mass), self.inertia]) + self.armature


class FreeBodyData:
    def __init__(self, m):
        self.qpos = np.concatenate([m.body_pos0, [1.0, 0.0, 0.0, 0.0]])
        self.qvel = np.zeros(6)
        self.xpos = np.zeros(3)
        self.xmat = np.eye(3).reshape(9)           # row-major, like mjData.xmat[body]
        self.qfrc_bias = np.zeros(6)
        self.qfrc_passive = np.zeros(6)
        self.qfrc_applied = np.zeros(6)
        self.xfrc_applied = np.zeros(6)            # world-frame (force, torque) at the body CoM
        self.qacc = np.zeros(6)
        self.site_xpos = np.zeros((10, 3))
        self.site_xmat = np.zeros((10, 9))
        self.time = 0.0


def quat2mat(q):
    """mju_quat2Mat (engine_util_spatial.c)."""
    q00, q01, q02, q03 = q[0] * q[0], q[0] * q[1], q[0] * q[2], q[0] * q[3]
    q11, q12, q13 = q[1] * q[1], q[1] * q[2], q[1] * q[3]
    q22, q23, q33 = q[2] * q[2], q[2] * q[3], q[3] * q[3]
    return np.array([
        [q00 + q11 - q22 - q33, 2 * (q12 - q03), 2 * (q13 + q02)],
        [2 * (q12 + q03), q00 - q11 + q22 - q33, 2 * (q23 - q01)],
        [2 * (q13 - q02), 2 * (q23 + q01), q00 - q11 - q22 + q33],
    ])


def mul_quat(a, b):
    """mju_mulQuat."""
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_integrate(quat, vel, scale):
    """mju_quatIntegrate: quat <- normalize(quat) * quat(axis=vel/|vel|, angle=scale*|vel|)."""
    n = np.sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2])
    if n < MJ_MINVAL:
        axis = np.array([1.0, 0.0, 0.0])
        n = 0.0
    else:
        axis = vel / n
    angle = scale * n
    if angle == 0.0:
        qrot = np.array([1.0, 0.0, 0.0, 0.0])
    else:
        s = np.sin(angle * 0.5)
        qrot = np.array([np.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s])
    q = quat / np.sqrt(quat.dot(quat))
    return mul_quat(q, qrot)


def mj_kinematics(m, d):
    """Position stage: xpos/xmat of the body (quaternion normalised on the copy) and model-defined sites."""
    q = d.qpos[3:7]
    q = q / np.sqrt(q.dot(q))
    R = quat2mat(q)
    d.xpos = d.qpos[:3].copy()
    d.xmat = R.reshape(9).copy()
    d.site_xpos = d.xpos[None, :] + m.site_pos @ R.T
    d.site_xmat = np.tile(R.reshape(9), (10, 1))


def mj_step1(m, d):
    """mj_fwdPosition + mj_fwdVelocity for one free body whose CoM is the joint origin.
    qfrc_bias = RNE with zero acceleration: -m*g on the world-frame translational dofs,
    w x (I w) on the body-frame rotational dofs (no armature in the bias).
    qfrc_passive = -damping * qvel (mj_passive runs in the velocity stage)."""
    mj_kinematics(m, d)
    w = d.qvel[3:6]
    Iw = m.inertia * w
    d.qfrc_bias = np.concatenate([-m.mass * m.gravity, np.cross(w, Iw)])
    d.qfrc_passive = -m.damping * d.qvel


def mj_forward(m, d):
    mj_step1(m, d)
    d.qacc = _qfrc_smooth(m, d) / m.fullM_diag()


def _qfrc_smooth(m, d):
    """mj_fwdAcceleration: passive - bias + applied + J^T xfrc_applied (mj_xfrcAccumulate)."""
    R = d.xmat.reshape(3, 3)
    xf = np.concatenate([d.xfrc_applied[:3], R.T @ d.xfrc_applied[3:6]])
    return d.qfrc_passive - d.qfrc_bias + d.qfrc_applied + xf


def mj_step2(m, d):
    """mj_fwdAcceleration + mj_Euler (integrator RK4 is not available in the split-step API and
    defaults to Euler).  Euler integrates joint damping implicitly:
        (M + h*diag(damping)) qacc = qfrc_smooth ; qvel += h*qacc ; mj_integratePos(qpos, qvel, h).
    xpos/xmat/site_* are NOT refreshed here (that is what makes the orientation read by the env one
    sub-step stale)."""
    h = m.timestep
    f = _qfrc_smooth(m, d)
    d.qacc = f / m.fullM_diag()
    qacc_damped = f / (m.fullM_diag() + h * m.damping)
    d.qvel = d.qvel + h * qacc_damped
    d.qpos = d.qpos.copy()
    d.qpos[:3] = d.qpos[:3] + h * d.qvel[:3]
    d.qpos[3:7] = quat_integrate(d.qpos[3:7], d.qvel[3:6], h)
    d.time += h


def mj_jacp(m, d, point):
    """Translational Jacobian of mj_jac for a world point attached to the body:
    columns 0..2 = I3 (world-frame translation), columns 3..5 = (R e_j) x (point - xpos)."""
    R = d.xmat.reshape(3, 3)
    jacp = np.zeros((3, 6))
    jacp[:, :3] = np.eye(3)
    r = point - d.xpos
    for j in range(3):
        jacp[:, 3 + j] = np.cross(R[:, j], r)
    return jacp


def mj_applyFT(m, d, force, torque, point, qfrc_target):
    R = d.xmat.reshape(3, 3)
    jacp = mj_jacp(m, d, point)
    qfrc_target += jacp.T @ np.asarray(force).reshape(3)
    qfrc_target[3:6] += R.T @ np.asarray(torque).reshape(3)
