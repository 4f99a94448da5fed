"""ORACLE (test infrastructure only -- never imported by the product package).

NumPy fp64 restatement of the SRBTrack per-environment step of the reference:

  env/SRBEnv5_mocap_list_window.py            (variant "test": SDS swing-foot filter, force clamp)
  env/SRBEnv5_mocap_list_window_training.py   (variant "training": no filter/clamp, other weights,
                                               obs taken after current_frame += 1)
  env/testSRB_mujoco_original_viewer.py:76-155,167-201,297-327,471-599  (frames, log_se3, QP)

Third-party arithmetic that is NOT in /root/reference and is restated instead:
  * mujoco (unpinned)  -> oracle/mj_freebody.py
  * clarabel (unpinned, interior point, default tol ~1e-8) -> the same QP is solved EXACTLY here:
    qddot is eliminated through the equality constraint and the remaining ridge-regularised
    non-negative least squares problem is solved with scipy.optimize.nnls (Lawson-Hanson active set).
  * scipy.linalg.logm is available in this image and is called exactly as the reference does.

PARITY: PINNED TO THE REFERENCE'S OWN ENVIRONMENT CODE for everything the reference itself wrote -- tests/tools/ref_srbenv_harness.py
imports env/SRBEnv5_mocap_list_window.py, ..._window_training.py and ..._terrain_window.py from /root/reference and runs them with
mujoco / clarabel / controlmodule.SDS / SRBdata answered by stand-ins (oracle/mj_freebody.py, an exact QP solve, oracle/sds_filter.py,
the committed clip tables, oracle/terrain.py for mj_ray); their roll-outs (tests/golden/refenv_*.npz, generator
tests/tools/make_golden_refenv.py) are reproduced by this file at 1e-13 (tests/test_refenv_pins.py), incl. the contact hysteresis, the
6-component impulse and the terrain variant on all four playground patches.  UNPINNED: the third-party semantics behind those stand-ins
(MuJoCo's mj_step1 / mj_step2 / mj_ray, Clarabel, taesooLib's SDS) -- checked analytically only (KKT residual of the un-eliminated QP,
free-flight invariants; see tests/test_oracle_cpu.py).

Every quirk listed in SURVEY.md section 8 (Q1,Q2,Q3,Q6,Q11,Q12,Q14) is reproduced deliberately.
"""
import math
import numpy as np
from scipy.linalg import logm
from scipy.optimize import nnls

from . import mj_freebody as mj
from .sds_filter import SDS


# ----------------------------------------------------------------------------- frame helpers
def _ff_frame(SRB_ori):
    """testSRB_mujoco_original_viewer.py:83-96 (shared by all global2* helpers)."""
    x_SRB = SRB_ori[:, 0]
    x_proj = np.array([x_SRB[0], x_SRB[1], 0])
    if np.linalg.norm(x_proj) != 0:
        x_ff = x_proj / np.linalg.norm(x_proj)
    else:
        x_ff = x_proj
    z_ff = np.array([0, 0, 1])
    y_ff = np.cross(z_ff, x_ff)
    return np.column_stack((x_ff, y_ff, z_ff))


def global2forward_facing(SRB_ori, SRB_origin, global_pos):
    ff = _ff_frame(np.array(SRB_ori))
    return ff.T @ (global_pos - SRB_origin), ff


def global2projected(SRB_ori, SRB_origin, global_pos):
    ff = _ff_frame(np.array(SRB_ori))
    o = SRB_origin.copy()
    o[2] = 0
    return ff.T @ (global_pos - o), ff


def transformation_matrix(R, p):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = p
    return T


def log_se3(T_t, T_hat):
    """viewer.py:175-201"""
    T_rel = np.dot(np.linalg.inv(T_t), T_hat)
    log_term = logm(T_rel)
    log_term = np.real(log_term)
    omega_hat = log_term[:3, :3]
    v = log_term[:3, 3]
    omega = np.array([omega_hat[2, 1], omega_hat[0, 2], omega_hat[1, 0]])
    twist = np.zeros(6)
    twist[:3] = v
    twist[3:] = omega
    return twist


def z_axis_rotation_matrix(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def axis_angle_to_rotation_matrix(axis_angle):
    """viewer.py:310-327"""
    angle = np.linalg.norm(axis_angle)
    if angle < 1e-8:
        return np.eye(3)
    axis = axis_angle / angle
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * np.dot(K, K)


def foot_contact_to_one_hot(foot_in_contact):
    idx = (1 if foot_in_contact[0] else 0) + (2 if foot_in_contact[1] else 0)
    one_hot = np.zeros(4, dtype=int)
    one_hot[idx] = 1
    return one_hot


def quaternion_multiply(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def angle_between_vectors(v1, v2, normal):
    """SRBEnv5_mocap_list_window.py:1302-1313 -- normalises its arguments IN PLACE (Q11)."""
    v1 /= np.sqrt(v1.dot(v1))
    v2 /= np.sqrt(v2.dot(v2))
    normal /= np.sqrt(normal.dot(normal))
    cross = np.cross(v1, v2)
    dot = np.dot(v1, v2)
    det = np.dot(normal, cross)
    return np.arctan2(det, dot)


# ----------------------------------------------------------------------------- feet sites
REL = [np.array([0.12, 0.06, 0]), np.array([0.12, -0.06, 0]), np.array([-0.12, 0.06, 0]), np.array([-0.12, -0.06, 0])]


def set_feet_site_this_step_env(d, feet_pos_this_step, feet_ori_this_step, des_left, des_right):
    """SRBEnv5_mocap_list_window.py:19-89: overwrite centre + 4 corner sites of both feet
    (swing or not) and the centre site's xmat."""
    from_action = np.vstack((des_left, des_right))
    for i in range(2):
        j = 5 * i
        d.site_xmat[j] = feet_ori_this_step[i].flatten()
        d.site_xpos[j] = from_action[i]
        for k in range(4):
            d.site_xpos[j + 1 + k] = from_action[i] + feet_ori_this_step[i] @ REL[k]


def return_site_to_apply_force_on(feet_pos_this_step):
    """:92-118 -- a foot whose z is exactly 0 is in stance (NaN != 0 -> swing, Q6)."""
    out = [-1] * 10
    for i in range(2):
        if feet_pos_this_step[i][2] != 0:
            continue
        for k in range(5):
            out[5 * i + k] = 5 * i + k
    return out


# ----------------------------------------------------------------------------- contact-force QP
def qp_matrices(m, d, curr_T_mocap, target_vel, site_to_apply_force_on, mu):
    """Assemble exactly what viewer.py:515-573 assembles (dense)."""
    nv = m.nv
    M = np.diag(m.fullM_diag())
    b = d.qfrc_bias.copy()
    xmat = d.xmat.reshape(3, 3)
    _, proj = global2projected(xmat, d.qpos[:3], d.qpos[:3])
    b_list = [np.array([1.2 * mu, 0, 1]), np.array([-1.2 * mu, 0, 1]), np.array([0, 1.2 * mu, 1]), np.array([0, -1.2 * mu, 1])]
    B = np.vstack([proj @ bb for bb in b_list]).T
    Tt = transformation_matrix(xmat.copy(), d.qpos[:3])
    local_twist = log_se3(Tt, curr_T_mocap)
    local_twist = np.concatenate([xmat @ local_twist[:3], local_twist[3:]])
    q_ddot_d = 120 * local_twist + 35 * (target_vel - d.qvel.copy())
    jacp_list = [mj.mj_jacp(m, d, d.site_xpos[s]) for s in site_to_apply_force_on if s != -1]
    nl = 4 * len(jacp_list)
    C_eq = np.zeros((nv, nl))
    for i, jacp in enumerate(jacp_list):
        C_eq[:, 4 * i:4 * (i + 1)] = jacp.T @ B
    return M, b, B, C_eq, q_ddot_d


def solve_qp_for_contact_forces(m, d, curr_T_mocap, target_vel, site_to_apply_force_on, mu, w_lambda=0.001):
    """min |qdd - qdd_d|^2 + w |lam|^2  s.t.  M qdd - C lam = -b, lam >= 0   (viewer.py:549-573).
    Exact solve: qdd = Minv (C lam - b)  =>  min_{lam>=0} |A lam - y|^2 + w|lam|^2,
    A = Minv C, y = qdd_d + Minv b, i.e. NNLS on [[A],[sqrt(w) I]] lam ~ [[y],[0]]."""
    M, b, B, C_eq, q_ddot_d = qp_matrices(m, d, curr_T_mocap, target_vel, site_to_apply_force_on, mu)
    Minv = 1.0 / np.diag(M)
    A = Minv[:, None] * C_eq
    y = q_ddot_d + Minv * b
    nl = A.shape[1]
    Aaug = np.vstack([A, math.sqrt(w_lambda) * np.eye(nl)])
    yaug = np.concatenate([y, np.zeros(nl)])
    lam, _ = nnls(Aaug, yaug, maxiter=50 * nl)
    lam = _polish_active_set(A, y, lam, w_lambda)
    qdd = Minv * (C_eq @ lam - b)
    lambda_list = lam.reshape(-1, 4)
    forces = [B @ l for l in lambda_list]
    return qdd, lambda_list, forces


def _polish_active_set(A, y, lam, w):
    """Re-solve the normal equations on the NNLS support (removes the solver's own tolerance);
    keeps the NNLS answer if the polished point is not primal/dual feasible."""
    S = lam > 0
    if not S.any():
        return lam
    As = A[:, S]
    ls = np.linalg.solve(As.T @ As + w * np.eye(As.shape[1]), As.T @ y)
    out = np.zeros_like(lam)
    out[S] = ls
    grad = A.T @ (A @ out - y) + w * out
    if (ls > 0).all() and (grad[~S] >= -1e-9 * (1 + np.abs(grad).max())).all():
        return out
    return lam


# ----------------------------------------------------------------------------- the environment
WEIGHTS = {
    #            w_s w_m  w_p  w_e   w_p1 w_p2 w_p3 w_p4 w_c
    "test":     (5, 0.1, 0.5, 0,    50, 20,  700, 50,  1),      # SRBEnv5_mocap_list_window.py:561-569
    "training": (5, 1,   0.5, 0.01, 5,  1.5, 25,  1.5, 0.1),    # ..._training.py:473-483
}


class OracleSRBEnv:
    """State-for-state restatement of `SRBEnv` (flat ground).  Clip tables are passed in as a dict
    with the keys loadMocap builds (SRBEnv5_mocap_list_window.py:1187-1197)."""

    def __init__(self, clips, variant="test", terrain=None):
        """variant "terrain" = env/SRBEnv5_mocap_list_terrain_window.py: `clips` must be terrain-lifted tables
        (lift_clip_to_terrain) and `terrain` an oracle.terrain.Terrain."""
        assert variant in ("test", "training", "terrain")
        assert (variant == "terrain") == (terrain is not None)
        self.terrain = terrain
        self.variant = variant
        self.model = mj.FreeBodyModel()
        self.data = mj.FreeBodyData(self.model)
        self.clips = clips
        self.minLogQ = 9
        self.current_frame = 0
        self.window_size = 5
        self.left_foot_in_contact = False
        self.right_foot_in_contact = False
        self.left_site_id, self.right_site_id = 0, 5
        self.site_id_list = [0, 5]
        self.mu = 1.2 * self.model.floor_friction[0]
        self.testing_mode = False
        self.contact_prob = [None, None]     # Policy_prior_posterior_MoE_window._last_contact_prob side channel
        self.is_stand = False
        self.frame_counter = 0
        if variant in ("test", "terrain"):
            self.impulse_interval = 1e9
            self.force = [700, 700, -350]
        else:
            self.impulse_interval = 100000000
            self.force = [1000 / math.sqrt(2) * 0.5, -1000 / math.sqrt(2) * 0.5, 0]
        self.impulse_duration = 24
        self.impulse = 0
        self.f = np.zeros(6)
        self.trace = None                    # optional dict filled with intermediates for parity tests

    # -- :191-206
    def update_impulse(self):
        if self.variant in ("test", "terrain"):
            trig = self.frame_counter % self.impulse_interval == self.impulse_interval - 1
        else:
            trig = self.frame_counter % self.impulse_interval == 0
        if trig:
            self.impulse = self.impulse_duration
        f = np.zeros(6)
        if self.impulse > 0:
            self.impulse -= 1
            f[0:len(self.force)] = self.force      # :199 -- a 6-vector carries a world-frame torque (driver :117-119)
        self.data.xfrc_applied = f.copy()
        self.f = f
        self.frame_counter += 1

    def _xmat(self):
        return self.data.xmat.reshape(3, 3)

    # -- :208-494
    def step(self, action):
        d, m = self.data, self.model
        action = np.asarray(action)
        _, ffSRB_ori = global2forward_facing(self._xmat().copy(), d.qpos[:3].copy(), d.qpos[:3].copy())
        SRB_ori = self._xmat()
        ffSRB_origin = d.qpos[:3].copy()
        ffSRB_origin[2] = 0

        desired_com_vel_global = np.concatenate([SRB_ori @ action[4:7], action[7:10]])
        delta = action[10:16]
        target_pos_com_global = SRB_ori @ delta[:3].copy() + d.qpos[:3]
        aa = delta[3:].copy()
        if np.linalg.norm(aa) < 1e-6:
            target_rot_com_local = np.eye(3)
        else:
            target_rot_com_local = axis_angle_to_rotation_matrix(aa)
        target_rot_com_global = SRB_ori @ target_rot_com_local
        next_T = transformation_matrix(target_rot_com_global, target_pos_com_global)
        target_vel = desired_com_vel_global

        self.next_contact_one_hot = action[18:22]
        next_contact_signal = int(np.argmax(self.next_contact_one_hot))
        if self.variant == "training" and self.is_stand:
            next_contact_signal = 3
        contact_prob = self.contact_prob if self.variant in ("test", "terrain") else [None, None]
        last_left_foot_in_contact = self.left_foot_in_contact
        last_right_foot_in_contact = self.right_foot_in_contact
        if contact_prob[0] is None:
            left_now = next_contact_signal in (1, 3)
            right_now = next_contact_signal in (2, 3)
        else:
            confidence = 0.7
            if self.left_foot_in_contact and contact_prob[0] < (1.0 - confidence):
                left_now = False
            elif not self.left_foot_in_contact and contact_prob[0] > confidence:
                left_now = True
            else:
                left_now = self.left_foot_in_contact
            if self.right_foot_in_contact and contact_prob[1] < (1.0 - confidence):
                right_now = False
            elif not self.right_foot_in_contact and contact_prob[1] > confidence:
                right_now = True
            else:
                right_now = self.right_foot_in_contact
        self.left_foot_in_contact = bool(left_now)
        self.right_foot_in_contact = bool(right_now)

        if self.variant == "terrain":     # ..._terrain_window.py:320-411
            if self.contactFilter[0] is None:
                self.initFilter(0, action[16], action[0:2], ffSRB_ori, ffSRB_origin)
                self.initFilter(1, action[17], action[2:4], ffSRB_ori, ffSRB_origin)
            for idx in (0, 1):
                in_contact = self.left_foot_in_contact if idx == 0 else self.right_foot_in_contact
                last = last_left_foot_in_contact if idx == 0 else last_right_foot_in_contact
                if not in_contact:
                    self.filterFoot(idx, action[16 + idx], action[2 * idx:2 * idx + 2], ffSRB_ori, ffSRB_origin)
                    pos = self.left_foot_contact_pos_global if idx == 0 else self.right_foot_contact_pos_global
                    h, _ = self.terrain.height_ray(pos[0], pos[1])
                    pos[2] = h
                    # Q9: left yaw = EMA filter, right yaw = raw action
                    yaw = self.contactFilter[0] if idx == 0 else action[17]
                    ori = ffSRB_ori @ z_axis_rotation_matrix(yaw)
                    pred = pos.copy(); pred[2] = np.nan
                else:
                    ori = self.left_foot_contact_ori if idx == 0 else self.right_foot_contact_ori
                    self.updateFilter(idx, ori, ffSRB_ori)
                    pos = self.left_foot_contact_pos_global if idx == 0 else self.right_foot_contact_pos_global
                    h, geom_id = self.terrain.height_ray(pos[0], pos[1])
                    box = self.terrain.box_by_geom(geom_id)
                    body = self.terrain.body(box["body_id"]) if box is not None else None
                    if body is not None and body["snap"] and not last:      # pyramid edge snap (:346-380, :407-442)
                        foot_xy = pos[:2]
                        gs = box["size"]; ngs = self.terrain.box_by_geom(geom_id + 1)["size"]
                        center = body["xpos"]
                        half = (gs[:2] + ngs[:2]) / 2
                        x_min, x_max = center[0] - half[0], center[0] + half[0]
                        y_min, y_max = center[1] - half[1], center[1] + half[1]
                        dx_min, dx_max = abs(foot_xy[0] - x_min), abs(foot_xy[0] - x_max)
                        dy_min, dy_max = abs(foot_xy[1] - y_min), abs(foot_xy[1] - y_max)
                        cx = np.clip(foot_xy[0], x_min, x_max); cy = np.clip(foot_xy[1], y_min, y_max)
                        md = min(dx_min, dx_max, dy_min, dy_max)
                        if md == dx_min:
                            proj = np.array([x_min, cy])
                        elif md == dx_max:
                            proj = np.array([x_max, cy])
                        elif md == dy_min:
                            proj = np.array([cx, y_min])
                        else:
                            proj = np.array([cx, y_max])
                        pos[:2] = proj
                    pred = pos.copy(); pred[2] = 0
                if idx == 0:
                    self.left_foot_contact_ori, self.next_predicted_left_foot_pos = ori, pred
                else:
                    self.right_foot_contact_ori, self.next_predicted_right_foot_pos = ori, pred
        elif self.variant == "test":
            if self.contactFilter[0] is None:
                self.initFilter(0, action[16], action[0:2], ffSRB_ori, ffSRB_origin)
                self.initFilter(1, action[17], action[2:4], ffSRB_ori, ffSRB_origin)
            if not self.left_foot_in_contact:
                self.filterFoot(0, action[16], action[0:2], ffSRB_ori, ffSRB_origin)
                self.left_foot_contact_ori = ffSRB_ori @ z_axis_rotation_matrix(self.contactFilter[0])
                self.next_predicted_left_foot_pos = self.left_foot_contact_pos_global.copy()
                self.next_predicted_left_foot_pos[2] = np.nan
            else:
                self.updateFilter(0, self.left_foot_contact_ori, ffSRB_ori)
                self.next_predicted_left_foot_pos = self.left_foot_contact_pos_global.copy()
                self.next_predicted_left_foot_pos[2] = 0
            if not self.right_foot_in_contact:
                self.filterFoot(1, action[17], action[2:4], ffSRB_ori, ffSRB_origin)
                self.right_foot_contact_ori = ffSRB_ori @ z_axis_rotation_matrix(self.contactFilter[1])
                self.next_predicted_right_foot_pos = self.right_foot_contact_pos_global.copy()
                self.next_predicted_right_foot_pos[2] = np.nan
            else:
                self.updateFilter(1, self.right_foot_contact_ori, ffSRB_ori)
                self.next_predicted_right_foot_pos = self.right_foot_contact_pos_global.copy()
                self.next_predicted_right_foot_pos[2] = 0
        else:   # ..._training.py:231-258
            if not self.left_foot_in_contact:
                ff = np.concatenate([action[0:2], np.array([0])])
                self.left_foot_contact_pos_global = ffSRB_ori @ ff + ffSRB_origin
                self.left_foot_contact_ori = ffSRB_ori @ z_axis_rotation_matrix(action[16])
                self.next_predicted_left_foot_pos = self.left_foot_contact_pos_global.copy()
                self.next_predicted_left_foot_pos[2] = np.nan
            else:
                self.next_predicted_left_foot_pos = self.left_foot_contact_pos_global.copy()
                self.next_predicted_left_foot_pos[2] = 0
            if not self.right_foot_in_contact:
                ff = np.concatenate([action[2:4], np.array([0])])
                self.right_foot_contact_pos_global = ffSRB_ori @ ff + ffSRB_origin
                self.right_foot_contact_ori = ffSRB_ori @ z_axis_rotation_matrix(action[17])
                self.next_predicted_right_foot_pos = self.right_foot_contact_pos_global.copy()
                self.next_predicted_right_foot_pos[2] = np.nan
            else:
                self.next_predicted_right_foot_pos = self.right_foot_contact_pos_global.copy()
                self.next_predicted_right_foot_pos[2] = 0

        next_predicted_feet_pos = np.stack((self.next_predicted_left_foot_pos, self.next_predicted_right_foot_pos))
        self.site_to_apply_force_on = return_site_to_apply_force_on(next_predicted_feet_pos)
        des_l = self.left_foot_contact_pos_global
        des_r = self.right_foot_contact_pos_global
        curr_feet_ori = np.stack((self.left_foot_contact_ori, self.right_foot_contact_ori))
        self.contact_forces_total_list = []
        self.qdd_qp_list = []
        for _ in range(round(self.mocap_dt / m.timestep)):
            d.qfrc_applied = np.zeros(6)
            mj.mj_step1(m, d)
            self.xquat_lagged = d.qpos[3:7].copy()        # the quaternion data.xmat now corresponds to (Q14)
            if self.variant in ("test", "terrain"):
                self.update_impulse()
            set_feet_site_this_step_env(d, next_predicted_feet_pos, curr_feet_ori, des_l, des_r)
            if not self.left_foot_in_contact and not self.right_foot_in_contact:
                self.qdd_qp_list.append(None)
            else:
                qdd, _, self.contact_forces_list = solve_qp_for_contact_forces(
                    m, d, next_T, target_vel, self.site_to_apply_force_on, self.mu)
                self.qdd_qp_list.append(qdd)
                self.contact_forces_total_list.append(self.contact_forces_list)
                if self.variant == "test":      # :402-443 force clamp (Q3)
                    forceSum = [np.zeros(3), np.zeros(3)]
                    j = 0
                    for site_id in self.site_to_apply_force_on:
                        if site_id == -1:
                            continue
                        cf = self.contact_forces_list[j]
                        j += 1
                        if site_id < 5:
                            forceSum[0] = forceSum[0] + cf
                        else:
                            forceSum[1] = forceSum[1] + cf
                    forceSum[0] = np.sqrt(forceSum[0].dot(forceSum[0]))
                    forceSum[1] = np.sqrt(forceSum[1].dot(forceSum[1]))
                j = 0
                for site_id in self.site_to_apply_force_on:
                    if site_id == -1:
                        continue
                    cf = self.contact_forces_list[j]
                    if self.variant == "test":
                        thr = 10000.0
                        forceLen = forceSum[0]
                        if site_id > 5:
                            forceLen = forceSum[1]
                        if forceLen > thr:
                            cf *= (thr / forceLen)
                    mj.mj_applyFT(m, d, cf, np.zeros(3), d.site_xpos[site_id], d.qfrc_applied)
                    j += 1
            mj.mj_step2(m, d)
        set_feet_site_this_step_env(d, next_predicted_feet_pos, curr_feet_ori, des_l, des_r)
        self.SRB_pos_list.append(d.qpos[:3].copy())
        self.SRB_ori_list.append(self._xmat().copy())            # stale by one sub-step (Q14)
        self.SRB_quat_list.append(self.xquat_lagged.copy())
        self.simulated_feet_pos_list.append([d.site_xpos[s].copy() for s in self.site_id_list])
        self.simulated_feet_ori_list.append([d.site_xmat[s].copy().reshape(3, 3) for s in self.site_id_list])
        self.contact_signal_list.append(next_contact_signal)
        self.action = action
        if self.variant in ("test", "terrain"):
            observation = self._get_obs()
            reward, reward_info = self._get_reward()
            done = self._get_done()
            self.current_frame += 1
        else:
            reward, reward_info = self._get_reward()
            done = self._get_done()
            self.current_frame += 1
            observation = self._get_obs()
        return observation, reward, done, False, {"reward_info": reward_info}

    # -- history accessors: test variant indexes from the end (capped list), training by frame
    def _hist(self, lst, back):
        """entry of step (current_frame - back): test: lst[-2-back]; training: lst[current_frame-back]."""
        return lst[len(lst) - 2 - back]

    # -- :496-803
    def _get_reward(self):
        if self.variant == "terrain":       # ..._terrain_window.py:525-526
            return 0, {}
        d = self.data
        cf, rf = self.current_frame, self.random_initial_frame
        curr_ref_pos = self.mocap_pos_com[cf + rf].copy()
        curr_ref_ori = self.mocap_ori_com[cf + rf]
        next_ref_pos = self.mocap_pos_com[cf + 1 + rf].copy()
        next_ref_ori = self.mocap_ori_com[cf + 1 + rf]
        xmat = self._xmat()
        _, curr_proj = global2projected(xmat, d.qpos[:3].copy(), d.qpos[:3].copy())
        curr_proj_origin = d.qpos[:3].copy()
        curr_proj_origin[2] = 0
        _, last_proj = global2projected(self._hist(self.SRB_ori_list, 0), self._hist(self.SRB_pos_list, 0), self._hist(self.SRB_pos_list, 0))
        if cf >= 15:
            _, proj_p15 = global2projected(self._hist(self.SRB_ori_list, 15), self._hist(self.SRB_pos_list, 15), self._hist(self.SRB_pos_list, 15))
            _, mproj_p15 = global2projected(self.mocap_ori_com[cf + rf - 15], self.mocap_pos_com[cf + rf - 15].copy(), self.mocap_pos_com[cf + rf - 15].copy())
        if cf >= 5:
            _, proj_p5 = global2projected(self._hist(self.SRB_ori_list, 5), self._hist(self.SRB_pos_list, 5), self._hist(self.SRB_pos_list, 5))
            _, mproj_p5 = global2projected(self.mocap_ori_com[cf + rf - 5], self.mocap_pos_com[cf + rf - 5].copy(), self.mocap_pos_com[cf + rf - 5].copy())
        _, curr_mproj = global2projected(curr_ref_ori, curr_ref_pos, curr_ref_pos)
        _, next_mproj = global2projected(next_ref_ori, next_ref_pos, next_ref_pos)
        next_mproj_origin = next_ref_pos.copy()
        next_mproj_origin[2] = 0

        w_s, w_m, w_p, w_e, w_p1, w_p2, w_p3, w_p4, w_c = WEIGHTS[self.variant]
        r_s_t = 1
        r_c_t = 0
        nl = self.feet_pos_list[cf + 1 + rf, 0, :]
        nr = self.feet_pos_list[cf + 1 + rf, 1, :]
        mocap_next_contact = (1 if nl[2] == 0 else 0) + (2 if nr[2] == 0 else 0)
        nl_xy = nl.copy(); nl_xy[2] = 0
        nr_xy = nr.copy(); nr_xy[2] = 0
        pred = int(np.argmax(self.next_contact_one_hot))
        r_c_t += 0 if pred == mocap_next_contact else 10

        r_e = 0
        lw = 1.9 if pred in (1, 3) else 0.1
        rw = 1.9 if pred in (2, 3) else 0.1
        r_e += lw * np.linalg.norm(curr_proj.T @ (d.site_xpos[0] - curr_proj_origin) - next_mproj.T @ (nl_xy - next_mproj_origin)) ** 2
        r_e += rw * np.linalg.norm(curr_proj.T @ (d.site_xpos[5] - curr_proj_origin) - next_mproj.T @ (nr_xy - next_mproj_origin)) ** 2
        f6 = lambda Mx: Mx[:, :2].flatten(order='F')
        r_e += np.linalg.norm(f6(curr_proj.T @ d.site_xmat[0].copy().reshape(3, 3)) - f6(next_mproj.T @ self.feet_ori_list[cf + 1 + rf][0]), ord=1)
        r_e += np.linalg.norm(f6(curr_proj.T @ d.site_xmat[5].copy().reshape(3, 3)) - f6(next_mproj.T @ self.feet_ori_list[cf + 1 + rf][1]), ord=1)

        dp = 0
        dR = 0
        if cf >= 15:
            dp += np.linalg.norm(proj_p15.T @ (d.qpos[:3].copy() - self._hist(self.SRB_pos_list, 15)) - mproj_p15.T @ (next_ref_pos - self.mocap_pos_com[cf + rf - 15]))
        if cf >= 5:
            dp += np.linalg.norm(proj_p5.T @ (d.qpos[:3].copy() - self._hist(self.SRB_pos_list, 5)) - mproj_p5.T @ (next_ref_pos - self.mocap_pos_com[cf + rf - 5].copy()))
        dp += np.linalg.norm(last_proj.T @ (d.qpos[:3].copy() - self._hist(self.SRB_pos_list, 0)) - curr_mproj.T @ (next_ref_pos - curr_ref_pos))
        if cf >= 15:
            dR += np.linalg.norm(f6(self._hist(self.SRB_ori_list, 15).T @ xmat) - f6(self.mocap_ori_com[cf + rf - 15].T @ next_ref_ori), ord=1)
        if cf >= 5:
            dR += np.linalg.norm(f6(proj_p5.T @ xmat) - f6(mproj_p5.T @ next_ref_ori), ord=1)
        dR += np.linalg.norm(f6(last_proj.T @ xmat.copy()) - f6(curr_mproj.T @ next_ref_ori), ord=1)

        p_term = np.linalg.norm(curr_proj.T @ (d.qpos[:3].copy() - curr_proj_origin) - next_mproj.T @ (next_ref_pos - next_mproj_origin))
        R_term = 1 * np.linalg.norm(f6(curr_proj.T @ xmat.copy()) - f6(next_mproj.T @ next_ref_ori), ord=1)

        r_p_t = w_p1 * dp + w_p2 * dR + w_p3 * p_term + w_p4 * R_term
        r_m_t = w_p * r_p_t + w_e * r_e + w_c * r_c_t
        r_t = w_s * r_s_t - w_m * r_m_t
        info = {
            'reward_survive': w_s * r_s_t,
            'reward_contact': -w_m * w_c * r_c_t,
            'reward_end_effector': -w_m * w_e * r_e,
            'reward_delta_pos': -w_m * w_p * w_p1 * dp,
            'reward_delta_ori': -w_m * w_p * w_p2 * dR,
            'reward_pos': -w_m * w_p * w_p3 * p_term,
            'reward_ori': -w_m * w_p * w_p4 * R_term,
            'reward_total': r_t,
        }
        return r_t, info

    # -- :805-845 ; the RNG draws are made by the caller (explicit reset) so that the CUDA path can be
    #    fed identical values; `reset()` below draws them with the same distributions as the reference.
    def _bind_clip(self, clip_id):
        c = self.clips[clip_id]
        self.clip_id = clip_id
        self.mocap_dt = c["mocap_dt"]
        self.cycle_length = c["cycle_length"]
        self.mocap_qpos = c["mocap_qpos"]
        self.mocap_pos_com = c["mocap_pos_com"]
        self.mocap_ori_com = c["mocap_ori_com"]
        self.mocap_vel_com = c["mocap_vel_com"]
        self.feet_pos_list = c["feet_pos_list"]
        self.feet_ori_list = c["feet_ori_list"]
        self.is_stand = bool(c.get("is_stand", False))
        if self.variant == "terrain":
            self.feet_pos_list_original = c["feet_pos_list_original"]
            self.mocap_height_com_terrain = c["mocap_height_com_terrain"]

    def reset_explicit(self, clip_id, initial_frame, vel_noise=None, noise_quat=None):
        """`vel_noise` (3,), `noise_quat` (4,) are the values np.random would have produced at :860-867;
        None for both reproduces testing_mode (no noise)."""
        self.contactFilter = [None] * 4
        self._bind_clip(clip_id)
        self.random_initial_frame = int(initial_frame)
        self.current_frame = 0
        self.set_state(self.random_initial_frame, vel_noise, noise_quat)
        return self._get_obs(), {}

    def reset(self, rng_py, rng_np, mocap_id=0):
        if not self.testing_mode:
            clip = rng_py.randrange(len(self.clips))
            cyc = self.clips[clip]["cycle_length"]
            frame = rng_py.randint(0, cyc - 2) if self.variant == "terrain" else rng_py.randint(0, max(cyc - 2, 512))   # Q12
            vel_noise = rng_np.uniform(low=-5e-2, high=5e-2, size=3)
            axis = rng_np.standard_normal(3)
            axis /= np.linalg.norm(axis)
            angle = rng_np.uniform(-5e-2, 5e-2)
            nq = np.array([np.cos(angle / 2.0), *(axis * np.sin(angle / 2.0))])
            return self.reset_explicit(clip, frame, vel_noise, nq)
        return self.reset_explicit(mocap_id, 0, None, None)

    # -- :847-911
    def set_state(self, initial_frame, vel_noise, noise_quat):
        d, m = self.data, self.model
        self.SRB_pos_list, self.SRB_ori_list, self.SRB_quat_list = [], [], []
        self.simulated_feet_pos_list, self.simulated_feet_ori_list = [], []
        self.contact_signal_list = []
        if vel_noise is not None:
            vn = np.zeros(6)
            vn[:3] = vel_noise
            pq = quaternion_multiply(noise_quat, self.mocap_qpos[initial_frame][3:7].copy())
            pq /= np.linalg.norm(pq)
            d.qpos = np.concatenate([self.mocap_qpos[initial_frame][:3], pq])
            d.qvel = self.mocap_vel_com[initial_frame] + vn
        else:
            d.qpos = self.mocap_qpos[initial_frame].copy()
            d.qvel = self.mocap_vel_com[initial_frame].copy()
        feet_pos = self.feet_pos_list[initial_frame]
        feet_ori = self.feet_ori_list[initial_frame]
        if self.variant == "terrain":       # :429-443: flags from the original (planar) z, positions from the lifted table
            feet_xy = feet_pos                                   # views into the table, as in the reference
            feet_pos = self.feet_pos_list_original[initial_frame]
        else:
            feet_xy = feet_pos.copy()
            feet_xy[:, 2] = 0
        self.left_foot_contact_ori = feet_ori[0]        # views into the clip table (mutated by Q11)
        self.right_foot_contact_ori = feet_ori[1]
        self.site_to_apply_force_on = return_site_to_apply_force_on(feet_pos)
        self.left_foot_contact_pos_global = feet_xy[0]
        self.right_foot_contact_pos_global = feet_xy[1]
        self.left_foot_in_contact = not all(s == -1 for s in self.site_to_apply_force_on[:3])     # Q2
        self.right_foot_in_contact = not all(s == -1 for s in self.site_to_apply_force_on[3:])    # Q2
        d.xfrc_applied = np.zeros(6)
        d.qfrc_applied = np.zeros(6)
        mj.mj_forward(m, d)
        self.xquat_lagged = d.qpos[3:7].copy()
        set_feet_site_this_step_env(d, feet_pos, feet_ori, self.left_foot_contact_pos_global, self.right_foot_contact_pos_global)
        self.SRB_pos_list.append(d.qpos[:3].copy())
        self.SRB_ori_list.append(self._xmat().copy())
        self.SRB_quat_list.append(self.xquat_lagged.copy())
        self.simulated_feet_pos_list.append([d.site_xpos[s].copy() for s in self.site_id_list])
        self.simulated_feet_ori_list.append([d.site_xmat[s].copy().reshape(3, 3) for s in self.site_id_list])
        sig = (1 if self.left_foot_in_contact else 0) + (2 if self.right_foot_in_contact else 0)
        self.contact_signal_list.append(sig)

    def _get_obs(self):
        return self.get_obs_window(self.window_size)

    # -- :916-1003
    def _get_done(self):
        d = self.data
        cf, rf = self.current_frame, self.random_initial_frame
        height = d.qpos[2]
        ref_height = self.mocap_pos_com[cf + 1 + rf, 2].copy()
        srb_ori = self._xmat()
        srb_vertical_axis = srb_ori[:, 2]
        gz = np.array([0, 0, 1])
        cos_angle = np.dot(srb_vertical_axis, gz) / (np.linalg.norm(srb_vertical_axis) * np.linalg.norm(gz))
        angle_degrees = np.degrees(np.arccos(np.clip(cos_angle, -1.0, 1.0)))
        done = False
        if ref_height - height > 0.2 or height - ref_height > 0.2 or height < 0.50 or height > 1.2:
            done = True
        elif angle_degrees > 65:
            done = True
        elif self.current_frame >= 512 - 1:
            done = True
        return done

    # -- :1055-1168
    def get_obs_window(self, window_size):
        d = self.data
        cf, rf = self.current_frame, self.random_initial_frame
        least_episode_length = max(self.cycle_length - 2, 512)
        assert cf + window_size < least_episode_length * 2
        xmat = self._xmat()
        SRB_pos_proj, projSRB_ori = global2projected(xmat, d.qpos[:3].copy(), d.qpos[:3].copy())
        if self.variant == "terrain":
            curr_height_terrain, _ = self.terrain.height_ray(d.qpos[0], d.qpos[1])
            SRB_pos_proj[2] -= curr_height_terrain
        height_proj = np.array([SRB_pos_proj[2]])
        world_mat = xmat.copy()
        projSRB_mat_6D = (projSRB_ori.T @ world_mat)[:, :2].flatten(order='F')
        local_vel_lin = projSRB_ori.T @ d.qvel[:3].copy()
        local_vel_ang = projSRB_ori.T @ (world_mat @ d.qvel[3:].copy())
        left_ff, ffSRB_ori = global2forward_facing(xmat, d.qpos[:3].copy(), d.site_xpos[0])
        left_mat_ff = ffSRB_ori.T @ d.site_xmat[0].reshape(3, 3)
        left_ang = np.array([np.arctan2(left_mat_ff[1, 0], left_mat_ff[0, 0])])
        right_ff, _ = global2forward_facing(xmat, d.qpos[:3].copy(), d.site_xpos[5])
        right_mat_ff = ffSRB_ori.T @ d.site_xmat[5].reshape(3, 3)
        right_ang = np.array([np.arctan2(right_mat_ff[1, 0], left_mat_ff[0, 0])])          # Q1
        contact_signal = foot_contact_to_one_hot([self.left_foot_in_contact, self.right_foot_in_contact])
        window_obs = []
        for wid in range(1, window_size + 1):
            window_obs.append(reference_frame_features(self, cf + wid + rf))
        window_mocap_obs = np.concatenate(window_obs, axis=-1)
        return np.concatenate([height_proj, projSRB_mat_6D, local_vel_lin, local_vel_ang, left_ff, left_ang,
                               right_ff, right_ang, contact_signal, window_mocap_obs], dtype=np.float32)

    # -- :1209-1299 (global position filter branch, useLocalPosFilter=False)
    def initFilter(self, index, action_ori, action_pos, ffSRB_ori, ffSRB_origin):
        self.contactFilter[index] = action_ori
        g = ffSRB_ori @ np.concatenate([action_pos, np.array([0])]) + ffSRB_origin
        Q = math.pow(10, 9)
        ballX = SDS(60.0, 0.001, 1, Q, 0, 1 / 120.0, True)
        ballY = SDS(60.0, 0.001, 1, Q, 0, 1 / 120.0, True)
        ballX.x[0] = g[0]; ballX.xd[0] = g[0]
        ballY.x[0] = g[1]; ballY.xd[0] = g[1]
        self.contactFilter[index + 2] = [ballX, ballY]

    def filterFoot(self, index, action_ori, action_pos, ffSRB_ori, ffSRB_origin):
        a = 0.5
        self.contactFilter[index] = self.contactFilter[index] * a + action_ori * (1.0 - a)
        g = ffSRB_ori @ np.concatenate([action_pos, np.array([0])]) + ffSRB_origin
        ballX, ballY = self.contactFilter[index + 2]
        ballX.xd[0] = g[0]
        ballY.xd[0] = g[1]
        v = self.data.qvel[:2]
        speed = math.sqrt(v.dot(v))
        if speed < 3:
            logQ = 9
        elif speed > 6:
            logQ = 11
        else:
            logQ = 9.0 + ((speed - 3.0) / 3.0) * (11.0 - 9.0)
        logQ = max(self.minLogQ, logQ)
        assert 5 <= logQ <= 11.000001
        Q = math.pow(10, logQ)
        ballX.updateCoef(0.001, 1, Q)
        ballY.updateCoef(0.001, 1, Q)
        for _ in range(4):
            ballX.singleStep()
            ballY.singleStep()
        pos = np.array([ballX.x[0], ballY.x[0], 0.0])
        if index == 0:
            self.left_foot_contact_pos_global = pos
        else:
            self.right_foot_contact_pos_global = pos

    def updateFilter(self, index, foot_contact_ori, ffSRB_ori):
        self.contactFilter[index] = angle_between_vectors(foot_contact_ori[0, :], ffSRB_ori[0, :], np.array([0., 0., 1.0]))
        for b in self.contactFilter[index + 2]:
            b.x[1] = 0
            b.xd[1] = 0


def reference_frame_features(env, f):
    """The 25 look-ahead features of reference frame `f` (SRBEnv5_mocap_list_window.py:1096-1161).
    They depend on the clip tables only, never on the simulated state."""
    terrain = env.variant == "terrain"
    pos = env.mocap_pos_com[f].copy()
    ori = env.mocap_ori_com[f]
    linvel = env.mocap_vel_com[f, :3]
    angvel = env.mocap_vel_com[f, 3:]
    pos_proj, proj = global2projected(ori, pos, pos)
    if terrain:
        pos_proj[2] -= env.mocap_height_com_terrain[f]
    h = np.array([pos_proj[2]])
    mat6 = (proj.T @ ori)[:, :2].flatten(order='F')
    vlin = proj.T @ linvel
    vang = proj.T @ (ori @ angvel)
    out = [h, mat6, vlin, vang]
    contact = []
    ff = None
    for i in range(2):
        fp = env.feet_pos_list[f, i, :]
        fxy = fp.copy()
        if terrain:                         # ..._terrain_window.py:589-606
            fxy[2], _ = env.terrain.height_ray(fxy[0], fxy[1])
            contact.append(bool(env.feet_pos_list_original[f, i, 2] == 0))
        else:
            fxy[2] = 0
            contact.append(bool(fp[2] == 0))
        p_ff, ff = global2forward_facing(ori, pos, fxy)
        mat_ff = ff.T @ env.feet_ori_list[f][i]
        out += [p_ff, np.array([np.arctan2(mat_ff[1, 0], mat_ff[0, 0])])]
    out.append(foot_contact_to_one_hot(contact))
    return np.concatenate(out, dtype=np.float32)


def lift_clip_to_terrain(clip, terrain, initial_pos_planar=(0.0, 0.0)):
    """Terrain variant of the reference tables (testSRB_mujoco_original_viewer_terrain.py:1043-1065 applied
    after `SRBdata.mocap_data[:, :2] += initial_pos_planar`, SRBEnv5_mocap_list_terrain_window.py:636).
    The planar shift is applied to the finished flat tables: humanoid FK, CoM and the yaw-aligned cycle
    stitching are translation-equivariant, so this equals re-running humanoid2SRB on the shifted clip up to
    round-off."""
    out = dict(clip)
    sh = np.array([initial_pos_planar[0], initial_pos_planar[1], 0.0])
    qpos = np.array(clip["mocap_qpos"], dtype=float).copy()
    qpos[:, :3] += sh
    feet = np.array(clip["feet_pos_list"], dtype=float).copy()
    feet += sh
    original = feet.copy()                       # z == 0 marks reference contact
    feet[:, :, 2] = 0
    hts = []
    for i in range(qpos.shape[0]):
        hc, _ = terrain.height_ray(qpos[i, 0], qpos[i, 1])
        hl, _ = terrain.height_ray(feet[i, 0, 0], feet[i, 0, 1])
        hr, _ = terrain.height_ray(feet[i, 1, 0], feet[i, 1, 1])
        qpos[i, 2] += hc
        feet[i, 0, 2] += hl
        feet[i, 1, 2] += hr
        hts.append(hc)
    out["mocap_qpos"] = qpos
    out["mocap_pos_com"] = qpos[:, :3]
    out["feet_pos_list"] = feet
    out["feet_pos_list_original"] = original
    out["mocap_height_com_terrain"] = np.array(hts)
    out["initial_pos_planar"] = np.array(initial_pos_planar, dtype=float)
    return out
