"""ORACLE (test infrastructure only).  CPU restatement of `SRBdata` -- the one-time mocap -> SRB
reference-table builder of the reference:

  env/testSRB_mujoco_original_viewer.py:342-404   get_feet_state_forward / get_z_rotation_matrix_forward
  env/testSRB_mujoco_original_viewer.py:612-636   project_to_forward_facing
  env/testSRB_mujoco_original_viewer.py:750-762   read_contact_npz
  env/testSRB_mujoco_original_viewer.py:765-1031  SRBdata.__init__/get_frame_data/humanoid2SRB

Third-party arithmetic restated from the published algorithms (packages absent / unpinned):
  * mujoco: mj_kinematics (joint composition in document order), mj_comPos (subtree_com of the
    world body = mass-weighted mean of body inertial frames, inertiafromgeom with explicit geom
    masses), mj_differentiatePos / mju_subQuat / mju_quat2Vel for the free joint.
  * transformations (Gohlke): quaternion_matrix, quaternion_from_matrix (isprecise=False, eigh),
    euler_from_matrix / euler_matrix with axes 'rxyz'.
PARITY UNPINNED (no golden vectors in the reference; mujoco/transformations not installable here).
"""
import math
import xml.etree.ElementTree as ET
import numpy as np

_EPS = np.finfo(float).eps * 4.0


# ----------------------------------------------------------------------------- transformations.py
def quaternion_matrix3(quaternion):
    q = np.array(quaternion, dtype=np.float64, copy=True)
    n = np.dot(q, q)
    if n < _EPS:
        return np.identity(3)
    q *= math.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([
        [1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0]],
        [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0]],
        [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2]]])


def quaternion_from_matrix3(M):
    m00, m01, m02 = M[0]
    m10, m11, m12 = M[1]
    m20, m21, m22 = M[2]
    K = np.array([[m00 - m11 - m22, 0.0, 0.0, 0.0],
                  [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
                  [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0],
                  [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22]])
    K /= 3.0
    w, V = np.linalg.eigh(K)
    q = V[[3, 0, 1, 2], np.argmax(w)]
    if q[0] < 0.0:
        np.negative(q, q)
    return q


def project_to_forward_facing(R):
    """euler_from_matrix(R,'rxyz')[2] -> euler_matrix(0,0,yaw,'rxyz') == Rz(yaw).
    'rxyz' = (firstaxis 2, parity 1, repetition 0, frame 1): i=2, j=1, k=0."""
    cy = math.sqrt(R[2, 2] * R[2, 2] + R[1, 2] * R[1, 2])
    if cy > _EPS:
        yaw = -math.atan2(R[0, 1], R[0, 0])
    else:
        yaw = -math.atan2(-R[1, 0], R[1, 1])
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


# ----------------------------------------------------------------------------- MuJoCo pieces
def _quat2mat(q):
    from .mj_freebody import quat2mat
    return quat2mat(q)


def _axis_rot(axis, angle):
    """rotation about a unit axis (mju_axisAngle2Quat + mju_quat2Mat)."""
    from .mj_freebody import quat2mat
    s = math.sin(angle * 0.5)
    return quat2mat(np.array([math.cos(angle * 0.5), axis[0] * s, axis[1] * s, axis[2] * s]))


def quat2vel(quat, dt):
    axis = np.array(quat[1:4], dtype=float)
    n = math.sqrt(axis.dot(axis))
    if n < 1e-15:
        axis = np.array([1.0, 0.0, 0.0]); n = 0.0
    else:
        axis = axis / n
    speed = 2 * math.atan2(n, quat[0])
    if speed > math.pi:
        speed -= 2 * math.pi
    return axis * (speed / dt)


def differentiate_pos_free(dt, qpos1, qpos2):
    from .mj_freebody import mul_quat
    lin = (qpos2[:3] - qpos1[:3]) / dt
    qb = qpos1[3:7]
    qneg = np.array([qb[0], -qb[1], -qb[2], -qb[3]])
    qdif = mul_quat(qneg, qpos2[3:7])
    ang = quat2vel(qdif, 1.0) / dt
    return lin, ang


class HumanoidModel:
    """Minimal MJCF reader for Characters/humanoid_deepmimic_withpalm_Tpose.xml (hinge joints at the
    body origin, one massive geom per body, sites)."""

    def __init__(self, xml_path):
        root = ET.parse(xml_path).getroot()
        self.bodies = []      # dict(name,parent,pos,joints=[(type,axis,qadr)],geoms=[(name,pos,mass)],sites=[(name,pos)])
        self.nq = 0
        wb = root.find("worldbody")
        for b in wb.findall("body"):
            self._add_body(b, -1)
        self.body_id = {b["name"]: i for i, b in enumerate(self.bodies)}

    @staticmethod
    def _vec(s, n=3):
        v = [float(x) for x in s.split()]
        assert len(v) == n
        return np.array(v)

    def _add_body(self, el, parent):
        b = dict(name=el.get("name"), parent=parent, pos=self._vec(el.get("pos", "0 0 0")), joints=[], geoms=[], sites=[])
        idx = len(self.bodies)
        self.bodies.append(b)
        for ch in el:
            if ch.tag == "joint":
                if ch.get("type") == "free":
                    b["joints"].append(("free", None, self.nq)); self.nq += 7
                else:
                    assert ch.get("type", "hinge") == "hinge" and self._vec(ch.get("pos", "0 0 0")).tolist() == [0, 0, 0]
                    ax = self._vec(ch.get("axis"))
                    b["joints"].append(("hinge", ax / np.linalg.norm(ax), self.nq)); self.nq += 1
            elif ch.tag == "geom":
                if ch.get("fromto") is not None:
                    ft = self._vec(ch.get("fromto"), 6)
                    gpos = 0.5 * (ft[:3] + ft[3:])
                else:
                    gpos = self._vec(ch.get("pos", "0 0 0"))
                b["geoms"].append((ch.get("name"), gpos, float(ch.get("mass"))))
            elif ch.tag == "site":
                b["sites"].append((ch.get("name"), self._vec(ch.get("pos", "0 0 0"))))
        for ch in el.findall("body"):
            self._add_body(ch, idx)

    def forward(self, qpos):
        """returns dict with xpos[nb,3], xmat[nb,3,3], com[3], geom/site world positions."""
        nb = len(self.bodies)
        xpos = np.zeros((nb, 3)); xmat = np.zeros((nb, 3, 3))
        for i, b in enumerate(self.bodies):
            if b["parent"] < 0:
                p, R = b["pos"].copy(), np.eye(3)
            else:
                p = xpos[b["parent"]] + xmat[b["parent"]] @ b["pos"]
                R = xmat[b["parent"]].copy()
            for (jt, ax, adr) in b["joints"]:
                if jt == "free":
                    p = np.array(qpos[adr:adr + 3], dtype=float)
                    q = np.array(qpos[adr + 3:adr + 7], dtype=float)
                    q = q / math.sqrt(q.dot(q))
                    R = _quat2mat(q)
                else:
                    R = R @ _axis_rot(ax, qpos[adr])
            xpos[i], xmat[i] = p, R
        msum = 0.0; com = np.zeros(3)
        geoms, sites = {}, {}
        for i, b in enumerate(self.bodies):
            for (gn, gp, gm) in b["geoms"]:
                w = xpos[i] + xmat[i] @ gp
                geoms[gn] = (w, xmat[i])
                com += gm * w; msum += gm
            for (sn, sp) in b["sites"]:
                sites[sn] = xpos[i] + xmat[i] @ sp
        return dict(xpos=xpos, xmat=xmat, com=com / msum, geoms=geoms, sites=sites)


def get_z_rotation_matrix_forward(xmat):
    """viewer.py:354-404"""
    foot_x = xmat[:, 0].copy()
    foot_y = xmat[:, 1].copy()
    x_proj = foot_x.copy(); x_proj[2] = 0
    x_proj_norm = np.linalg.norm(x_proj)
    x_proj /= x_proj_norm
    y_proj = foot_y.copy(); y_proj[2] = 0
    y_proj /= np.linalg.norm(y_proj)
    if x_proj_norm < 1e-6:
        use_y = True
    else:
        use_y = np.dot(foot_x, x_proj) < 0.3
    z = np.array([0, 0, 1])
    x_proj_new = np.cross(y_proj, z)
    x_proj_new /= np.linalg.norm(x_proj_new + 1e-8)
    if use_y:
        x_proj = 0.9 * x_proj_new + 0.1 * x_proj
    else:
        x_proj = 0.7 * x_proj_new + 0.3 * x_proj
    theta = np.arctan2(x_proj[1], x_proj[0])
    return np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]])


def read_contact_npz(path):
    d = np.load(path)
    n = len(d['leftHeel'])
    left = np.zeros(n); right = np.zeros(n)
    for i in range(n):
        if d['leftHeel'][i] or d['leftToe'][i]:
            left[i] = 1
        if d['rightHeel'][i] or d['rightToe'][i]:
            right[i] = 1
    return left, right


class SRBdata:
    def __init__(self, mocap_data, contact_npz_path, humanoid_xml):
        self.dt = 0.033333
        self.mocap_data = np.array(mocap_data, dtype=float)
        self.mocap_contact_path = contact_npz_path
        self.model = HumanoidModel(humanoid_xml)

    def get_frame_data(self, frame_idx, prev_cycle_root=None):
        cycle_frame = self.initial_frame + (frame_idx - self.initial_frame) % self.cycle_length
        current_pose = self.mocap_data[cycle_frame].copy()
        if prev_cycle_root is not None:
            prev_pos = prev_cycle_root[:3]
            projected_prev_pos = prev_pos.copy(); projected_prev_pos[2] = 0
            prev_mat = quaternion_matrix3(prev_cycle_root[3:7])
            start_pos = self.mocap_data[self.initial_frame][:3]
            projected_start_pos = start_pos.copy(); projected_start_pos[2] = 0
            start_mat = quaternion_matrix3(self.mocap_data[self.initial_frame][3:7])
            local_pos = project_to_forward_facing(start_mat).T @ (current_pose[:3] - projected_start_pos)
            final_pos = project_to_forward_facing(prev_mat) @ local_pos + projected_prev_pos
            current_mocap_mat = quaternion_matrix3(current_pose[3:7])
            local_mat = project_to_forward_facing(start_mat).T @ current_mocap_mat
            global_mat = project_to_forward_facing(prev_mat) @ local_mat
            current_pose[:3] = final_pos
            current_pose[3:7] = quaternion_from_matrix3(global_mat)
        return current_pose

    def humanoid2SRB(self, initial_frame=0, last_frame=None):
        self.initial_frame = initial_frame
        frame_idx = initial_frame
        if last_frame is None:
            last_frame = len(self.mocap_data) - 1
        self.last_frame = last_frame
        self.cycle_length = last_frame - initial_frame + 1
        prev_cycle_root = None
        root_id = self.model.body_id["root"]
        feet_pos, feet_ori, pos_com, poses, ori_com, lin_vel, ang_vel = [], [], [], [], [], [], []
        i = 0
        while True:
            pose = self.get_frame_data(frame_idx, prev_cycle_root)
            poses.append(pose)
            fk = self.model.forward(pose)
            com = fk["com"].copy()
            fp, fo = [], []
            for g in ("left_foot", "right_foot"):
                w, R = fk["geoms"][g]
                fp.append(w.copy()); fo.append(get_z_rotation_matrix_forward(R.copy()))
            if i == 0:
                lv = np.array([np.nan] * 3); av = np.array([np.nan] * 3)
            else:
                _, av = differentiate_pos_free(self.dt, poses[i - 1], poses[i])
                lv = (com - pos_com[i - 1]) / self.dt
            feet_pos.append(fp); feet_ori.append(fo); pos_com.append(com)
            lin_vel.append(lv); ang_vel.append(av); ori_com.append(fk["xmat"][root_id].copy())
            if (frame_idx - self.initial_frame) % self.cycle_length == self.cycle_length - 1:
                prev_cycle_root = pose.copy()
            frame_idx += 1
            i += 1
            if i >= self.cycle_length * 2 - 1:
                break
        self.mocap_pos_humanoid = np.array(poses)
        self.mocap_com_humanoid = np.hstack((np.array(pos_com), self.mocap_pos_humanoid[:, 3:7]))
        lin_vel[0] = lin_vel[1].copy(); ang_vel[0] = ang_vel[1].copy()
        self.mocap_vel_com = np.hstack((lin_vel, ang_vel))
        self.mocap_ori_com = ori_com
        self.feet_pos_list = np.array(feet_pos)
        self.feet_ori_list = feet_ori
        left, right = read_contact_npz(self.mocap_contact_path)
        left = left[self.initial_frame:self.last_frame].copy()
        right = right[self.initial_frame:self.last_frame].copy()
        n = self.mocap_pos_humanoid.shape[0]
        while True:
            left = np.concatenate([left, left]); right = np.concatenate([right, right])
            if left.shape[0] >= n:
                left, right = left[:n], right[:n]
                break
        self.left_foot_step, self.right_foot_step = left, right
        self.feet_pos_list[:, 0, 2] = self.feet_pos_list[:, 0, 2] * (1 - left)
        self.feet_pos_list[:, 1, 2] = self.feet_pos_list[:, 1, 2] * (1 - right)

    def tables(self, is_stand=False):
        """dict with the keys SRBEnv.loadMocap builds (SRBEnv5_mocap_list_window.py:1187-1197)."""
        return dict(mocap_dt=self.dt, cycle_length=self.cycle_length, mocap_qpos=self.mocap_com_humanoid,
                    mocap_pos_com=self.mocap_com_humanoid[:, :3], mocap_ori_com=self.mocap_ori_com,
                    mocap_vel_com=self.mocap_vel_com, feet_pos_list=self.feet_pos_list,
                    feet_ori_list=self.feet_ori_list, is_stand=is_stand)


def load_mocap_txt(path):
    with open(path, 'r') as f:
        return np.array(eval(f.read()))
