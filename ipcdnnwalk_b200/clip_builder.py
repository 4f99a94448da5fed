"""Mocap clip -> SRB reference tables on the device.

Reference: `SRBdata.humanoid2SRB` / `get_frame_data` (env/testSRB_mujoco_original_viewer.py:820-1031): for every frame
of the doubled cycle, MuJoCo FK of the 41-dof humanoid -> whole-body CoM (data.subtree_com[0]), root orientation,
foot-geom position and yaw (`get_feet_state_forward`, :342-404), finite-difference CoM velocity and
`mj_differentiatePos` root angular velocity, the yaw-aligned stitching of the second cycle (:835-868) and the contact
mask of `read_contact_npz` (:750-762, :995-1012).

The per-frame FK + CoM is the batched CUDA kernel of bone_fk (the MJCF humanoid is a hinge tree with Euler channels);
the remaining per-frame closed forms are a few torch tensor expressions on the same device (one-time preprocessing,
not on the step path).  Output: the dict `SRBEnv.loadMocap` builds, ready for `SRBVecEnv` / `srb_upload_clip`."""
import math

import numpy as np

from . import mjcf_humanoid
from .bone_fk import Skeleton


def _yaw_of_quat_rxyz(q):
    """yaw of project_to_forward_facing (viewer.py:612-636): third 'rxyz' Euler angle = -atan2(R01, R00)."""
    w, x, y, z = q
    r00 = w * w + x * x - y * y - z * z
    r01 = 2 * (x * y - w * z)
    n = w * w + x * x + y * y + z * z          # transformations.quaternion_matrix normalises by |q|^2
    return -math.atan2(r01 / n, r00 / n)


class ClipBuilder:
    def __init__(self, humanoid=None, device=0, dt=0.033333):
        import torch
        self.torch = torch
        h = dict(humanoid or mjcf_humanoid.builtin())
        h["offset"] = [list(o) for o in h["offset"]]
        h["offset"][0] = [0.0, 0.0, 0.0]        # a free joint's qpos replaces the body's model position
        self.h = h
        self.skel = Skeleton(h, device=device)
        self.device = self.skel.device
        self.dt = dt
        sign = [1.0] * 7
        for s in h["sign"]:
            sign += [float(v) for v in s]
        self.sign = torch.tensor(sign, dtype=torch.float64, device=self.device)

    def build(self, mocap_qpos, left_contact, right_contact, initial_frame=0, last_frame=None, is_stand=False):
        """mocap_qpos [n,41] (root xyz, root quaternion wxyz, 34 hinge angles); left/right_contact: per-frame
        (heel or toe) contact flags of the clip (read_contact_npz)."""
        t = self.torch
        md = np.asarray(mocap_qpos, dtype=np.float64)
        if last_frame is None:
            last_frame = md.shape[0] - 1
        cyc = last_frame - initial_frame + 1
        n = 2 * cyc - 1
        idx = initial_frame + (np.arange(n) % cyc)
        poses = md[idx].copy()
        # ---- second cycle: continue from the last root pose of the first cycle, yaw-aligned (get_frame_data :835-868)
        prev, start = poses[cyc - 1].copy(), md[initial_frame]
        yaw_p, yaw_s = _yaw_of_quat_rxyz(prev[3:7]), _yaw_of_quat_rxyz(start[3:7])
        P = t.tensor(poses, dtype=t.float64, device=self.device)
        if n > cyc:
            sec = P[cyc:]
            cs, ss, cp, sp = math.cos(yaw_s), math.sin(yaw_s), math.cos(yaw_p), math.sin(yaw_p)
            dx, dy = sec[:, 0] - start[0], sec[:, 1] - start[1]
            lx, ly = cs * dx + ss * dy, -ss * dx + cs * dy                 # Rz(yaw_s)^T (pos - projected start)
            fx, fy = cp * lx - sp * ly + prev[0], sp * lx + cp * ly + prev[1]
            # rotation: Rz(yaw_p) Rz(yaw_s)^T R(q/|q|) as a quaternion, w >= 0 (quaternion_from_matrix convention)
            q = sec[:, 3:7] / sec[:, 3:7].norm(dim=1, keepdim=True)
            half = 0.5 * (yaw_p - yaw_s)
            c, s = math.cos(half), math.sin(half)
            w2 = c * q[:, 0] - s * q[:, 3]
            x2 = c * q[:, 1] - s * q[:, 2]
            y2 = c * q[:, 2] + s * q[:, 1]
            z2 = c * q[:, 3] + s * q[:, 0]
            q2 = t.stack([w2, x2, y2, z2], 1)
            q2 = t.where(q2[:, :1] < 0, -q2, q2)
            P[cyc:, 0] = fx; P[cyc:, 1] = fy; P[cyc:, 3:7] = q2
        # ---- batched FK + CoM (CUDA kernel): mj_kinematics normalises the root quaternion
        fkp = P.clone()
        qn = fkp[:, 3:7] / fkp[:, 3:7].norm(dim=1, keepdim=True)
        fkp[:, 3:7] = qn
        fkp = (fkp * self.sign).contiguous()
        out = self.skel.forward(fkp, None, want_xform=True, want_com=True, want_momentum=False)
        xf, com = out["xform"], out["com"]
        # root orientation matrix (mju_quat2Mat of the normalised quaternion)
        R = self._quat2mat(qn)
        # ---- feet: geom position + forward-looking yaw (get_feet_state_forward / get_z_rotation_matrix_forward)
        feet_pos, feet_yaw = [], []
        for g in ("left_foot", "right_foot"):
            b, lp = self.h["geoms"][g]
            Rb = self._quat2mat(xf[:, b, 0:4])
            lp_t = t.tensor(lp, dtype=t.float64, device=self.device)
            feet_pos.append(xf[:, b, 4:7] + (Rb @ lp_t))
            feet_yaw.append(self._foot_yaw(Rb))
        feet_pos = t.stack(feet_pos, 1)                                  # [n,2,3]
        feet_yaw = t.stack(feet_yaw, 1)                                  # [n,2] angle
        # ---- velocities (:925-954)
        lin = t.empty_like(com)
        lin[1:] = (com[1:] - com[:-1]) / self.dt
        lin[0] = lin[1]
        q1, q0 = P[1:, 3:7], P[:-1, 3:7]                                 # mj_differentiatePos takes the poses as given
        qd = self._qmul(t.stack([q0[:, 0], -q0[:, 1], -q0[:, 2], -q0[:, 3]], 1), q1)
        axis = qd[:, 1:4]
        sn = axis.norm(dim=1)
        speed = 2 * t.atan2(sn, qd[:, 0])
        speed = t.where(speed > math.pi, speed - 2 * math.pi, speed)
        unit = t.where(sn[:, None] < 1e-15, t.tensor([1.0, 0, 0], dtype=t.float64, device=self.device).expand_as(axis), axis / sn.clamp_min(1e-300)[:, None])
        ang = t.empty_like(com)
        ang[1:] = unit * (speed / 1.0)[:, None] / self.dt
        ang[0] = ang[1]
        # ---- contact mask (:995-1012): period cycle-1 (sic), zero the z of feet in contact
        lc = np.asarray(left_contact, dtype=np.float64)[initial_frame:last_frame]
        rc = np.asarray(right_contact, dtype=np.float64)[initial_frame:last_frame]
        while lc.shape[0] < n:
            lc = np.concatenate([lc, lc]); rc = np.concatenate([rc, rc])
        lc, rc = t.tensor(lc[:n], device=self.device), t.tensor(rc[:n], device=self.device)
        feet_pos[:, 0, 2] = feet_pos[:, 0, 2] * (1 - lc)
        feet_pos[:, 1, 2] = feet_pos[:, 1, 2] * (1 - rc)
        qpos = t.cat([com, P[:, 3:7]], 1)
        return dict(mocap_dt=self.dt, cycle_length=int(cyc), mocap_qpos=qpos.cpu().numpy(), mocap_pos_com=com.cpu().numpy(),
                    mocap_ori_com=R.cpu().numpy(), mocap_vel_com=t.cat([lin, ang], 1).cpu().numpy(),
                    feet_pos_list=feet_pos.cpu().numpy(),
                    feet_yaw=t.stack([t.cos(feet_yaw), t.sin(feet_yaw)], 2).cpu().numpy(), is_stand=bool(is_stand))

    # ---- small closed forms ---------------------------------------------------------------------------
    def _quat2mat(self, q):
        t = self.torch
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        return t.stack([t.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], 1),
                        t.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], 1),
                        t.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], 1)], 1)

    def _qmul(self, a, b):
        t = self.torch
        return t.stack([a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3],
                        a[:, 0] * b[:, 1] + a[:, 1] * b[:, 0] + a[:, 2] * b[:, 3] - a[:, 3] * b[:, 2],
                        a[:, 0] * b[:, 2] - a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0] + a[:, 3] * b[:, 1],
                        a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1] + a[:, 3] * b[:, 0]], 1)

    def _foot_yaw(self, Rb):
        """get_z_rotation_matrix_forward (viewer.py:354-404) -> yaw angle."""
        t = self.torch
        fx, fy = Rb[:, :, 0], Rb[:, :, 1]
        xp = fx.clone(); xp[:, 2] = 0
        xn = xp.norm(dim=1)
        xp = xp / xn[:, None]
        yp = fy.clone(); yp[:, 2] = 0
        yp = yp / yp.norm(dim=1, keepdim=True)
        use_y = (xn < 1e-6) | ((fx * xp).sum(1) < 0.3)
        z = t.tensor([0.0, 0.0, 1.0], dtype=t.float64, device=Rb.device).expand_as(yp)
        xnew = t.linalg.cross(yp, z)
        xnew = xnew / (xnew + 1e-8).norm(dim=1, keepdim=True)
        mix = t.where(use_y[:, None], 0.9 * xnew + 0.1 * xp, 0.7 * xnew + 0.3 * xp)
        return t.atan2(mix[:, 1], mix[:, 0])

    def close(self):
        self.skel.close()
