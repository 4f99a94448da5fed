"""Shared test infrastructure: golden clip tables, oracle construction, action generators."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_clip(name):
    """tests/golden/clip_<name>.npz -> dict with the keys SRBEnv.loadMocap builds."""
    d = np.load(os.path.join(GOLDEN, f"clip_{name}.npz"))
    n = d["mocap_qpos"].shape[0]
    yaw = d["feet_yaw"]
    fo = np.zeros((n, 2, 3, 3))
    fo[:, :, 0, 0] = yaw[:, :, 0]; fo[:, :, 0, 1] = -yaw[:, :, 1]
    fo[:, :, 1, 0] = yaw[:, :, 1]; fo[:, :, 1, 1] = yaw[:, :, 0]
    fo[:, :, 2, 2] = 1.0
    return dict(mocap_dt=float(d["mocap_dt"]), cycle_length=int(d["cycle_length"]), mocap_qpos=d["mocap_qpos"].copy(),
                mocap_pos_com=d["mocap_qpos"][:, :3].copy(), mocap_ori_com=d["mocap_ori_com"].copy(),
                mocap_vel_com=d["mocap_vel_com"].copy(), feet_pos_list=d["feet_pos_list"].copy(),
                feet_ori_list=fo, feet_yaw=yaw.copy(), is_stand=bool(d["is_stand"]))


def fresh_clips(names=("stand1", "aiming_show")):
    return [load_clip(n) for n in names]


def log_so3(R):
    c = (np.trace(R) - 1) / 2
    a = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.linalg.norm(a)
    th = math.atan2(s, c)
    if th < 1e-9:
        return a
    return a * (th / s)


def tracking_action(env, rng, sigma, dtype=np.float32):
    """Reference-tracking action + N(0, sigma^2) noise from the oracle env's current state:
    feet / yaw / contact from the next reference frame, pose and velocity targets = next reference frame."""
    from oracle.srb_env import global2forward_facing
    f = env.current_frame + env.random_initial_frame + 1
    a = np.zeros(22)
    xm = env.data.xmat.reshape(3, 3)
    p = env.data.qpos[:3]
    o = p.copy(); o[2] = 0
    for i in range(2):
        fp = env.feet_pos_list[f, i].copy(); fp[2] = 0
        pff, ff = global2forward_facing(xm, o, fp)
        a[2 * i:2 * i + 2] = pff[:2]
        mff = ff.T @ env.feet_ori_list[f][i]
        a[16 + i] = math.atan2(mff[1, 0], mff[0, 0])
    Rr = env.mocap_ori_com[f]; pr = env.mocap_pos_com[f]; v = env.mocap_vel_com[f]
    a[4:7] = xm.T @ v[:3]; a[7:10] = v[3:]
    a[10:13] = xm.T @ (pr - p); a[13:16] = log_so3(xm.T @ Rr)
    c = (1 if env.feet_pos_list[f, 0, 2] == 0 else 0) + (2 if env.feet_pos_list[f, 1, 2] == 0 else 0)
    a[18 + c] = 1
    a += rng.normal(0, sigma, 22)
    return a.astype(dtype)


def random_plan(rng, clips, noise=True):
    clip = int(rng.integers(0, len(clips)))
    cyc = clips[clip]["cycle_length"]
    frame = int(rng.integers(0, max(cyc - 2, 512) + 1))
    if noise:
        vn = rng.uniform(-5e-2, 5e-2, 3)
        ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(-5e-2, 5e-2)
        q = np.array([math.cos(ang / 2), *(ax * math.sin(ang / 2))])
    else:
        vn = np.zeros(3); q = np.array([np.nan, 0, 0, 0])
    return np.concatenate([[clip, frame], vn, q])


def oracle_reset(env, plan):
    if np.isnan(plan[5]):
        return env.reset_explicit(int(plan[0]), int(plan[1]), None, None)
    return env.reset_explicit(int(plan[0]), int(plan[1]), plan[2:5].copy(), plan[5:9].copy())


def oracle_state_vector(env):
    """[qpos 7, qvel 6, left foot 3, right foot 3, yaw cs 4, flags 2, frame 1]"""
    d = env.data
    lo, ro = env.left_foot_contact_ori, env.right_foot_contact_ori
    return np.concatenate([d.qpos, d.qvel, env.left_foot_contact_pos_global, env.right_foot_contact_pos_global,
                           [lo[0, 0], lo[1, 0], ro[0, 0], ro[1, 0]],
                           [float(env.left_foot_in_contact), float(env.right_foot_in_contact)], [float(env.current_frame)]])


def cuda_state_vector(named):
    return np.concatenate([named["qpos"], named["qvel"], named["left_foot_contact_pos_global"], named["right_foot_contact_pos_global"],
                           named["left_foot_yaw_cs"], named["right_foot_yaw_cs"],
                           [float(named["left_foot_in_contact"]), float(named["right_foot_in_contact"])], [float(named["current_frame"])]])


def oracle_to_cuda_state(env, n_real=46, n_int=7, slots=32, width=8):
    """Pack an OracleSRBEnv's full state in the layout of srb_get_state / srb_set_state
    (ipcdnnwalk_b200/csrc/srb_types.h)."""
    d = env.data
    r = np.zeros(n_real)
    r[0:7] = d.qpos; r[7:13] = d.qvel; r[13:17] = env.xquat_lagged
    r[17:20] = env.left_foot_contact_pos_global; r[20:23] = env.right_foot_contact_pos_global
    lo, ro = env.left_foot_contact_ori, env.right_foot_contact_ori
    r[23:27] = [lo[0, 0], lo[1, 0], ro[0, 0], ro[1, 0]]
    inited = env.variant in ("test", "terrain") and env.contactFilter[0] is not None
    if inited:
        r[27:29] = [float(env.contactFilter[0]), float(env.contactFilter[1])]
        balls = [env.contactFilter[2][0], env.contactFilter[2][1], env.contactFilter[3][0], env.contactFilter[3][1]]
        r[29:33] = [b.x[0] for b in balls]; r[33:37] = [b.x[1] for b in balls]
    k = np.zeros(n_int, dtype=np.int32)
    k[0] = (1 if env.left_foot_in_contact else 0) | (2 if env.right_foot_in_contact else 0) | (4 if inited else 0) | (8 if env.is_stand else 0)
    k[1] = env.current_frame; k[2] = env.random_initial_frame; k[3] = env.clip_id
    k[4] = env.frame_counter; k[5] = env.impulse; k[6] = 1
    h = np.zeros((slots, width))
    n = len(env.SRB_pos_list)
    for e in range(max(0, n - slots), n):
        h[e % slots, 0:3] = env.SRB_pos_list[e]; h[e % slots, 3:7] = env.SRB_quat_list[e]
    return dict(reals=r, ints=k, hist=h)
