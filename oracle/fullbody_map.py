"""ORACLE (test infrastructure only).  CPU restatement of gym_trackSRB/taesooLib/DeltaExtractor.py:577-643
`_extractSingleFrameFullbody`: one decoded full-body observation row + the SRB state of its frame ->
full-body pose, Euler velocity / acceleration targets, foot marker targets, CoM and centroidal-momentum targets for MMSTO.

  DeltaExtractor.py:17-22             column layout: toPelvis 6 | toFeet 24 | toMomentum 9 | poses 53 | dx 59 | ddx 59 | con 4
  DeltaExtractor.py:1038-1045         unpack_obs_extra: pose_tl 7 | dq 6 | foot frames 2 x 7 (translation, quaternion)
  BaseLib/motion/GMBS/liegroup.inl:131-151   Exp(se3) (Rodrigues + the translational part; identity rotation below 1e-9)
  BaseLib/motion/GMBS/liegroup2.inl:226-268  toBaseR: rotation matrix -> quaternion (Jehee Lee's branches)
  BaseLib/math/transf.cpp:98-103,178-183     transf::mult, toGlobalPos
  BaseLib/motion/IK_sdls/NodeWrap.cpp:325-371,548-573   calcMomentumCOM / setDQ of the one-body SRB tree (as oracle/bone_fk.py)
PARITY: the primitives (Exp(se3), toBaseR, transf, inv_dAd / one-body momentum) are pinned to the reference's compiled BaseLib
(tests/test_refpins.py: test_oracle_se3_exp_equals_reference, test_oracle_primitives_equal_reference); their composition is pinned to the
reference's own unpack_obs_extra / _extractSingleFrameFullbody / pack_obs_extra run from /root/reference (tests/tools/ref_delta_harness.py,
tests/test_refenv_pins.py: exact, momentum 3e-14).  The SRB body's mass / inertia / local CoM are inputs: in the reference they come from
`RE.MujocoLoader('gym_trackSRB/Characters/SRB4.xml', {convertToYUP=true})` (walk_SRBTrack2025.py:243-246), a file that is not
in the reference tree."""
import math

import numpy as np

from . import bone_fk as fk

TO_PELVIS, TO_FEET, TO_MOM, POSES, NDOF = 6, 24, 9, 53, 59
ROW_W = TO_PELVIS + TO_FEET + TO_MOM + POSES + 2 * NDOF + 4        # 214
EXTRA_W = 7 + 6 + 14 + 2                                            # 29


def se3_exp(S):
    """Exp(se3) -> (R row-major 3x3, p)."""
    w = np.asarray(S[0:3], dtype=float); v = np.asarray(S[3:6], dtype=float)
    theta = math.sqrt(float(w.dot(w)))
    if abs(theta) < 1e-9:
        return np.eye(3), v.copy()
    it = 1.0 / theta
    s = np.concatenate([w, v]) * it
    st, ct = math.sin(theta), math.cos(theta)
    vt = 1.0 - ct
    ut = (theta - st) * (s[0] * s[3] + s[1] * s[4] + s[2] * s[5])
    t0, t1, t2 = s[2] * st, s[1] * st, s[0] * st
    R = np.array([[s[0] * s[0] * vt + ct, s[0] * s[1] * vt - t0, s[0] * s[2] * vt + t1],
                  [s[0] * s[1] * vt + t0, s[1] * s[1] * vt + ct, s[1] * s[2] * vt - t2],
                  [s[0] * s[2] * vt - t1, s[1] * s[2] * vt + t2, s[2] * s[2] * vt + ct]])
    p = np.array([st * s[3] + ut * s[0] + vt * (s[1] * s[5] - s[2] * s[4]),
                  st * s[4] + ut * s[1] + vt * (s[2] * s[3] - s[0] * s[5]),
                  st * s[5] + ut * s[2] + vt * (s[0] * s[4] - s[1] * s[3])])
    return R, p


def mat_to_quat(m):
    """toBaseR / quater::setRotation(matrix)."""
    q = np.zeros(4)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0.0:
        s = math.sqrt(tr + 1.0)
        q[0] = s * 0.5
        s = 0.5 / s
        q[1] = (m[2, 1] - m[1, 2]) * s; q[2] = (m[0, 2] - m[2, 0]) * s; q[3] = (m[1, 0] - m[0, 1]) * s
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        s = math.sqrt((m[i, i] - (m[j, j] + m[k, k])) + 1.0)
        q[i + 1] = s * 0.5
        s = 0.5 / s
        q[0] = (m[k, j] - m[j, k]) * s
        q[j + 1] = (m[j, i] + m[i, j]) * s
        q[k + 1] = (m[k, i] + m[i, k]) * s
    return q


def extract_single_frame_fullbody(extra, row, srb_mass, srb_inertia, srb_local_com):
    """extra[29] (pack_obs_extra), row[214] (decoder output) -> dict(humanPose 60, dx_ddx 118, feetpos 24, com 3, momentum 6, con_bin 4, con 4).
    srb_inertia: 3x3 body-frame inertia of the SRB segment, srb_local_com: its local centre of mass."""
    extra = np.asarray(extra, dtype=float); row = np.asarray(row, dtype=float)
    pose_tl = extra[0:7]; dq = extra[7:13]
    foot = [(extra[13:16], extra[16:20]), (extra[20:23], extra[23:27])]
    rt, rq = pose_tl[0:3], pose_tl[3:7]
    c = 0
    toPelvis = row[c:c + TO_PELVIS]; c += TO_PELVIS
    toFeet = row[c:c + TO_FEET]; c += TO_FEET
    toMom = row[c:c + TO_MOM]; c += TO_MOM
    R, p = se3_exp(toPelvis)
    q2 = mat_to_quat(R)
    human = np.zeros(7 + POSES)
    human[7:] = row[c:c + POSES]; c += POSES
    human[3:7] = fk.qmult(rq, q2)
    human[0:3] = fk.vrotate(rq, p) + rt
    # one-body tree: setPoseDOF(pose_tl), setDQ(dq), calcMomentumCOM
    skel = dict(mass=[srb_mass], com=[list(srb_local_com)], inertia=[np.asarray(srb_inertia, dtype=float).ravel().tolist()])
    invR = fk.qinv(rq)
    W = fk.vrotate(invR, dq[0:3])[None]; V = fk.vrotate(invR, dq[3:6])[None]
    _, srb_mom = fk.calc_momentum_com(skel, rq[None], rt[None], W, V)
    com_error = fk.vrotate(rq, toMom[0:3])
    mom = np.concatenate([srb_mom[0:3] + fk.vrotate(rq, toMom[3:6]), srb_mom[3:6] + fk.vrotate(rq, toMom[6:9])])
    dx = row[c:c + NDOF].copy(); c += NDOF
    ddx = row[c:c + NDOF].copy(); c += NDOF
    con = row[c:c + 4].copy(); c += 4
    assert c == ROW_W
    feetpos = np.zeros(24)
    for ic in range(4):
        ft, fq = foot[ic // 2]
        feetpos[ic * 3:ic * 3 + 3] = fk.vrotate(fq, toFeet[ic * 3:ic * 3 + 3]) + ft
    for ic in range(4):
        feetpos[12 + ic * 3:15 + ic * 3] = fk.vrotate(rq, toFeet[12 + ic * 3:15 + ic * 3]) * 10.0 + dq[3:6]
    dx *= 10.0; ddx *= 10.0
    dx[0:3] = fk.vrotate(rq, dx[0:3]) + dq[3:6]
    ddx[0:3] = fk.vrotate(rq, ddx[0:3])
    return dict(humanPose=human, dx_ddx=np.concatenate([dx, ddx]), feetpos=feetpos, com=rt + com_error, momentum=mom,
                con_bin=(con > 0).astype(float), con=con)


def pack_obs_extra(env):
    """DeltaExtractor.py:1015-1036 on an OracleSRBEnv.  data.site_xpos / site_xmat of the two foot-centre sites are what
    set_feet_site_this_step_env wrote last (env/SRBEnv5_mocap_list_window.py:58-87): the feet of the action, i.e.
    left/right_foot_contact_pos_global and left/right_foot_contact_ori.  ZtoY: work/rendermodule.py:631-636."""
    d = env.data
    qpos = np.asarray(d.qpos, dtype=float); qvel = np.asarray(d.qvel, dtype=float)

    def v_ztoy(v):
        return np.array([v[1], v[2], v[0]])

    def q_ztoy(q):
        return np.array([q[0], q[2], q[3], q[1]])
    pose_q = q_ztoy(qpos[3:7]); pose_t = v_ztoy(qpos[0:3])
    dq = np.concatenate([fk.vrotate(pose_q, v_ztoy(qvel[3:6])), v_ztoy(qvel[0:3])])
    feet = []
    for pos, ori in ((env.left_foot_contact_pos_global, env.left_foot_contact_ori), (env.right_foot_contact_pos_global, env.right_foot_contact_ori)):
        feet.append(np.concatenate([v_ztoy(np.asarray(pos, dtype=float)), q_ztoy(mat_to_quat(np.asarray(ori, dtype=float)))]))
    con = np.array([1.0 if env.left_foot_in_contact else 0.0, 1.0 if env.right_foot_in_contact else 0.0])
    return np.concatenate([pose_t, pose_q, dq, feet[0], feet[1], con])
