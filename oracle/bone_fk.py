"""ORACLE (test infrastructure only).  CPU restatement of taesooLib's full-body kinematics used by the
SRB -> full-body momentum mapping:

  BaseLib/motion/BoneKinematics.cpp:18-47,97-136   BoneForwardKinematics::setPoseDOF / forwardKinematics
  BaseLib/motion/MotionLoader.cpp:49-131           Bone::getLocalTrans / getLocalOri
  BaseLib/math/quater.cpp:52-65,327-340,428-470    quater::mult / setRotation(axis,angle) / setRotation(channels)
  BaseLib/math/vector3.cpp:344-373                 vector3::rotate(quater, v)
  BaseLib/math/transf.cpp:98-103                   transf::mult
  BaseLib/motion/IK_sdls/NodeWrap.cpp:325-386,496-600  LoaderToTree::calcMomentumCOM / calcCOM / setQ / setDQ
  BaseLib/motion/IK_sdls/Node.cpp:58-99            Node::ComputeS / ComputeDS (one node per hinge dof)
  BaseLib/motion/Liegroup.cpp:45-75, Liegroup.h:131-147  dse3::dAd / inv_dAd, Inertia * se3 (Q13: no parallel-axis term)

PARITY PINNED to the reference's own code: BaseLib math / motion / IK_sdls compile from /root/reference (oracle/ref_build/Makefile ->
oracle/_ref/taesoo_ref); tests/test_refpins.py compares this file with what that binary computes on seeded poses (frames 1e-14,
link velocities / CoM 1e-12, momentum 1e-11; fixture tests/golden/ref_taesoolib.npz, generator tests/tools/make_golden_ref.py).
Analytic checks in addition: tests/test_oracle_fk_cpu.py (finite-difference velocities, momentum of a rigidly moving body,
independence from the choice of the node chain).

Pose layout (MotionDOF): [root xyz, root quaternion w x y z, joint angles in tree order]; dq layout
(LoaderToTree): [root angular velocity (world), root linear velocity (world), joint rates]."""
import math
import numpy as np

AXIS = {"X": np.array([1.0, 0, 0]), "Y": np.array([0, 1.0, 0]), "Z": np.array([0, 0, 1.0])}


def qmult(a, b):
    """quater::mult (w,x,y,z)."""
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                     a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def qaxis(axis, angle):
    h = 0.5 * angle
    s = math.sin(h)
    return np.array([math.cos(h), s * axis[0], s * axis[1], s * axis[2]])


def qinv(q):
    """quater::inverse (quater.h:103): conjugate, then normalize() (times 1/length)."""
    c = np.array([q[0], q[1] * -1.0, q[2] * -1.0, q[3] * -1.0])
    return c * (1.0 / math.sqrt(c.dot(c)))


def vrotate(q, v):
    """vector3::rotate(quater, v): (2 u.v) u + (s^2 - u.u) v + 2 s u x v."""
    u = q[1:4]
    s = q[0]
    return (2.0 * u.dot(v)) * u + (s * s - u.dot(u)) * v + (2.0 * s) * np.cross(u, v)


def local_rotation(channels, angles):
    """quater::setRotation(channels, values): right-multiplied single-axis rotations."""
    q = np.array([1.0, 0, 0, 0])
    for c, a in zip(channels, angles):
        q = qmult(q, qaxis(AXIS[c], a))
    return q


def bone_forward_kinematics(skel, pose):
    """-> global rotation [nb,4] (wxyz) and translation [nb,3] (BoneForwardKinematics::global(bone))."""
    nb = len(skel["parent"])
    rot = np.zeros((nb, 4)); tr = np.zeros((nb, 3))
    start = 7
    for b in range(nb):
        off = np.array(skel["offset"][b])
        if b == 0:
            lq = np.array(pose[3:7], dtype=float)                # toQuater: not normalised
            lt = np.array(pose[0:3], dtype=float) + off
            rot[b], tr[b] = lq, lt
        else:
            nc = len(skel["channels"][b])
            lq = local_rotation(skel["channels"][b], pose[start:start + nc])
            start += nc
            p = skel["parent"][b]
            rot[b] = qmult(rot[p], lq)
            tr[b] = vrotate(rot[p], off) + tr[p]
    assert start == len(pose)
    return rot, tr


def loader_to_tree(skel, pose, dq=None):
    """LoaderToTree with one node per dof.  Returns per-bone (last node) global rotation/translation and
    body-frame angular / linear velocity."""
    nb = len(skel["parent"])
    rot = np.zeros((nb, 4)); tr = np.zeros((nb, 3)); W = np.zeros((nb, 3)); V = np.zeros((nb, 3))
    start = 7
    for b in range(nb):
        off = np.array(skel["offset"][b])
        if b == 0:
            R = np.array(pose[3:7], dtype=float)
            rot[b] = R
            tr[b] = np.array(pose[0:3], dtype=float) + off
            if dq is not None:
                invR = qinv(R)
                W[b] = vrotate(invR, np.array(dq[0:3], dtype=float))     # fj->SetJointVel(invR*v, invR*w)
                V[b] = vrotate(invR, np.array(dq[3:6], dtype=float))
            continue
        p = skel["parent"][b]
        g_rot, g_tr, w, v = rot[p], tr[p], W[p], V[p]
        for k, c in enumerate(skel["channels"][b]):
            th = pose[start + k]
            lq = qaxis(AXIS[c], th)
            lt = off if k == 0 else np.zeros(3)
            # Node::ComputeS: global = parent.global * local
            new_rot = qmult(g_rot, lq)
            new_tr = vrotate(g_rot, lt) + g_tr
            if dq is not None:
                # Node::ComputeDS
                t_rel = qinv(lq)
                v1 = np.cross(w, lt)
                nv = vrotate(t_rel, v1) + vrotate(t_rel, v)               # rel_lin_vel = 0 for a hinge
                nw = vrotate(t_rel, w) + AXIS[c] * dq[6 + (start - 7) + k]
                w, v = nw, nv
            g_rot, g_tr = new_rot, new_tr
        start += len(skel["channels"][b])
        rot[b], tr[b], W[b], V[b] = g_rot, g_tr, w, v
    return rot, tr, W, V


def calc_com(skel, rot, tr):
    m = np.array(skel["mass"])
    com = np.zeros(3)
    for b in range(len(m)):
        com += (vrotate(rot[b], np.array(skel["com"][b])) + tr[b]) * m[b]
    return com * (1.0 / m.sum())


def calc_momentum_com(skel, rot, tr, W, V):
    """LoaderToTree::calcMomentumCOM -> dse3 [M(3) angular about the CoM, F(3) linear], world frame."""
    nb = len(skel["mass"])
    com = np.zeros(3); H = np.zeros(6); total = 0.0
    for b in range(nb):
        mass = skel["mass"][b]
        c = np.array(skel["com"][b])
        com += (vrotate(rot[b], c) + tr[b]) * mass
        total += mass
        I = np.array(skel["inertia"][b]).reshape(3, 3)
        Ixx, Iyy, Izz, Ixy, Ixz, Iyz = I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]
        r = mass * c
        w, v = W[b], V[b]
        # Inertia * se3 (Liegroup.cpp:66-74)
        M = np.array([Ixx * w[0] + Ixy * w[1] + Ixz * w[2] + r[1] * v[2] - r[2] * v[1],
                      Ixy * w[0] + Iyy * w[1] + Iyz * w[2] + r[2] * v[0] - r[0] * v[2],
                      Ixz * w[0] + Iyz * w[1] + Izz * w[2] + r[0] * v[1] - r[1] * v[0]])
        F = np.array([w[1] * r[2] - w[2] * r[1] + mass * v[0],
                      w[2] * r[0] - w[0] * r[2] + mass * v[1],
                      w[0] * r[1] - w[1] * r[0] + mass * v[2]])
        # inv_dAd(T) = dAd(T^-1): T^-1 = (R^-1, -R^-1 t)
        Rinv = qinv(rot[b])
        tinv = -vrotate(Rinv, tr[b])
        tmp = M - np.cross(tinv, F)
        Rinvinv = qinv(Rinv)
        H[:3] += vrotate(Rinvinv, tmp)
        H[3:] += vrotate(Rinvinv, F)
    com /= total
    out = np.zeros(6)
    out[:3] = H[:3] - np.cross(com, H[3:])      # dAd(transf(identity, com), H)
    out[3:] = H[3:]
    return com, out
