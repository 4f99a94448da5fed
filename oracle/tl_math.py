"""ORACLE (test infrastructure only).  taesooLib math primitives restated in scalar Python/NumPy for the
AdaptiveSRB (gym_cdm2) path.  Quaternions are (w, x, y, z), Y is up.  Sources (relative to the reference root):

  work/taesooLib/BaseLib/math/quater.cpp:52-65 (mult), :154-168 (normalize), :185-210 (exp), :217-272 (slerp /
      safeSlerp), :307-325 (axisToAxis), :327-340 (setRotation(axis, angle)), :777-788 (difference),
      :891-929 (decompose / rotationAngleAboutAxis);  quater.h:103-135 (inverse, rotationY, align, rotationAngle)
  work/taesooLib/BaseLib/math/vector3.cpp:155-163 (ln), :344-373 (rotate), :381-386 (rotationVector)
  work/taesooLib/BaseLib/math/Operator.cpp:15-36 (sop::map / clampMap)
  work/taesooLib/BaseLib/math/matrixn.cpp:781-805 (matrixn::sampleRow, float32 `t`)
  work/taesooLib/BaseLib/motion/MotionDOF.cpp:593-620 (samplePose), :237-257 (MotionDOFinfo::blend)
  work/taesooLib/Resource/scripts/module.lua:2876-2913 (smoothClamp / smoothClampVec3), :4407-4437
      (vectorn:smoothClamp), :5312-5324 (sop.mapSin / mapCos), :5847-5851 (MotionDOF:sample), :6062-6105
      (dposeToDQ / calcDerivative), :7531-7544 (transf:twist)
"""
import math
import numpy as np

EPS = 1e-10        # BaseLib/math/mathclass.h EPS
Y = np.array([0.0, 1.0, 0.0])
QI = np.array([1.0, 0.0, 0.0, 0.0])


def qmult(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                     a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def qnormalize(q):
    return q * (1.0 / math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]))


def qinv(q):
    return qnormalize(np.array([q[0], q[1] * -1.0, q[2] * -1.0, q[3] * -1.0]))


def qrot(q, v):
    """vector3::rotate(quater, v)."""
    u = q[1:4]
    s = q[0]
    return (2.0 * u.dot(v)) * u + (s * s - u.dot(u)) * v + (2.0 * s) * np.cross(u, v)


def qaxis(axis, angle):
    h = 0.5 * angle
    s = math.sin(h)
    return np.array([math.cos(h), s * axis[0], s * axis[1], s * axis[2]])


def qalign(q, other):
    return q * -1.0 if q.dot(other) < 0 else q


def vnormalize(v):
    return v / math.sqrt(v.dot(v))


def axis_to_axis(vfrom, vto):
    va = vnormalize(np.asarray(vfrom, float))
    vb = vnormalize(np.asarray(vto, float))
    if va.dot(vb) < -0.999:
        q = axis_to_axis(va, vb * -1.0)
        return qmult(qaxis(np.array([0.0, 0.0, 1.0]), math.radians(180)), q)
    vh = vnormalize(va + vb)
    ax = np.cross(va, vh)
    return np.array([va.dot(vh), ax[0], ax[1], ax[2]])


def decompose_notwist_twist(q, axis):
    """q = noTwist * twist (quater::decomposeNoTwistTimesTwist)."""
    rotated = qrot(q, axis)
    nt = axis_to_axis(axis, rotated)
    tw = qmult(qinv(nt), q)
    return nt, tw


def rotation_y(q):
    """quater::rotationY(): q = rotY * offset, rotY = twist of decomposeTwistTimesNoTwist about Y."""
    return decompose_notwist_twist(q, Y)[1]


def offset_q(q):
    ry = rotation_y(q)
    return qmult(qinv(ry), q)


def qln(q):
    sc = math.sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    theta = math.atan2(sc, q[0])
    sc = theta / sc if sc > 1e-10 else 1.0          # vector3.cpp eps
    return np.array([sc * q[1], sc * q[2], sc * q[3]])


def rotation_vector(q):
    return qln(q) * 2.0


def rotation_angle(q):
    r = rotation_vector(q)
    return math.sqrt(r.dot(r))


def rotation_angle_about_axis(q, axis=Y):
    tw = decompose_notwist_twist(q, axis)[1]
    r = rotation_vector(tw)
    n = math.sqrt(r.dot(r))
    return n * -1.0 if r.dot(axis) < 0 else n


def qexp(w):
    theta = math.sqrt(w.dot(w))
    sc = 1.0 if theta < EPS else math.sin(theta) / theta
    return np.array([math.cos(theta), sc * w[0], sc * w[1], sc * w[2]])


def q_from_rotvec(v):
    return qexp(np.asarray(v, float) / 2.0)


def qdifference(q1, q2):
    """delta * q1 = q2."""
    return qnormalize(qmult(qalign(q2.copy(), q1), qinv(q1)))


def slerp(a, b, t):
    c = a.dot(b)
    if 1.0 + c > EPS:
        if 1.0 - c > EPS:
            theta = math.acos(c)
            return (a * math.sin((1 - t) * theta) + b * math.sin(t * theta)) * (1.0 / math.sin(theta))
        return qnormalize(a * (1 - t) + b * t)
    return a * math.sin((0.5 - t) * math.pi) + b * math.sin(t * math.pi)


def safe_slerp(a, b, t):
    return slerp(a * -1.0 if a.dot(b) < 0 else a, b, t)


def q2mat(q):
    """matrix4::setRotation(quater) for a unit quaternion (rows)."""
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def twist(q1, p1, q2, p2, timestep):
    """transf:twist (module.lua:7531): unskew(T1^-1 (T2 - T1)/dt) -> (w, v) in the body frame of T1."""
    R1, R2 = q2mat(q1), q2mat(q2)
    M = R1.T.dot(R2 - R1) / timestep
    v = R1.T.dot(p2 - p1) / timestep
    return np.array([M[1, 2] * -1.0, M[0, 2], M[0, 1] * -1.0]), v


def sop_map(t, lo, hi, v1, v2):
    t = (t - lo) / (hi - lo)
    return v1 * (1.0 - t) + v2 * t


def clamp_map(t, lo, hi, v1, v2):
    t = min(max(t, lo), hi)
    return sop_map(t, lo, hi, v1, v2)


def map_sin(i, a, b, c, d):
    return math.sin(sop_map(i, a, b, 0.0, math.pi / 2)) * (d - c) + c


def map_cos(i, a, b, c, d):
    return sop_map(math.cos(sop_map(i, a, b, 0.0, math.pi / 2)), 1.0, 0.0, c, d)


def smooth_clamp(v, vmax):
    return math.sin(v / vmax) * vmax if v < math.radians(90) * vmax else vmax


def smooth_clamp_vec3(vec, max_len):
    ln = math.sqrt(vec.dot(vec))
    out = vec.copy()
    if ln > 0.001:
        out = out * (smooth_clamp(ln, max_len) / ln)
    return out


def smooth_clamp_sym(x, mag):
    """vectorn:smoothClamp(-mag, mag) element."""
    return smooth_clamp(x, mag) if x > 0 else smooth_clamp(x * -1.0, mag) * -1.0


def sample_row(mat, t):
    """matrixn::sampleRow: linear, the fractional part is rounded to float32, snaps within 0.005 of a row."""
    a = int(math.floor(t))
    f = float(np.float32(t - float(np.float32(a))))
    if f < 0.005:
        return mat[a].copy()
    if f > 0.995:
        return mat[a + 1].copy()
    if a < 0:
        f, r0, r1 = f - 1.0, mat[a + 1], mat[a + 2]
    elif a + 1 >= mat.shape[0]:
        f, r0, r1 = f + 1.0, mat[a - 1], mat[a]
    else:
        r0, r1 = mat[a], mat[a + 1]
    return r0 * (1.0 - f) + r1 * f                 # v::interpolate


def motiondof_sample(mat, t):
    """MotionDOF:sample (module.lua:5847): sampleRow + normalised root quaternion."""
    r = sample_row(mat, t)
    r[3:7] = qnormalize(r[3:7])
    return r


def sample_pose(mat, t):
    """MotionDOF::samplePose: double `t`, root quaternion by safeSlerp."""
    a = int(math.floor(t))
    f = t - a
    if f < 0.005:
        return mat[a].copy()
    if f > 0.995:
        return mat[a + 1].copy()
    if a < 0:
        f, r0, r1 = f - 1.0, mat[a + 1], mat[a + 2]
    elif a + 1 >= mat.shape[0]:
        f, r0, r1 = f + 1.0, mat[a - 1], mat[a]
    else:
        r0, r1 = mat[a], mat[a + 1]
    out = (1.0 - f) * r0 + f * r1
    out[3:7] = safe_slerp(r0[3:7], r1[3:7], f)
    return out


def calc_derivative(mot, frame_rate):
    """MotionDOF.calcDerivative (module.lua:6077): forward difference, root as body twist [v(0:3), 0, w(4:7)]."""
    n = mot.shape[0]
    d = np.zeros_like(mot)
    for i in range(n - 1):
        d[i] = (mot[i + 1] - mot[i]) * frame_rate
        w, v = twist(mot[i, 3:7], mot[i, 0:3], mot[i + 1, 3:7], mot[i + 1, 0:3], 1.0 / frame_rate)
        d[i, 0:3] = v
        d[i, 3] = 0.0
        d[i, 4:7] = w
    d[n - 1] = d[n - 2]
    return d
