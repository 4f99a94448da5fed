"""Reference tables of the AdaptiveSRB environment (gym_cdm2 `walk2-v1`), built on the host once per clip.

The reference builds them with taesooLib preprocessing that cannot run here (SURVEY 7 "Hard parts"):
`RagdollSim:prepareMotionForSim` (gym_cdm2/RagdollSim.lua:18-95) -> `fc.CDMTraj` (module/CDMTraj.lua:611-839, pendulum
branch :709-775) and `fc.prepareMotionsForSim` (:1400-1610), plus the lazily built integer tables of
defaultRLsetting.lua:291-346,395-418.  This module restates that pipeline in NumPy:

  exact restatements   .dof reader (BinaryFile: int tags, SPFLOATMN float32 matrix), contact flags
                       (fc.buildContactConstraints :576-603), loop stitching with planar alignment (:1451-1513),
                       DMotionDOF / CDM_DMotionDOF (MotionDOF.calcDerivative), rttTouchDown / rttTouchOff /
                       swingDuration, mocapFeetOffset (:1564-1606), CoM by FK over the WRL skeleton.
                       `math.gaussFilter` / `math.filter` (BaseLib/math/Filter.cpp:25-31,257-277,347-389: kernel
                       exp(-i^2 / size) over an odd size, end samples repeated) and `math.NonuniformSpline` with zero end
                       accelerations (BaseLib/math/BSpline.cpp:143-414 = the natural cubic spline; SciPy's agrees with the
                       compiled reference to 5e-15) -- both pinned against oracle/_ref in tests/test_refpins.py.
  STAND-IN (stated)    `makeCyclic` (module/info_hyunwooLowDOF.lua:23-35: MotionDOF:Stitch twice, taesooLib's C++ stitch
                       operators over an Eigen solve) is replaced by a linear blend of the end-of-cycle discontinuity over
                       `spread` frames.
The step kernel and its oracle consume the tables as inputs, so parity of the hot path does not depend on them.

Table row (TAB_W reals per frame, see `pack_rows`): CDM pose 7, CDM body twist (v, w) 6, mocap feet offset 2x3,
leg chain pose 21 (root xyz + quat, L hip/knee/ankle, R hip/knee/ankle), cycle frame 1, swing duration 2,
rttTouchDown 2, rttTouchOff 2, contact bits 1.
"""
import math
import struct
import numpy as np

TAB_W = 48
WALK4 = dict(touchDown=[[0, 16], [0, 32]], touchOff=[[2], [18]], start=0, strideThr=0.45, maxContactLen=0.9,
             minSwingPhase=0.25, filterFD=16, spread=15)
FEET_CONFIG_TOE = np.array([0.0, -0.12, 0.15])       # input.limbs[1][3]  (module/info_hyunwooLowDOF.lua:38)
FEET_CONFIG_HEEL = np.array([0.0, -0.15, -0.03])     # input.limbs[2][3]  (:39)
# motionInfos.run4 (gym_cdm2/module/motionInfo.lua:133-183): no strideThr / maxContactLen / minSwingPhase, no makeCyclic
# (fixDOF_part2 is empty), the CoM acceleration of the pendulum lean is low-passed (option_refCDMtraj.filterCDMacc = 12)
RUN4 = dict(touchDown=[[11], [0, 23]], touchOff=[[20], [8]], start=0, strideThr=None, maxContactLen=None, minSwingPhase=None,
            filterFD=16, filterCDMacc=12, spread=0)
AX = {"X": 0, "Y": 1, "Z": 2}

# The registered environments of gym_cdm2/__init__.py:9-11 whose clips ship with the reference: motion file (under
# work/taesooLib/Resource/motion/MOB1/), motionInfo, and the overrides of the walk2-v1 defaults that cdm_create starts from
# (cdm_set_param names).  `params` = the TRAINING configuration (no GUI: defaultRLsetting.lua:183-187 forces touchDownHeight 0
# and unlimitedLeglen); `gui_params` = what the spec file itself sets and the Lua main-loop viewer keeps (test_cdm_deepmimic.lua: isMainloopInLua).
#   run2-v1: spec/run2-v1.lua:14-19 (maxTurning, facingWeight, maxLegLen, contactThr, healthy_y_range), defaults where the spec
#   and run4 are silent: k_p_ID 340 (defaultRLsetting.lua:43), EFdif_scale 1/8 (deepmimiccdm-v1.lua:2108), maxContactLen 2.0
#   (:1246), minSwingPhase 0.4 (:615).  Not overridden although run4 names them: filterEndPhase (RagdollSim:filterContact hard-codes 0.4,
#   deepmimiccdm-v1.lua:524 -- the recorded run reproduces only with 0.4) and maxContactY (never copied to the model: QPservo keeps 1.1); stride_thr < 0 = none.
SPECS = {
    "walk2-v1": dict(motion="hanyang_lowdof_T_lafan1_walk_2915.dof", info=WALK4, params={}, terrain=False,
                     gui_params=dict(unlimited_leglen=0.0, touch_down_height=0.83)),
    "walkterrain-v1": dict(motion="hanyang_lowdof_T_lafan1_walk_2915.dof", info=WALK4, params={}, terrain=True,
                           gui_params=dict(unlimited_leglen=0.0, touch_down_height=0.83)),
    "run2-v1": dict(motion="hanyang_lowdof_T_lafan1_run_3750.dof", info=RUN4, terrain=False,
                    params=dict(k_p=340.0, ef_scale=0.125, max_leg_len=1.5, contact_thr=0.1, stride_thr=-1.0, max_contact_len=2.0,
                                min_swing_phase=0.4, healthy_min=0.8, healthy_max=1.1, facing_weight=10.0,
                                max_turning=0.4),
                    gui_params=dict(unlimited_leglen=0.0, touch_down_height=0.7)),      # defaultRLsetting.lua:22 (the spec does not override it)
}


def read_dof(path):
    """taesooLib MotionDOFcontainer file: (INT version)(INT frames)(INT dofs)(SPFLOATMN rows cols float32...)."""
    b = open(path, "rb").read()
    tags = struct.unpack_from("<9i", b, 0)
    assert tags[0] == 0 and tags[2] == 0 and tags[4] == 0 and tags[6] == 14, "unexpected .dof layout"
    rows, cols = tags[7], tags[8]
    assert rows == tags[3] and cols == tags[5]
    return np.frombuffer(b, dtype="<f4", count=rows * cols, offset=36).reshape(rows, cols).astype(np.float64)


# ---- small quaternion helpers (w, x, y, z), taesooLib conventions -------------------------------------------------
def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                     a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def _qrot(q, v):
    u, s = q[1:4], q[0]
    return 2.0 * u.dot(v) * u + (s * s - u.dot(u)) * v + 2.0 * s * np.cross(u, v)


def _qinv(q):
    c = np.array([q[0], -q[1], -q[2], -q[3]])
    return c / math.sqrt(c.dot(c))


def _qaxis(axis, ang):
    s = math.sin(0.5 * ang)
    return np.array([math.cos(0.5 * ang), s * axis[0], s * axis[1], s * axis[2]])


def _axis_to_axis(a, b):
    a = a / np.linalg.norm(a); b = b / np.linalg.norm(b)
    h = a + b
    h = h / np.linalg.norm(h)
    ax = np.cross(a, h)
    return np.array([a.dot(h), ax[0], ax[1], ax[2]])


def _rot_y(q):
    nt = _axis_to_axis(np.array([0.0, 1.0, 0.0]), _qrot(q, np.array([0.0, 1.0, 0.0])))
    return _qmul(_qinv(nt), q)


def _q2mat(q):
    s = 2.0 / math.sqrt(q.dot(q))
    w, x, y, z = q
    xs, ys, zs = x * s, y * s, z * s
    return np.array([[1 - (y * ys + z * zs), x * ys - w * zs, x * zs + w * ys],
                     [x * ys + w * zs, 1 - (x * xs + z * zs), y * zs - w * xs],
                     [x * zs - w * ys, y * zs + w * xs, 1 - (x * xs + y * ys)]])


def fk(skel, pose):
    nb = len(skel["parent"])
    rot = np.zeros((nb, 4)); tr = np.zeros((nb, 3))
    start = 7
    for b in range(nb):
        off = np.array(skel["offset"][b])
        if b == 0:
            rot[b] = pose[3:7]; tr[b] = pose[0:3] + off
            continue
        lq = np.array([1.0, 0, 0, 0])
        for c in skel["channels"][b]:
            e = np.zeros(3); e[AX[c]] = 1.0
            lq = _qmul(lq, _qaxis(e, pose[start])); start += 1
        p = skel["parent"][b]
        rot[b] = _qmul(rot[p], lq)
        tr[b] = _qrot(rot[p], off) + tr[p]
    return rot, tr


def _com(skel, rot, tr):
    m = np.array(skel["mass"])
    c = np.array([_qrot(rot[b], np.array(skel["com"][b])) + tr[b] for b in range(len(m))])
    return (m[:, None] * c).sum(0) / m.sum()


def _twist(q1, p1, q2, p2, dt):
    R1, R2 = _q2mat(q1), _q2mat(q2)
    M = R1.T.dot(R2 - R1) / dt
    return np.array([-M[1, 2], M[0, 2], -M[0, 1]]), R1.T.dot(p2 - p1) / dt


def calc_derivative(mot, frame_rate=30.0):
    d = np.zeros_like(mot)
    for i in range(len(mot) - 1):
        d[i] = (mot[i + 1] - mot[i]) * frame_rate
        w, v = _twist(mot[i, 3:7], mot[i, :3], mot[i + 1, 3:7], mot[i + 1, :3], 1.0 / frame_rate)
        d[i, :3] = v; d[i, 3] = 0.0; d[i, 4:7] = w
    d[-1] = d[-2]
    return d


def _project2d(q, p):
    return _rot_y(q), np.array([p[0], 0.0, p[2]])


def _transform_rows(rows, dq, dp):
    out = rows.copy()
    for r in out:
        r[:3] = _qrot(dq, r[:3]) + dp
        r[3:7] = _qmul(dq, r[3:7])
    return out


def leg_chain_columns(skel):
    """columns of the 60-dof pose needed for the ankle markers: root 7 + L hip/knee/ankle + R hip/knee/ankle."""
    names = skel["names"]
    start = {}
    s = 7
    for b, n in enumerate(names):
        start[n] = (s, len(skel["channels"][b]))
        s += len(skel["channels"][b])
    cols = list(range(7))
    for n in ("LeftHip", "LeftKnee", "LeftAnkle", "RightHip", "RightKnee", "RightAnkle"):
        cols += list(range(start[n][0], start[n][0] + start[n][1]))
    return cols


def gauss_kernel(kernel_size):
    """Filter::GetGaussFilter (BaseLib/math/Filter.cpp:257-277): an even size becomes the next odd one, w_i = exp(-i^2 / size), sum 1."""
    if kernel_size % 2 == 0:
        kernel_size = kernel_size // 2 * 2 + 1
    i = np.arange(-(kernel_size // 2), kernel_size // 2 + 1, dtype=np.float64)
    w = np.exp(-(i * i) / float(kernel_size))
    return w / w.sum()


def gauss_filter(kernel_size, x):
    """math.gaussFilter / math.filter: Filter::LTIFilter(1, GetGaussFilter(size), x) per column (Filter.cpp:347-389,569-630) --
    convolution with the first / last sample repeated beyond the ends."""
    w = gauss_kernel(kernel_size)
    k = len(w) // 2
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    assert len(w) < n, "LTIFilter: kernel is too large"
    idx = np.clip(np.arange(n)[:, None] + np.arange(-k, k + 1)[None, :], 0, n - 1)      # [n, 2k+1]
    return np.einsum("nm...,m->n...", x[idx], w)


def build_tables(mot, skel, info=WALK4, frame_rate=30.0, limit_length=300, default_lean=0.28 * 0.5):
    mot = mot.copy()
    nfr = mot.shape[0]
    names = skel["names"]
    # ---- makeCyclic stand-in: blend the end-of-cycle discontinuity of joint angles / height over `spread` frames
    sp = info["spread"]                                            # 0: the clip is used as it is (run4)
    err = mot[-1] - mot[0]
    for i in range(nfr - sp if sp else nfr, nfr):
        w = (i - (nfr - 1 - sp)) / float(sp)
        mot[i, 7:] -= err[7:] * w
        mot[i, 1] -= err[1] * w
    # ---- fc.CDMTraj (pendulum branch)
    com = np.zeros((nfr, 3)); fwd = np.zeros((nfr, 3))
    iLS, iRS, iLH, iRH = (names.index(n) for n in ("LeftShoulder", "RightShoulder", "LeftHip", "RightHip"))
    for i in range(nfr):
        rot, tr = fk(skel, mot[i])
        com[i] = _com(skel, rot, tr)
        across = (tr[iLS] - tr[iRS]) + (tr[iLH] - tr[iRH])
        v = np.cross(across, np.array([0.0, 1.0, 0.0])); v[1] = 0.0
        fwd[i] = v / np.linalg.norm(v)
    sm = gauss_filter(info["filterFD"], fwd)                       # math.gaussFilter(filterFD, forwardDir) (CDMTraj.lua:875)
    rot_y = np.array([_qaxis(np.array([0.0, 1.0, 0.0]), math.atan2(f[0], f[2])) for f in sm])
    from scipy.interpolate import CubicSpline
    knots = np.arange(0, nfr, 7)
    spl = CubicSpline(knots.astype(float), com[knots], bc_type="natural")
    tt = np.arange(knots[-1] + 1).astype(float)
    curve, acc = spl(tt), spl(tt, 2)
    if info.get("filterCDMacc"):                                   # math.filter(newAcc, filterCDMacc) (CDMTraj.lua:738)
        acc = gauss_filter(info["filterCDMacc"], acc)
    cdm = np.zeros((nfr, 7))
    lean0 = _qaxis(np.array([1.0, 0.0, 0.0]), default_lean)
    for i in range(nfr):
        ii = min(i, len(tt) - 1)
        foot = curve[ii] + acc[ii] * -40.0; foot[1] = 0.0
        lean = _axis_to_axis(np.array([0.0, 1.0, 0.0]), curve[ii] - foot)
        q = _qmul(_qmul(lean, rot_y[i]), lean0)
        cdm[i, :3] = com[i]; cdm[i, 3:7] = q / np.linalg.norm(q)
    # ---- fc.prepareMotionsForSim
    def flags(lst):
        out = np.zeros((2, nfr), bool)
        for k in range(2):
            for v in lst[k]:
                if v < nfr: out[k, v] = True
        return out
    td0, to0 = flags(info["touchDown"]), flags(info["touchOff"])
    con0 = np.zeros((2, nfr), bool)
    for k in range(2):
        c = False
        for f in range(nfr):
            if f in info["touchDown"][k]: c = True
            elif f in info["touchOff"][k]: c = False
            con0[k, f] = c
    dm0 = calc_derivative(cdm, frame_rate)
    cdm_l, full_l, ifr, dm_l = cdm.copy(), mot.copy(), np.arange(nfr, dtype=float), dm0.copy()
    td, to, con = td0.copy(), to0.copy(), con0.copy()
    while True:
        qa, pa = _project2d(full_l[-1, 3:7], full_l[-1, :3])
        qb, pb = _project2d(mot[0, 3:7], mot[0, :3])
        dq = _qmul(qa, _qinv(qb)); dq /= np.linalg.norm(dq)         # delta * B0 = A_last (planar)
        dp = pa - _qrot(dq, pb); dp[1] = 0.0
        full_l = np.concatenate([full_l, _transform_rows(mot[1:], dq, dp)])
        cdm_l = np.concatenate([cdm_l, _transform_rows(cdm[1:], dq, dp)])
        dm_l = np.concatenate([dm_l, dm0[1:]])
        ifr = np.concatenate([ifr, np.arange(1, nfr, dtype=float)])
        pnf = td.shape[1]
        td = np.concatenate([td, td0[:, 1:]], 1)
        if td0[0, 0]: td[0, pnf] = True
        if td0[1, 0]: td[0, pnf] = True                            # sic: the reference sets touchDown[1] for both (:1506)
        to = np.concatenate([to, to0[:, 1:]], 1)
        con = np.concatenate([con, con0[:, 1:]], 1)
        if cdm_l.shape[0] - 1 > (limit_length + 2) * 3:
            break
    F = cdm_l.shape[0]
    # ---- lazily built integer tables (defaultRLsetting.lua:291-346, 395-418)
    def find(arr, start, val=True):
        j = start
        while j < len(arr) and arr[j] != val: j += 1
        return j
    def find_prev(arr, start, val):
        j = start - 1
        while j >= 0 and arr[j] != val: j -= 1
        return j
    rtt_down = np.zeros((2, F), np.int32); rtt_off = np.zeros((2, F), np.int32); swing_dur = np.zeros((2, F))
    for k in range(2):
        for i in range(F):
            rtt_down[k, i] = (find_prev(con[k], i, False) - i + 1) if con[k, i] else (find(td[k], i) - i)
            rtt_off[k, i] = (find_prev(con[k], i, True) - i + 1) if not con[k, i] else (find(to[k], i) - i)
        sw = ~con[k]
        runs = []
        i = 0
        while i < F:
            if sw[i]:
                j = i
                while j < F and sw[j]: j += 1
                runs.append((i, j)); i = j
            else:
                i += 1
        prev_e = None
        for n, (s, e) in enumerate(runs):
            dur = e - s - 1
            s2, e2 = s, e
            if n == 0 and s != 0: s2 = 0
            elif n == len(runs) - 1: e2 = F
            swing_dur[k, s2:e2] = dur
            if n > 0:
                pe = runs[n - 1][1]
                pd = swing_dur[k, pe - 1]
                m = (s2 - 1) - pe
                if m > 0: swing_dur[k, pe:s2 - 1] = np.linspace(pd, dur, m)
                # the reference's range(prev_e, s-1) leaves entry s-1 of an uninitialised vectorn untouched
                # (undefined value); it is given the following swing's duration here
                if s2 - 1 >= pe: swing_dur[k, s2 - 1] = dur
    # ---- mocapFeetOffset (:1564-1606): offset of the ankle mid-point in the planar CDM frame at each touch-down
    iRA, iLA = names.index("RightAnkle"), names.index("LeftAnkle")
    lpos = FEET_CONFIG_TOE * 0.5 + FEET_CONFIG_HEEL * 0.5
    feet_off = np.zeros((2, F, 3))
    for k in range(2):
        segs = []
        i = 0
        while i < F:
            if con[k, i]:
                j = i
                while j < F and con[k, j]: j += 1
                segs.append((i, j)); i = j
            else:
                i += 1
        for n, (s, e) in enumerate(segs):
            rot, tr = fk(skel, full_l[s])
            b = iRA if k == 0 else iLA
            gpos = _qrot(rot[b], lpos) + tr[b]
            gq, gp = _project2d(cdm_l[s, 3:7], cdm_l[s, :3])
            off = _qrot(_qinv(gq), gpos - gp); off[1] = 0.0
            feet_off[k, s:e] = off
            if n == 0:
                feet_off[k, :s] = off
            else:
                ps = segs[n - 1][1]
                prev = feet_off[k, ps - 1].copy()
                for i in range(ps, s):
                    t = min(max((i - ps) / float(s - ps), 0.0), 1.0)
                    t = -2.0 * t ** 3 + 3.0 * t * t
                    feet_off[k, i] = prev * (1 - t) + off * t
                if n == len(segs) - 1:
                    feet_off[k, e:] = off
    return dict(cdm=cdm_l, dcdm=dm_l, cdm_d=calc_derivative(cdm_l, frame_rate), fullbody=full_l, iframe_tab=ifr,
                con=con, touch_down=td, touch_off=to, rtt_down=rtt_down, rtt_off=rtt_off, swing_dur=swing_dur,
                feet_off=feet_off, nf=np.array(nfr - 1), leg_cols=np.array(leg_chain_columns(skel)))


def pack_rows(tab):
    """-> float64 [F, TAB_W] rows for the device table (see module docstring)."""
    F = tab["cdm"].shape[0]
    rows = np.zeros((F, TAB_W))
    rows[:, 0:7] = tab["cdm"]
    rows[:, 7:10] = tab["cdm_d"][:, 0:3]
    rows[:, 10:13] = tab["cdm_d"][:, 4:7]
    rows[:, 13:16] = tab["feet_off"][0]
    rows[:, 16:19] = tab["feet_off"][1]
    rows[:, 19:40] = tab["fullbody"][:, tab["leg_cols"]]
    rows[:, 40] = tab["iframe_tab"]
    rows[:, 41:43] = tab["swing_dur"].T
    rows[:, 43:45] = tab["rtt_down"].T
    rows[:, 45:47] = tab["rtt_off"].T
    rows[:, 47] = tab["con"][0] * 1 + tab["con"][1] * 2
    return rows
