"""ORACLE (test infrastructure only).  CPU restatement of the foot-sliding removal that follows the MMSTO solve in the online loop
(SURVEY 8f rank 2):

  gym_trackSRB/taesooLib/DeltaExtractor.py:650-760   removeFootSliding_online       (call site walk_SRBTrack2025.py:631)
  gym_trackSRB/taesooLib/DeltaExtractor.py:763-893   removeFootSliding_online_stair + avoidHull
                                                      (call site walk_SRBTrack2025_terrain_singlethreaded.py:624)
  Resource/scripts/module.lua:2108-2207              QuadraticFunctionHardCon (add / addWeighted / addSquared / con / solve): restated by
                                                      oracle/mmsto.hardcon_solve (KKT system solved by LU, as the reference does)
  gym_cdm2/ConvexHull2D.lua:150-186,362-455          math.projectTrajectoryToItsHull over ConvexHull2D
  work/taesooLib/PhysicsLib/convexhull/graham.cpp:81-209,252-270   GrahamScan: partition by the left-right chord (strictly above ->
                                                      upper), half hulls that drop collinear points (`<= 0`), closed polygon

Per foot and marker (toe, heel) and coordinate, a trajectory x[0..nvar) over the planning horizon minimises
    sum_v (x_v - x_{v-1} - desiredVel_{v-1}/30)^2 + 1.5 sum_v (-x_{v-1} + 2 x_v - x_{v+1})^2 + (c sum_{f>=2} x_f - c sum_{f>=2} marker_f)^2
(c = 1 for the vertical coordinate, 0.1 otherwise) subject to x_0 = initialFeet and, on every contact interval, x_f = the marker's
contact position while that marker is down, or the position on a circle of radius footLen about the OTHER marker's contact position
while only the other one is down.

NOT IN THE REFERENCE TREE (restated from their call sites, stated stand-ins):
  W5  fc.getContactPositions_v2(footLen, conToe, conHeel, toeTraj, heelTraj) -> conPos, conDir, intervals.  Restated on the pattern of
      fc.getContactPositions (gym_cdm2/module/CDMTraj.lua:376-435): intervals = run-length encoding of conToe | conHeel; conPos = mean over
      the interval of the mid-sole point, where a frame with only the toe (heel) down rebuilds the other marker at footLen along the
      horizontal toe-heel direction; conDir = normalised horizontal sum of (toe - heel) over the interval.
  W6  fc.getContactPositions_stair(conToe, conHeel, footCenterTraj) -> conPos = mean of the foot-centre trajectory over each interval.
PARITY UNPINNED for these two.  The quadratic programmes, their hard constraints, the ground clamp and avoidHull are PINNED to the reference's own
functions run from /root/reference with W5 / W6 / the hull projection plugged in as its Lua helpers (tests/tools/ref_delta_harness.py,
tests/test_refenv_pins.py: 1e-15 over 40 windows)."""
import math

import numpy as np

from .mmsto import hardcon_solve


def run_length_encode(flags):
    """intIntervals:runLengthEncode(boolN): maximal runs of true as (start, end) with end exclusive."""
    out, s = [], None
    for i, f in enumerate(list(flags) + [False]):
        if f and s is None:
            s = i
        elif not f and s is not None:
            out.append((s, i)); s = None
    return out


def _hnorm(v):
    v = np.array([v[0], 0.0, v[2]])
    n = math.sqrt(float(v.dot(v)))
    return v / n if n > 0 else v


def get_contact_positions_v2(foot_len, con_toe, con_heel, toe, heel):
    """W5 (see header).  toe / heel: [nvar, 3] (Y up)."""
    ii = run_length_encode(np.logical_or(con_toe, con_heel))
    con_pos, con_dir = [], []
    for s, e in ii:
        avg = np.zeros(3); dsum = np.zeros(3)
        for i in range(s, e):
            t, h = toe[i].copy(), heel[i].copy()
            if con_toe[i] and not con_heel[i]:
                h = t + _hnorm(h - t) * foot_len
            elif con_heel[i] and not con_toe[i]:
                t = h + _hnorm(t - h) * foot_len
            avg += t * 0.5 + h * 0.5
            d = toe[i] - heel[i]
            dsum += np.array([d[0], 0.0, d[2]])
        con_pos.append(avg * (1.0 / (e - s)))
        con_dir.append(_hnorm(dsum))
    return np.array(con_pos).reshape(-1, 3), np.array(con_dir).reshape(-1, 3), ii


def get_contact_positions_stair(con_toe, con_heel, foot_center):
    """W6 (see header)."""
    ii = run_length_encode(np.logical_or(con_toe, con_heel))
    return np.array([foot_center[s:e].mean(axis=0) for s, e in ii]).reshape(-1, 3), ii


def _marker_qps(foot_len, con_all, initial_feet, marker_traj, marker, desired_vel, ground, stair, foot_center):
    nvar = con_all.shape[0]
    traj = [np.asarray(t, dtype=float) for t in marker_traj]
    marker = np.asarray(marker, dtype=float)
    if desired_vel is None:                                   # matrixn:derivative_forward(30) of the four smooth marker trajectories
        allm = np.concatenate(traj, axis=1)
        desired_vel = np.zeros_like(allm)
        desired_vel[:-1] = (allm[1:] - allm[:-1]) * 30.0
        desired_vel[-1] = desired_vel[-2]
    opt = np.zeros((nvar, 12))
    for ifoot in range(2):
        con_toe = con_all[:, ifoot * 2] > 0.5
        con_heel = con_all[:, ifoot * 2 + 1] > 0.5
        con_pos, con_dir, ii = get_contact_positions_v2(foot_len, con_toe, con_heel, marker[:, ifoot * 6:ifoot * 6 + 3], marker[:, ifoot * 6 + 3:ifoot * 6 + 6])
        if stair:
            con_pos, _ = get_contact_positions_stair(con_toe, con_heel, np.asarray(foot_center[ifoot], dtype=float))
        con_pos = con_pos.copy()
        opt[:, ifoot * 6:ifoot * 6 + 3] = traj[ifoot * 2]
        opt[:, ifoot * 6 + 3:ifoot * 6 + 6] = traj[ifoot * 2 + 1]
        if con_toe[0]:
            con_pos[0] = initial_feet[ifoot * 6:ifoot * 6 + 3] - con_dir[0] * foot_len * 0.5
        elif con_heel[0]:
            con_pos[0] = initial_feet[ifoot * 6 + 3:ifoot * 6 + 6] + con_dir[0] * foot_len * 0.5
        for i in range(2):
            imarker = ifoot * 2 + i
            cur = traj[imarker]
            if i == 0:
                con = con_toe
                m_con = con_pos + con_dir * foot_len * 0.5; o_con = con_pos - con_dir * foot_len * 0.5
                mdir = cur - traj[ifoot * 2 + 1]
            else:
                con = con_heel
                m_con = con_pos - con_dir * foot_len * 0.5; o_con = con_pos + con_dir * foot_len * 0.5
                mdir = cur - traj[ifoot * 2]
            m_con = m_con.copy(); o_con = o_con.copy()
            if ground is not None and not stair:
                for j in range(m_con.shape[0]):
                    m_con[j, 1] = ground(m_con[j]); o_con[j, 1] = ground(o_con[j])
            mdir = mdir / np.linalg.norm(mdir, axis=1, keepdims=True)
            init = initial_feet[imarker * 3:imarker * 3 + 3]
            for dim in range(3):
                sq = []
                for v in range(1, nvar):
                    sq.append((1.0, [(1.0, v), (-1.0, v - 1)], -desired_vel[v - 1, imarker * 3 + dim] / 30.0))
                for v in range(1, nvar - 1):
                    sq.append((1.5, [(-1.0, v - 1), (2.0, v), (-1.0, v + 1)], 0.0))
                c = 1.0 if dim == 1 else 0.1
                sq.append((1.0, [(c, f) for f in range(2, nvar)], c * (-1.0 * float(marker[2:nvar, imarker * 3 + dim].sum()))))
                cons = [([(1.0, 0)], -init[dim])]
                for icon, (s, e) in enumerate(ii):
                    if s == 0:
                        s += 1
                    for f in range(s, e):
                        if con[f]:
                            cons.append(([(1.0, f)], -m_con[icon, dim]))
                        else:
                            cons.append(([(1.0, f)], -(o_con[icon, dim] + foot_len * mdir[f, dim])))
                opt[:, imarker * 3 + dim] = hardcon_solve(nvar, sq, cons)
    return opt


def remove_foot_sliding_online(foot_len, con_all, initial_feet, marker_traj, marker, desired_vel=None, ground=None):
    """-> optMarkerTraj [nvar, 12] (toe L, heel L, toe R, heel R).  marker_traj: four [nvar, 3] smooth trajectories; marker [nvar, 12]: the
    jerky predictions; ground(pos) -> height under pos (projectToGround; None = no projection)."""
    con_all = np.asarray(con_all, dtype=float); initial_feet = np.asarray(initial_feet, dtype=float)
    opt = _marker_qps(foot_len, con_all, initial_feet, marker_traj, marker, desired_vel, ground, False, None)
    if ground is not None:
        for i in range(4):
            for f in range(opt.shape[0]):
                g = ground(opt[f, i * 3:i * 3 + 3])
                if opt[f, i * 3 + 1] < g:
                    opt[f, i * 3 + 1] = g
    return opt


# ---- math.projectTrajectoryToItsHull -----------------------------------------------------------------------------------------------
def _direction(p0, p1, p2):
    return (p0[0] - p1[0]) * (p2[1] - p1[1]) - (p2[0] - p1[0]) * (p0[1] - p1[1])


def _half_hull(left, right, pts, factor):
    inp = list(pts) + [right]
    out = [left]
    while inp:
        out.append(inp.pop(0))
        while len(out) >= 3:
            e = len(out) - 1
            if factor * _direction(out[e - 2], out[e], out[e - 1]) <= 0:
                del out[e - 1]
            else:
                break
    return out


def graham_hull(points):
    """GrahamScan partition_points + build_hull + get_hull: closed polygon (first vertex repeated), lower hull left to right then the
    upper hull right to left."""
    raw = sorted((float(x), float(y)) for x, y in points)
    left, right = raw[0], raw[-1]
    mid = raw[1:-1]
    upper = [p for p in mid if _direction(left, right, p) < 0]
    lower = [p for p in mid if not _direction(left, right, p) < 0]
    lo = _half_hull(left, right, lower, 1.0)
    up = _half_hull(left, right, upper, -1.0)
    return lo + up[::-1][1:]


def project_trajectory_to_its_hull(y):
    """y[n] (heights over the frame index) -> (changed?, y raised onto the upper side of the convex hull of the points (i, y_i)).
    The reference projects only when the hull has more than three edges (`supportPlanes:size() > 3`)."""
    y = np.array(y, dtype=float)
    n = y.shape[0]
    hull = graham_hull([(i, y[i]) for i in range(n)])
    if len(hull) - 1 <= 3:
        return False, y
    idx = [int(round(p[0])) for p in hull]                    # search_vtx: hull vertices are input points, x = frame index
    err = 1e-3
    out = y.copy()
    for i in range(n):
        def lines(z):
            return [k for k in range(len(idx) - 1) if min(idx[k], idx[k + 1]) - err < z < max(idx[k], idx[k + 1]) + err]
        ln = lines(i)
        if not ln:
            ln = lines(0) if i < err else lines(n - 1)
        ymax = -10000.0
        for k in ln:
            z1, z2 = idx[k], idx[k + 1]
            t = (i - z1) / (z2 - z1)
            ymax = max(ymax, y[z1] * (1.0 - t) + y[z2] * t)
        out[i] = max(out[i], ymax)
    return True, out


def avoid_hull(opt, ground, toe_weight):
    """avoidHull (DeltaExtractor.py:869-893): lift both markers of a foot where their weighted mean dips below the hull of the ground
    heights along the trajectory."""
    n = opt.shape[0]
    for ii in range(2):
        ty = np.zeros(n)
        for f in range(n):
            p = opt[f, ii * 6:ii * 6 + 3] * toe_weight + opt[f, ii * 6 + 3:ii * 6 + 6] * (1.0 - toe_weight)
            ty[f] = ground(p)
        _, ty = project_trajectory_to_its_hull(ty)
        for f in range(n):
            py = opt[f, ii * 6 + 1] * toe_weight + opt[f, ii * 6 + 4] * (1.0 - toe_weight)
            if py < ty[f]:
                d = ty[f] - py
                opt[f, ii * 6 + 1] += d
                opt[f, ii * 6 + 4] += d


def remove_foot_sliding_online_stair(foot_len, con_all, initial_feet, marker_traj, marker, foot_center, desired_vel=None, ground=None):
    """foot_center: two [nvar, 3] foot-centre trajectories (from the SRB observation); contact positions are not projected to the ground,
    the result is kept above the hull of the ground profile instead (three passes, toe weights 0.5 / 0.75 / 0.25)."""
    con_all = np.asarray(con_all, dtype=float); initial_feet = np.asarray(initial_feet, dtype=float)
    opt = _marker_qps(foot_len, con_all, initial_feet, marker_traj, marker, desired_vel, ground, True, foot_center)
    for w in (0.5, 0.75, 0.25):
        avoid_hull(opt, ground, w)
    return opt
