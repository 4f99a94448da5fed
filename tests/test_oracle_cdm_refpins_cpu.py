"""Pins of the AdaptiveSRB oracle against OUTPUTS OF THE REFERENCE ITSELF (SURVEY 8c "golden vectors / fixtures"):

  * tests/golden/cdm_reference_recordings.npz holds the first 700 rows of gym_cdm2/spec/refTraj_walk2-v1.dat and
    refTraj_run2-v1.dat -- roll-outs recorded by the reference simulator (deepmimiccdm-v1.lua:1095-1146; per RL step:
    root pose 7, filtered L / R foot 2x3, dq 6 = [w_world, v_world], R / L foot yaw, and rawFootInfo = support flags,
    raw contact targets, swing phases) -- and the LQR solution the reference cached in work/_LQR_CACHE_2.DAT.
  * Each test drives the ORACLE'S OWN methods (oracle/cdm_env.py) with the recorded inputs and compares with what the
    reference recorded one row later.  Row r holds the root state at the START of step r and the foot state AFTER step r's
    calcContactPos, so the filter update of step r sees the root state of row r.

Recorded positions carry float32 rounding (1e-6 at the 10-100 m the roll-outs reach), velocities and angles do not."""
import math

import numpy as np
import pytest

import helpers
import helpers_cdm as hc
from oracle import tl_math as tl
from oracle import cdm_env as ce


@pytest.fixture(scope="module")
def rec():
    import os
    return dict(np.load(os.path.join(helpers.GOLDEN, "cdm_reference_recordings.npz")))


@pytest.fixture(scope="module")
def env():
    tab, skel = hc.load_tables()
    return ce.OracleCDMEnv(tab, skel)


def test_lqr_known_answer(rec):
    """work/_LQR_CACHE_2.DAT: (K, A, B, Q, R) as written by Physics.LQR (PhysicsLib/SDRE.cpp:233-243)."""
    A, B, Q = rec["lqr_A"], rec["lqr_B"], rec["lqr_Q"]
    M = 1.0 / B[1, 0]
    K = ce.lqr_gain(M, -A[1, 1] * M, -A[1, 0] * M, Q[0, 0], Q[1, 1])
    assert np.allclose(K, rec["lqr_K"].ravel(), rtol=1e-12)
    assert np.allclose(rec["lqr_K"].ravel(), ce.PARAMS["filter_K"], rtol=1e-12)      # Q15: the gain every 2x2 LQR call returns


def _new_leg(env, is_left, pos, yaw):
    li = ce.Leg()
    li.isLeft = is_left
    li.isSwing = False
    li.swingPhase = 0.0
    li.spprtPhase = 0.0
    li.contactPos = pos.copy(); li.prevContact = pos.copy(); li.contactPreFilter = pos.copy(); li.contactFilter = pos.copy()
    li.dotContactFilter = np.zeros(3)
    li.ball = [ce.SDS(30.0, 0.001, 1.0, 1e7, ce.RL_STEP / 2) for _ in range(3)]
    for k in range(3):
        li.ball[k].x[:] = (pos[k], 0.0)
    li.filtQ = ce.SDS(30.0, 0.001, 1.0, 1e7, ce.RL_STEP / 2)
    li.filtQorigin = tl.qaxis(tl.Y, yaw)
    li.rotY = li.filtQorigin.copy()
    li.contactFilterQ = tl.QI.copy()
    li.earlyTouchDown = None; li.lateTouchDown = None
    return li


@pytest.mark.parametrize("clip", ["walk2", "run2"])
def test_contact_filter_replays_the_recorded_rollouts(rec, env, clip):
    """startSwing / filterContact('sw') / getFilteredFootPos (deepmimiccdm-v1.lua:441-599,1865-1903) through the oracle:
    every swing row of the recorded walk and of the recorded run (run tables, run2-v1 parameters; filterEndPhase is 0.4 for both, :524), both
    feet, positions < 5e-6 m (float32 storage), foot yaw < 1e-7 rad."""
    g, f, rf = rec[clip + "_globalTraj"], rec[clip + "_rawFootInfo"], rec[clip + "_refFrames"].ravel().copy()
    if clip == "run2":
        tab, skel = hc.load_tables("run")
        env = ce.OracleCDMEnv(tab, skel, params=dict(ce.RUN2_PARAMS, **ce.RUN2_GUI))
        rf = np.where(np.concatenate([[False], np.cumsum(np.diff(rf) < 0) > 0]), rf + int(tab["nf"]), rf)     # stitched frame index
    n = len(g)
    worst_p = worst_y = 0.0
    rows = 0
    for leg in (0, 1):
        sup, cpos, sw = f[:, leg], f[:, 2 + 3 * leg:5 + 3 * leg], f[:, 8 + leg]
        mine = g[:, 10:13] if leg == 0 else g[:, 7:10]
        other = g[:, 7:10] if leg == 0 else g[:, 10:13]
        yaw = g[:, 19 + leg]
        r = 1
        while r < n:
            if not (sup[r - 1] == 1 and sup[r] == 0):
                r += 1
                continue
            li = _new_leg(env, leg == 1, mine[r - 1], yaw[r - 1])
            oli = _new_leg(env, leg == 0, other[r - 1], 0.0)
            env.leg = {1 + leg: li, 2 - leg: oli}
            first = True
            while r < n and sup[r] == 0:
                Q, p = g[r, 3:7], g[r, 0:3]
                rotY = tl.rotation_y(Q)
                gv = g[r, 16:19]
                env.iframe = float(rf[r])                         # cycle frame; the rtt tables of the first cycle are indexed by it
                oli.isSwing = f[r, 1 - leg] == 0
                oli.contactFilter = other[r].copy()
                if first:
                    env.start_swing(li)                            # :441-468 -> filterContact 'ssw'
                    first = False
                li.swingPhase = sw[r]
                li.contactPos = cpos[r].copy()                     # recorded after the leg-length clamp
                env.filter_contact_sw(li, rotY, gv)
                pos, ori = env.filtered_foot_pos(Q, 1 + leg)
                worst_p = max(worst_p, float(np.abs(pos - mine[r]).max()))
                worst_y = max(worst_y, abs(tl.rotation_angle_about_axis(ori, tl.Y) - yaw[r]))
                rows += 1
                r += 1
    assert rows > 500
    assert worst_p < 5e-6 and worst_y < 1e-7, (worst_p, worst_y)


def test_gait_state_machine_matches_the_recorded_walk(rec, env):
    """Swing starts exactly where rttTouchOff <= 0 (Q8 tables, defaultRLsetting.lua:291-314,395-418); touch-down where
    rttTouchDown <= 0 (the recorded GUI run has touchDownHeight / limited leg length enabled, which moves a few
    touch-downs by a step, :1150-1200; those are counted and bounded)."""
    f, rf = rec["walk2_rawFootInfo"], rec["walk2_refFrames"].astype(int)
    tab = env.tab
    early = 0
    for leg in (0, 1):
        sup = f[:, leg]
        for r in range(1, len(f)):
            if sup[r - 1] == 1:
                assert (tab["rtt_off"][leg][rf[r]] <= 0) == (sup[r] == 0), (leg, r)
            elif (tab["rtt_down"][leg][rf[r]] <= 0) != (sup[r] == 1):
                early += 1
        sw = f[:, 8 + leg]
        inc = np.diff(sw)[(sup[1:] == 0) & (sup[:-1] == 0)]
        assert np.allclose(inc, ce.RL_STEP, atol=1e-12)            # swingPhase advances by RL_step
    assert early <= 4


def test_gait_state_machine_matches_the_recorded_run(rec):
    """run2-v1: every swing start and touch-down of the roll-out the reference recorded (refTraj_run2-v1.dat, 700 rows) falls
    where the rtt tables of the running clip (cdm_tables.RUN4 through build_tables) say -- including the right foot's
    one-step touch-down at each loop boundary, which comes from the stitching code flagging touchDown[1] for both legs
    (module/CDMTraj.lua:1506, kept as it is in cdm_tables.build_tables).  refFrames are recorded cycle-relative; the table
    index is the stitched frame (cycle frame + cycle length after the first wrap, defaultRLsetting.lua:239-262)."""
    tab, _ = hc.load_tables("run")
    f, rf = rec["run2_rawFootInfo"], rec["run2_refFrames"].ravel().astype(int)
    nf = int(tab["nf"])
    wrapped = np.concatenate([[False], np.cumsum(np.diff(rf) < 0) > 0])
    idx = np.where(wrapped, rf + nf, rf)
    n_off = n_down = 0
    for leg in (0, 1):
        sup = f[:, leg]
        for r in range(1, len(f)):
            if sup[r - 1] == 1:
                assert (tab["rtt_off"][leg][idx[r]] <= 0) == (sup[r] == 0), (leg, r)
                n_off += int(sup[r] == 0)
            else:
                assert (tab["rtt_down"][leg][idx[r]] <= 0) == (sup[r] == 1), (leg, r)
                n_down += int(sup[r] == 1)
    assert n_off >= 45 and n_down >= 45
    boundary = [r for r in range(1, len(f)) if rf[r] < rf[r - 1]]
    assert len(boundary) >= 15 and all(f[r, 0] == 1 and f[r - 1, 0] == 0 and f[r + 1, 0] == 0 for r in boundary[:-1])


def test_flight_substeps_replay_the_recorded_run(rec, env):
    """Rows of the recorded running clip with both feet in swing: QPsim:frameMove without contacts (gravity 9.8, body-frame
    bias, two semi-implicit sub-steps of 1/120 s, q <- normalize(q + q' dt)) through the oracle's frame_move."""
    g, f = rec["run2_globalTraj"], rec["run2_rawFootInfo"]
    fl = np.nonzero((f[:-1, 0] == 0) & (f[:-1, 1] == 0))[0]
    assert len(fl) > 100
    worst = dict(v=0.0, w=0.0, q=0.0, p=0.0)
    for r in fl:
        env.p, env.Q = g[r, 0:3].copy(), g[r, 3:7].copy()
        env.v = g[r, 16:19].copy()
        env.wb = tl.qrot(tl.qinv(env.Q), g[r, 13:16])             # dq = [w_world, v_world]  (:1064)
        env.frame_move([], np.concatenate([env.p, env.Q]), np.zeros(7), np.zeros(6))
        worst["v"] = max(worst["v"], float(np.abs(env.v - g[r + 1, 16:19]).max()))
        worst["w"] = max(worst["w"], float(np.abs(tl.qrot(env.Q, env.wb) - g[r + 1, 13:16]).max()))
        worst["q"] = max(worst["q"], float(np.abs(tl.qalign(env.Q, g[r + 1, 3:7]) - g[r + 1, 3:7]).max()))
        worst["p"] = max(worst["p"], float(np.abs(env.p - g[r + 1, 0:3]).max()))
    assert worst["v"] < 2e-7 and worst["w"] < 2e-7 and worst["q"] < 2e-7 and worst["p"] < 2e-5, worst


def test_stance_substeps_are_consistent_with_the_recorded_walk(rec):
    """With contacts the action is unknown, but the integrator is still visible: p' - p = h (v_mid + v'), and the
    semi-implicit order (velocity first) fits the recording an order of magnitude better than the explicit one."""
    g = rec["walk2_globalTraj"]
    h = ce.TIMESTEP_QP
    p, v = g[:, 0:3], g[:, 16:19]
    dp = p[1:] - p[:-1]
    semi = np.abs(dp - h * (1.5 * v[1:] + 0.5 * v[:-1])).mean()
    expl = np.abs(dp - h * (1.5 * v[:-1] + 0.5 * v[1:])).mean()
    swapped = np.abs(dp - h * (1.5 * g[1:, 13:16] + 0.5 * g[:-1, 13:16])).mean()
    assert semi < 6e-5 and expl > 4 * semi and swapped > 100 * semi


def test_binary_table_reader_against_the_dof_reader(tmp_path):
    """tl_binary (BinaryFile / util.saveTable reader) on a stream written here in the same format."""
    import struct
    from ipcdnnwalk_b200 import tl_binary as tb
    m = np.arange(6, dtype=np.float64).reshape(2, 3)
    b = struct.pack("<ii", 0, 1)                                   # numRef = 1
    b += struct.pack("<ii", 0, 1) + struct.pack("<ii", 0, 0)      # entry follows, not a reference
    b += struct.pack("<ii", 0, 1) + struct.pack("<ii", 8, 2) + b"mm"          # key: string "mm"
    b += struct.pack("<ii", 0, 5) + struct.pack("<iii", 5, 2, 3) + m.tobytes()    # value: matrixn
    b += struct.pack("<ii", 0, 1) + struct.pack("<ii", 0, 0)
    b += struct.pack("<ii", 0, 0) + struct.pack("<i", 1) + struct.pack("<d", 3.0)       # key: number 3
    b += struct.pack("<ii", 0, 0) + struct.pack("<i", 1) + struct.pack("<d", 0.25)      # value: number
    b += struct.pack("<ii", 0, 0)                                  # end of table
    fn = tmp_path / "t.dat"
    fn.write_bytes(b)
    t = tb.load_table(str(fn))
    assert np.array_equal(t["mm"], m) and t[3] == 0.25
