"""CPU pins of the AdaptiveSRB oracle (oracle/cdm_env.py, oracle/tl_math.py): the reference ships no expected outputs
for this path (SURVEY 4, 8c -- PARITY UNPINNED), so the restatement is pinned by analytic properties and by a regression
check against the committed golden roll-out."""
import math

import numpy as np
import pytest

import helpers_cdm as hc
from oracle import tl_math as tl
from oracle import cdm_env as ce


@pytest.fixture(scope="module")
def data():
    tab, skel = hc.load_tables()
    return tab, skel, dict(np.load(hc.ROLLOUT))


def test_lqr_gain_solves_the_riccati_equation():
    """Physics.LQR (PhysicsLib/SDRE.cpp:100-245): K = R^-1 B^T P with P the stabilising solution of A^T P + P A - P B B^T P + Q = 0."""
    M, b, k, q = 30.0, 0.001, 1.0, 1e7
    K0, K1 = ce.lqr_gain(M, b, k, q)
    A = np.array([[0.0, 1.0], [-k / M, -b / M]]); B = np.array([[0.0], [1.0 / M]]); Q = np.diag([q, 0.0])
    import scipy.linalg
    P = scipy.linalg.solve_continuous_are(A, B, Q, np.eye(1))
    K = (B.T @ P).ravel()
    assert np.allclose([K0, K1], K, rtol=1e-9)
    assert (np.linalg.eigvals(A - B @ K[None, :]).real < 0).all()


def test_qp_kkt_of_the_uneliminated_problem():
    """_QPsolve as stated (QPservo_v3.lua:1058-1463): variables [qdd, tau, lam], cost w_ddq|qdd-a|^2 + w_tau|tau|^2 +
    w_lam|lam|^2, M qdd - tau - JtV lam = -bias, tau = 0, lam >= 0.  Check stationarity / complementarity of the
    oracle's answer in that un-eliminated form, with every stance point present twice (Q7)."""
    rng = np.random.default_rng(3)
    P = ce.PARAMS
    Md = np.array([ce.INERTIA[0], ce.INERTIA[1], ce.INERTIA[2], ce.MASS, ce.MASS, ce.MASS])
    for trial in range(20):
        Q = tl.qnormalize(np.array([1.0, 0, 0, 0]) + rng.uniform(-0.2, 0.2, 4))
        p = np.array([0.0, 0.9, 0.0]) + rng.uniform(-0.05, 0.05, 3)
        a_des = rng.uniform(-1, 1, 6) * np.array([20, 20, 20, 5, 5, 5])
        bias = np.concatenate([rng.uniform(-1, 1, 3), tl.qrot(tl.qinv(Q), np.array([0, 60 * 9.8, 0]))])
        pts = [np.array([0.1, 0, 0.1]), np.array([0.1, 0, -0.14]), np.array([-0.12, 0, 0.3]), np.array([-0.12, 0, 0.06])][:2 + 2 * (trial % 2)]
        contacts = [c for c in pts for _ in range(2)]
        qdd, lam = ce.solve_contact_qp(Q, p, a_des, bias, contacts, P)
        lam = lam.ravel()
        normals = ce.contact_basis(P["inv_fric"])
        qi = tl.qinv(Q)
        JtV = np.array([np.concatenate([tl.qrot(qi, np.cross(c - p, n)), tl.qrot(qi, n)]) for c in contacts for n in normals]).T
        assert np.allclose(Md * qdd - JtV @ lam, -bias, atol=1e-8)                      # equality rows with tau = 0
        mu = 2 * P["w_ddq"] * (qdd - a_des) / Md                                     # multiplier of the equality rows
        g = 2 * P["w_lam"] * lam + JtV.T @ mu                                        # d/d lam of the Lagrangian
        assert (lam >= 0).all()
        scale = 1 + np.abs(JtV.T @ mu).max()
        assert (g >= -1e-7 * scale).all() and np.abs(g[lam > 1e-9]).max(initial=0.0) <= 1e-7 * scale
        assert np.allclose(lam.reshape(-1, 5)[0::2], lam.reshape(-1, 5)[1::2], rtol=1e-6, atol=1e-6 * (1 + lam.max()))   # duplicates share


def test_free_fall_without_contacts(data):
    tab, skel, _ = data
    env = ce.OracleCDMEnv(tab, skel)
    env.reset()
    env.p[1] = 5.0                                   # above max_contact_y: the QP has no columns
    v0, y0 = env.v.copy(), env.p[1]
    theta_d = np.concatenate([env.p, env.Q]); dtheta_d = np.zeros(7)
    env.P["k_p"] = 0.0; env.P["k_d"] = 0.0
    env.wb[:] = 0.0
    env.frame_move([], theta_d, dtheta_d, np.zeros(6))
    h = ce.TIMESTEP_QP
    assert np.allclose(env.v, v0 + np.array([0, -9.8 * 2 * h, 0]), atol=1e-12)
    assert abs(env.p[1] - (y0 + h * (v0[1] - 9.8 * h) + h * (v0[1] - 2 * 9.8 * h))) < 1e-12


def test_markers_agree_with_full_body_fk(data):
    """The leg-chain columns the device table keeps reproduce the ankle markers of the full-body FK."""
    from ipcdnnwalk_b200 import cdm_tables as ct
    from oracle.bone_fk import bone_forward_kinematics
    tab, skel, _ = data
    rows = ct.pack_rows(tab)
    names = skel["names"]
    for f in (0, 7, 100, 500):
        rot, tr = bone_forward_kinematics(skel, tab["fullbody"][f])
        r = rows[f, 19:40]
        q = r[3:7]; t = r[0:3] + np.array(skel["offset"][0])
        col = 7
        for chain, ankle in ((("LeftHip", "LeftKnee", "LeftAnkle"), "LeftAnkle"), (("RightHip", "RightKnee", "RightAnkle"), "RightAnkle")):
            qq, tt = q.copy(), t.copy()
            for n in chain:
                b = names.index(n)
                tt = tl.qrot(qq, np.array(skel["offset"][b])) + tt
                lq = np.array([1.0, 0, 0, 0])
                for c in skel["channels"][b]:
                    ax = np.zeros(3); ax["XYZ".index(c)] = 1.0
                    lq = tl.qmult(lq, tl.qaxis(ax, r[col])); col += 1
                qq = tl.qmult(qq, lq)
            b = names.index(ankle)
            assert np.allclose(tt, tr[b], atol=1e-12) and np.allclose(qq, rot[b], atol=1e-12)


def test_rotation_y_decomposition():
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = tl.qnormalize(rng.normal(size=4))
        ry = tl.rotation_y(q)
        assert abs(ry[1]) < 1e-12 and abs(ry[3]) < 1e-12                       # pure yaw
        off = tl.offset_q(q)
        assert np.allclose(tl.qmult(ry, off), q, atol=1e-12)
        assert abs(tl.qrot(off, tl.Y).dot(np.cross(tl.Y, tl.qrot(off, tl.Y)))) < 1e-12


def test_sample_row_float32_fraction_and_snapping():
    m = np.arange(20, dtype=float).reshape(10, 2) ** 2
    assert np.array_equal(tl.sample_row(m, 3.004), m[3]) and np.array_equal(tl.sample_row(m, 3.996), m[4])
    f = float(np.float32(3.3 - 3.0))
    assert np.array_equal(tl.sample_row(m, 3.3), m[3] * (1.0 - f) + m[4] * f)


def test_integer_tables_use_truncation(data):
    """Q8: rttTouchDown / rttTouchOff are indexed with (int) of the real-valued frame."""
    tab, skel, _ = data
    env = ce.OracleCDMEnv(tab, skel)
    env.reset()
    env.iframe = 17.999
    assert env.rtt_off(2) == int(tab["rtt_off"][1][17]) and env.rtt_down(1) == int(tab["rtt_down"][0][17])


def test_golden_rollout_regression(data):
    tab, skel, roll = data
    n = roll["plans0"].shape[0]
    envs = [ce.OracleCDMEnv(tab, skel) for _ in range(n)]
    for i, e in enumerate(envs):
        assert np.allclose(hc.oracle_reset(e, roll["plans0"][i]), roll["obs0"][i], atol=1e-12)
    for t in range(40):
        for i, e in enumerate(envs):
            o, r, d = e.step(roll["acts"][t, i])
            assert np.allclose(o, roll["obs"][t, i], rtol=1e-9, atol=1e-10) and abs(r - roll["rew"][t, i]) < 1e-10 and d == roll["done"][t, i]
            if d:
                hc.oracle_reset(e, roll["plans"][t, i])
    assert roll["done"].sum() >= 5 and (roll["ints"][:, :, 0] != 0).sum() > 100       # the fixture exercises resets and swing phases


def test_zero_action_episode_survives_a_gait_cycle(data):
    """Property: without noise and with zero action the tracking controller keeps the body up for more than a cycle."""
    tab, skel, roll = data
    assert not roll["done"][:64, 0].any()
    assert (roll["reals"][:64, 0, 1] > 0.75).all() and (roll["reals"][:64, 0, 1] < 1.0).all()
