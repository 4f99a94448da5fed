"""CPU tests that pin the oracle (analytic properties -- the reference ships no golden vectors)."""
import math
import os

import numpy as np
import pytest
from scipy.linalg import expm

import helpers
from oracle import mj_freebody as mj
from oracle import srb_env as E
from oracle import sds_filter as S


def _rollout_env(variant="test", steps=6, seed=3, sigma=0.1):
    rng = np.random.default_rng(seed)
    clips = helpers.fresh_clips()
    env = E.OracleSRBEnv(helpers.fresh_clips(), variant)
    helpers.oracle_reset(env, helpers.random_plan(rng, clips))
    return env, rng


def test_qp_kkt_of_uneliminated_problem():
    """The NNLS-after-elimination solve satisfies the KKT system of the QP exactly as the reference
    states it (viewer.py:549-573): variables [qdd, lam], equality M qdd - C lam = -b, lam >= 0."""
    env, rng = _rollout_env()
    checked = 0
    orig = E.solve_qp_for_contact_forces

    def spy(m, d, T, tv, sites, mu, w_lambda=0.001):
        nonlocal checked
        qdd, lam_list, forces = orig(m, d, T, tv, sites, mu, w_lambda)
        M, b, B, C, qdd_d = E.qp_matrices(m, d, T, tv, sites, mu)
        lam = np.concatenate(lam_list)
        assert np.allclose(M @ qdd - C @ lam, -b, rtol=0, atol=1e-8)           # primal equality
        assert (lam >= 0).all()
        # stationarity: 2(qdd-qdd_d) + M^T eta = 0 ; 2 w lam - C^T eta - z = 0, z >= 0, z.lam = 0
        eta = -np.linalg.solve(M.T, 2 * (qdd - qdd_d))
        z = 2 * w_lambda * lam - C.T @ eta
        scale = 1 + np.abs(z).max()
        assert (z >= -1e-8 * scale).all(), z.min()
        assert abs(z @ lam) <= 1e-7 * scale * (1 + lam.sum())
        for f, l in zip(forces, lam_list):
            assert np.allclose(f, B @ l)
        checked += 1
        return qdd, lam_list, forces
    E.solve_qp_for_contact_forces = spy
    try:
        for _ in range(8):
            env.step(helpers.tracking_action(env, rng, 0.2))
    finally:
        E.solve_qp_for_contact_forces = orig
    assert checked >= 8


def test_qp_independent_solver_dual_newton():
    """Second, independent algorithm (semismooth Newton on the 6-dim dual, the one the CUDA kernel uses)
    agrees with the Lawson-Hanson solve."""
    env, rng = _rollout_env(seed=5)
    w = 1e-3

    def dual_newton(A, y):
        nu = -y.copy()
        phi = lambda v: 0.5 * v @ v + v @ y + (np.minimum(A.T @ v, 0) ** 2).sum() / (2 * w)
        S = (A.T @ nu) < 0
        for _ in range(50):
            As = A[:, S]
            nn = np.linalg.solve(np.eye(6) + As @ As.T / w, -y)
            Sn = (A.T @ nn) < 0
            if (Sn == S).all():
                nu = nn
                break
            a = 1.0
            while phi(nu + a * (nn - nu)) > phi(nu) and a > 1e-8:
                a *= 0.5
            nu = nu + a * (nn - nu)
            S = (A.T @ nu) < 0
        return np.maximum(0, -A.T @ nu) / w

    orig = E.solve_qp_for_contact_forces
    n = 0

    def spy(m, d, T, tv, sites, mu, w_lambda=0.001):
        nonlocal n
        out = orig(m, d, T, tv, sites, mu, w_lambda)
        M, b, B, C, qdd_d = E.qp_matrices(m, d, T, tv, sites, mu)
        Minv = 1 / np.diag(M)
        lam2 = dual_newton(Minv[:, None] * C, qdd_d + Minv * b)
        lam = np.concatenate(out[1])
        assert np.abs(lam - lam2).max() <= 1e-9 * (1 + np.abs(lam).max())
        n += 1
        return out
    E.solve_qp_for_contact_forces = spy
    try:
        for _ in range(6):
            env.step(helpers.tracking_action(env, rng, 0.3))
    finally:
        E.solve_qp_for_contact_forces = orig
    assert n >= 6


def test_free_flight_matches_closed_form():
    """No contact: z acceleration is -g*m/(m+armature) with linear damping (Q5), quaternion stays unit,
    and xmat lags qpos by exactly one sub-step (Q14)."""
    m = mj.FreeBodyModel()
    d = mj.FreeBodyData(m)
    d.qvel = np.array([0.3, -0.2, 0.5, 0.4, -0.7, 0.2])
    mj.mj_forward(m, d)
    h = m.timestep
    vz = d.qvel[2]
    for k in range(20):
        q_before = d.qpos[3:7].copy()
        d.qfrc_applied = np.zeros(6)
        mj.mj_step1(m, d)
        mj.mj_step2(m, d)
        vz = vz + h * ((-m.mass * 9.81 - 1.0 * vz) / (61.0 + h * 1.0))
        assert abs(d.qvel[2] - vz) < 1e-13
        assert abs(np.linalg.norm(d.qpos[3:7]) - 1) < 1e-14
        assert np.allclose(d.xmat.reshape(3, 3), mj.quat2mat(q_before / np.linalg.norm(q_before)), atol=1e-15)


def test_jacobian_matches_finite_difference():
    m = mj.FreeBodyModel()
    d = mj.FreeBodyData(m)
    rng = np.random.default_rng(0)
    q = rng.standard_normal(4); q /= np.linalg.norm(q)
    d.qpos = np.concatenate([rng.standard_normal(3), q])
    mj.mj_kinematics(m, d)
    point = d.xpos + np.array([0.3, -0.2, -0.9])
    J = mj.mj_jacp(m, d, point)
    R = d.xmat.reshape(3, 3)
    local = R.T @ (point - d.xpos)
    eps = 1e-7
    for j in range(6):
        v = np.zeros(6); v[j] = 1.0
        d2 = mj.FreeBodyData(m)
        d2.qpos = d.qpos.copy(); d2.qvel = v
        d2.qpos[:3] = d.qpos[:3] + eps * v[:3]
        d2.qpos[3:7] = mj.quat_integrate(d.qpos[3:7], v[3:6], eps)
        mj.mj_kinematics(m, d2)
        p2 = d2.xpos + d2.xmat.reshape(3, 3) @ local
        assert np.allclose((p2 - point) / eps, J[:, j], atol=1e-6)


def test_log_se3_roundtrip():
    rng = np.random.default_rng(1)
    for _ in range(20):
        w = rng.standard_normal(3) * rng.uniform(0, 2.5)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        T_hat = np.eye(4); T_hat[:3, :3] = expm(K); T_hat[:3, 3] = rng.standard_normal(3)
        tw = E.log_se3(np.eye(4), T_hat)
        X = np.zeros((4, 4)); X[:3, :3] = np.array([[0, -tw[5], tw[4]], [tw[5], 0, -tw[3]], [-tw[4], tw[3], 0]]); X[:3, 3] = tw[:3]
        assert np.allclose(expm(X), T_hat, atol=1e-9)


def test_sds_preset_gain_lookup():
    assert np.allclose(S.k_preset(1e9), [31563.832, 1887.117])
    assert np.allclose(S.k_preset(1e11), [316168.772, 6099.859])
    assert np.allclose(S.k_preset(1e10), [99941.017, 3403.601])
    k = S.k_preset(10 ** 9.5)                       # linear in Q between decades, float32-rounded weight
    t = float(np.float32((10 ** 9.5 - 1e9) / 9e9))
    assert np.allclose(k, S.K_PRESET[4] * (1 - t) + S.K_PRESET[5] * t, rtol=1e-12)
    b = S.SDS(60.0, 0.001, 1, 1e9, 0, 1 / 120.0, True)
    b.xd[0] = 1.0
    for _ in range(400):
        b.singleStep()
    assert abs(b.x[0] - 1.0) < 1e-3 and abs(b.x[1]) < 1e-3


@pytest.mark.parametrize("variant", ["test", "training"])
def test_oracle_reproduces_golden_rollout(variant):
    """Regression pin: the committed golden vectors are what the oracle computes today."""
    g = np.load(os.path.join(helpers.GOLDEN, f"rollout_{variant}.npz"))
    for e in (0, 1):
        env = E.OracleSRBEnv(helpers.fresh_clips(), variant)
        obs0, _ = helpers.oracle_reset(env, g["plans"][e])
        assert np.abs(obs0 - g["obs0"][e]).max() < 1e-12
        for k in range(8):
            obs, r, done, _, info = env.step(g["actions"][e, k])
            assert np.allclose(obs, g["obs"][e, k], rtol=0, atol=1e-6)
            assert abs(r - g["rew"][e, k]) < 1e-9 * max(1, abs(r))
            assert done == g["done"][e, k]


def test_reset_quirk_q2_right_flag_is_left_or_right():
    clips = helpers.fresh_clips()
    env = E.OracleSRBEnv(helpers.fresh_clips(), "test")
    fp = clips[1]["feet_pos_list"]
    idx = [f for f in range(600) if fp[f, 0, 2] == 0 and fp[f, 1, 2] != 0]
    assert idx, "need a frame with left stance / right swing"
    env.reset_explicit(1, idx[0], None, None)
    assert env.left_foot_in_contact and env.right_foot_in_contact


@pytest.mark.skipif(not os.path.exists("/root/reference/gym_trackSRB"), reason="reference mount not present")
def test_clip_tables_regenerate_from_reference_data():
    from oracle.srbdata import SRBdata, load_mocap_txt
    R = "/root/reference/gym_trackSRB/"
    md = load_mocap_txt(R + "motiondata_mujoco_refined/aiming_show.txt")
    s = SRBdata(md[:40], R + "ref_contact/aiming_show.bvh.constraint.npz", R + "Characters/humanoid_deepmimic_withpalm_Tpose.xml")
    s.humanoid2SRB()
    assert s.cycle_length == 40 and s.mocap_com_humanoid.shape == (79, 7)
    full = helpers.load_clip("aiming_show")
    # first cycle of the truncated clip equals the first 40 frames of the full table
    assert np.allclose(s.mocap_com_humanoid[:40], full["mocap_qpos"][:40], atol=1e-12)
    # humanoid total mass = 45 kg (explicit geom masses), CoM inside the body
    assert abs(sum(g[2] for b in s.model.bodies for g in b["geoms"]) - 45.0) < 1e-12
    # stitched second cycle continues from the end of the first (yaw-aligned): no position jump
    jump = np.linalg.norm(s.mocap_com_humanoid[40, :2] - s.mocap_com_humanoid[39, :2])
    assert jump < 0.2
