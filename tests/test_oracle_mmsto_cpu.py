"""Analytic pins of the MMSTO oracle (oracle/mmsto.py; PARITY UNPINNED against the reference, see its header)."""
import numpy as np
import pytest

import helpers_mmsto as hm
from oracle import bone_fk as fk
from oracle import mmsto as mm


@pytest.fixture(scope="module")
def skel():
    return hm.skeleton()


def test_euler_tree_equals_the_quaternion_root_tree(skel):
    """Same bone frames, CoM and momentum as LoaderToTree with a quaternion root (oracle/bone_fk.py, rows a-22/a-23) when the
    root pose / velocity are converted: R = Ry Rz Rx, w = sum of the hinge axes times their rates."""
    from ipcdnnwalk_b200.short_traj_opt import _yzx_to_quat
    rng = np.random.default_rng(0)
    t = mm.EulerTree(skel)
    assert t.ndof == 59
    q = rng.uniform(-.6, .6, t.ndof); dq = rng.uniform(-1, 1, t.ndof)
    t.set_q(q); t.set_dq(dq)
    pose = np.concatenate([q[0:3], _yzx_to_quat(q[3:6]), q[6:]])
    w_world = sum(t.W[j] * dq[j] for j in (3, 4, 5))
    rot, tr, W, V = fk.loader_to_tree(skel, pose, np.concatenate([w_world, dq[0:3], dq[6:]]))
    for b in range(t.nb):
        assert np.allclose(fk.vrotate(rot[b], np.array([0.3, -0.2, 0.5])) + tr[b], t.global_pos(b, [0.3, -0.2, 0.5]), atol=1e-13)
    com, mom = fk.calc_momentum_com(skel, rot, tr, W, V)
    assert np.allclose(com, t.calc_com(), atol=1e-13)
    assert np.allclose(mom, t.calc_momentum_com(), atol=1e-11)


def test_momentum_jacobian_is_the_momentum_map(skel):
    """calcMomentumJacobianTranspose (NodeWrap.cpp:107-160) restated literally must reproduce calcMomentumCOM: J^T' dq = momentum."""
    rng = np.random.default_rng(1)
    t = mm.EulerTree(skel)
    for _ in range(3):
        q = rng.uniform(-.8, .8, t.ndof); dq = rng.uniform(-2, 2, t.ndof)
        t.set_q(q); t.set_dq(dq)
        assert np.allclose(t.calc_momentum_jacobian_T().T.dot(dq), t.calc_momentum_com(), atol=1e-11)


def test_point_and_com_jacobians_by_finite_differences(skel):
    rng = np.random.default_rng(2)
    t = mm.EulerTree(skel); t2 = mm.EulerTree(skel)
    q = rng.uniform(-.5, .5, t.ndof)
    t.set_q(q)
    Jc = t.calc_com_jacobian_T(); Jp = t.jacobian_T_at(13, [0.05, 0.01, -0.02])
    e = 1e-6
    for j in range(t.ndof):
        qq = q.copy(); qq[j] += e; t2.set_q(qq)
        assert np.allclose((t2.calc_com() - t.calc_com()) / e, Jc[j], atol=2e-6)
        assert np.allclose((t2.global_pos(13, [0.05, 0.01, -0.02]) - t.global_pos(13, [0.05, 0.01, -0.02])) / e, Jp[j], atol=2e-6)
    g = np.zeros(t.ndof); dS = np.array([0.3, -1.0, 2.0])
    t.update_grad_S_JT(g, dS, 13, [0.05, 0.01, -0.02])
    assert np.allclose(g, Jp.dot(dS), atol=1e-14)


def test_objective_gradient_by_finite_differences(skel):
    """Every term except the momentum one has an exact gradient; the momentum gradient goes through the cached mean Jacobian
    (`useMomentumCache`, "slightly inaccurate but much faster", shortTermOptimizer.lua:380) and is only close."""
    p = hm.make_problem(1, skel=skel)
    rng = np.random.default_rng(3)
    x = (p["x_original"] + rng.normal(0, 0.01, p["x_original"].shape)).ravel()
    idx = rng.choice(x.size, 24, replace=False)

    def fd(o):
        out = []
        for j in idx:
            xp = x.copy(); xp[j] += 1e-6; xm = x.copy(); xm[j] -= 1e-6
            out.append((o.evaluate(xp)[0] - o.evaluate(xm)[0]) / 2e-6)
        return np.array(out)
    o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel), weightEEcon=1000.0, weightMomentum=0.0)
    f, g = o.evaluate(x)
    assert np.abs(fd(o) - g[idx]).max() < 1e-8 * np.abs(g).max()
    o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel), weightEEcon=1000.0)
    f, g = o.evaluate(x)
    assert np.abs(fd(o) - g[idx]).max() < 2e-2 * np.abs(g).max()


def test_gauss_filter_known_values():
    """Filter::GetGaussFilter(7): exp(-i^2/7) normalised, symmetric; even sizes round up to the next odd size."""
    a = mm.gauss_filter(7)
    ref = np.exp(-np.arange(-3, 4) ** 2 / 7.0)
    assert np.allclose(a, ref / ref.sum(), rtol=1e-15) and abs(a.sum() - 1) < 1e-15
    assert len(mm.gauss_filter(6)) == 7


def test_lbfgs_on_known_problems():
    """libLBFGS restatement: exact minimiser of a convex quadratic, Rosenbrock (the sample of lbfgs.h), iteration cap code."""
    rng = np.random.default_rng(4)
    A = rng.normal(size=(12, 12)); A = A.dot(A.T) + np.eye(12); b = rng.normal(size=12)
    x, fx, ret, k = mm.lbfgs(np.zeros(12), lambda z: (0.5 * z.dot(A).dot(z) - b.dot(z), A.dot(z) - b), epsilon=1e-6)
    assert ret == mm.LBFGS_SUCCESS and np.allclose(x, np.linalg.solve(A, b), atol=1e-6)

    def rosen(z):
        f = 0.0; g = np.zeros_like(z)
        for i in range(0, len(z), 2):
            t1 = 1.0 - z[i]; t2 = 10.0 * (z[i + 1] - z[i] * z[i])
            g[i + 1] = 20.0 * t2; g[i] = -2.0 * (z[i] * g[i + 1] + t1); f += t1 * t1 + t2 * t2
        return f, g
    z0 = np.array([-1.2, 1.0] * 5)
    x, fx, ret, k = mm.lbfgs(z0, rosen, epsilon=1e-7)
    assert ret == mm.LBFGS_SUCCESS and np.allclose(x, 1.0, atol=1e-5) and fx < 1e-10
    x, fx, ret, k = mm.lbfgs(z0, rosen, epsilon=1e-7, max_iterations=3)
    assert ret == mm.LBFGSERR_MAXIMUMITERATION and k == 3


def test_solve_tracks_consistent_targets(skel):
    """Targets generated from the desired motion itself: the optimum is the desired motion; frames 0 and 1 are the (clamped) input."""
    p = hm.make_problem(5, target_noise=0.0, velcon=False, skel=skel)
    o = hm.oracle_opt(p, skel)
    start = p["x_original"].copy(); start[2:] += np.random.default_rng(6).normal(0, 0.003, start[2:].shape)
    o.x_original = start                                   # the initial guess is x_original (:826)
    o.dx = p["dx"]; o.ddx = p["ddx"]
    mot = p["x_original"].copy()
    f0 = o.evaluate(o.initial_guess(mot).ravel())[0]
    X, info = o.solve(mot, max_iterations=25)
    assert info["fx"] < 0.05 * f0
    assert np.array_equal(X[:2], mot[:2])
    # the 25 degree clamp of the constant frames
    mot2 = mot.copy(); mot2[0, 10] += 1.0
    assert abs(o.initial_guess(mot2)[0, 10] - (start[0, 10] + np.radians(25))) < 1e-15


def _com_slide_case(skel, seed, nvar=8):
    from ipcdnnwalk_b200.short_traj_opt import _yzx_to_quat
    rng = np.random.default_rng(seed)
    q = hm.make_problem(seed, nf=nvar, skel=skel)["x_original"]
    pose = np.concatenate([q[:, 0:3], _yzx_to_quat(q[:, 3:6]), q[:, 6:]], 1)
    com_srb = np.cumsum(rng.normal(0, 0.02, (nvar, 3)), 0) + [0, 0.9, 0]
    des = np.cumsum(rng.normal(0, 0.02, (nvar, 3)), 0) + [0.1, 0.9, 0]
    return pose, com_srb, des


def test_remove_com_sliding_satisfies_its_constraints(skel):
    """removeCOMsliding_online: first two CoMs kept, sum of the rest equals the sum of the desired CoMs, the acceleration of
    the optimised CoM follows com_srb up to the constraint forces (KKT), and the moved poses have the optimised CoM."""
    pose, com_srb, des = _com_slide_case(skel, 3)
    out, curr, opt = mm.remove_com_sliding_online(skel, com_srb, des, pose)
    nvar = len(pose)
    assert np.allclose(opt[:2], curr[:2], atol=1e-13)
    assert np.allclose(opt[1:].sum(0), des[:nvar - 1].sum(0), atol=1e-12)
    # stationarity on the unconstrained-by-position variables 2..nvar-1: gradient of the smoothness objective is the same
    # constant (the multiplier of the sum constraint) for all of them
    acc = lambda c: -c[:-2] + 2 * c[1:-1] - c[2:]
    e = acc(opt) - acc(com_srb)                                      # [nvar-2, 3]
    g = np.zeros((nvar, 3))
    for v in range(1, nvar - 1):
        g[v - 1] += -e[v - 1]; g[v] += 2 * e[v - 1]; g[v + 1] += -e[v - 1]
    assert np.allclose(g[2:] - g[2], 0, atol=1e-9)
    for i in (0, 3, nvar - 1):
        rot, tr = fk.bone_forward_kinematics(skel, out[i])
        assert np.allclose(fk.calc_com(skel, rot, tr), opt[i], atol=1e-12)
    # fix_y_also = False keeps the heights; a ground height changes only the lean angle
    out2, _, opt2 = mm.remove_com_sliding_online(skel, com_srb, des, pose, fix_y_also=False)
    assert np.allclose(opt2[:, 1], curr[:, 1]) and np.allclose(opt2[:, [0, 2]], opt[:, [0, 2]])
    out3, _, _ = mm.remove_com_sliding_online(skel, com_srb, des, pose, ground_h=np.full(nvar, 0.3))
    assert not np.allclose(out3[4, 3:7], out[4, 3:7]) and np.allclose(out3[0], out[0])


def test_extract_fullbody_building_blocks():
    """Exp(se3) against scipy's matrix exponential (also across the small-angle branch), matrix -> quaternion on all four
    branches, and the one-body momentum of _extractSingleFrameFullbody against m v / R I R^T w."""
    from scipy.linalg import expm
    from oracle import fullbody_map as fm
    rng = np.random.default_rng(0)
    for scale in (0.5, 3.0, 1e-12):
        S = rng.normal(0, scale, 6)
        R, p = fm.se3_exp(S)
        X = np.zeros((4, 4)); w = S[:3]
        X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; X[:3, 3] = S[3:]
        E = expm(X)
        assert np.allclose(E[:3, :3], R, atol=1e-9) and np.allclose(E[:3, 3], p, atol=1e-9)
    branches = set()
    for _ in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = np.array([fk.vrotate(q, e) for e in np.eye(3)]).T
        q2 = fm.mat_to_quat(R)
        tr = np.trace(R)
        branches.add(-1 if tr > 0 else int(np.argmax(np.diag(R))))
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12
    assert branches == {-1, 0, 1, 2}
    extra = np.zeros(fm.EXTRA_W); row = np.zeros(fm.ROW_W)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    extra[0:3] = [0.3, 0.9, -0.2]; extra[3:7] = q; extra[7:13] = rng.normal(size=6)
    extra[16] = extra[23] = 1.0
    I = np.diag([3.0, 1.5, 2.5])
    out = fm.extract_single_frame_fullbody(extra, row, 60.0, I, [0, 0, 0])
    R = np.array([fk.vrotate(q, e) for e in np.eye(3)]).T
    assert np.allclose(out["momentum"][3:], 60.0 * extra[10:13], atol=1e-12)
    assert np.allclose(out["momentum"][:3], R.dot(I).dot(R.T).dot(extra[7:10]), atol=1e-12)
    assert np.allclose(out["humanPose"][:7], extra[:7]) and np.allclose(out["com"], extra[0:3])
