"""CPU pins of the FK / momentum oracle (taesooLib cannot be built here: analytic checks)."""
import numpy as np

import helpers
from oracle import bone_fk as F
from ipcdnnwalk_b200 import vrml_skeleton


def _random_pose(rng, skel, unit=True):
    q = rng.standard_normal(4); q /= np.linalg.norm(q)
    pose = np.concatenate([rng.uniform(-5, 5, 3), q, rng.uniform(-1, 1, skel["n_pose"] - 7)])
    dq = rng.uniform(-2, 2, skel["n_dq"])
    return pose, dq


def test_bone_fk_equals_node_chain_fk():
    skel = vrml_skeleton.builtin()
    rng = np.random.default_rng(0)
    for _ in range(5):
        pose, dq = _random_pose(rng, skel)
        r1, t1 = F.bone_forward_kinematics(skel, pose)
        r2, t2, _, _ = F.loader_to_tree(skel, pose, dq)
        assert np.allclose(r1, r2, atol=1e-13) and np.allclose(t1, t2, atol=1e-13)
        # bone lengths are preserved and quaternions stay unit
        for b in range(1, 20):
            p = skel["parent"][b]
            assert abs(np.linalg.norm(t1[b] - t1[p]) - np.linalg.norm(skel["offset"][b])) < 1e-12
        assert np.allclose(np.linalg.norm(r1, axis=1), 1, atol=1e-12)


def test_body_velocities_match_finite_differences():
    skel = vrml_skeleton.builtin()
    rng = np.random.default_rng(1)
    pose, dq = _random_pose(rng, skel)
    rot, tr, W, V = F.loader_to_tree(skel, pose, dq)
    eps = 1e-7
    p2 = pose.copy()
    p2[0:3] += eps * dq[3:6]
    wq = np.concatenate([[0.0], dq[0:3]])
    p2[3:7] = pose[3:7] + 0.5 * eps * F.qmult(wq, pose[3:7])          # world-frame angular velocity
    p2[3:7] /= np.linalg.norm(p2[3:7])
    p2[7:] += eps * dq[6:]
    rot2, tr2, _, _ = F.loader_to_tree(skel, p2, None)
    for b in range(20):
        v_world = (tr2[b] - tr[b]) / eps
        assert np.allclose(F.vrotate(rot[b], V[b]), v_world, atol=2e-6), b
        dqb = F.qmult(F.qinv(rot[b]), rot2[b])                          # body-frame rotation increment
        w_body = 2 * dqb[1:4] / eps
        assert np.allclose(W[b], w_body, atol=2e-6), b


def test_momentum_of_rigid_translation_and_independent_formula():
    skel = vrml_skeleton.builtin()
    rng = np.random.default_rng(2)
    pose, dq = _random_pose(rng, skel)
    # pure translation of the whole body: P = M v, L_com = 0
    dq0 = np.zeros_like(dq); dq0[3:6] = [0.3, -0.2, 0.5]
    rot, tr, W, V = F.loader_to_tree(skel, pose, dq0)
    com, H = F.calc_momentum_com(skel, rot, tr, W, V)
    assert np.allclose(H[3:], sum(skel["mass"]) * dq0[3:6], atol=1e-12) and np.allclose(H[:3], 0, atol=1e-12)
    assert np.allclose(com, F.calc_com(skel, rot, tr), atol=1e-13)
    # general motion: linear momentum = sum m_i v_com_i, angular momentum from the same (Q13) joint-origin inertia
    rot, tr, W, V = F.loader_to_tree(skel, pose, dq)
    com, H = F.calc_momentum_com(skel, rot, tr, W, V)
    P = np.zeros(3); L0 = np.zeros(3)
    for b in range(20):
        m = skel["mass"][b]; c = np.array(skel["com"][b]); I = np.array(skel["inertia"][b]).reshape(3, 3)
        R = lambda x: F.vrotate(rot[b], x)
        p_b = m * (V[b] + np.cross(W[b], c))
        P += R(p_b)
        L0 += R(I @ W[b] + m * np.cross(c, V[b])) + np.cross(tr[b], R(p_b))
    assert np.allclose(H[3:], P, atol=1e-11)
    assert np.allclose(H[:3], L0 - np.cross(com, P), atol=1e-11)
