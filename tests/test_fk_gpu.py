"""GPU parity of the batched FK / CoM / centroidal-momentum kernel vs the oracle (through the C-ABI)."""
import numpy as np
import pytest

import helpers
from oracle import bone_fk as F

pytestmark = pytest.mark.gpu


def _batch(rng, skel, n):
    q = rng.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pose = np.concatenate([rng.uniform(-5, 5, (n, 3)), q, rng.uniform(-1, 1, (n, skel["n_pose"] - 7))], axis=1)
    dq = rng.uniform(-2, 2, (n, skel["n_dq"]))
    return pose, dq


@pytest.mark.parametrize("generic", [False, True, "no_tma", "tma1", "tma3", "tma4"])
@pytest.mark.parametrize("dtype,tol_pos,tol_mom", [("float64", 1e-11, 1e-10), ("float32", 1e-5, 2e-4)])
def test_fk_com_momentum_match_oracle(built_lib, dtype, tol_pos, tol_mom, generic):
    """FK joint positions within 1e-5 m (BASELINE.json); fp64 to round-off.  Both kernels: the skeleton-specialised
    instantiation (default for this skeleton) and the structure-agnostic one."""
    import torch
    from ipcdnnwalk_b200 import vrml_skeleton, Skeleton, BoneForwardKinematics, LoaderToTree
    skel = vrml_skeleton.builtin()
    sk = Skeleton(skel)
    assert sk.is_specialised()
    if generic == "no_tma":                               # specialised kernel, rows accessed directly (no TMA tiles)
        sk.set_tma(0)
    elif isinstance(generic, str):                        # the other TMA-tile variants
        sk.set_tma(int(generic[3:]))
    else:
        sk.set_generic(generic)
        assert sk.is_specialised() == (not generic)
    rng = np.random.default_rng(4)
    n = 600                                              # ragged vs the 64/128/256-pose tiles
    pose, dq = _batch(rng, skel, n)
    td = getattr(torch, dtype)
    tp = torch.as_tensor(pose, dtype=td, device="cuda"); tq = torch.as_tensor(dq, dtype=td, device="cuda")
    bfk = BoneForwardKinematics(sk); bfk.setPoseDOF(tp)
    tree = LoaderToTree(sk); tree.setPoseDOF(tp); tree.setDQ(tq)
    xf = bfk.global_frames().double().cpu().numpy()
    com = tree.calcCOM().double().cpu().numpy(); mom = tree.calcMomentumCOM().double().cpu().numpy()
    # FK-only and FK + momentum may run different kernel variants: same arithmetic, not the same instruction order
    assert np.abs(tree.globalFrame(5).cpu().numpy() - bfk.globalFrame(5).cpu().numpy()).max() <= (1e-12 if dtype == "float64" else 2e-6)
    for i in range(0, n, 7):
        rot, tr = F.bone_forward_kinematics(skel, pose[i])
        assert np.abs(xf[i, :, 4:7] - tr).max() <= tol_pos, "joint positions"
        assert np.abs(xf[i, :, 0:4] - rot).max() <= max(tol_pos, 1e-6 if dtype == "float32" else 0)
        r2, t2, W, V = F.loader_to_tree(skel, pose[i], dq[i])
        c_ref, h_ref = F.calc_momentum_com(skel, r2, t2, W, V)
        assert np.abs(com[i] - c_ref).max() <= tol_pos
        assert np.abs(mom[i] - h_ref).max() <= tol_mom * max(1.0, np.abs(h_ref).max())
    sk.close()


def test_fk_edge_cases(built_lib):
    import torch
    from ipcdnnwalk_b200 import vrml_skeleton, Skeleton
    skel = vrml_skeleton.builtin()
    sk = Skeleton(skel)
    # empty batch
    out = sk.forward(torch.zeros(0, 60, dtype=torch.float64, device="cuda"), torch.zeros(0, 59, dtype=torch.float64, device="cuda"))
    assert out["xform"].shape == (0, 20, 7)
    # T-pose: every joint sits at the sum of the offsets along its chain, CoM = mass-weighted segment centres
    pose = torch.zeros(3, 60, dtype=torch.float64, device="cuda"); pose[:, 3] = 1.0; pose[:, 1] = 1.0
    out = sk.forward(pose, None)
    xf = out["xform"].cpu().numpy()
    off = np.array(skel["offset"]); exp = np.zeros((20, 3))
    for b in range(20):
        exp[b] = (exp[skel["parent"][b]] if b else np.array([0, 1.0, 0])) + off[b]
    assert np.allclose(xf[0, :, 4:7], exp, atol=1e-14) and np.allclose(xf[0, :, 0], 1) and np.allclose(xf[0, :, 1:4], 0)
    m = np.array(skel["mass"]); c = np.array(skel["com"])
    assert np.allclose(out["com"][0].cpu().numpy(), ((exp + c) * m[:, None]).sum(0) / m.sum(), atol=1e-13)
    assert out["momentum"] is None
    # single pose, velocities only through the root: momentum of a rigid body
    dq = torch.zeros(1, 59, dtype=torch.float64, device="cuda"); dq[0, 3:6] = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    o2 = sk.forward(pose[:1].contiguous(), dq)
    assert np.allclose(o2["momentum"][0, 3:].cpu().numpy(), m.sum() * np.array([1.0, 2.0, 3.0]), atol=1e-12)
    assert np.allclose(o2["momentum"][0, :3].cpu().numpy(), 0, atol=1e-12)
    sk.close()


def test_fk_large_batch_properties(built_lib):
    """BASELINE config 5 size (1M poses): size-independent properties -- bone lengths, unit quaternions,
    linearity of the momentum in dq, and equality with a small-batch evaluation of sampled rows."""
    import torch
    from ipcdnnwalk_b200 import vrml_skeleton, Skeleton
    skel = vrml_skeleton.builtin()
    sk = Skeleton(skel)
    n = 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    q = torch.randn(n, 4, device="cuda", generator=g); q = q / q.norm(dim=1, keepdim=True)
    pose = torch.cat([torch.rand(n, 3, device="cuda", generator=g) * 10 - 5, q, torch.rand(n, 53, device="cuda", generator=g) * 2 - 1], 1).contiguous()
    dq = (torch.rand(n, 59, device="cuda", generator=g) * 2 - 1).contiguous()
    out = sk.forward(pose, dq)
    xf = out["xform"]
    assert torch.isfinite(xf).all() and torch.isfinite(out["momentum"]).all()
    assert (xf[:, :, 0:4].norm(dim=2) - 1).abs().max() < 2e-5
    off = torch.tensor(skel["offset"], device="cuda")
    par = torch.tensor(skel["parent"], device="cuda")
    seg = (xf[:, 1:, 4:7] - xf[:, par[1:], 4:7]).norm(dim=2)
    assert (seg - off[1:].norm(dim=1)).abs().max() < 2e-5
    o2 = sk.forward(pose, (2.0 * dq).contiguous(), want_xform=False)
    assert torch.allclose(o2["momentum"], 2 * out["momentum"], rtol=1e-4, atol=1e-3)
    idx = torch.tensor([0, 1, 12345, n // 2, n - 1], device="cuda")
    o3 = sk.forward(pose[idx].contiguous(), dq[idx].contiguous())
    # the 5-row batch runs the tail kernel, the large one the tile kernel: same arithmetic, different instruction order
    assert torch.allclose(o3["xform"], xf[idx], rtol=0, atol=2e-5) and torch.allclose(o3["momentum"], out["momentum"][idx], rtol=1e-4, atol=1e-3)
    sk.close()


def test_fk_other_skeleton_uses_generic_kernel(built_lib):
    """A skeleton with a different tree (the last bone removed) does not match the compile-time table: the generic kernel
    runs and still agrees with the oracle; on the matching skeleton both kernels agree with each other."""
    import copy
    import torch
    from ipcdnnwalk_b200 import vrml_skeleton, Skeleton
    skel = vrml_skeleton.builtin()
    cut = copy.deepcopy(skel)
    for k in ("names", "parent", "depth", "channels", "offset", "mass", "com", "inertia"):
        cut[k] = cut[k][:-1]
    cut["n_pose"] -= 3; cut["n_dq"] -= 3
    sk = Skeleton(cut)
    assert not sk.is_specialised()
    rng = np.random.default_rng(9)
    pose, dq = _batch(rng, skel, 40)
    pc = np.ascontiguousarray(pose[:, :57]); dc = np.ascontiguousarray(dq[:, :56])
    out = sk.forward(torch.as_tensor(pc, device="cuda"), torch.as_tensor(dc, device="cuda"))
    for i in range(0, 40, 9):
        rot, tr = F.bone_forward_kinematics(cut, pc[i])
        assert np.abs(out["xform"][i, :, 4:7].cpu().numpy() - tr).max() < 1e-11
    sk.close()
    a = Skeleton(skel); b = Skeleton(skel); b.set_generic(True)
    tp = torch.as_tensor(pose, device="cuda"); tq = torch.as_tensor(dq, device="cuda")
    oa, ob = a.forward(tp, tq), b.forward(tp, tq)
    for k in ("xform", "com", "momentum"):
        assert (oa[k] - ob[k]).abs().max() < 1e-11
    oa, ob = a.forward(tp.float(), None), b.forward(tp.float(), None)
    assert (oa["xform"] - ob["xform"]).abs().max() < 1e-5 and oa["momentum"] is None
    a.close(); b.close()
