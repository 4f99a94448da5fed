"""Parity of the batched MMSTO kernel (through the C-ABI, include/mmsto_b200.h) with oracle/mmsto.py.

Tolerances: one evaluation (objective, gradient) 1e-11 relative; L-BFGS iterates after k iterations 1e-9 (k <= 10; the iterate
map amplifies rounding because the momentum gradient is inexact by design and the Hessian scale is 1e5..1e7); integer outputs
(libLBFGS return code, iteration and evaluation counts) bit exact."""
import numpy as np
import pytest

import helpers_mmsto as hm

pytestmark = pytest.mark.gpu

KEYS = ("gpos_target", "x_original", "dx", "ddx", "com", "momentum")


@pytest.fixture(scope="module")
def skel():
    return hm.skeleton()


@pytest.fixture()
def opt(skel):
    from ipcdnnwalk_b200 import ShortTrajOpt
    o = ShortTrajOpt(skel, hm.NF, hm.MARKERS)
    o.set_dof_weights(hm.ankle_dof_weights(skel))
    yield o
    o.close()


def test_evaluate_matches_oracle(opt, skel):
    """Objective and full gradient (all 7 x 59 variables) at a perturbed point, every term switched on, ragged batch of 5."""
    ps = [hm.make_problem(s, skel=skel) for s in range(5)]
    opt.set("weightEEcon", 1000.0)
    rng = np.random.default_rng(9)
    xs = np.stack([p["x_original"] + rng.normal(0, 0.01, p["x_original"].shape) for p in ps])
    obj, grad = opt.evaluate(xs, *[hm.stack(ps, k) for k in KEYS], con=hm.stack(ps, "con"), velcon=hm.stack(ps, "velcon"))
    obj = obj.cpu().numpy(); grad = grad.cpu().numpy()
    active = 0
    for i, p in enumerate(ps):
        o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel), weightEEcon=1000.0)
        f, g = o.evaluate(xs[i].ravel())
        assert abs(f - obj[i]) <= 1e-11 * abs(f)
        assert np.abs(g.reshape(hm.NF, -1) - grad[i]).max() <= 1e-11 * np.abs(g).max()
        o.hs_vel = {}
        active += o.evaluate(xs[i].ravel())[0] != f
    assert active >= 1                                     # the velocity half-space term was active somewhere


def test_evaluate_term_by_term(opt, skel):
    p = hm.make_problem(11, skel=skel)
    x = p["x_original"] + np.random.default_rng(2).normal(0, 0.02, p["x_original"].shape)
    names = ("weightDDX", "weightDX", "weightAngleOriginal", "weightEE", "weightEEcon", "weightVelHalfSpace", "weightMomentum", "weightCOM")
    for on in names:
        w = {n: 0.0 for n in names}; w[on] = 1234.5
        for n, v in w.items():
            opt.set(n, v)
        obj, grad = opt.evaluate(x[None], *[p[k][None] for k in KEYS], con=p["con"][None], velcon=p["velcon"][None])
        o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel), **w)
        f, g = o.evaluate(x.ravel())
        assert f > 0, on
        assert abs(f - obj.item()) <= 1e-11 * abs(f), on
        assert np.abs(g.reshape(hm.NF, -1) - grad[0].cpu().numpy()).max() <= 1e-11 * np.abs(g).max(), on


@pytest.mark.parametrize("iters", [1, 4, 10])
def test_lbfgs_iterates_match_oracle(opt, skel, iters):
    ps = [hm.make_problem(20 + s, skel=skel) for s in range(3)]
    opt.set("max_iterations", iters)
    x, info = opt.solveEndEffectors_euler(hm.stack(ps, "euler_mot"), *[hm.stack(ps, k) for k in KEYS], velcon=hm.stack(ps, "velcon"))
    x = x.cpu().numpy(); info = info.cpu().numpy()
    for i, p in enumerate(ps):
        o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel))
        X, inf = o.solve(p["euler_mot"], max_iterations=iters)
        assert (int(info[i, 1]), int(info[i, 2]), int(info[i, 3])) == (inf["ret"], inf["iterations"], inf["evaluations"])
        assert np.abs(X - x[i]).max() < 1e-9
        assert abs(info[i, 0] - inf["fx"]) <= 1e-9 * abs(inf["fx"])
        assert np.array_equal(x[i, :2], X[:2])             # frames 0, 1: the clamped input motion, untouched


def test_exact_gradient_problem_converges_to_the_oracle_minimiser(opt, skel):
    """Without the momentum term the gradient is exact and L-BFGS converges (ret 0 or a line-search stop at the round-off floor):
    both implementations must land on the same minimiser although their iterates drift apart by rounding."""
    p = hm.make_problem(31, target_noise=0.005, skel=skel)
    opt.set("weightMomentum", 0.0); opt.set("epsilon", 1e-7)
    x, info = opt.solveEndEffectors_euler(p["euler_mot"][None], *[p[k][None] for k in KEYS], velcon=p["velcon"][None])
    o = hm.oracle_opt(p, skel, dof_weights=hm.ankle_dof_weights(skel), weightMomentum=0.0)
    X, inf = o.solve(p["euler_mot"], epsilon=1e-7, max_iterations=400)
    assert abs(info[0, 0].item() - inf["fx"]) <= 1e-6 * abs(inf["fx"])
    assert np.abs(X - x[0].cpu().numpy()).max() < 1e-5


def test_batch_is_independent_of_its_neighbours(opt, skel):
    """Problems share a warp four at a time and stop after different numbers of evaluations: each result must equal the
    result of solving it alone (bit exact), also in a batch that does not fill the last warp."""
    ps = [hm.make_problem(40 + s, target_noise=0.002 * (1 + s), skel=skel) for s in range(7)]
    opt.set("max_iterations", 12)
    args = [hm.stack(ps, k) for k in KEYS]
    x, info = opt.solveEndEffectors_euler(hm.stack(ps, "euler_mot"), *args, velcon=hm.stack(ps, "velcon"))
    assert len(set(info[:, 3].cpu().numpy().tolist())) > 1
    for i in (0, 3, 6):
        xi, ii = opt.solveEndEffectors_euler(ps[i]["euler_mot"][None], *[a[i:i + 1] for a in args], velcon=ps[i]["velcon"][None])
        assert np.array_equal(xi[0].cpu().numpy(), x[i].cpu().numpy()) and np.array_equal(ii[0].cpu().numpy(), info[i].cpu().numpy())


def test_host_conversions_round_trip(opt, skel):
    """calculateOriginalEulerPoses / eulerToPoseDOF: pose -> Euler-root q -> pose, and the bone frames of the FK kernel agree with
    the oracle's Euler tree at q."""
    from ipcdnnwalk_b200 import Skeleton
    from oracle import mmsto as mm
    rng = np.random.default_rng(5)
    q = rng.uniform(-1.0, 1.0, (4, hm.NF, opt.ndof)); q[..., 3] += np.linspace(0, 7, hm.NF)      # ry beyond pi: alignment matters
    pose = opt.eulerToPoseDOF(q)
    q2 = opt.calculateOriginalEulerPoses(pose, prev_ry=q[:, 0, 3])
    assert np.abs(q2 - q).max() < 1e-12
    import torch
    sk = Skeleton(skel)
    com = sk.forward(torch.as_tensor(pose.reshape(-1, pose.shape[-1]), device="cuda").contiguous(), None)["com"].cpu().numpy().reshape(4, hm.NF, 3)
    t = mm.EulerTree(skel)
    for i, f in ((0, 0), (1, 2), (3, 6)):
        t.set_q(q[i, f])
        assert np.abs(com[i, f] - t.calc_com()).max() < 1e-12
    sk.close()


def test_rejects_unsupported_configuration(opt):
    with pytest.raises(NotImplementedError):
        opt.set("weightEEvel", 1.0)
    from ipcdnnwalk_b200.short_traj_opt import MmstoError
    with pytest.raises(MmstoError):
        opt.set("noSuchWeight", 1.0)


@pytest.mark.parametrize("nvar,fix_y,ground", [(8, True, False), (8, False, True), (6, True, True)])
def test_remove_com_sliding_matches_oracle(opt, skel, nvar, fix_y, ground):
    """removeCOMsliding_online (DeltaExtractor.py:421-512) on a ragged batch of 19 windows: poses, currCOM, optCOM at 1e-10."""
    import torch
    from oracle import mmsto as mm
    from test_oracle_mmsto_cpu import _com_slide_case
    cases = [_com_slide_case(skel, 50 + s, nvar) for s in range(19)]
    pose = torch.as_tensor(np.stack([c[0] for c in cases]), device="cuda").contiguous()
    gh = np.random.default_rng(1).uniform(0, 0.4, (19, nvar)) if ground else None
    com0 = opt.windowCOM(pose.clone()).cpu().numpy()
    curr, op = opt.removeCOMsliding_online(np.stack([c[1] for c in cases]), np.stack([c[2] for c in cases]), pose, fix_y_also=fix_y, ground_h=gh)
    pose = pose.cpu().numpy(); curr = curr.cpu().numpy(); op = op.cpu().numpy()
    for i in (0, 7, 18):
        out, c, o = mm.remove_com_sliding_online(skel, cases[i][1], cases[i][2], cases[i][0], fix_y_also=fix_y, ground_h=None if gh is None else gh[i])
        assert np.abs(c - curr[i]).max() < 1e-12 and np.abs(c - com0[i]).max() < 1e-12
        assert np.abs(o - op[i]).max() < 1e-10
        assert np.abs(out - pose[i]).max() < 1e-10


def test_nan_window_is_contained(opt, skel):
    """A window with a NaN target stops with LBFGSERR_UNKNOWNERROR (-1024) at once; its warp neighbours are solved as if alone."""
    ps = [hm.make_problem(60 + s, skel=skel) for s in range(4)]
    opt.set("max_iterations", 6)
    args = [hm.stack(ps, k) for k in KEYS]
    x0, i0 = opt.solveEndEffectors_euler(hm.stack(ps, "euler_mot"), *args, velcon=hm.stack(ps, "velcon"))
    bad = [a.copy() for a in args]
    bad[0][2, 3, 5] = np.nan                                  # gpos_target of window 2
    x1, i1 = opt.solveEndEffectors_euler(hm.stack(ps, "euler_mot"), *bad, velcon=hm.stack(ps, "velcon"))
    i0 = i0.cpu().numpy(); i1 = i1.cpu().numpy()
    assert int(i1[2, 1]) == -1024 and int(i1[2, 3]) == 1
    for w in (0, 1, 3):
        assert np.array_equal(x0[w].cpu().numpy(), x1[w].cpu().numpy()) and np.array_equal(i0[w], i1[w])


def test_extract_fullbody_matches_oracle():
    """_extractSingleFrameFullbody (DeltaExtractor.py:577-643) on 37 items incl. a zero-rotation pelvis offset (small-angle branch
    of Exp) and a half-turn one (the non-trace branches of the matrix -> quaternion conversion): every output at 1e-12."""
    from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
    from oracle import fullbody_map as fm
    rng = np.random.default_rng(3)
    n = 37
    extra = rng.normal(0, 0.5, (n, fm.EXTRA_W)); row = rng.normal(0, 0.3, (n, fm.ROW_W))
    for c in (3, 16, 23):
        extra[:, c:c + 4] /= np.linalg.norm(extra[:, c:c + 4], axis=1, keepdims=True)
    row[1, 0:3] = 0.0
    row[2, 0:3] = np.array([0.0, 3.1, 0.2]); row[3, 0:3] = np.array([3.0, 0.1, 0.0]); row[4, 0:3] = np.array([0.1, 0.0, 3.05])
    I = np.array([[3.0, 0.1, 0.0], [0.1, 1.5, -0.2], [0.0, -0.2, 2.5]]); lc = [0.01, -0.03, 0.02]
    out = extractSingleFrameFullbody(extra, row, 60.0, I, lc)
    for i in range(n):
        ref = fm.extract_single_frame_fullbody(extra[i], row[i], 60.0, I, lc)
        for k, v in ref.items():
            assert np.abs(out[k][i].cpu().numpy() - v).max() < 1e-12, (i, k)


def test_full_batch_properties(opt, skel):
    """4096 windows (the size of the throughput runs) built from 8 distinct problems: copies of a problem give bit-identical
    solutions wherever they sit in the batch, the objective never ends above its starting value, frames 0 and 1 keep the clamped
    input, and the solution of the copies equals the solution of the problem solved alone."""
    import torch
    base = [hm.make_problem(70 + s, skel=skel) for s in range(8)]
    n = 4096
    idx = np.arange(n) % 8
    args = [torch.as_tensor(hm.stack(base, k), device="cuda")[idx].contiguous() for k in KEYS]
    em = torch.as_tensor(hm.stack(base, "euler_mot"), device="cuda")[idx].contiguous()
    vc = torch.as_tensor(hm.stack(base, "velcon"), device="cuda")[idx].contiguous()
    opt.set("max_iterations", 8)
    f0, _ = opt.evaluate(args[1], *args, velcon=vc)
    x, info = opt.solveEndEffectors_euler(em, *args, velcon=vc)
    x = x.cpu().numpy(); info = info.cpu().numpy()
    assert np.isfinite(x).all()
    for s in range(8):
        sel = np.nonzero(idx == s)[0]
        assert (x[sel] == x[sel[0]]).all() and (info[sel] == info[sel[0]]).all()
    assert (info[:, 0] <= f0.cpu().numpy() * (1 + 1e-12)).all()
    xo = args[1].cpu().numpy(); e = em.cpu().numpy()
    assert np.array_equal(x[:, :2], np.clip(e[:, :2], xo[:, :2] - np.radians(25), xo[:, :2] + np.radians(25)))
    x1, i1 = opt.solveEndEffectors_euler(em[3:4], *[a[3:4] for a in args], velcon=vc[3:4])
    assert np.array_equal(x1[0].cpu().numpy(), x[3]) and np.array_equal(i1[0].cpu().numpy(), info[3])
