"""removeFootSliding_online / removeFootSliding_online_stair (gym_trackSRB/taesooLib/DeltaExtractor.py:650-893).

CPU: properties that pin oracle/foot_sliding.py -- the hard constraints of every programme, the optimality conditions on the free
frames, the Graham-scan hull against the REFERENCE's own graham.cpp (recorded in tests/golden/ref_taesoolib.npz), and
math.projectTrajectoryToItsHull as the upper hull.  GPU: the batched kernels against the oracle at 1e-10, through the C-ABI."""
import os

import numpy as np
import pytest

import helpers
from oracle import foot_sliding as fs

FOOT_LEN = 0.21


def make_window(seed, nvar=7, stairs=False):
    """One planning window in the shapes walk_SRBTrack2025.py:620-631 builds: a foot pair moving forward, toe / heel markers."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(-1, 1, 3) * [1.0, 0.0, 1.0]
    heading = rng.uniform(-np.pi, np.pi)
    fwd = np.array([np.sin(heading), 0.0, np.cos(heading)])
    traj = []
    for foot in range(2):
        side = np.array([fwd[2], 0.0, -fwd[0]]) * (0.1 if foot == 0 else -0.1)
        speed = rng.uniform(0.0, 1.5)
        lift = rng.uniform(0.0, 0.12)
        pitch = rng.uniform(-0.4, 0.4, nvar)
        centre = np.stack([base + side + fwd * speed * (f / 30.0) + np.array([0, 0.05 + lift * np.sin(np.pi * f / (nvar - 1)), 0]) for f in range(nvar)])
        d = np.stack([fwd * np.cos(p) + np.array([0, -np.sin(p), 0]) for p in pitch])
        traj += [centre + d * FOOT_LEN / 2, centre - d * FOOT_LEN / 2]
    marker_traj = [t + rng.normal(0, 0.002, t.shape) for t in traj]
    marker = np.concatenate([t + rng.normal(0, 0.01, t.shape) for t in traj], axis=1)
    con = np.zeros((nvar, 4))
    pattern = seed % 6
    for foot in range(2):
        if pattern == 0:
            pass                                                   # no contact at all
        elif pattern == 1:
            con[:, foot * 2:foot * 2 + 2] = 1                      # full contact
        elif pattern == 2:
            s = rng.integers(1, nvar - 2); con[s:, foot * 2] = 1; con[min(s + 1, nvar - 1):, foot * 2 + 1] = 1     # toe strike, then flat
        elif pattern == 3:
            e = rng.integers(2, nvar - 1); con[:e, foot * 2 + 1] = 1; con[:max(e - 2, 1), foot * 2] = 1            # flat, then heel only, then off
        elif pattern == 4:
            con[0:2, foot * 2:foot * 2 + 2] = 1; con[nvar - 2:, foot * 2 + foot] = 1                               # two intervals
        else:
            con[:, foot * 2:foot * 2 + 2] = (rng.uniform(size=(nvar, 2)) > 0.5)
    initial = np.concatenate([t[0] for t in marker_traj]) + rng.normal(0, 0.003, 12)
    dv = np.concatenate(marker_traj, axis=1)
    desired_vel = np.zeros_like(dv); desired_vel[:-1] = (dv[1:] - dv[:-1]) * 30 + rng.normal(0, 0.05, (nvar - 1, 12))
    foot_center = [0.5 * (traj[0] + traj[1]), 0.5 * (traj[2] + traj[3])]
    return dict(con=con, initial=initial, marker_traj=marker_traj, marker=marker, desired_vel=desired_vel, foot_center=foot_center)


# ------------------------------------------------------------------------------------------------------------------- CPU
def test_hull_equals_reference_graham_scan():
    ref = np.load(os.path.join(helpers.GOLDEN, "ref_taesoolib.npz"))
    n = int(ref["hull_count"])
    for k in range(n):
        pts = ref[f"hull_in_{k}"]
        assert np.array_equal(np.array(fs.graham_hull(pts)), ref[f"hull_out_{k}"]), k
    assert n >= 100


def test_project_trajectory_to_its_hull_is_the_upper_hull():
    rng = np.random.default_rng(0)
    seen = set()
    for k in range(300):
        n = int(rng.integers(4, 9))
        y = np.round(rng.uniform(0, 1, n) * 3) / 3 if k % 2 else rng.uniform(0, 1, n)
        changed, out = fs.project_trajectory_to_its_hull(y)
        edges = len(fs.graham_hull([(i, y[i]) for i in range(n)])) - 1
        seen.add(edges)
        assert changed == (edges > 3)
        if not changed:
            assert np.array_equal(out, y)
            continue
        assert (out >= y - 1e-15).all() and out[0] == y[0] and out[-1] == y[-1]
        d2 = out[:-2] - 2 * out[1:-1] + out[2:]
        assert (d2 <= 1e-12).all()                                 # concave: the least concave majorant of the points
        assert np.isclose(out, y).sum() >= 2
    assert {2, 3, 4, 5} <= seen or {3, 4, 5} <= seen


@pytest.mark.parametrize("seed", range(12))
def test_oracle_constraints_and_optimality(seed):
    p = make_window(seed)
    nvar = p["con"].shape[0]
    opt = fs.remove_foot_sliding_online(FOOT_LEN, p["con"], p["initial"], p["marker_traj"], p["marker"], p["desired_vel"], ground=None)
    assert np.abs(opt[0] - p["initial"]).max() < 1e-12              # continuity with the previous plan
    for foot in range(2):
        ct = p["con"][:, foot * 2] > 0.5; ch = p["con"][:, foot * 2 + 1] > 0.5
        toe, heel = opt[:, foot * 6:foot * 6 + 3], opt[:, foot * 6 + 3:foot * 6 + 6]
        for s, e in fs.run_length_encode(ct | ch):
            fr = [f for f in range(max(s, 1), e)]
            for a, b in zip(fr[:-1], fr[1:]):                       # a marker that stays down does not slide
                if ct[a] and ct[b]:
                    assert np.abs(toe[a] - toe[b]).max() < 1e-10
                if ch[a] and ch[b]:
                    assert np.abs(heel[a] - heel[b]).max() < 1e-10
            for f in fr:                                            # rigid foot while in contact: |toe - heel| = footLen
                assert abs(np.linalg.norm(toe[f] - heel[f]) - FOOT_LEN) < 1e-9
    # free frames: gradient of the quadratic objective vanishes (checked by finite perturbation of one free coordinate)
    def objective(x, im, dim):
        dvc = p["desired_vel"][:, im * 3 + dim]
        c = 1.0 if dim == 1 else 0.1
        o = sum((x[v] - x[v - 1] - dvc[v - 1] / 30.0) ** 2 for v in range(1, nvar))
        o += 1.5 * sum((-x[v - 1] + 2 * x[v] - x[v + 1]) ** 2 for v in range(1, nvar - 1))
        return o + (c * x[2:].sum() - c * p["marker"][2:, im * 3 + dim].sum()) ** 2
    for im in range(4):
        pinned = np.zeros(nvar, bool); pinned[0] = True
        cc = p["con"][:, (im // 2) * 2] > 0.5; hh = p["con"][:, (im // 2) * 2 + 1] > 0.5
        pinned[1:] = (cc | hh)[1:]
        for dim in range(3):
            x = opt[:, im * 3 + dim].copy()
            f0 = objective(x, im, dim)
            for f in np.nonzero(~pinned)[0]:
                for eps in (1e-4, -1e-4):
                    y = x.copy(); y[f] += eps
                    assert objective(y, im, dim) >= f0 - 1e-12


def test_oracle_ground_projection_and_stair_variant():
    p = make_window(7)
    g = lambda pos: 0.02                                            # noqa: E731
    opt = fs.remove_foot_sliding_online(FOOT_LEN, p["con"], p["initial"], p["marker_traj"], p["marker"], p["desired_vel"], ground=g)
    assert (opt[:, 1::3] >= 0.02 - 1e-15).all()
    steps = lambda pos: 0.15 * np.floor(max(pos[2] + pos[0], 0.0) / 0.05)      # noqa: E731  a staircase along x + z
    st = fs.remove_foot_sliding_online_stair(FOOT_LEN, p["con"], p["initial"], p["marker_traj"], p["marker"], p["foot_center"], p["desired_vel"], ground=steps)
    base = fs._marker_qps(FOOT_LEN, np.asarray(p["con"]), np.asarray(p["initial"]), p["marker_traj"], p["marker"], p["desired_vel"], steps, True, p["foot_center"])
    assert (st[:, 1::3] >= base[:, 1::3] - 1e-15).all()             # the hull pass only lifts
    assert np.array_equal(st[:, 0::3], base[:, 0::3]) and np.array_equal(st[:, 2::3], base[:, 2::3])


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("nvar", [7, 8, 5])
def test_gpu_matches_oracle(built_lib, nvar):
    import torch
    from ipcdnnwalk_b200.short_traj_opt import removeFootSliding_online
    ws = [make_window(s, nvar=nvar) for s in range(48)]
    stack = lambda k: np.stack([w[k] for w in ws])                  # noqa: E731
    mt = np.stack([np.concatenate(w["marker_traj"], axis=1) for w in ws])
    for ground, use_dv in ((None, True), (0.0, True), (0.03, False)):
        out = removeFootSliding_online(FOOT_LEN, stack("con"), stack("initial"), mt, stack("marker"), stack("desired_vel") if use_dv else None, projectToGround=ground)
        torch.cuda.synchronize()
        out = out.cpu().numpy()
        for k, w in enumerate(ws):
            ref = fs.remove_foot_sliding_online(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["desired_vel"] if use_dv else None,
                                                ground=None if ground is None else (lambda pos, g=ground: g))
            assert np.abs(out[k] - ref).max() < 1e-10, (k, ground)


@pytest.mark.gpu
def test_gpu_stair_variant_on_the_playground_terrain(built_lib):
    """_stair variant with projectToGround = the SRBTrack playground (pyramid steps): contact positions from the foot-centre
    trajectories, three hull passes.  The oracle queries the same terrain through oracle/terrain.py (Z-up <-> Y-up)."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from ipcdnnwalk_b200.short_traj_opt import removeFootSliding_online
    from oracle.terrain import Terrain
    T = Terrain.load(os.path.join(helpers.GOLDEN, "terrain_playground.npz"))
    env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), 1, variant="terrain", precision=64, terrain=TerrainModel.from_npz(os.path.join(helpers.GOLDEN, "terrain_playground.npz")))

    def ground(pos):                                               # Y-up (a, b, c) <-> Z-up (c, a, b)
        h, _ = T.height_ray(pos[2], pos[0])
        return h
    ws = []
    for s in range(24):
        w = make_window(100 + s)
        shift = np.array([4.0 + 0.37 * s, 0.0, 20.0])               # Y-up x = Z-up y, Y-up z = Z-up x: on the slope of pyramid_1
        for key in ("marker_traj", "foot_center"):
            w[key] = [t + shift for t in w[key]]
        w["marker"] = w["marker"] + np.tile(shift, 4); w["initial"] = w["initial"] + np.tile(shift, 4)
        ws.append(w)
    stack = lambda k: np.stack([w[k] for w in ws])                  # noqa: E731
    mt = np.stack([np.concatenate(w["marker_traj"], axis=1) for w in ws])
    fc = np.stack([np.stack(w["foot_center"]) for w in ws])
    lifted = 0
    for stair in (True, False):
        out = removeFootSliding_online(FOOT_LEN, stack("con"), stack("initial"), mt, stack("marker"), stack("desired_vel"), projectToGround=env,
                                       footCenterTraj=fc if stair else None)
        torch.cuda.synchronize()
        out = out.cpu().numpy()
        for k, w in enumerate(ws):
            if stair:
                ref = fs.remove_foot_sliding_online_stair(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"], w["desired_vel"], ground=ground)
                base = fs._marker_qps(FOOT_LEN, np.asarray(w["con"]), np.asarray(w["initial"]), w["marker_traj"], w["marker"], w["desired_vel"], ground, True, w["foot_center"])
                lifted += int((ref[:, 1::3] > base[:, 1::3] + 1e-9).any())
            else:
                ref = fs.remove_foot_sliding_online(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["desired_vel"], ground=ground)
            assert np.abs(out[k] - ref).max() < 1e-10, (stair, k)
    assert lifted >= 3, "the hull pass should act on some windows"
    env.close()
