"""The online SRB -> full-body chain of walk_SRBTrack2025.py:440-620 on the device, end to end through the C-ABI pieces:
SRBVecEnv.step -> pack_obs_extra -> [decoder rows: synthetic here, the FullbodyVAE is PyTorch and out of scope] ->
fullbody_extract -> mmsto_solve -> mmsto_remove_com_sliding -> mmsto_remove_foot_sliding.  Each piece has its own parity test against the oracle; this test
checks that their layouts plug together (Y-up records, Euler-root conversion, window assembly) and the chain's invariants."""
import numpy as np
import pytest

import helpers
import helpers_mmsto as hm

pytestmark = pytest.mark.gpu


def test_chain_runs_and_keeps_its_invariants(built_lib):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, ShortTrajOpt
    from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
    from oracle import mmsto as mm
    skel = hm.skeleton()
    n, nf = 6, hm.NF
    clips = helpers.fresh_clips()
    env = SRBVecEnv(clips, n, variant="test", precision=64, auto_reset=False, seed=2)
    env.reset()
    act = torch.zeros(n, 22, device="cuda")
    extras = []
    for k in range(nf + 1):
        env.tracking_action(act, sigma=0.02, seed=5, step=k)
        env.step(act)
        extras.append(env.pack_obs_extra().clone())
    extras = torch.stack(extras, 1)                                   # [n, nf + 1, 29]
    assert torch.isfinite(extras[..., :13]).all()
    # synthetic decoder rows: a smooth full-body motion around the T pose, small SRB -> pelvis / momentum offsets
    rng = np.random.default_rng(0)
    rows = np.zeros((n, nf + 1, 214))
    base = rng.normal(0, 0.15, (n, 1, 53)); drift = rng.normal(0, 0.02, (n, 1, 53))
    rows[..., 39:92] = base + drift * np.arange(nf + 1)[None, :, None]
    rows[..., 0:6] = rng.normal(0, 0.02, (n, nf + 1, 6))
    rows[..., 6:30] = rng.normal(0, 0.05, (n, nf + 1, 24))
    rows[..., 30:39] = rng.normal(0, 0.01, (n, nf + 1, 9))
    rows[..., 92:210] = rng.normal(0, 0.01, (n, nf + 1, 118))
    rows[..., 210:214] = rng.normal(0, 1.0, (n, nf + 1, 4))
    ex = extras.reshape(-1, 29).clone()
    ex[:, 13:27] = torch.nan_to_num(ex[:, 13:27], nan=0.0)            # swing feet carry NaN heights in the flat env (Q6)
    out = extractSingleFrameFullbody(ex, rows.reshape(-1, 214), 60.0, np.diag([10.6, 9.7, 2.1]))
    human = out["humanPose"].reshape(n, nf + 1, 60).cpu().numpy()
    assert np.isfinite(human).all() and np.allclose(np.linalg.norm(human[..., 3:7], axis=-1), 1.0, atol=1e-9)
    opt = ShortTrajOpt(skel, nf, hm.MARKERS)
    opt.set_dof_weights(hm.ankle_dof_weights(skel))
    q = opt.calculateOriginalEulerPoses(human)                        # [n, nf + 1, 59]
    assert np.abs(opt.eulerToPoseDOF(q)[..., :3] - human[..., :3]).max() < 1e-12
    x_original = q[:, 1:]                                             # the last planningHorizon frames (walk_SRBTrack2025.py:538)
    dxd = out["dx_ddx"].reshape(n, nf + 1, 118).cpu().numpy()
    dx = dxd[:, 2:, :59]; ddx = dxd[:, 3:, 59:]                       # [-(H-1):], [-(H-1):-1] of the buffers (:513-514)
    gpos = out["feetpos"].reshape(n, nf + 1, 24)[:, 1:, :12].cpu().numpy()
    com = out["com"].reshape(n, nf + 1, 3)[:, 1:].cpu().numpy()
    mom = out["momentum"].reshape(n, nf + 1, 6)[:, 1:nf].cpu().numpy()
    euler_mot = x_original + rng.normal(0, 0.002, x_original.shape)
    opt.set("max_iterations", 6)
    f0, _ = opt.evaluate(x_original, gpos, x_original, dx, ddx, com, mom)
    x, info = opt.solveEndEffectors_euler(euler_mot, gpos, x_original, dx, ddx, com, mom)
    info = info.cpu().numpy(); x = x.cpu().numpy()
    assert np.isfinite(x).all() and (info[:, 3] >= 6).all()
    assert (info[:, 0] < f0.cpu().numpy()).all()                      # six iterations lower the objective of every window
    assert np.array_equal(x[:, :2], np.clip(euler_mot[:, :2], x_original[:, :2] - np.radians(25), x_original[:, :2] + np.radians(25)))
    # the step after the solve on the poses of the solution; com_srb = the SRB positions of the same frames (Y-up)
    poses = torch.as_tensor(opt.eulerToPoseDOF(x), device="cuda").contiguous()
    before = poses.clone()
    com_srb = extras[:, 1:, 0:3].contiguous()
    curr, optc = opt.removeCOMsliding_online(com_srb, com, poses)
    curr = curr.cpu().numpy(); optc = optc.cpu().numpy()
    assert np.allclose(optc[:, :2], curr[:, :2], atol=1e-12)
    assert np.allclose(optc[:, 1:].sum(1), com[:, :nf - 1].sum(1), atol=1e-9)
    after = opt.windowCOM(poses).cpu().numpy()
    assert np.abs(after - optc).max() < 1e-9                          # the moved poses carry the optimised CoM
    assert np.array_equal(poses[..., 7:].cpu().numpy(), before[..., 7:].cpu().numpy())   # only the roots move
    # spot check of one window against the oracle chain
    o_out, o_curr, o_opt = mm.remove_com_sliding_online(skel, com_srb[0].cpu().numpy(), com[0], before[0].cpu().numpy())
    assert np.abs(o_out - poses[0].cpu().numpy()).max() < 1e-9
    # limb step (walk_SRBTrack2025.py:620-631): marker trajectories of the moved plan (fc.getMarkerTraj = FK of the four foot markers),
    # contact targets and marker targets of the same frames -> sliding-free marker trajectories
    from ipcdnnwalk_b200 import Skeleton
    from ipcdnnwalk_b200.short_traj_opt import removeFootSliding_online
    from oracle import foot_sliding as fs, bone_fk as bk
    S = Skeleton(skel)
    xf = S.forward(poses.reshape(-1, 60), None, want_com=False, want_momentum=False)["xform"].reshape(n, nf, 20, 7)
    mk = []
    for m in hm.MARKERS:                                              # frame * local position of every marker
        qb = xf[:, :, m["bone"], 0:4].cpu().numpy(); tb = xf[:, :, m["bone"], 4:7].cpu().numpy()
        mk.append(np.stack([[bk.vrotate(qb[i, f], np.array(m["lpos"])) + tb[i, f] for f in range(nf)] for i in range(n)]))
    marker_traj = np.concatenate(mk, axis=-1)                         # [n, nf, 12]
    con_all = (out["con"].reshape(n, nf + 1, 4)[:, 1:].cpu().numpy() > 0).astype(float)
    marker_tgt = out["feetpos"].reshape(n, nf + 1, 24)[:, 1:].cpu().numpy()
    initial = marker_traj[:, 0].copy()
    foot_len = float(np.linalg.norm(np.array(hm.MARKERS[0]["lpos"]) - np.array(hm.MARKERS[1]["lpos"])))
    feet = removeFootSliding_online(foot_len, con_all, initial, marker_traj, marker_tgt[..., :12], marker_tgt[..., 12:], projectToGround=0.0)
    feet = feet.cpu().numpy()
    assert np.isfinite(feet).all() and np.abs(feet[:, 0, 0::3] - initial[:, 0::3]).max() < 1e-12 and (feet[..., 1::3] >= -1e-15).all()
    for i in (0, n - 1):
        ref = fs.remove_foot_sliding_online(foot_len, con_all[i], initial[i], [marker_traj[i, :, 3 * k:3 * k + 3] for k in range(4)], marker_tgt[i, :, :12],
                                            marker_tgt[i, :, 12:], ground=lambda pos: 0.0)
        assert np.abs(feet[i] - ref).max() < 1e-10
    S.close(); opt.close(); env.close()
