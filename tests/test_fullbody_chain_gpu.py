"""The online SRB -> full-body chain of walk_SRBTrack2025.py:440-620 on the device, end to end through the C-ABI pieces:
SRBVecEnv.step -> pack_obs_extra -> [decoder rows: synthetic here, the FullbodyVAE is PyTorch and out of scope] ->
fullbody_extract -> mmsto_solve -> mmsto_remove_com_sliding.  Each piece has its own parity test against the oracle; this test
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
    opt.close(); env.close()
