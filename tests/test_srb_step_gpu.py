"""GPU parity of the CUDA SRBTrack step against the oracle (through the C-ABI)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

RTOL64 = 1e-5     # BASELINE.json: single-step state / GRF / QP within 1e-5 relative in fp64


def _golden(variant):
    import os
    return np.load(os.path.join(helpers.GOLDEN, f"rollout_{variant}.npz"))


def _close(a, b, rtol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(1.0, np.abs(b))
    err = np.abs(a - b) / scale
    assert np.nanmax(err) <= rtol, f"{what}: max rel err {np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}"


@pytest.mark.parametrize("variant", ["test", "training"])
@pytest.mark.parametrize("lanes", [8, 32, 1])
def test_golden_rollout_fp64(built_lib, variant, lanes):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    g = _golden(variant)
    n_env, n_steps = g["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(), n_env, variant=variant, precision=64, lanes_per_env=lanes, auto_reset=False)
    env.attach(dbg=True)
    obs0 = env.reset(plan=g["plans"]).cpu().numpy()
    _close(obs0, g["obs0"], 2e-6, "reset obs")
    for k in range(n_steps):
        env.dbg.zero_()
        obs, rew, done, rinfo = env.step(torch.as_tensor(g["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(i)) for i in range(n_env)])
        # integer / boolean outputs: bit exact
        assert (done.cpu().numpy().astype(bool) == g["done"][:, k]).all(), f"done mismatch at step {k}"
        assert (st[:, 23:26] == g["state"][:, k, 23:26]).all(), f"contact flags / frame mismatch at step {k}"
        _close(st[:, :23], g["state"][:, k, :23], RTOL64, f"state step {k}")
        _close(obs.cpu().numpy(), g["obs"][:, k], RTOL64, f"obs step {k}")
        _close(rew.cpu().numpy(), g["rew"][:, k], RTOL64, f"reward step {k}")
        _close(rinfo.cpu().numpy(), g["rinfo"][:, k], RTOL64, f"reward_info step {k}")
        dbg = env.dbg.cpu().numpy().reshape(n_env, 4, 64)
        qdd_ref = g["qdd"][:, k]
        m = ~np.isnan(qdd_ref[..., 0])
        _close(dbg[..., :6][m], qdd_ref[m], RTOL64, f"QP qdd step {k}")
    env.close()


def _make_oracles(n, variant, seed, noise=True):
    from oracle.srb_env import OracleSRBEnv
    rng = np.random.default_rng(seed)
    clips = helpers.fresh_clips()
    plans = np.stack([helpers.random_plan(rng, clips, noise=noise) for _ in range(n)])
    envs = []
    for i in range(n):
        o = OracleSRBEnv(helpers.fresh_clips(), variant)
        helpers.oracle_reset(o, plans[i])
        envs.append(o)
    return envs, plans, rng


@pytest.mark.parametrize("variant", ["test", "training"])
def test_single_step_fp32_within_1e3(built_lib, variant):
    """BASELINE.json: single-step state / QP solution within 1e-3 relative in fp32.  Every step the fp32 CUDA
    environment is loaded with the oracle's fp64 state (srb_set_state), both take the same float32 action."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 6
    oracles, plans, rng = _make_oracles(n, variant, seed=21)
    env = SRBVecEnv(helpers.fresh_clips(), n, variant=variant, precision=32, auto_reset=False)
    env.attach(dbg=True)
    env.reset(plan=plans)
    for k in range(25):
        for i, o in enumerate(oracles):
            env.set_state(i, helpers.oracle_to_cuda_state(o))
        acts = np.stack([helpers.tracking_action(o, rng, 0.1) for o in oracles])
        env.dbg.zero_()
        obs, rew, done, rinfo = env.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        dbg = env.dbg.cpu().numpy().reshape(n, 4, 64)
        for i, o in enumerate(oracles):
            oo, rr, dd, _, info = o.step(acts[i])
            st = helpers.cuda_state_vector(env.named_state(i))
            ref = helpers.oracle_state_vector(o)
            assert (st[23:26] == ref[23:26]).all(), "contact flags / frame must be bit exact"
            _close(st[:23], ref[:23], 1e-3, f"fp32 state step {k}")
            _close(obs[i].cpu().numpy(), oo, 1e-3, f"fp32 obs step {k}")
            _close(float(rew[i]), rr, 2e-3, f"fp32 reward step {k}")
            for sub, q in enumerate(o.qdd_qp_list):
                if q is not None:
                    _close(dbg[i, sub, :6], q, 1e-3, f"fp32 QP qdd step {k}")
            if dd:
                helpers.oracle_reset(o, plans[i])
    env.close()


def test_set_state_roundtrip_and_restore_fp64(built_lib):
    """getFullState / restoreFullState: loading the oracle's state into the CUDA env and stepping reproduces the
    oracle (tight tolerance: same arithmetic, fp64)."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 4
    oracles, plans, rng = _make_oracles(n, "test", seed=33)
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=64, auto_reset=False)
    env.reset(plan=plans[::-1].copy())                      # deliberately different initial states
    for k in range(12):
        acts = np.stack([helpers.tracking_action(o, rng, 0.05) for o in oracles])
        if k % 4 == 0:
            for i, o in enumerate(oracles):
                stt = helpers.oracle_to_cuda_state(o)
                env.set_state(i, stt)
                back = env.get_state(i)
                assert np.array_equal(back["reals"][:38], stt["reals"][:38]) and np.array_equal(back["ints"], stt["ints"])
        obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        for i, o in enumerate(oracles):
            oo, rr, dd, _, _ = o.step(acts[i])
            _close(obs[i].cpu().numpy(), oo, 1e-6, f"obs step {k}")
            _close(float(rew[i]), rr, 1e-6, f"reward step {k}")
            assert bool(done[i]) == dd
    env.close()


def test_autoreset_rollout_matches_oracle_fp64(built_lib):
    """Whole episodes with the fused auto-reset: the CUDA env is driven by the device-side tracking controller,
    the oracle receives the very same actions and reset draws.  Trajectories (hence all reward statistics) agree."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n, steps = 12, 70
    oracles, plans, rng = _make_oracles(n, "test", seed=44)
    clips = helpers.fresh_clips()
    env = SRBVecEnv(clips, n, variant="test", precision=64, auto_reset=True)
    env.attach(reset_plan=True, term_obs=True)
    env.reset(plan=plans)
    nxt = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    env.reset_plan.copy_(torch.as_tensor(nxt))
    act = torch.zeros(n, 22, device="cuda")
    n_done = 0
    ret_c = np.zeros(n); ret_o = np.zeros(n); finished_c, finished_o = [], []
    for k in range(steps):
        env.tracking_action(act, sigma=0.15, seed=5, step=k)
        a = act.cpu().numpy()
        obs, rew, done, _ = env.step(act)
        torch.cuda.synchronize()
        obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); done = done.cpu().numpy().astype(bool)
        term = env.term_obs.cpu().numpy()
        for i, o in enumerate(oracles):
            oo, rr, dd, _, _ = o.step(a[i])
            assert dd == done[i], f"done mismatch env {i} step {k}"
            _close(rew[i], rr, RTOL64, "reward")
            ret_c[i] += rew[i]; ret_o[i] += rr
            if dd:
                _close(term[i], oo, RTOL64, "terminal obs")
                o0, _ = helpers.oracle_reset(o, nxt[i])
                _close(obs[i], o0, RTOL64, "obs after auto-reset")
                finished_c.append(ret_c[i]); finished_o.append(ret_o[i]); ret_c[i] = ret_o[i] = 0
                nxt[i] = helpers.random_plan(rng, clips)
                n_done += 1
            else:
                _close(obs[i], oo, RTOL64, "obs")
        env.reset_plan.copy_(torch.as_tensor(nxt))
    assert n_done >= 5, "the roll-out should contain several terminations"
    _close(np.array(finished_c), np.array(finished_o), 1e-5, "episode returns")
    ep = env.episode_stats()
    assert int(ep[2]) == n_done and abs(ep[0] - sum(finished_o)) < 1e-4 * max(1, abs(sum(finished_o)))
    env.close()


def test_rollout_statistics_fp32_vs_fp64_1000_episodes(built_lib):
    """BASELINE.json: roll-out reward statistics statistically indistinguishable over >= 1000 episodes."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    stats = {}
    for prec in (64, 32):
        env = SRBVecEnv(helpers.fresh_clips(), 2048, variant="test", precision=prec, auto_reset=True, seed=9)
        env.reset()
        act = torch.zeros(2048, 22, device="cuda")
        rews = []
        for k in range(120):
            env.tracking_action(act, sigma=0.1, seed=77, step=k)
            _, rew, _, _ = env.step(act)
            rews.append(rew.double().mean().item())
        ep = env.episode_stats()
        stats[prec] = (ep, np.mean(rews))
        env.close()
    (e64, r64), (e32, r32) = stats[64], stats[32]
    assert e64[2] >= 1000 and e32[2] >= 1000
    mean_len64, mean_len32 = e64[1] / e64[2], e32[1] / e32[2]
    mean_ret64, mean_ret32 = e64[0] / e64[2], e32[0] / e32[2]
    assert abs(e64[2] - e32[2]) <= 0.02 * e64[2]
    assert abs(mean_len64 - mean_len32) <= 0.02 * mean_len64
    assert abs(mean_ret64 - mean_ret32) <= 0.02 * abs(mean_ret64)
    assert abs(r64 - r32) <= 0.01 * abs(r64)


def test_device_reset_draws_follow_reference_distributions(built_lib):
    """random.randrange / random.randint / np.random.uniform / random_quaternion_perturbation (:810-867)."""
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 4096
    clips = helpers.fresh_clips()
    env = SRBVecEnv(clips, n, variant="test", precision=64, auto_reset=False, seed=3)
    env.reset()
    st = [env.named_state(i) for i in range(0, n, 8)]
    clip = np.array([s["clip"] for s in st]); rif = np.array([s["random_initial_frame"] for s in st])
    assert set(np.unique(clip)) == {0, 1} and abs(clip.mean() - 0.5) < 0.1
    for c, cl in enumerate(clips):
        hi = max(cl["cycle_length"] - 2, 512)
        r = rif[clip == c]
        assert r.min() >= 0 and r.max() <= hi and r.max() > 0.9 * hi
        assert abs(r.mean() - hi / 2) < 0.1 * hi
    # velocity noise within +-0.05 of the table, quaternion within 0.05 rad of the table's
    for s in st[:64]:
        cl = clips[s["clip"]]; f = s["random_initial_frame"]
        dv = s["qvel"][:3] - cl["mocap_vel_com"][f, :3]
        assert np.abs(dv).max() <= 0.05 + 1e-9 and np.abs(s["qvel"][3:] - cl["mocap_vel_com"][f, 3:]).max() < 1e-12
        q0 = cl["mocap_qpos"][f, 3:7] / np.linalg.norm(cl["mocap_qpos"][f, 3:7])
        ang = 2 * np.arccos(min(1.0, abs(np.dot(q0, s["qpos"][3:7]))))
        assert ang <= 0.05 + 1e-6
        assert abs(np.linalg.norm(s["qpos"][3:7]) - 1) < 1e-12
    env.close()


def test_tracking_action_kernel_matches_host_formula(built_lib):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 5
    oracles, plans, rng = _make_oracles(n, "test", seed=55)
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=64, auto_reset=False)
    env.reset(plan=plans)
    act = torch.zeros(n, 22, device="cuda")
    env.tracking_action(act, sigma=0.0, seed=1, step=0)
    a = act.cpu().numpy()
    for i, o in enumerate(oracles):
        ref = helpers.tracking_action(o, np.random.default_rng(0), 0.0, dtype=np.float64)
        assert np.allclose(a[i], ref, atol=2e-6)
    env.close()


@pytest.mark.parametrize("mode,n,lanes", [(3, 64, 8), (3, 61, 8), (3, 37, 4), (2, 64, 8), (2, 61, 8), (2, 37, 4), (1, 64, 8), (0, 64, 8), (1, 61, 8), (1, 37, 4), (1, 16, 2)])
def test_host_buffer_step_equals_device_step(built_lib, mode, n, lanes):
    """srb_step_host (host-buffer path used for the e2e number) returns what srb_step returns: zero-copy mode (the
    coalesced-I/O kernel writing mapped pinned memory, incl. ragged last warps and auto-resets), staged copies, and the
    fallback to staged copies for lane counts the coalesced kernel is not built for."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(8)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    plans[::5, 4] = -8.0                                 # some environments start falling and terminate inside the loop
    e1 = SRBVecEnv(clips, n, precision=64, lanes_per_env=lanes, auto_reset=True, seed=1)
    e2 = SRBVecEnv(clips, n, precision=64, lanes_per_env=lanes, auto_reset=True, seed=1)
    e2.set_host_mode(mode)
    e1.reset(plan=plans); e2.reset(plan=plans)
    act = torch.zeros(n, 22, device="cuda")
    ndone = 0
    for k in range(8):
        e1.tracking_action(act, sigma=0.1, seed=2, step=k)
        o1, r1, d1, i1 = e1.step(act)
        o2, r2, d2, i2 = e2.step_host(act.cpu().numpy())
        torch.cuda.synchronize()
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2) and np.array_equal(i1.cpu().numpy(), i2)
        ndone += int(d2.sum())
    assert ndone > 0
    e1.close(); e2.close()


def test_obs_stats_kernel(built_lib):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    n = 1000
    env = SRBVecEnv(helpers.fresh_clips(), n, precision=32, auto_reset=True, seed=4)
    env.reset()
    acc = torch.zeros(301, dtype=torch.float64, device="cuda")
    env.obs_stats(acc)
    torch.cuda.synchronize()
    o = env.obs.double()
    assert acc[0].item() == n
    assert torch.allclose(acc[1:151], o.sum(0), rtol=1e-12, atol=1e-9) and torch.allclose(acc[151:], (o * o).sum(0), rtol=1e-12, atol=1e-9)
    env.close()


@pytest.mark.parametrize("n", [1, 3, 5, 33])
def test_ragged_batch_sizes(built_lib, n):
    """Batch sizes that do not fill a warp / a 4-environment lane group: every env equals its N=1 evaluation."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(n)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    acts = rng.normal(0, 0.3, (4, n, 22)).astype(np.float32)
    env = SRBVecEnv(clips, n, precision=64, auto_reset=False)
    env.reset(plan=plans)
    outs = []
    for k in range(4):
        o, r, d, _ = env.step(torch.as_tensor(acts[k], device="cuda"))
        outs.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().copy()))
    env.close()
    for i in (0, n - 1):
        e1 = SRBVecEnv(clips, 1, precision=64, auto_reset=False)
        e1.reset(plan=plans[i:i + 1])
        for k in range(4):
            o, r, d, _ = e1.step(torch.as_tensor(acts[k, i:i + 1], device="cuda"))
            assert np.array_equal(o.cpu().numpy()[0], outs[k][0][i]) and r.cpu().numpy()[0] == outs[k][1][i] and d.cpu().numpy()[0] == outs[k][2][i]
        e1.close()


def test_nan_action_is_contained_and_terminates(built_lib):
    """A NaN action poisons only its own environment and the Newton loop still terminates (iteration bounded)."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(2)
    n = 8
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    a = rng.normal(0, 0.1, (n, 22)).astype(np.float32)
    a[:, 21] = 2.0                                      # both feet in contact: the QP runs
    ref = SRBVecEnv(clips, n, precision=32, auto_reset=False); ref.reset(plan=plans)
    o_ref = ref.step(torch.as_tensor(a, device="cuda"))[0].cpu().numpy().copy()
    bad = SRBVecEnv(clips, n, precision=32, auto_reset=False); bad.reset(plan=plans)
    a2 = a.copy(); a2[5, 10:16] = np.nan
    o_bad = bad.step(torch.as_tensor(a2, device="cuda"))[0].cpu().numpy()
    torch.cuda.synchronize()
    ok = [i for i in range(n) if i != 5]
    assert np.array_equal(o_bad[ok], o_ref[ok]) and np.isnan(o_bad[5, :13]).any()
    ref.close(); bad.close()


def test_single_env_proxy_has_reference_surface(built_lib):
    """SRBEnv(...) proxy: reset/step signatures, info keys and the attributes the reference drivers read."""
    from ipcdnnwalk_b200 import SRBEnv, REWARD_INFO_KEYS
    from oracle.srb_env import OracleSRBEnv
    env = SRBEnv(helpers.fresh_clips(), variant="test", precision=64)
    env.testing_mode = True
    obs, info = env.reset(mocap_id=1)
    o = OracleSRBEnv(helpers.fresh_clips(), "test"); o.testing_mode = True
    oo, _ = o.reset_explicit(1, 0, None, None)
    assert obs.shape == (150,) and obs.dtype == np.float32 and info == {} and np.abs(obs - oo).max() < 2e-6
    rng = np.random.default_rng(0)
    for k in range(5):
        a = helpers.tracking_action(o, rng, 0.05)
        obs, r, done, trunc, info = env.step(a)
        oo, rr, dd, _, ii = o.step(a)
        assert trunc is False and set(info["reward_info"]) == set(REWARD_INFO_KEYS) == set(ii["reward_info"])
        assert np.abs(obs - oo).max() < 1e-5 and abs(r - rr) < 1e-5 * max(1, abs(rr)) and done == dd
        assert env.current_frame == o.current_frame and env.left_foot_in_contact == o.left_foot_in_contact
        assert env.site_to_apply_force_on == o.site_to_apply_force_on
        assert np.abs(env.data.qpos - o.data.qpos).max() < 1e-6 and np.abs(env.data.site_xpos - o.data.site_xpos).max() < 1e-6
    st = env.getFullState()
    env.step(helpers.tracking_action(o, rng, 0.05))
    env.restoreFullState(st)
    assert env.current_frame == o.current_frame
    env.close()


def test_pack_obs_extra_matches_oracle(built_lib):
    """pack_obs_extra (DeltaExtractor.py:1015-1036): the Y-up SRB record of the SRB -> full-body mapping after a few steps."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    from oracle.srb_env import OracleSRBEnv
    from oracle import fullbody_map as fm
    rng = np.random.default_rng(11)
    clips = helpers.fresh_clips()
    n = 6
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    env = SRBVecEnv(clips, n, variant="test", precision=64, auto_reset=False)
    env.reset(plan=plans)
    oracles = []
    for i in range(n):
        o = OracleSRBEnv(helpers.fresh_clips(), "test")
        helpers.oracle_reset(o, plans[i])
        oracles.append(o)
    for k in range(6):
        acts = np.stack([helpers.tracking_action(o, rng, 0.05) for o in oracles])
        env.step(torch.as_tensor(acts, device="cuda"))
        ex = env.pack_obs_extra().cpu().numpy()
        for i, o in enumerate(oracles):
            o.step(acts[i])
            ref = fm.pack_obs_extra(o)
            both = np.isnan(ref) & np.isnan(ex[i])
            assert ex[i].shape == (29,) and (np.isnan(ref) == np.isnan(ex[i])).all()
            assert np.abs(np.where(both, 0.0, ex[i] - ref)).max() < 1e-5, (k, i)
            assert (ex[i, 27:29] == ref[27:29]).all()
    env.close()
