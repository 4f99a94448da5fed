"""GPU parity of the batched AdaptiveSRB step (`deepmimiccdm-v1`, SURVEY 8a rows a-16 ... a-20) against the oracle and
the committed golden roll-out, all through the C-ABI (include/cdmenv_b200.h).

Tolerances (BASELINE north star): fp64 1e-5 relative on state / observation / reward / QP accelerations, fp32 1e-3
single-step; integer outputs (swing flags, replay length, done) bit exact."""
import numpy as np
import pytest

import helpers_cdm as hc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data():
    tab, skel = hc.load_tables()
    roll = dict(np.load(hc.ROLLOUT))
    return tab, skel, roll


def _close(a, b, rtol, atol):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)) <= atol + rtol * np.abs(np.asarray(b, float))


def _assert_close(a, b, rtol, atol, what):
    ok = _close(a, b, rtol, atol)
    if not ok.all():
        idx = np.argwhere(~ok)[:8]
        a = np.asarray(a, float); b = np.asarray(b, float)
        raise AssertionError(f"{what}: {(~ok).sum()} mismatches, first at {idx.tolist()} got {[float(a[tuple(i)]) for i in idx]} want {[float(b[tuple(i)]) for i in idx]}")


def _run_golden(tab, skel, roll, lanes, precision=64, rtol=1e-5, atol=1e-7, steps=None, terrain=None, params=None):
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    n = roll["plans0"].shape[0]
    T = roll["acts"].shape[0] if steps is None else steps
    env = CDMVecEnv(tab, skel, n, precision=precision, lanes_per_env=lanes, auto_reset=False, terrain=terrain, params=params)
    env.attach(dbg=True)
    obs = env.reset(plan=roll["plans0"])
    torch.cuda.synchronize()
    _assert_close(obs.cpu().numpy(), roll["obs0"], rtol, 10 * atol, "reset observation")
    for i in range(n):
        _assert_close(env.get_state(i)["reals"][:22], roll["reals0"][i][:22], rtol, atol, f"reset state env {i}")
    for t in range(T):
        obs, rew, done, _ = env.step(torch.as_tensor(roll["acts"][t], device="cuda"))
        torch.cuda.synchronize()
        what = f"step {t}"
        assert (done.cpu().numpy().astype(bool) == roll["done"][t]).all(), what + ": done flags differ"
        dbg = env.dbg.cpu().numpy()
        _assert_close(dbg[:, 0:12].reshape(n, 2, 6), roll["qdd"][t], rtol, 1e-6, what + " qdd")
        _assert_close(dbg[:, 64:68], roll["scal"][t], rtol, 1e-6, what + " pose_dif/com_lvdif/ref_comvel/timing_mod")
        assert (dbg[:, 77].astype(int) == np.array([2 * bin(int(m)).count("1") for m in roll["cmask"][t]])).all(), what + ": contact count differs"
        # multipliers: weakly determined individually (ridge 1e-8), the net wrench is what qdd checks; compare loosely
        _assert_close(dbg[:, 24:64].reshape(n, 2, 20), roll["lam"][t], 1e-3, 1e-3 * max(1.0, np.abs(roll["lam"][t]).max()), what + " lambda")
        _assert_close(obs.cpu().numpy(), roll["obs"][t], rtol, 10 * atol, what + " observation")
        _assert_close(rew.cpu().numpy(), roll["rew"][t], rtol, 1e-7, what + " reward")
        for i in range(n):
            st = env.get_state(i)
            assert int(st["ints"][0]) == int(roll["ints"][t, i, 0]), f"{what} env {i}: swing flags differ"
            assert int(st["ints"][1]) == int(roll["ints"][t, i, 1]), f"{what} env {i}: replay length differs"
            sel = np.ones(61, bool); sel[22] = False
            _assert_close(st["reals"][sel], roll["reals"][t, i][sel], rtol, atol, f"{what} env {i} state")
        dn = np.nonzero(roll["done"][t])[0]
        if len(dn):
            env.reset(env_idx=dn, plan=roll["plans"][t][dn])
            torch.cuda.synchronize()
            _assert_close(env.obs.cpu().numpy()[dn], roll["reset_obs"][t][dn], rtol, 10 * atol, what + " reset observation")
    env.close()


@pytest.mark.parametrize("lanes", [4, 8, 1, 2])
def test_golden_rollout_fp64(data, lanes):
    tab, skel, roll = data
    _run_golden(tab, skel, roll, lanes)


@pytest.mark.parametrize("lanes", [1, 2])
def test_golden_rollout_terrain_fp64(data, lanes):
    """walkterrain-v1 variant (deepmimiccdmterrain-v1.lua): 31-float observation, projectToGround on the height field,
    terrain-following targets -- golden roll-out of the terrain oracle incl. resets."""
    tab, skel, _ = data
    roll = dict(np.load(hc.ROLLOUT.replace("cdm_rollout.npz", "cdm_rollout_terrain.npz")))
    assert roll["obs"].shape[2] == 31
    _run_golden(tab, skel, roll, lanes, terrain=hc.load_terrain())


@pytest.mark.parametrize("mode,lanes", [("train", 1), ("train", 4), ("gui", 1), ("gui", 8)])
def test_golden_rollout_run2_fp64(mode, lanes):
    """run2-v1 (gym_cdm2/spec/run2-v1.lua over motionInfos.run4: k_p_ID 340, EFdif_scale 1/8, no strideThr, healthy range
    0.8-1.1, ...) on the tables of the running clip, in the training configuration and in the GUI one (limited leg length: the
    roll-out holds 9 late touch-downs, deepmimiccdm-v1.lua:702-709,1150-1159; touchDownHeight 0.7)."""
    from ipcdnnwalk_b200 import cdm_tables as ct
    tab, skel = hc.load_tables("run")
    roll = dict(np.load(hc.ROLLOUT.replace("cdm_rollout.npz", f"cdm_rollout_run2_{mode}.npz")))
    params = dict(ct.SPECS["run2-v1"]["params"])
    if mode == "gui":
        params.update(ct.SPECS["run2-v1"]["gui_params"], healthy_max=2.0)
    _run_golden(tab, skel, roll, lanes, params=params)


def test_fp32_single_step(data):
    """fp32 state mode (QP kept in fp64): one step from the oracle's fp64 state, 1e-3."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    tab, skel, roll = data
    n = roll["plans0"].shape[0]
    env = CDMVecEnv(tab, skel, n, precision=32, auto_reset=False)
    env.attach(dbg=True)
    env.reset(plan=roll["plans0"])
    nr, ni, hs, hw = env.state_dims()
    prev_reals = roll["reals0"]
    from oracle.cdm_env import OracleCDMEnv
    oracles = [OracleCDMEnv(tab, skel) for _ in range(n)]
    for i, o in enumerate(oracles):
        hc.oracle_reset(o, roll["plans0"][i])
    worst = 0.0
    for t in range(40):
        for i, o in enumerate(oracles):
            r, ints, hist = hc.oracle_state_vectors(o)
            env.set_state(i, dict(reals=r, ints=ints, hist=hist))
        obs, rew, done, _ = env.step(torch.as_tensor(roll["acts"][t], device="cuda"))
        torch.cuda.synchronize()
        for i, o in enumerate(oracles):
            oo, rr, dd = o.step(roll["acts"][t, i])
            assert bool(done[i]) == dd, f"step {t} env {i}: done differs"
            _assert_close(obs[i].cpu().numpy(), oo, 1e-3, 1e-3, f"fp32 step {t} env {i} obs")
            _assert_close(float(rew[i]), rr, 1e-3, 1e-4, f"fp32 step {t} env {i} reward")
            st = env.get_state(i)
            r, ints, _ = hc.oracle_state_vectors(o)
            assert int(st["ints"][0]) == int(ints[0])
            _assert_close(st["reals"][:13], r[:13], 1e-3, 1e-3, f"fp32 step {t} env {i} rigid-body state")
            _assert_close(env.dbg[i, 0:12].cpu().numpy(), np.concatenate(o.debug["qdd"]), 1e-3, 2e-2, f"fp32 step {t} env {i} qdd")
            worst = max(worst, float(np.abs(st["reals"][:13] - r[:13]).max()))
            if dd:
                hc.oracle_reset(o, roll["plans"][t, i])
    assert worst < 1e-3
    env.close()


def test_auto_reset_with_plans(data):
    """Fused auto-reset: terminal observation kept aside, reset observation returned, same stream as explicit resets."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    tab, skel, roll = data
    n = roll["plans0"].shape[0]
    env = CDMVecEnv(tab, skel, n, precision=64, auto_reset=True)
    env.attach(term_obs=True, reset_plan=True)
    env.reset(plan=roll["plans0"])
    ndone = 0
    for t in range(roll["acts"].shape[0]):
        env.reset_plan.copy_(torch.as_tensor(roll["plans"][t]))
        obs, rew, done, _ = env.step(torch.as_tensor(roll["acts"][t], device="cuda"))
        torch.cuda.synchronize()
        d = roll["done"][t]
        assert (done.cpu().numpy().astype(bool) == d).all()
        o = obs.cpu().numpy()
        _assert_close(o[~d], roll["obs"][t][~d], 1e-5, 1e-6, f"step {t} obs")
        if d.any():
            _assert_close(o[d], roll["reset_obs"][t][d], 1e-5, 1e-6, f"step {t} reset obs")
            _assert_close(env.term_obs.cpu().numpy()[d], roll["obs"][t][d], 1e-5, 1e-6, f"step {t} terminal obs")
            ndone += int(d.sum())
    st = env.episode_stats()
    assert int(st[2]) == ndone and ndone > 0
    env.close()


def test_host_path_equals_device_path(data):
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    tab, skel, roll = data
    n = roll["plans0"].shape[0]
    a = CDMVecEnv(tab, skel, n, auto_reset=False); b = CDMVecEnv(tab, skel, n, auto_reset=False)
    a.reset(plan=roll["plans0"]); b.reset(plan=roll["plans0"])
    for t in range(10):
        o1, r1, d1, _ = a.step(torch.as_tensor(roll["acts"][t], device="cuda"))
        o2, r2, d2 = b.step_host(roll["acts"][t])
        torch.cuda.synchronize()
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(d1.cpu().numpy(), d2)
    a.close(); b.close()


def test_device_reset_draws_and_nan_containment(data):
    """Device RNG draws follow the reference's distributions; a NaN action is replaced by a random one as the
    reference's step() wrapper does (deepmimiccdm-v1.lua:858-863); a NaN state ends the episode with a zero observation
    and zero reward (:1557-1561)."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    tab, skel, roll = data
    n = 4096
    env = CDMVecEnv(tab, skel, n, precision=64, auto_reset=True, seed=5)
    env.reset()
    torch.cuda.synchronize()
    pos = np.stack([env.get_state(i)["reals"][:3] for i in range(0, n, 64)])
    assert np.abs(pos[:, 0]).max() <= 0.05 + 1e-12 and np.abs(pos[:, 2]).max() <= 0.05 + 1e-12 and pos[:, 0].std() > 0.01
    act = torch.zeros(n, 10, device="cuda")
    act[7] = float("nan")
    obs, rew, done, _ = env.step(act)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    st = env.get_state(9)                                 # poison one state: NaN -> zero observation, reward 0, done
    st["reals"][0] = float("nan")
    env.set_state(9, st)
    env.attach(term_obs=True)
    obs, rew, done, _ = env.step(torch.zeros(n, 10, device="cuda"))
    torch.cuda.synchronize()
    assert bool(done[9]) and float(rew[9]) == 0.0 and float(env.term_obs[9].abs().max()) == 0.0
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    for _ in range(20):
        obs, rew, done, _ = env.step(torch.rand(n, 10, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert 0.0 <= float(rew.min()) and float(rew.max()) <= 1.0
    env.close()


def test_luaenv_proxy(data):
    from types import SimpleNamespace
    from ipcdnnwalk_b200 import LuaEnv
    tab, skel, roll = data
    e = LuaEnv(SimpleNamespace(tables=tab, skel=skel, testMode=True))
    o, info = e.reset()
    assert o.dtype == np.float32 and o.shape == (27,) and info == {}
    s, r, d, tr, info = e.step(np.zeros(10, dtype=np.float32))
    assert s.shape == (27,) and isinstance(r, float) and tr is False and info == {}
    e.close()


@pytest.mark.parametrize("precision,rtol,atol", [(64, 1e-5, 1e-6), (32, 5e-3, 5e-3)])
def test_sample_at_bench_size_16384(data, precision, rtol, atol):
    """BASELINE configs[3] batch (16384 envs, default lane count = one thread per environment): 48 random environments
    of the full batch against the oracle over 8 steps with U(-1,1) actions (the workload of tools/bench_cdm.py).
    fp64 at the single-step tolerance; fp32 free-running, so the accumulated bound (5e-3); swing flags / replay length /
    done bit exact in both."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    from oracle.cdm_env import OracleCDMEnv
    tab, skel, _ = data
    n = 16384
    rng = np.random.default_rng(16384)
    plans = np.stack([hc.random_plan(rng, 0.3) for _ in range(n)])      # 0.3 x the reset noise: most episodes outlive the 8 steps
    env = CDMVecEnv(tab, skel, n, precision=precision, auto_reset=False)
    env.reset(plan=plans)
    sample = sorted(set(rng.choice(n, 46, replace=False).tolist() + [0, n - 1]))
    oracles = {}
    for i in sample:
        o = OracleCDMEnv(tab, skel)
        hc.oracle_reset(o, plans[i])
        oracles[i] = o
    alive = set(sample)
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    idx = torch.as_tensor(sample, device="cuda")
    for t in range(8):
        act = torch.rand(n, 10, device="cuda", generator=gen) * 2 - 1
        a = act[idx].cpu().numpy()
        obs, rew, done, _ = env.step(act)
        torch.cuda.synchronize()
        o_s = obs[idx].cpu().numpy(); r_s = rew[idx].cpu().numpy(); d_s = done[idx].cpu().numpy()
        for j, i in enumerate(sample):
            if i not in alive:
                continue
            oo, rr, dd = oracles[i].step(a[j])
            assert bool(d_s[j]) == dd, f"step {t} env {i}: done differs"
            if dd:
                alive.discard(i)              # NaN / terminal state: the reference returns zeros; not compared further
                continue
            _assert_close(o_s[j], oo, rtol, atol, f"step {t} env {i} obs")
            _assert_close(float(r_s[j]), rr, rtol, atol, f"step {t} env {i} reward")
            hc.compare_state(env.get_state(i), oracles[i], rtol, atol, f"step {t} env {i}")
    assert len(alive) >= 24
    env.close()
