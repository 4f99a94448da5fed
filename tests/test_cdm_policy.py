"""Closed-loop check of the AdaptiveSRB environment with the policy THE REFERENCE TRAINED on it
(trained_models/ppo/walk2-v1.pt, flattened into tests/golden/cdm_policy_walk2.npz by tests/tools/make_golden_cdm.py): a policy
trained against the reference's environment only keeps walking in an environment that presents the same observation
layout, action semantics, contact logic and dynamics.  Gait statistics are set beside those of the roll-out the
reference recorded with the same policy (gym_cdm2/spec/refTraj_walk2-v1.dat); the reference tables used here are the
NumPy restatement with stated stand-ins (cdm_tables.py) and the recording was made in GUI mode, so the comparison is a
band, not an equality."""
import os

import numpy as np
import pytest

import helpers
import helpers_cdm as hc


@pytest.fixture(scope="module")
def policy():
    d = dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_walk2.npz")))
    return d


@pytest.fixture(scope="module")
def policy_terrain():
    return dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_walkterrain.npz")))


def act_numpy(pol, obs):
    """gym_gang.model.Policy.act(deterministic=True) behind SB3 VecNormalize(norm_obs=True) (gym_cdm2/test_gym.py:89-145)."""
    h = np.clip((obs - pol["obs_mean"]) / np.sqrt(pol["obs_var"] + 1e-8), -10.0, 10.0).astype(np.float32).astype(np.float64)
    n = int(pol["n_layers"])
    for i in range(n):
        h = h @ pol[f"W{i}"].T + pol[f"b{i}"]
        if i < n - 1 or int(pol["final_tanh"]):
            h = np.tanh(h)
    return h.astype(np.float32)


def gait_stats(p, swing):
    n = len(p)
    dist = float(np.linalg.norm(p[-1, [0, 2]] - p[0, [0, 2]]))
    d = (p[-1, [0, 2]] - p[0, [0, 2]]) / dist
    lat = float(np.abs((p[:, [0, 2]] - p[0, [0, 2]]) @ np.array([-d[1], d[0]])).max())
    durs = []
    for leg in range(2):
        s = swing[:, leg]
        i = 0
        while i < n:
            if s[i]:
                j = i
                while j < n and s[j]:
                    j += 1
                if i > 0 and j < n:
                    durs.append(j - i)
                i = j
            else:
                i += 1
    return dist / (n / 60.0), float(p[:, 1].mean()), lat / dist, float(np.mean(durs)), float((~swing).all(1).mean())


def test_reference_policy_walks_in_the_oracle(policy):
    from oracle.cdm_env import OracleCDMEnv
    tab, skel = hc.load_tables()
    env = OracleCDMEnv(tab, skel)
    obs = env.reset()
    P, S, R = [], [], []
    for t in range(480):                                  # 8 s, ~14 gait cycles
        obs, r, d = env.step(act_numpy(policy, obs))
        assert not d, f"the reference's policy fell at step {t}"
        P.append(env.p.copy()); S.append([env.leg[1].isSwing, env.leg[2].isSwing]); R.append(r)
    speed, height, lat, swing_steps, dsup = gait_stats(np.array(P), np.array(S))
    rs, rh, rl, rsw, rds = policy["recorded_gait"]
    assert abs(speed - rs) < 0.15 * rs, (speed, rs)             # forward speed of the recorded roll-out: 1.75 m/s
    assert abs(height - rh) < 0.03                               # CoM height
    assert lat < 0.05                                            # walks straight
    assert abs(swing_steps - rsw) < 0.2 * rsw                   # swing duration in RL steps
    assert np.mean(R) > 0.2 and min(np.array(P)[:, 1]) > 0.78


def test_reference_run_policy_runs_in_the_oracle():
    """run2-v1: trained_models/ppo/run2-v1.pt on the tables of the running clip with the run2-v1 parameter set; the roll-out the
    reference recorded with it (gym_cdm2/spec/refTraj_run2-v1.dat) runs at 2.91 m/s, CoM height 0.889, 17.6-step swings, 2.3 %
    double support."""
    from oracle import cdm_env as ce
    pol = dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_run2.npz")))
    tab, skel = hc.load_tables("run")
    env = ce.OracleCDMEnv(tab, skel, params=ce.RUN2_PARAMS)
    obs = env.reset()
    P, S, R = [], [], []
    for t in range(360):
        obs, r, d = env.step(act_numpy(pol, obs))
        assert not d, f"the reference's run policy fell at step {t}"
        P.append(env.p.copy()); S.append([env.leg[1].isSwing, env.leg[2].isSwing]); R.append(r)
    speed, height, lat, swing_steps, dsup = gait_stats(np.array(P), np.array(S))
    rs, rh, rl, rsw, rds = pol["recorded_gait"]
    assert abs(speed - rs) < 0.05 * rs, (speed, rs)
    assert abs(height - rh) < 0.01, (height, rh)
    assert lat < 0.03 and abs(swing_steps - rsw) < 0.1 * rsw and abs(dsup - rds) < 0.02
    assert np.mean(R) > 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_reference_run_policy_on_the_gpu(precision):
    """run2-v1 closed loop on the device through CDMVecEnv.from_spec: the noise-free environment follows the oracle, the ensemble
    keeps running at the recorded speed."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    from oracle import cdm_env as ce
    pol = dict(np.load(os.path.join(helpers.GOLDEN, "cdm_policy_run2.npz")))
    tab, skel = hc.load_tables("run")
    n, T = 256, 300
    rng = np.random.default_rng(7)
    plans = np.stack([hc.random_plan(rng, 0.2) for _ in range(n)])
    plans[0] = 0.0
    env = CDMVecEnv.from_spec("run2-v1", n, tables=tab, skel=skel, precision=precision, auto_reset=False)
    assert env.get_param("k_p") == 340.0 and env.get_param("stride_thr") < 0 and env.get_param("healthy_max") == 1.1
    obs = env.reset(plan=plans)
    nl = int(pol["n_layers"])
    W = [torch.as_tensor(pol[f"W{i}"], dtype=torch.float32, device="cuda") for i in range(nl)]
    B = [torch.as_tensor(pol[f"b{i}"], dtype=torch.float32, device="cuda") for i in range(nl)]
    mean = torch.as_tensor(pol["obs_mean"], dtype=torch.float32, device="cuda")
    istd = torch.as_tensor(1.0 / np.sqrt(pol["obs_var"] + 1e-8), dtype=torch.float32, device="cuda")
    oracle = ce.OracleCDMEnv(tab, skel, params=ce.RUN2_PARAMS)
    o_obs = hc.oracle_reset(oracle, plans[0])
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    worst = 0.0
    for t in range(T):
        h = torch.clamp((obs - mean) * istd, -10.0, 10.0)
        for i, (w, b) in enumerate(zip(W, B)):
            h = h @ w.T + b
            if i < nl - 1 or int(pol["final_tanh"]):
                h = torch.tanh(h)
        obs, rew, done, _ = env.step(h)
        alive &= done == 0
        if t < 100:
            o_obs, o_r, o_d = oracle.step(act_numpy(pol, o_obs))
            worst = max(worst, float(np.abs(obs[0].cpu().numpy() - o_obs).max()))
    torch.cuda.synchronize()
    assert worst < (2e-3 if precision == 64 else 5e-2), worst
    assert float(alive.float().mean()) > 0.9
    z = np.array([env.get_state(i)["reals"][2] for i in range(0, n, 8)])
    speed = z[alive.cpu().numpy()[::8]] / (T / 60.0)
    assert abs(speed.mean() - pol["recorded_gait"][0]) < 0.1 * pol["recorded_gait"][0], speed.mean()
    env.close()


def test_reference_terrain_policy_walks_over_the_terrain_in_the_oracle(policy_terrain):
    """walkterrain-v1 (deepmimiccdmterrain-v1.lua): the reference's terrain policy (31-float observation with 4 height
    samples) crosses the height field in the terrain oracle -- 8 s, ground under the path between 0.1 and 0.6 m."""
    from oracle.cdm_env import OracleCDMEnv
    tab, skel = hc.load_tables()
    T, tpos = hc.load_terrain_oracle()
    env = OracleCDMEnv(tab, skel, terrain=T, terrain_pos=tpos)
    obs = env.reset()
    assert obs.shape == (31,)
    P, S, H = [], [], []
    for t in range(480):
        obs, r, d = env.step(act_numpy(policy_terrain, obs))
        assert not d, f"the reference's terrain policy fell at step {t}"
        P.append(env.p.copy()); S.append([env.leg[1].isSwing, env.leg[2].isSwing]); H.append(env.G(env.p))
    P, H = np.array(P), np.array(H)
    speed, _, lat, swing_steps, _ = gait_stats(P, np.array(S))
    rs, rh, rl, rsw, rds = policy_terrain["recorded_gait"]         # recorded on flat ground (collectRefTraj sets terrainHeight = 0)
    assert abs(speed - rs) < 0.15 * rs and abs(swing_steps - rsw) < 0.2 * rsw and lat < 0.05
    assert H.max() - H.min() > 0.25                                 # the path really climbs and descends
    above = P[:, 1] - H
    assert above.min() > 0.75 and above.max() < 0.95               # body height above the terrain stays in the walking band


@pytest.mark.gpu
def test_reference_terrain_policy_on_the_gpu(policy_terrain):
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    from oracle.cdm_env import OracleCDMEnv
    tab, skel = hc.load_tables()
    n, T = 128, 300
    rng = np.random.default_rng(6)
    plans = np.stack([hc.random_plan(rng, 0.2) for _ in range(n)])
    plans[0] = 0.0
    env = CDMVecEnv(tab, skel, n, precision=64, auto_reset=False, terrain=hc.load_terrain())
    obs = env.reset(plan=plans)
    assert obs.shape == (n, 31)
    pol = policy_terrain
    nl = int(pol["n_layers"])
    W = [torch.as_tensor(pol[f"W{i}"], dtype=torch.float32, device="cuda") for i in range(nl)]
    B = [torch.as_tensor(pol[f"b{i}"], dtype=torch.float32, device="cuda") for i in range(nl)]
    mean = torch.as_tensor(pol["obs_mean"], dtype=torch.float32, device="cuda")
    istd = torch.as_tensor(1.0 / np.sqrt(pol["obs_var"] + 1e-8), dtype=torch.float32, device="cuda")
    Tt, tpos = hc.load_terrain_oracle()
    oracle = OracleCDMEnv(tab, skel, terrain=Tt, terrain_pos=tpos)
    o_obs = hc.oracle_reset(oracle, plans[0])
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    worst = 0.0
    for t in range(T):
        h = torch.clamp((obs - mean) * istd, -10.0, 10.0)
        for i, (w, b) in enumerate(zip(W, B)):
            h = h @ w.T + b
            if i < nl - 1 or int(pol["final_tanh"]):
                h = torch.tanh(h)
        obs, rew, done, _ = env.step(h)
        alive &= done == 0
        if t < 100:
            o_obs, o_r, o_d = oracle.step(act_numpy(pol, o_obs))
            worst = max(worst, float(np.abs(obs[0].cpu().numpy() - o_obs).max()))
    torch.cuda.synchronize()
    assert worst < 2e-3, worst
    assert float(alive.float().mean()) > 0.9
    z = np.array([env.get_state(i)["reals"][2] for i in range(0, n, 8)])
    assert abs(z[alive.cpu().numpy()[::8]].mean() / (T / 60.0) - pol["recorded_gait"][0]) < 0.2 * pol["recorded_gait"][0]
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_reference_policy_walks_on_the_gpu(policy, precision):
    """The same closed loop on the device: 256 environments started with the reference's reset noise (scaled to 20 %, the
    policy is not expected to recover from the full +-0.5 quaternion noise), policy evaluated in torch on the GPU.  The noise-free
    environment must follow the oracle closely; the ensemble must keep walking at the recorded speed."""
    import torch
    from ipcdnnwalk_b200 import CDMVecEnv
    from oracle.cdm_env import OracleCDMEnv
    tab, skel = hc.load_tables()
    n, T = 256, 360
    rng = np.random.default_rng(5)
    plans = np.stack([hc.random_plan(rng, 0.2) for _ in range(n)])
    plans[0] = 0.0
    env = CDMVecEnv(tab, skel, n, precision=precision, auto_reset=False)
    obs = env.reset(plan=plans)
    W = [torch.as_tensor(policy[f"W{i}"], dtype=torch.float32, device="cuda") for i in range(int(policy["n_layers"]))]
    B = [torch.as_tensor(policy[f"b{i}"], dtype=torch.float32, device="cuda") for i in range(int(policy["n_layers"]))]
    mean = torch.as_tensor(policy["obs_mean"], dtype=torch.float32, device="cuda")
    istd = torch.as_tensor(1.0 / np.sqrt(policy["obs_var"] + 1e-8), dtype=torch.float32, device="cuda")
    oracle = OracleCDMEnv(tab, skel)
    o_obs = hc.oracle_reset(oracle, plans[0])
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    rsum = torch.zeros(n, device="cuda")
    z0 = None
    worst = 0.0
    for t in range(T):
        h = torch.clamp((obs - mean) * istd, -10.0, 10.0)
        for w, b in zip(W, B):
            h = torch.tanh(h @ w.T + b)                   # walk2-v1 head ends in Tanh (final_tanh = 1)
        obs, rew, done, _ = env.step(h)
        alive &= done == 0
        rsum += rew * alive
        if t < 120:                                       # environment 0 (no noise) against the oracle driven by ITS OWN policy output
            o_obs, o_r, o_d = oracle.step(act_numpy(policy, o_obs))
            worst = max(worst, float(np.abs(obs[0].cpu().numpy() - o_obs).max()))
    torch.cuda.synchronize()
    assert worst < (2e-3 if precision == 64 else 5e-2), worst      # closed loop amplifies float32 policy round-off; stays small
    frac = float(alive.float().mean())
    assert frac > 0.9, f"only {frac:.2f} of the environments kept walking"
    z = np.array([env.get_state(i)["reals"][2] for i in range(0, n, 8)])
    al = alive.cpu().numpy()[::8]
    speed = z[al] / (T / 60.0)
    assert abs(speed.mean() - policy["recorded_gait"][0]) < 0.15 * policy["recorded_gait"][0], speed.mean()
    assert float((rsum[alive] / T).mean()) > 0.2
    env.close()
