"""Generate the committed AdaptiveSRB fixtures under tests/golden/ from the oracle and the reference's data files.

Run HERE (needs /root/reference, which does not exist on the GPU box):
    python tests/tools/make_golden_cdm.py
Writes:
  tests/golden/cdm_walk_tables.npz   reference tables of walk2-v1 (ipcdnnwalk_b200/cdm_tables.py over
                                     work/taesooLib/Resource/motion/MOB1/hanyang_lowdof_T_lafan1_walk_2915.dof and
                                     hanyang_lowdof_T.wrl) + the parsed skeleton as JSON
  tests/golden/cdm_rollout.npz       seeded oracle roll-outs: reset draws, float32 actions, per-step observation /
                                     reward / done / state vectors / QP accelerations and multipliers
  tests/golden/cdm_reference_recordings.npz   the first 700 rows of gym_cdm2/spec/refTraj_{walk2,run2}-v1.dat (roll-outs
                                     recorded by the reference simulator) + the LQR known answer of work/_LQR_CACHE_2.DAT
  tests/golden/cdm_terrain_heights.npz   mHeight of the gym_cdm2 terrain (heightmap1_256_256_2bytes.raw through the loader)
  tests/golden/cdm_policy_walk2.npz  the reference's trained walk2-v1 policy (actor weights + observation normaliser) and gait
                                     statistics of the roll-out the reference recorded with it
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ipcdnnwalk_b200 import cdm_tables as ct                 # noqa: E402
from ipcdnnwalk_b200.vrml_skeleton import parse_wrl          # noqa: E402
from oracle.cdm_env import OracleCDMEnv                      # noqa: E402

MOB1 = "/root/reference/work/taesooLib/Resource/motion/MOB1/"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_tables(spec="walk2-v1", out="cdm_walk_tables.npz"):
    skel = parse_wrl(MOB1 + "hanyang_lowdof_T.wrl")
    mot = ct.read_dof(MOB1 + ct.SPECS[spec]["motion"])
    tab = ct.build_tables(mot, skel, info=ct.SPECS[spec]["info"])
    np.savez_compressed(os.path.join(GOLDEN, out), skeleton_json=np.array(json.dumps(skel)), **tab)
    print("tables:", spec, tab["cdm"].shape[0], "frames, cycle", int(tab["nf"]))


def unique_lambda(lam, cmask):
    """oracle multipliers [2*nc, 5] (every point twice) -> [20] in device column order (leg, point, direction)."""
    out = np.zeros(20)
    k = 0
    for ci in range(4):
        if (cmask >> ci) & 1:
            out[5 * ci:5 * ci + 5] = lam[2 * k]
            k += 1
    return out


def make_rollout(n_envs=6, n_steps=90, seed=11, terrain=False, run2=None):
    """run2: None = walk2-v1; "train" / "gui" = run2-v1 tables with oracle.cdm_env.RUN2_PARAMS (+ RUN2_GUI: limited leg length -> late
    touch-downs, touchDownHeight 0.7 -> early ones)."""
    import helpers_cdm as hc
    from oracle import cdm_env as ce
    tab, skel = hc.load_tables("run" if run2 else "walk")
    rng = np.random.default_rng(seed)
    scales = [0.0, 0.05, 0.1, 0.1, 0.3, 1.0][:n_envs]
    kw = {}
    if terrain:
        T, tpos = hc.load_terrain_oracle()
        kw = dict(terrain=T, terrain_pos=tpos)
    if run2:
        kw["params"] = dict(ce.RUN2_PARAMS, **(dict(ce.RUN2_GUI, healthy_max=2.0) if run2 == "gui" else {}))     # 2.0: high starts must live to a touch-down
    envs = [OracleCDMEnv(tab, skel, **kw) for _ in range(n_envs)]
    late = 0
    OW = 31 if terrain else 27
    plans0 = np.stack([hc.random_plan(rng, scales[i]) for i in range(n_envs)])
    if run2 == "gui":                                   # two bodies start high and are pushed up (late touch-downs), one starts low (touchDownHeight)
        plans0[1, 5] += 0.2; plans0[2, 5] += 0.4; plans0[3, 5] -= 0.15
        plans0[1, 14] = plans0[2, 14] = 2.0
    obs0 = np.stack([hc.oracle_reset(e, plans0[i]) for i, e in enumerate(envs)])
    reals0 = np.stack([hc.oracle_state_vectors(e)[0] for e in envs])
    T = n_steps
    acts = np.zeros((T, n_envs, 10), np.float32); plans = np.zeros((T, n_envs, 16))
    obs = np.zeros((T, n_envs, OW)); rew = np.zeros((T, n_envs)); done = np.zeros((T, n_envs), bool)
    reals = np.zeros((T, n_envs, 61)); ints = np.zeros((T, n_envs, 4), np.int32)
    qdd = np.zeros((T, n_envs, 2, 6)); lam = np.zeros((T, n_envs, 2, 20)); cm = np.zeros((T, n_envs), np.int32)
    scal = np.zeros((T, n_envs, 4)); reset_obs = np.zeros((T, n_envs, OW))
    for t in range(T):
        for i, e in enumerate(envs):
            amp = [0.0, 0.2, 0.5, 1.0, 0.5, 0.5][i]
            a = (rng.uniform(-1, 1, 10) * amp).astype(np.float32)
            if run2 == "gui" and i in (1, 2):
                a[5] = 1.0                                         # desired body position raised
            acts[t, i] = a
            o, r, d = e.step(a)
            late += int(e.debug["late"])
            obs[t, i] = o; rew[t, i] = r; done[t, i] = d
            rr, ii, _ = hc.oracle_state_vectors(e)
            reals[t, i] = rr; ints[t, i] = ii
            cm[t, i] = e.debug["cmask"]
            for it in range(2):
                qdd[t, i, it] = e.debug["qdd"][it]
                lam[t, i, it] = unique_lambda(e.debug["lam"][it], e.debug["cmask"]) if len(e.debug["lam"][it]) else 0.0
            scal[t, i] = (e.debug["pose_dif"], e.debug["com_lvdif"], e.debug["ref_comvel"], e.debug["timing_mod"])
            plans[t, i] = hc.random_plan(rng, scales[i])
            if d:
                reset_obs[t, i] = hc.oracle_reset(e, plans[t, i])
    name = "cdm_rollout_terrain.npz" if terrain else (f"cdm_rollout_run2_{run2}.npz" if run2 else "cdm_rollout.npz")
    np.savez_compressed(os.path.join(GOLDEN, name), plans0=plans0, obs0=obs0, reals0=reals0, acts=acts, plans=plans,
                        obs=obs, rew=rew, done=done, reals=reals, ints=ints, qdd=qdd, lam=lam, cmask=cm, scal=scal, reset_obs=reset_obs)
    print("rollout:", name, "late touch-downs", late, "episodes finished", int(done.sum()), "mean reward", rew.mean(), "swing steps", int((ints[:, :, 0] != 0).sum()))




def make_reference_recordings(rows=700):
    """Slices of the roll-outs the reference simulator itself recorded (gym_cdm2/spec/refTraj_*.dat, written by
    deepmimiccdm-v1.lua:1095-1146) and the reference's cached LQR solution (work/_LQR_CACHE_2.DAT)."""
    from ipcdnnwalk_b200 import tl_binary as tb
    out = {}
    for name in ("walk2", "run2"):
        t = tb.load_table(f"/root/reference/gym_cdm2/spec/refTraj_{name}-v1.dat")
        out[name + "_globalTraj"] = t["globalTraj"][:rows]
        out[name + "_rawFootInfo"] = t["rawFootInfo"][:rows]
        out[name + "_refFrames"] = t["refFrames"][:rows]
    r = tb.BinaryReader(open("/root/reference/work/_LQR_CACHE_2.DAT", "rb").read())
    for k in ("lqr_K", "lqr_A", "lqr_B", "lqr_Q", "lqr_R"):
        out[k] = r.unpack_any()
    np.savez_compressed(os.path.join(GOLDEN, "cdm_reference_recordings.npz"), **out)
    print("reference recordings:", {k: v.shape for k, v in out.items()})


def make_terrain():
    """Height table of the gym_cdm2 terrain spec (deepmimiccdm-v1.lua:368-382): 64 x 64 after the loader's decimation, in cm."""
    from oracle import tl_terrain as tt
    img = tt.load_raw("/root/reference/work/taesooLib/Resource/motion/crowdEditingScripts/terrain/heightmap1_256_256_2bytes.raw", 256, 256)
    np.savez_compressed(os.path.join(GOLDEN, "cdm_terrain_heights.npz"), heights=tt.load_heights(img, 100.0), size_x=4800.0, size_z=4800.0,
                        terrain_pos=np.array([-24.0, 0.1, -24.0]))
    print("terrain heights written")


def _gait_stats(p, sup):
    """forward speed (m/s), mean height, lateral drift per metre, mean swing duration (RL steps) from root positions / support flags."""
    n = len(p)
    dist = float(np.linalg.norm(p[-1, [0, 2]] - p[0, [0, 2]]))
    d = (p[-1, [0, 2]] - p[0, [0, 2]]) / dist
    lat = float(np.abs((p[:, [0, 2]] - p[0, [0, 2]]) @ np.array([-d[1], d[0]])).max())
    durs = []
    for leg in range(2):
        s = sup[:, leg]
        i = 0
        while i < n:
            if s[i] == 0:
                j = i
                while j < n and s[j] == 0:
                    j += 1
                if i > 0 and j < n:
                    durs.append(j - i)
                i = j
            else:
                i += 1
    return np.array([dist / (n / 60.0), float(p[:, 1].mean()), lat / dist, float(np.mean(durs)), float((sup.sum(1) == 2).mean())])


def make_policy_fixture(name="walk2"):
    """The policy the reference trained on walk2-v1 / walkterrain-v1 (trained_models/ppo/walk2-v1.pt = [gym_gang.model.Policy, obs_rms],
    gym_cdm2/train_gym.py:230-233), imported here from /root/reference and flattened to plain arrays, plus gait statistics
    of the roll-out the reference recorded with it (gym_cdm2/spec/refTraj_walk2-v1.dat)."""
    import types
    import torch
    sys.path.insert(0, "/root/reference")
    for mod in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.running_mean_std",
                 "stable_baselines3.common.vec_env", "stable_baselines3.common.vec_env.vec_normalize",
                 "baselines", "baselines.common", "baselines.common.running_mean_std"):      # not installed: stubs for unpickling
        m = types.ModuleType(mod); m.__path__ = []
        sys.modules.setdefault(mod, m)
    sys.modules["stable_baselines3.common.running_mean_std"].RunningMeanStd = type("RunningMeanStd", (), {})
    sys.modules["baselines.common.running_mean_std"].RunningMeanStd = type("RunningMeanStd", (), {})
    sys.modules["stable_baselines3.common.vec_env.vec_normalize"].VecNormalize = type("VecNormalize", (), {})
    sys.modules["stable_baselines3.common.vec_env"].VecNormalize = sys.modules["stable_baselines3.common.vec_env.vec_normalize"].VecNormalize
    ac, rms = torch.load(f"/root/reference/trained_models/ppo/{name}-v1.pt", weights_only=False, map_location="cpu")
    nobs = int(np.asarray(rms.mean).shape[0])
    head = list(ac.dist.fc_mean) if isinstance(ac.dist.fc_mean, torch.nn.Sequential) else [ac.dist.fc_mean]   # older checkpoints: plain Linear
    mods = list(ac.base.actor) + head
    layers = [m for m in mods if isinstance(m, torch.nn.Linear)]
    acts = [type(m).__name__ for m in mods if not isinstance(m, torch.nn.Linear)]
    final_tanh = len(acts) == len(layers)
    assert all(a == "Tanh" for a in acts) and len(acts) in (len(layers), len(layers) - 1), "expected Linear/Tanh pairs"
    out = {f"W{i}": l.weight.detach().numpy().astype(np.float64) for i, l in enumerate(layers)}
    out.update({f"b{i}": l.bias.detach().numpy().astype(np.float64) for i, l in enumerate(layers)})
    out["obs_mean"] = np.asarray(rms.mean, np.float64); out["obs_var"] = np.asarray(rms.var, np.float64)
    # the flattened network reproduces Policy.act(deterministic=True)
    x = np.random.default_rng(0).normal(size=(16, nobs)).astype(np.float32)
    with torch.no_grad():
        _, a, _, _ = ac.act(torch.from_numpy(x), torch.zeros(16, ac.recurrent_hidden_state_size), torch.ones(16, 1), deterministic=True)
    h = x.astype(np.float64)
    for i in range(len(layers)):
        h = h @ out[f"W{i}"].T + out[f"b{i}"]
        if i < len(layers) - 1 or final_tanh:
            h = np.tanh(h)
    assert np.abs(h - a.numpy()).max() < 1e-5
    out["final_tanh"] = np.array(int(final_tanh))
    from ipcdnnwalk_b200 import tl_binary as tb
    t = tb.load_table(f"/root/reference/gym_cdm2/spec/refTraj_{name}-v1.dat")
    out["recorded_gait"] = _gait_stats(t["globalTraj"][:, 0:3], t["rawFootInfo"][:, 0:2])
    np.savez_compressed(os.path.join(GOLDEN, f"cdm_policy_{name}.npz"), n_layers=len(layers), **out)
    print("policy:", [out[f"W{i}"].shape for i in range(len(layers))], "recorded gait [speed, height, lateral/m, swing steps, double support]:", out["recorded_gait"])


def make_run2():
    """run2-v1 (gym_cdm2/spec/run2-v1.lua over motionInfos.run4): tables, roll-outs in the training and the GUI configuration, trained policy."""
    make_tables("run2-v1", "cdm_run_tables.npz")
    make_rollout(n_envs=6, n_steps=70, seed=21, run2="train")
    make_rollout(n_envs=6, n_steps=70, seed=22, run2="gui")
    make_policy_fixture("run2")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "run2":
        make_run2()
        sys.exit(0)
    make_tables()
    make_rollout()
    make_reference_recordings()
    make_terrain()
    make_rollout(n_envs=6, n_steps=60, seed=12, terrain=True)
    make_policy_fixture("walk2")
    make_policy_fixture("walkterrain")
    make_run2()
