"""Generate the committed AdaptiveSRB fixtures under tests/golden/ from the oracle and the reference's data files.

Run HERE (needs /root/reference, which does not exist on the GPU box):
    python tools/make_golden_cdm.py
Writes:
  tests/golden/cdm_walk_tables.npz   reference tables of walk2-v1 (ipcdnnwalk_b200/cdm_tables.py over
                                     work/taesooLib/Resource/motion/MOB1/hanyang_lowdof_T_lafan1_walk_2915.dof and
                                     hanyang_lowdof_T.wrl) + the parsed skeleton as JSON
  tests/golden/cdm_rollout.npz       seeded oracle roll-outs: reset draws, float32 actions, per-step observation /
                                     reward / done / state vectors / QP accelerations and multipliers
  tests/golden/cdm_reference_recordings.npz   the first 700 rows of gym_cdm2/spec/refTraj_{walk2,run2}-v1.dat (roll-outs
                                     recorded by the reference simulator) + the LQR known answer of work/_LQR_CACHE_2.DAT
  tests/golden/cdm_terrain_heights.npz   mHeight of the gym_cdm2 terrain (heightmap1_256_256_2bytes.raw through the loader)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ipcdnnwalk_b200 import cdm_tables as ct                 # noqa: E402
from ipcdnnwalk_b200.vrml_skeleton import parse_wrl          # noqa: E402
from oracle.cdm_env import OracleCDMEnv                      # noqa: E402

MOB1 = "/root/reference/work/taesooLib/Resource/motion/MOB1/"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_tables():
    skel = parse_wrl(MOB1 + "hanyang_lowdof_T.wrl")
    mot = ct.read_dof(MOB1 + "hanyang_lowdof_T_lafan1_walk_2915.dof")
    tab = ct.build_tables(mot, skel)
    np.savez_compressed(os.path.join(GOLDEN, "cdm_walk_tables.npz"), skeleton_json=np.array(json.dumps(skel)), **tab)
    print("tables:", tab["cdm"].shape[0], "frames, cycle", int(tab["nf"]))


def unique_lambda(lam, cmask):
    """oracle multipliers [2*nc, 5] (every point twice) -> [20] in device column order (leg, point, direction)."""
    out = np.zeros(20)
    k = 0
    for ci in range(4):
        if (cmask >> ci) & 1:
            out[5 * ci:5 * ci + 5] = lam[2 * k]
            k += 1
    return out


def make_rollout(n_envs=6, n_steps=90, seed=11):
    import helpers_cdm as hc
    tab, skel = hc.load_tables()
    rng = np.random.default_rng(seed)
    scales = [0.0, 0.05, 0.1, 0.1, 0.3, 1.0][:n_envs]
    envs = [OracleCDMEnv(tab, skel) for _ in range(n_envs)]
    plans0 = np.stack([hc.random_plan(rng, scales[i]) for i in range(n_envs)])
    obs0 = np.stack([hc.oracle_reset(e, plans0[i]) for i, e in enumerate(envs)])
    reals0 = np.stack([hc.oracle_state_vectors(e)[0] for e in envs])
    T = n_steps
    acts = np.zeros((T, n_envs, 10), np.float32); plans = np.zeros((T, n_envs, 16))
    obs = np.zeros((T, n_envs, 27)); rew = np.zeros((T, n_envs)); done = np.zeros((T, n_envs), bool)
    reals = np.zeros((T, n_envs, 61)); ints = np.zeros((T, n_envs, 4), np.int32)
    qdd = np.zeros((T, n_envs, 2, 6)); lam = np.zeros((T, n_envs, 2, 20)); cm = np.zeros((T, n_envs), np.int32)
    scal = np.zeros((T, n_envs, 4)); reset_obs = np.zeros((T, n_envs, 27))
    for t in range(T):
        for i, e in enumerate(envs):
            amp = [0.0, 0.2, 0.5, 1.0, 0.5, 0.5][i]
            a = (rng.uniform(-1, 1, 10) * amp).astype(np.float32)
            acts[t, i] = a
            o, r, d = e.step(a)
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
    np.savez_compressed(os.path.join(GOLDEN, "cdm_rollout.npz"), plans0=plans0, obs0=obs0, reals0=reals0, acts=acts, plans=plans,
                        obs=obs, rew=rew, done=done, reals=reals, ints=ints, qdd=qdd, lam=lam, cmask=cm, scal=scal, reset_obs=reset_obs)
    print("rollout: episodes finished", int(done.sum()), "mean reward", rew.mean(), "swing steps", int((ints[:, :, 0] != 0).sum()))




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


if __name__ == "__main__":
    make_tables()
    make_rollout()
    make_reference_recordings()
    make_terrain()
