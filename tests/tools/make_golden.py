"""Generate the committed fixtures under tests/golden/ from the oracle and the reference's data files.

Run HERE (needs /root/reference, which does not exist on the GPU box):
    python tests/tools/make_golden.py
Writes:
  tests/golden/clip_<name>.npz        reference tables of the shipped clips (oracle/srbdata.py over
                                      gym_trackSRB/motiondata_mujoco_refined/<name>.txt + ref_contact/*.npz)
  tests/golden/rollout_<variant>.npz  seeded oracle roll-outs: reset plans, float32 actions, per-step
                                      obs / reward / done / reward_info / state vector / QP accelerations
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.srbdata import SRBdata, load_mocap_txt          # noqa: E402
from oracle.srb_env import OracleSRBEnv                      # noqa: E402

REF = "/root/reference/gym_trackSRB/"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_clip(name, is_stand):
    md = load_mocap_txt(REF + f"motiondata_mujoco_refined/{name}.txt")
    s = SRBdata(md, REF + f"ref_contact/{name}.bvh.constraint.npz", REF + "Characters/humanoid_deepmimic_withpalm_Tpose.xml")
    s.humanoid2SRB()
    n = s.mocap_com_humanoid.shape[0]
    fo = np.array(s.feet_ori_list)                      # [n,2,3,3] pure yaw
    yaw = np.stack([fo[:, :, 0, 0], fo[:, :, 1, 0]], axis=-1)
    np.savez_compressed(os.path.join(GOLDEN, f"clip_{name}.npz"), mocap_dt=s.dt, cycle_length=s.cycle_length,
                        mocap_qpos=s.mocap_com_humanoid, mocap_ori_com=np.array(s.mocap_ori_com),
                        mocap_vel_com=s.mocap_vel_com, feet_pos_list=s.feet_pos_list, feet_yaw=yaw, is_stand=is_stand)
    print(name, "frames", n, "cycle", s.cycle_length)


def make_rollout(variant, n_envs=6, n_steps=40, seed=7):
    import helpers
    rng = np.random.default_rng(seed)
    clips = helpers.fresh_clips()
    out = dict(plans=[], actions=[], obs=[], rew=[], done=[], rinfo=[], state=[], qdd=[], obs0=[])
    for e in range(n_envs):
        env = OracleSRBEnv(helpers.fresh_clips(), variant)
        plan = helpers.random_plan(rng, clips, noise=(e % 3 != 0))
        obs0, _ = helpers.oracle_reset(env, plan)
        A, O, Rw, D, RI, S, Q = [], [], [], [], [], [], []
        sigma = [0.02, 0.1, 0.3][e % 3]
        for k in range(n_steps):
            a = helpers.tracking_action(env, rng, sigma)
            obs, r, done, _, info = env.step(a)
            A.append(a); O.append(obs); Rw.append(r); D.append(done)
            RI.append([info["reward_info"][k2] for k2 in ("reward_survive", "reward_contact", "reward_end_effector", "reward_delta_pos",
                                                          "reward_delta_ori", "reward_pos", "reward_ori", "reward_total")])
            S.append(helpers.oracle_state_vector(env))
            Q.append([np.full(6, np.nan) if q is None else q for q in env.qdd_qp_list])
        out["plans"].append(plan); out["obs0"].append(obs0)
        for key, v in zip(("actions", "obs", "rew", "done", "rinfo", "state", "qdd"), (A, O, Rw, D, RI, S, Q)):
            out[key].append(np.array(v))
    np.savez_compressed(os.path.join(GOLDEN, f"rollout_{variant}.npz"), **{k: np.array(v) for k, v in out.items()})
    print(variant, "rollout", {k: np.array(v).shape for k, v in out.items()})


def build_builder_case(name="aiming_show", frames=90):
    """raw mocap rows + contact flags + the oracle's tables for them: pins ipcdnnwalk_b200.clip_builder."""
    from oracle.srbdata import read_contact_npz
    md = load_mocap_txt(REF + f"motiondata_mujoco_refined/{name}.txt")[:frames]
    npz = REF + f"ref_contact/{name}.bvh.constraint.npz"
    s = SRBdata(md, npz, REF + "Characters/humanoid_deepmimic_withpalm_Tpose.xml")
    s.humanoid2SRB()
    left, right = read_contact_npz(npz)
    fo = np.array(s.feet_ori_list)
    np.savez_compressed(os.path.join(GOLDEN, "srbdata_case.npz"), mocap_raw=md, left=left[:frames + 5], right=right[:frames + 5],
                        cycle_length=s.cycle_length, mocap_qpos=s.mocap_com_humanoid, mocap_ori_com=np.array(s.mocap_ori_com),
                        mocap_vel_com=s.mocap_vel_com, feet_pos_list=s.feet_pos_list,
                        feet_yaw=np.stack([fo[:, :, 0, 0], fo[:, :, 1, 0]], axis=-1))
    print("builder case", md.shape, s.mocap_com_humanoid.shape)


def build_raw_clips():
    """tests/golden/gym_trackSRB/ref_contact/<name>.motion.npz + <name>.constraint.npz: the shipped clips' raw mocap rows (41-float
    qpos) and contact flags in the reference's own `.npz` clip layout ('motion' key; contact file = path[:-10] + 'constraint.npz',
    env/testSRB_mujoco_original_viewer.py:795-809), so that SRBEnv(mocap_path_list) / ClipBuilder can run where /root/reference is absent."""
    out = os.path.join(GOLDEN, "gym_trackSRB", "ref_contact")
    os.makedirs(out, exist_ok=True)
    for name in ("stand1", "aiming_show"):
        md = load_mocap_txt(REF + f"motiondata_mujoco_refined/{name}.txt")
        np.savez_compressed(os.path.join(out, f"{name}.motion.npz"), motion=np.asarray(md, dtype=np.float64))
        c = np.load(REF + f"ref_contact/{name}.bvh.constraint.npz")
        np.savez_compressed(os.path.join(out, f"{name}.constraint.npz"), **{k: c[k] for k in c})
        print("raw clip", name, np.asarray(md).shape)


def build_terrain():
    """tests/golden/terrain_playground.npz: SRB5_playground.xml + its PNG height fields through the ORACLE's parser (PIL)."""
    from oracle.terrain import load_playground
    T = load_playground(REF + "Characters/SRB5_playground.xml", REF + "terrain")
    T.save(os.path.join(GOLDEN, "terrain_playground.npz"))
    print("terrain: plane geom", T.plane_geom, "hfield geoms", [h["geom_id"] for h in T.hfields], "boxes", len(T.boxes), "bodies", [b["body_id"] for b in T.bodies])


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "terrain":
        build_terrain()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "raw_clips":
        build_raw_clips()
        sys.exit(0)
    build_clip("stand1", True)
    build_clip("aiming_show", False)
    make_rollout("test")
    make_rollout("training")
    build_builder_case()
