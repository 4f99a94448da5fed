"""Record roll-outs of the REFERENCE'S OWN SRBEnv code (run through tests/tools/ref_srbenv_harness.py: MuJoCo / Clarabel / SDS /
SRBdata answered by stated stand-ins, everything else the reference's code as it stands) as committed fixtures:

    python tests/tools/make_golden_refenv.py        ->  tests/golden/refenv_rollout_{test,training,terrain}.npz,
                                                        tests/golden/refenv_special.npz

Same layout as tests/golden/rollout_{test,training}.npz (tests/tools/make_golden.py, which records the ORACLE): per environment a
reset plan, the reset observation, 40 float32 actions and per step the observation, reward, done, the eight reward_info terms, the
state vector (helpers.oracle_state_vector) and the QP accelerations of the four sub-steps -- here read from the reference environment
object (`q_ddot_QP` as solve_qp_for_contact_forces returns it).  The actions come from helpers.tracking_action on an oracle environment
stepped alongside (it needs the state; the two agree to 1e-13, which this script asserts).
refenv_delta.npz: the reference's SRB -> full-body mapping code (gym_trackSRB/taesooLib/DeltaExtractor.py through tests/tools/ref_delta_harness.py):
pack_obs_extra on the reference environment while it steps, unpack_obs_extra + _extractSingleFrameFullbody on seeded decoder rows (SRB tree
momentum by the compiled reference), removeCOMsliding_online on seeded windows (full-body CoM by the compiled reference).
refenv_sds.npz: the reference's swing-foot filter class by itself (work/controlmodule.py SDS through tests/tools/ref_sds_harness.py) driven
the way SRBEnv.step drives it, over the whole gain range (logQ 5 ... 11, the environment uses 9 ... 11).
refenv_special.npz: the policy-probability contact hysteresis (`_last_contact_prob` side channel, SRBEnv5_mocap_list_window.py:239-275)
and the push-impulse schedule with a 6-component force (:191-206) -- branches the plain roll-outs do not reach."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import helpers                                   # noqa: E402
import ref_srbenv_harness as rh                  # noqa: E402
from oracle.srb_env import OracleSRBEnv          # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
RI_KEYS = ("reward_survive", "reward_contact", "reward_end_effector", "reward_delta_pos", "reward_delta_ori", "reward_pos", "reward_ori", "reward_total")


def _capture_qdd(mod):
    """wrap the reference's solve_qp_for_contact_forces so that q_ddot_QP of every call is kept"""
    log = []
    if not hasattr(mod, "_orig_solve_qp"):
        mod._orig_solve_qp = mod.solve_qp_for_contact_forces

    def wrapped(*a, **k):
        r = mod._orig_solve_qp(*a, **k)
        log.append(np.array(r[0], dtype=float))
        return r
    mod.solve_qp_for_contact_forces = wrapped
    return log


def rollout(variant, n_envs=6, n_steps=40, seed=17, contact_prob=False, impulse=None):
    rng = np.random.default_rng(seed)
    clips = helpers.fresh_clips()
    ref, mod = rh.make_env(variant, clips)
    qlog = _capture_qdd(mod)
    pmod = sys.modules["env.Policy_prior_posterior_MoE_window"]
    out = dict(plans=[], actions=[], obs=[], rew=[], done=[], rinfo=[], state=[], qdd=[], obs0=[], prob=[])
    worst = 0.0
    for e in range(n_envs):
        orc = OracleSRBEnv(helpers.fresh_clips(), variant)
        plan = helpers.random_plan(rng, clips, noise=(e % 3 != 0))
        if np.isnan(plan[5]):
            obs0 = rh.reset_explicit(ref, mod, int(plan[0]), int(plan[1]))
        else:
            obs0 = rh.reset_explicit(ref, mod, int(plan[0]), int(plan[1]), plan[2:5], plan[5:9])
        o0, _ = helpers.oracle_reset(orc, plan)
        worst = max(worst, float(np.abs(np.asarray(obs0, float) - o0).max()))
        if impulse is not None:
            for env_ in (ref, orc):
                env_.impulse_interval = impulse["interval"]; env_.impulse_duration = impulse["duration"]; env_.force = list(impulse["force"])
                env_.frame_counter = 0; env_.impulse = 0
        A, O, Rw, D, RI, S, Q, PR = [], [], [], [], [], [], [], []
        sigma = [0.02, 0.1, 0.3][e % 3]
        for k in range(n_steps):
            a = helpers.tracking_action(orc, rng, sigma)
            prob = [None, None]
            if contact_prob:
                prob = [float(x) for x in rng.uniform(0, 1, 2)]
            pmod._last_contact_prob = list(prob)
            orc.contact_prob = list(prob)
            del qlog[:]
            obs, r, done, _, info = ref.step(a)
            o2, r2, d2, _, _ = orc.step(a)
            worst = max(worst, float(np.abs(np.asarray(obs, float) - o2).max()), abs(float(r) - float(r2)))
            assert bool(done) == bool(d2)
            q = [np.full(6, np.nan)] * 4
            if qlog:
                assert len(qlog) == 4
                q = [x.copy() for x in qlog]
            A.append(a); O.append(np.asarray(obs, float)); Rw.append(float(r)); D.append(bool(done))
            RI.append([float(info["reward_info"][k2]) for k2 in RI_KEYS])
            S.append(helpers.oracle_state_vector(ref)); Q.append(q)
            PR.append([np.nan if p is None else p for p in prob])
        out["plans"].append(plan); out["obs0"].append(np.asarray(obs0, float))
        for key, v in zip(("actions", "obs", "rew", "done", "rinfo", "state", "qdd", "prob"), (A, O, Rw, D, RI, S, Q, PR)):
            out[key].append(np.array(v))
    pmod._last_contact_prob = [None, None]
    assert worst < 1e-9, worst
    return {k: np.array(v) for k, v in out.items()}, worst


def rollout_terrain(planars=((0.0, 0.0), (22.0, 24.0), (-42.0, 42.0), (42.0, -40.0)), n_steps=50, seed=31):
    """env/SRBEnv5_mocap_list_terrain_window.py on the playground: one environment per terrain patch (flat ground, the pyramid with
    its edge snapping, the hill and the rough height field); reset(mocap_id, initial_pos_planar) as the terrain driver calls it."""
    from oracle.srb_env import lift_clip_to_terrain
    from oracle.terrain import Terrain
    T = Terrain.load(os.path.join(GOLDEN, "terrain_playground.npz"))
    rng = np.random.default_rng(seed)
    out = dict(plans=[], planar=[], actions=[], obs=[], rew=[], done=[], state=[], obs0=[])
    worst = 0.0
    for planar in planars:
        ref, mod = rh.make_env("terrain", helpers.fresh_clips(("aiming_show",)))
        orc = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, planar)], "terrain", terrain=T)
        frame = int(rng.integers(0, 300))
        vn = rng.uniform(-.05, .05, 3); nq = np.array([1.0, 0.0, 0.0, 0.0])
        obs0 = rh.reset_explicit(ref, mod, 0, frame, vn, nq, planar=planar)
        o0, _ = orc.reset_explicit(0, frame, vn, nq)
        worst = max(worst, float(np.abs(np.asarray(obs0, float) - o0).max()))
        A, O, Rw, D, S = [], [], [], [], []
        for k in range(n_steps):
            a = helpers.tracking_action(orc, rng, 0.1)
            obs, r, done, _, info = ref.step(a)
            o2, r2, d2, _, _ = orc.step(a)
            worst = max(worst, float(np.abs(np.asarray(obs, float) - o2).max()), abs(float(r) - float(r2)))
            assert bool(done) == bool(d2)
            A.append(a); O.append(np.asarray(obs, float)); Rw.append(float(r)); D.append(bool(done)); S.append(helpers.oracle_state_vector(ref))
        out["plans"].append(np.concatenate([[0, frame], vn, nq])); out["planar"].append(planar); out["obs0"].append(np.asarray(obs0, float))
        for key, v in zip(("actions", "obs", "rew", "done", "state"), (A, O, Rw, D, S)):
            out[key].append(np.array(v))
    assert worst < 1e-9, worst
    return {k: np.array(v) for k, v in out.items()}, worst


def sds_plan(seed=5, n=80):
    rng = np.random.default_rng(seed)
    plan = [(10 ** rng.uniform(5, 10.99), rng.uniform(-1, 1), bool(rng.random() < 0.1)) for _ in range(n)]
    plan[3] = (1e9, 0.3, False); plan[4] = (1e5, 0.2, False); plan[5] = (1e11, 0.1, False); plan[6] = (10 ** 7.003, 0.1, False); plan[7] = (10 ** 9.996, -0.4, True)
    return plan


DELTA_SRB = dict(mass=60.0, inertia=np.array([[3.0, 0.1, 0.0], [0.1, 1.5, -0.2], [0.0, -0.2, 2.5]]), local_com=np.array([0.01, -0.03, 0.02]))


def playground_ground():
    from oracle.terrain import Terrain
    T = Terrain.load(os.path.join(GOLDEN, "terrain_playground.npz"))
    return lambda pos: T.height_ray(pos[2], pos[0])[0]              # Y-up (a, b, c) <-> Z-up (c, a, b)


def playground_window(j):
    from test_foot_sliding import make_window
    w = make_window(100 + j)
    shift = np.array([4.0 + 0.37 * j, 0.0, 20.0])                  # on the slope of pyramid_1 (as tests/test_foot_sliding.py)
    for key in ("marker_traj", "foot_center"):
        w[key] = [t + shift for t in w[key]]
    w["marker"] = w["marker"] + np.tile(shift, 4); w["initial"] = w["initial"] + np.tile(shift, 4)
    return w


def make_delta(n_items=24, n_steps=14):
    import ref_delta_harness as dh
    from oracle import fullbody_map as fm
    from ipcdnnwalk_b200 import vrml_skeleton
    from test_oracle_mmsto_cpu import _com_slide_case
    assert dh.available(), "needs /root/reference and oracle/_ref (make -C oracle/ref_build)"
    mod = dh.load()
    out = {}
    # ---- pack_obs_extra on the reference environment (test variant), two environments
    rng = np.random.default_rng(41)
    clips = helpers.fresh_clips()
    ref, emod = rh.make_env("test", clips)
    P, A, E = [], [], []
    for e in range(2):
        orc = OracleSRBEnv(helpers.fresh_clips(), "test")
        plan = helpers.random_plan(rng, clips, noise=True)
        rh.reset_explicit(ref, emod, int(plan[0]), int(plan[1]), plan[2:5], plan[5:9]); helpers.oracle_reset(orc, plan)
        acts, ex = [], []
        for k in range(n_steps):
            a = helpers.tracking_action(orc, rng, 0.1)
            ref.step(a); orc.step(a)
            acts.append(a); ex.append(mod.pack_obs_extra(ref).a.copy())
        P.append(plan); A.append(np.array(acts)); E.append(np.array(ex))
    out.update(pk_plans=np.array(P), pk_actions=np.array(A), pk_extra=np.array(E))
    # ---- unpack_obs_extra + _extractSingleFrameFullbody
    rng = np.random.default_rng(43)
    extra = rng.normal(0, 0.5, (n_items, fm.EXTRA_W)); row = rng.normal(0, 0.3, (n_items, fm.ROW_W))
    extra[:2] = E[0][3:5]                                    # two records the environment produced
    for c in (3, 16, 23):
        extra[:, c:c + 4] /= np.linalg.norm(extra[:, c:c + 4], axis=1, keepdims=True)
    row[1, 0:3] = 0.0                                        # small-angle branch of Exp
    row[2, 0:3] = [0.0, 3.1, 0.2]; row[3, 0:3] = [3.0, 0.1, 0.0]; row[4, 0:3] = [0.1, 0.0, 3.05]     # near half turns: the non-trace branches of matrix -> quaternion
    res = [dh.extract(mod, extra[i], row[i], DELTA_SRB["mass"], DELTA_SRB["inertia"], DELTA_SRB["local_com"]) for i in range(n_items)]
    out.update(ex_extra=extra, ex_row=row, ex_mass=np.array(DELTA_SRB["mass"]), ex_inertia=DELTA_SRB["inertia"], ex_local_com=DELTA_SRB["local_com"])
    for k in res[0]:
        out["ex_" + k] = np.stack([r[k] for r in res])
    # ---- removeCOMsliding_online
    skel = vrml_skeleton.builtin()
    for j, (seed, nvar, fy, ground) in enumerate(((3, 8, True, False), (4, 8, False, True), (5, 6, True, True), (6, 8, True, False))):
        pose, com_srb, des = _com_slide_case(skel, seed, nvar)
        gh = np.random.default_rng(100 + seed).uniform(0, 0.4, nvar) if ground else None
        new = dh.remove_com_sliding(mod, com_srb, des, pose, fix_y_also=fy, ground_h=gh)
        out.update({f"cs{j}_pose": pose, f"cs{j}_com_srb": com_srb, f"cs{j}_desired": des, f"cs{j}_fix_y": np.array(int(fy)),
                    f"cs{j}_ground_h": gh if gh is not None else np.zeros(0), f"cs{j}_out": new})
    out["cs_count"] = np.array(4)
    # ---- removeFootSliding_online / _stair: every contact pattern of tests/test_foot_sliding.make_window, flat (y = 0, what
    #      walk_SRBTrack2025.py:53 passes) and on a wavy ground, with the caller's desiredVel and with derivative_forward
    from test_foot_sliding import make_window, FOOT_LEN
    wavy = lambda p: 0.02 * np.sin(3 * p[0]) + 0.03 * np.cos(2 * p[2])          # noqa: E731
    k = 0
    for seed in range(12):
        w = make_window(seed)
        for stair in (False, True):
            kind = seed % 2                                   # 0: flat ground, 1: wavy
            use_dv = seed % 3 != 0
            g = wavy if kind else (lambda p: 0.0)
            r = dh.remove_foot_sliding(mod, FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"] if stair else None, g,
                                       w["desired_vel"] if use_dv else None)
            out.update({f"fs{k}_seed": np.array(seed), f"fs{k}_stair": np.array(int(stair)), f"fs{k}_wavy": np.array(kind), f"fs{k}_use_dv": np.array(int(use_dv)), f"fs{k}_out": r})
            k += 1
    out["fs_count"] = np.array(k)
    # ---- the _stair variant (and the plain one) on the SRBTrack playground: the slope of pyramid_1, ground = terrain_height_ray
    ground_pg = playground_ground()
    for j in range(8):
        w = playground_window(j)
        for stair in (True, False):
            r = dh.remove_foot_sliding(mod, FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"] if stair else None, ground_pg, w["desired_vel"])
            out[f"pg{j}_{'stair' if stair else 'plain'}_out"] = r
    out["pg_count"] = np.array(8)
    np.savez_compressed(os.path.join(GOLDEN, "refenv_delta.npz"), **out)
    print("DeltaExtractor of the reference:", {k: np.asarray(v).shape for k, v in out.items() if k.startswith(("pk_", "ex_h", "ex_m", "cs0_out"))})


def make_sds():
    import ref_sds_harness as sh
    plan = sds_plan()
    xs, Ks = sh.drive(sh.load().SDS, plan)
    np.savez_compressed(os.path.join(GOLDEN, "refenv_sds.npz"), plan=np.array([[q, t, float(z)] for q, t, z in plan]), x=xs, K=Ks)
    print("SDS of the reference: ", xs.shape, "states,", Ks.shape, "gains")


if __name__ == "__main__":
    assert rh.available(), "needs /root/reference"
    make_sds()
    make_delta()
    g, worst = rollout_terrain()
    np.savez_compressed(os.path.join(GOLDEN, "refenv_rollout_terrain.npz"), **g)
    print("terrain reference-code roll-out", {k: v.shape for k, v in g.items()}, "| max |reference - oracle|", worst)
    for variant in ("test", "training"):
        g, worst = rollout(variant)
        g.pop("prob")
        np.savez_compressed(os.path.join(GOLDEN, f"refenv_rollout_{variant}.npz"), **g)
        print(variant, "reference-code roll-out", {k: v.shape for k, v in g.items()}, "episodes done", int(g["done"].any(axis=1).sum()),
              "| max |reference - oracle|", worst)
    sp = {}
    g, worst = rollout("test", n_envs=4, n_steps=30, seed=23, contact_prob=True)
    sp.update({"hyst_" + k: v for k, v in g.items()})
    print("hysteresis roll-out: contact flag changes", int((np.diff(g["state"][:, :, 23:25], axis=1) != 0).sum()), "| max |reference - oracle|", worst)
    imp = dict(interval=8, duration=5, force=[300.0, -200.0, 100.0, 20.0, -30.0, 10.0])
    g, worst = rollout("test", n_envs=3, n_steps=30, seed=29, impulse=imp)
    sp.update({"imp_" + k: v for k, v in g.items()})
    sp["imp_interval"] = np.array(imp["interval"]); sp["imp_duration"] = np.array(imp["duration"]); sp["imp_force"] = np.array(imp["force"])
    print("impulse roll-out | max |reference - oracle|", worst)
    np.savez_compressed(os.path.join(GOLDEN, "refenv_special.npz"), **sp)
