"""Pins of the SRBTrack path against the REFERENCE'S OWN ENVIRONMENT CODE.

tests/golden/refenv_rollout_{test,training}.npz and refenv_special.npz are roll-outs of env/SRBEnv5_mocap_list_window.py /
..._training.py themselves, executed from /root/reference through tests/tools/ref_srbenv_harness.py -- MuJoCo, Clarabel and SRBdata
answered by stated stand-ins (restated MuJoCo free-body semantics, an exact QP solve, the committed clip tables), controlmodule.SDS being the
reference's own class over NumPy-backed matrixn / vectorn objects (tests/tools/ref_sds_harness.py), every other line the reference's code as it stands: action decoding, contact logic incl. the probability hysteresis, site
placement, the assembly of the QP in solve_qp_for_contact_forces, force clamp, impulse schedule, history, get_obs_window, _get_reward,
_get_done, reset / set_state.  Generator: tests/tools/make_golden_refenv.py.

  * CPU: oracle/srb_env.py replays the same plans / actions and reproduces the reference code's outputs at 1e-10.
  * `live` (where /root/reference is mounted): the harness re-runs the reference code and must print the committed fixture.
  * GPU: the CUDA step, through the C-ABI, is compared with the reference code's roll-outs DIRECTLY (fp64 1e-5; done, contact flags and
    frame counters bit exact), without the oracle in between."""
import os
import sys

import numpy as np
import pytest

import helpers

sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "tools"))
import ref_srbenv_harness as rh            # noqa: E402

RTOL64 = 1e-5
LIVE_RTOL = 1e-10          # live re-runs reproduce the fixtures bit for bit on the box that recorded them; the QP solves go through BLAS / LAPACK, whose kernels depend on the CPU
live = pytest.mark.skipif(not rh.available(), reason="/root/reference is not mounted: the committed fixture stands in")


def _fixture(variant):
    return dict(np.load(os.path.join(helpers.GOLDEN, f"refenv_rollout_{variant}.npz")))


def _special():
    return dict(np.load(os.path.join(helpers.GOLDEN, "refenv_special.npz")))


def _close(a, b, rtol, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert np.nanmax(err) <= rtol, f"{what}: max rel err {np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}"


def _replay_oracle(g, variant, prob=None, impulse=None):
    from oracle.srb_env import OracleSRBEnv
    n_env, n_steps = g["actions"].shape[:2]
    for e in range(n_env):
        o = OracleSRBEnv(helpers.fresh_clips(), variant)
        obs0, _ = helpers.oracle_reset(o, g["plans"][e])
        _close(obs0, g["obs0"][e], 1e-10, f"env {e} reset obs")
        if impulse is not None:
            o.impulse_interval, o.impulse_duration, o.force = impulse
            o.frame_counter = 0; o.impulse = 0
        for k in range(n_steps):
            if prob is not None:
                p = prob[e, k]
                o.contact_prob = [None, None] if np.isnan(p[0]) else [float(p[0]), float(p[1])]
            obs, r, done, _, info = o.step(g["actions"][e, k])
            what = f"env {e} step {k}"
            assert bool(done) == bool(g["done"][e, k]), what + " done"
            st = helpers.oracle_state_vector(o)
            assert (st[23:26] == g["state"][e, k, 23:26]).all(), what + " contact flags / frame"
            _close(st[:23], g["state"][e, k, :23], 1e-10, what + " state")
            _close(obs, g["obs"][e, k], 1e-10, what + " obs")
            _close(r, g["rew"][e, k], 1e-10, what + " reward")
            ri = [info["reward_info"][k2] for k2 in ("reward_survive", "reward_contact", "reward_end_effector", "reward_delta_pos",
                                                      "reward_delta_ori", "reward_pos", "reward_ori", "reward_total")]
            _close(ri, g["rinfo"][e, k], 1e-10, what + " reward_info")
            q = np.array([np.full(6, np.nan) if x is None else x for x in o.qdd_qp_list])
            assert (np.isnan(q[:, 0]) == np.isnan(g["qdd"][e, k][:, 0])).all(), what + " which sub-steps ran the QP"
            m = ~np.isnan(q[:, 0])
            if m.any():
                _close(q[m], g["qdd"][e, k][m], 1e-9, what + " QP accelerations")


@pytest.mark.parametrize("variant", ["test", "training"])
def test_oracle_equals_the_reference_environment_code(variant):
    g = _fixture(variant)
    assert g["done"].any() and not g["done"].all()
    _replay_oracle(g, variant)


def test_oracle_equals_the_reference_code_on_hysteresis_and_impulse():
    s = _special()
    h = {k[5:]: v for k, v in s.items() if k.startswith("hyst_")}
    assert (np.diff(h["state"][:, :, 23:25], axis=1) != 0).sum() > 20          # the flags really switch
    _replay_oracle(h, "test", prob=h["prob"])
    i = {k[4:]: v for k, v in s.items() if k.startswith("imp_") and k not in ("imp_interval", "imp_duration", "imp_force")}
    _replay_oracle(i, "test", impulse=(int(s["imp_interval"]), int(s["imp_duration"]), list(s["imp_force"])))


def test_oracle_equals_the_reference_terrain_environment_code():
    """env/SRBEnv5_mocap_list_terrain_window.py on the playground (flat ground, pyramid with edge snapping, hill, rough height field);
    mj_ray is answered by oracle/terrain.py, so this pins the environment logic (lifting, snapping, Q9 / Q10, terrain observations), not
    the height-field semantics of MuJoCo."""
    from oracle.srb_env import OracleSRBEnv, lift_clip_to_terrain
    from oracle.terrain import Terrain
    T = Terrain.load(os.path.join(helpers.GOLDEN, "terrain_playground.npz"))
    g = dict(np.load(os.path.join(helpers.GOLDEN, "refenv_rollout_terrain.npz")))
    for e in range(g["actions"].shape[0]):
        planar = tuple(g["planar"][e])
        o = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, planar)], "terrain", terrain=T)
        plan = g["plans"][e]
        obs0, _ = o.reset_explicit(0, int(plan[1]), plan[2:5].copy(), plan[5:9].copy())
        _close(obs0, g["obs0"][e], 1e-10, f"terrain env {e} reset obs")
        for k in range(g["actions"].shape[1]):
            obs, r, done, _, _ = o.step(g["actions"][e, k])
            st = helpers.oracle_state_vector(o)
            assert bool(done) == bool(g["done"][e, k]) and (st[23:26] == g["state"][e, k, 23:26]).all(), (e, k)
            _close(st[:23], g["state"][e, k, :23], 1e-10, f"terrain env {e} step {k} state")
            _close(obs, g["obs"][e, k], 1e-10, f"terrain env {e} step {k} obs")


def test_oracle_sds_equals_the_reference_filter_class():
    """oracle/sds_filter.py against work/controlmodule.py's SDS / NonlinearController run from /root/reference (tests/tools/ref_sds_harness.py):
    preset gains exact over logQ 5 ... 11, filter state after every singleStep at round-off."""
    import ref_sds_harness as sh
    from oracle.sds_filter import SDS
    g = dict(np.load(os.path.join(helpers.GOLDEN, "refenv_sds.npz")))
    plan = [(q, t, bool(z)) for q, t, z in g["plan"]]
    xs, Ks = sh.drive(SDS, plan)
    _close(Ks, g["K"], 1e-15, "SDS gains")
    assert np.abs(xs - g["x"]).max() < 1e-13 * max(1.0, np.abs(g["x"]).max())


@live
def test_live_reference_sds_class_reproduces_the_fixture():
    import ref_sds_harness as sh
    g = dict(np.load(os.path.join(helpers.GOLDEN, "refenv_sds.npz")))
    xs, Ks = sh.drive(sh.load().SDS, [(q, t, bool(z)) for q, t, z in g["plan"]])
    _close(Ks, g["K"], 1e-15, "SDS gains"); _close(xs, g["x"], 1e-13, "SDS states")     # identical on the box that recorded them; a BLAS with other kernels may differ in the last bit


@live
@pytest.mark.parametrize("name", ["aiming_show", "stand1"])
def test_live_reference_srbdata_code_reproduces_the_clip_tables(name):
    """Row a-13: the reference's own SRBdata(mocap_path).humanoid2SRB() (testSRB_mujoco_original_viewer.py:765-1031: frame stitching, the
    yaw-aligned continuation of the cycle, CoM / velocity / foot-orientation / contact processing), run on the shipped clip with MuJoCo's
    humanoid FK, subtree CoM, mj_differentiatePos and transformations.py answered by oracle/srbdata.py's restatements, yields the committed
    clip tables tests/golden/clip_<name>.npz -- the tables the product's ClipBuilder is tested against -- exactly."""
    s = rh.reference_srbdata(f"motiondata_mujoco_refined/{name}.txt")
    c = helpers.load_clip(name)
    assert s.cycle_length == c["cycle_length"] and abs(s.dt - c["mocap_dt"]) < 1e-15
    for a, b in ((s.mocap_com_humanoid, c["mocap_qpos"]), (np.array(s.mocap_ori_com), c["mocap_ori_com"]), (s.mocap_vel_com, c["mocap_vel_com"]),
                 (s.feet_pos_list, c["feet_pos_list"]), (np.array(s.feet_ori_list), c["feet_ori_list"])):
        assert np.abs(np.asarray(a, dtype=float) - b).max() < 1e-12


@live
def test_live_reference_terrain_code_reproduces_the_fixture():
    g = dict(np.load(os.path.join(helpers.GOLDEN, "refenv_rollout_terrain.npz")))
    e = 1                                                        # the pyramid
    ref, mod = rh.make_env("terrain", helpers.fresh_clips(("aiming_show",)))
    plan = g["plans"][e]
    obs0 = rh.reset_explicit(ref, mod, 0, int(plan[1]), plan[2:5], plan[5:9], planar=tuple(g["planar"][e]))
    _close(obs0, g["obs0"][e], LIVE_RTOL, "terrain reset obs")
    for k in range(15):
        obs, r, done, _, _ = ref.step(g["actions"][e, k])
        _close(obs, g["obs"][e, k], LIVE_RTOL, f"terrain obs step {k}")
        assert bool(done) == bool(g["done"][e, k])


@live
@pytest.mark.parametrize("variant", ["test", "training"])
def test_live_reference_code_reproduces_the_fixture(variant):
    """Re-runs the reference's environment code here on the first environments of the fixture: plans, actions in -> identical outputs."""
    g = _fixture(variant)
    clips = helpers.fresh_clips()
    ref, mod = rh.make_env(variant, clips)
    for e in (0, 1):
        plan = g["plans"][e]
        if np.isnan(plan[5]):
            obs0 = rh.reset_explicit(ref, mod, int(plan[0]), int(plan[1]))
        else:
            obs0 = rh.reset_explicit(ref, mod, int(plan[0]), int(plan[1]), plan[2:5], plan[5:9])
        _close(obs0, g["obs0"][e], LIVE_RTOL, "reset obs")
        for k in range(12):
            obs, r, done, _, info = ref.step(g["actions"][e, k])
            _close(obs, g["obs"][e, k], LIVE_RTOL, f"obs env {e} step {k}"); _close(r, g["rew"][e, k], LIVE_RTOL, "reward")
            assert bool(done) == bool(g["done"][e, k])


# --------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["test", "training"])
@pytest.mark.parametrize("lanes", [8, 4])
def test_cuda_equals_the_reference_environment_code(built_lib, variant, lanes):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    g = _fixture(variant)
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
        assert (done.cpu().numpy().astype(bool) == g["done"][:, k]).all(), f"done mismatch at step {k}"
        assert (st[:, 23:26] == g["state"][:, k, 23:26]).all(), f"contact flags / frame mismatch at step {k}"
        _close(st[:, :23], g["state"][:, k, :23], RTOL64, f"state step {k}")
        _close(obs.cpu().numpy(), g["obs"][:, k], RTOL64, f"obs step {k}")
        _close(rew.cpu().numpy(), g["rew"][:, k], RTOL64, f"reward step {k}")
        _close(rinfo.cpu().numpy(), g["rinfo"][:, k], RTOL64, f"reward_info step {k}")
        dbg = env.dbg.cpu().numpy().reshape(n_env, 4, 64)
        m = ~np.isnan(g["qdd"][:, k][..., 0])
        _close(dbg[..., :6][m], g["qdd"][:, k][m], RTOL64, f"QP qdd step {k}")
    env.close()


@pytest.mark.gpu
def test_cuda_equals_the_reference_code_on_hysteresis_and_impulse(built_lib):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    s = _special()
    # contact-probability hysteresis (SRBEnv5_mocap_list_window.py:239-275)
    h = {k[5:]: v for k, v in s.items() if k.startswith("hyst_")}
    n_env, n_steps = h["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(), n_env, variant="test", precision=64, auto_reset=False)
    env.attach(contact_prob=True)
    env.reset(plan=h["plans"])
    for k in range(n_steps):
        env.contact_prob.copy_(torch.as_tensor(h["prob"][:, k].astype(np.float32)))
        obs, rew, done, _ = env.step(torch.as_tensor(h["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(i)) for i in range(n_env)])
        assert (st[:, 23:26] == h["state"][:, k, 23:26]).all(), f"hysteresis: contact flags at step {k}"
        assert (done.cpu().numpy().astype(bool) == h["done"][:, k]).all()
        _close(obs.cpu().numpy(), h["obs"][:, k], RTOL64, f"hysteresis obs step {k}")
        _close(rew.cpu().numpy(), h["rew"][:, k], RTOL64, f"hysteresis reward step {k}")
    env.close()
    # push impulse with a 6-component force (:191-206)
    i = {k[4:]: v for k, v in s.items() if k.startswith("imp_") and k not in ("imp_interval", "imp_duration", "imp_force")}
    n_env, n_steps = i["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(), n_env, variant="test", precision=64, auto_reset=False)
    env.set_impulse(int(s["imp_interval"]), int(s["imp_duration"]), s["imp_force"])
    env.reset(plan=i["plans"])
    for k in range(n_steps):
        obs, rew, done, _ = env.step(torch.as_tensor(i["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(j)) for j in range(n_env)])
        assert (done.cpu().numpy().astype(bool) == i["done"][:, k]).all()
        _close(st[:, :23], i["state"][:, k, :23], RTOL64, f"impulse state step {k}")
        _close(obs.cpu().numpy(), i["obs"][:, k], RTOL64, f"impulse obs step {k}")
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [8, 4])
def test_cuda_equals_the_reference_terrain_environment_code(built_lib, lanes):
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    g = dict(np.load(os.path.join(helpers.GOLDEN, "refenv_rollout_terrain.npz")))
    n_env, n_steps = g["actions"].shape[:2]
    env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), n_env, variant="terrain", precision=64, lanes_per_env=lanes, auto_reset=False,
                    terrain=TerrainModel.from_npz(os.path.join(helpers.GOLDEN, "terrain_playground.npz")), initial_pos_planar=g["planar"])
    obs0 = env.reset(plan=g["plans"]).cpu().numpy()
    _close(obs0, g["obs0"], 2e-6, "terrain reset obs")
    for k in range(n_steps):
        obs, rew, done, _ = env.step(torch.as_tensor(g["actions"][:, k], device="cuda"))
        torch.cuda.synchronize()
        st = np.stack([helpers.cuda_state_vector(env.named_state(i)) for i in range(n_env)])
        assert (done.cpu().numpy().astype(bool) == g["done"][:, k]).all(), f"terrain done mismatch at step {k}"
        assert (st[:, 23:26] == g["state"][:, k, 23:26]).all(), f"terrain contact flags / frame mismatch at step {k}"
        _close(st[:, :23], g["state"][:, k, :23], RTOL64, f"terrain state step {k}")
        _close(obs.cpu().numpy(), g["obs"][:, k], RTOL64, f"terrain obs step {k}")
    env.close()


# ---------------------------------------------------------------------------------------------------------- SRB -> full-body mapping (f-2a, f-2b, f-2c)
def _delta():
    return dict(np.load(os.path.join(helpers.GOLDEN, "refenv_delta.npz")))


EX_KEYS = ("humanPose", "dx_ddx", "feetpos", "com", "momentum", "con_bin", "con")


def test_oracle_equals_the_reference_fullbody_mapping_code():
    """oracle/fullbody_map.py and oracle/mmsto.remove_com_sliding_online against gym_trackSRB/taesooLib/DeltaExtractor.py run from
    /root/reference (tests/tools/ref_delta_harness.py): pack_obs_extra on the stepping reference environment, unpack_obs_extra +
    _extractSingleFrameFullbody, removeCOMsliding_online (flat and with projectToGround)."""
    from oracle import fullbody_map as fm, mmsto as mm
    from oracle.srb_env import OracleSRBEnv
    from ipcdnnwalk_b200 import vrml_skeleton
    g = _delta()
    for e in range(g["pk_plans"].shape[0]):
        orc = OracleSRBEnv(helpers.fresh_clips(), "test")
        helpers.oracle_reset(orc, g["pk_plans"][e])
        for k in range(g["pk_actions"].shape[1]):
            orc.step(g["pk_actions"][e, k])
            _close(fm.pack_obs_extra(orc), g["pk_extra"][e, k], 1e-12, f"pack_obs_extra env {e} step {k}")
    for i in range(g["ex_extra"].shape[0]):
        o = fm.extract_single_frame_fullbody(g["ex_extra"][i], g["ex_row"][i], float(g["ex_mass"]), g["ex_inertia"], g["ex_local_com"])
        for k in EX_KEYS:
            _close(o[k], g["ex_" + k][i], 1e-12, f"extract item {i} {k}")
    skel = vrml_skeleton.builtin()
    for j in range(int(g["cs_count"])):
        gh = g[f"cs{j}_ground_h"]
        out, _, _ = mm.remove_com_sliding_online(skel, g[f"cs{j}_com_srb"], g[f"cs{j}_desired"], g[f"cs{j}_pose"], fix_y_also=bool(g[f"cs{j}_fix_y"]),
                                                 ground_h=gh if gh.size else None)
        _close(out, g[f"cs{j}_out"], 1e-12, f"removeCOMsliding_online case {j}")
        assert np.abs(out - g[f"cs{j}_pose"]).max() > 1e-3


def _delta_available():
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests", "tools"))
    import ref_delta_harness as dh
    return dh.available()


@pytest.mark.skipif(not _delta_available(), reason="/root/reference or oracle/_ref is missing: the committed fixture stands in")
def test_live_reference_fullbody_mapping_code_reproduces_the_fixture():
    import ref_delta_harness as dh
    g = _delta()
    mod = dh.load()
    for i in (0, 1, 4, 9):
        r = dh.extract(mod, g["ex_extra"][i], g["ex_row"][i], float(g["ex_mass"]), g["ex_inertia"], g["ex_local_com"])
        for k in EX_KEYS:
            _close(r[k], g["ex_" + k][i], 1e-13, f"extract {i} {k}")
    for j in (1, 2):
        gh = g[f"cs{j}_ground_h"]
        new = dh.remove_com_sliding(mod, g[f"cs{j}_com_srb"], g[f"cs{j}_desired"], g[f"cs{j}_pose"], fix_y_also=bool(g[f"cs{j}_fix_y"]), ground_h=gh if gh.size else None)
        _close(new, g[f"cs{j}_out"], 1e-13, f"removeCOMsliding_online case {j}")


@pytest.mark.gpu
def test_cuda_equals_the_reference_fullbody_mapping_code(built_lib):
    """The CUDA side of rows f-2a / f-2b / f-2c against the reference code's outputs directly: srb_pack_obs_extra after the same steps,
    fullbody_extract on the same rows, mmsto_remove_com_sliding on the same windows."""
    import torch
    import helpers_mmsto as hm
    from ipcdnnwalk_b200 import SRBVecEnv, ShortTrajOpt, vrml_skeleton
    from ipcdnnwalk_b200.short_traj_opt import extractSingleFrameFullbody
    g = _delta()
    n = g["pk_plans"].shape[0]
    env = SRBVecEnv(helpers.fresh_clips(), n, variant="test", precision=64, auto_reset=False)
    env.reset(plan=g["pk_plans"])
    for k in range(g["pk_actions"].shape[1]):
        env.step(torch.as_tensor(g["pk_actions"][:, k], device="cuda"))
        ex = env.pack_obs_extra().cpu().numpy()
        _close(ex, g["pk_extra"][:, k], RTOL64, f"pack_obs_extra step {k}")
        assert np.array_equal(ex[:, 27:29], g["pk_extra"][:, k, 27:29])
    env.close()
    out = extractSingleFrameFullbody(g["ex_extra"], g["ex_row"], float(g["ex_mass"]), g["ex_inertia"], g["ex_local_com"])
    for k in EX_KEYS:
        _close(out[k].cpu().numpy(), g["ex_" + k], 1e-11, f"fullbody_extract {k}")
    opt = ShortTrajOpt(vrml_skeleton.builtin(), hm.NF, hm.MARKERS)
    for j in range(int(g["cs_count"])):
        gh = g[f"cs{j}_ground_h"]
        pose = torch.as_tensor(g[f"cs{j}_pose"][None], device="cuda").contiguous()
        opt.removeCOMsliding_online(g[f"cs{j}_com_srb"][None], g[f"cs{j}_desired"][None], pose, fix_y_also=bool(g[f"cs{j}_fix_y"]), ground_h=gh[None] if gh.size else None)
        _close(pose.cpu().numpy()[0], g[f"cs{j}_out"], 1e-10, f"mmsto_remove_com_sliding case {j}")
    opt.close()


def _foot_cases(g):
    from test_foot_sliding import make_window
    wavy = lambda p: 0.02 * np.sin(3 * p[0]) + 0.03 * np.cos(2 * p[2])          # noqa: E731
    for k in range(int(g["fs_count"])):
        w = make_window(int(g[f"fs{k}_seed"]))
        yield k, w, bool(g[f"fs{k}_stair"]), (wavy if int(g[f"fs{k}_wavy"]) else (lambda p: 0.0)), bool(g[f"fs{k}_use_dv"]), g[f"fs{k}_out"]


def test_oracle_equals_the_reference_foot_sliding_code():
    """oracle/foot_sliding.py against removeFootSliding_online / removeFootSliding_online_stair / avoidHull of DeltaExtractor.py:650-897 run
    from /root/reference: 24 windows over all contact patterns, flat and wavy ground, with and without the caller's desiredVel.  The Lua
    helpers under it are stand-ins on both sides (W5 / W6, hull projection over the pinned GrahamScan); the programmes, their constraints,
    the ground clamp and the hull passes are the reference's."""
    from oracle import foot_sliding as fs
    from test_foot_sliding import FOOT_LEN
    for k, w, stair, ground, use_dv, want in _foot_cases(_delta()):
        dv = w["desired_vel"] if use_dv else None
        got = (fs.remove_foot_sliding_online_stair(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"], dv, ground) if stair
               else fs.remove_foot_sliding_online(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], dv, ground))
        _close(got, want, 1e-12, f"foot sliding case {k}")


def test_oracle_equals_the_reference_foot_sliding_code_on_the_playground():
    """Both variants on the slope of the playground's pyramid (ground = terrain_height_ray): the hull passes of avoidHull act here."""
    from oracle import foot_sliding as fs
    from test_foot_sliding import FOOT_LEN
    from make_golden_refenv import playground_ground, playground_window
    g = _delta()
    ground = playground_ground()
    lifted = 0
    for j in range(int(g["pg_count"])):
        w = playground_window(j)
        st = fs.remove_foot_sliding_online_stair(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"], w["desired_vel"], ground)
        _close(st, g[f"pg{j}_stair_out"], 1e-12, f"stair case {j}")
        base = fs._marker_qps(FOOT_LEN, np.asarray(w["con"]), np.asarray(w["initial"]), w["marker_traj"], w["marker"], w["desired_vel"], ground, True, w["foot_center"])
        lifted += int((g[f"pg{j}_stair_out"][:, 1::3] > base[:, 1::3] + 1e-9).any())
        pl = fs.remove_foot_sliding_online(FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["desired_vel"], ground)
        _close(pl, g[f"pg{j}_plain_out"], 1e-12, f"plain case {j}")
    assert lifted >= 1, "the hull pass should act on some window"


@pytest.mark.skipif(not _delta_available(), reason="/root/reference or oracle/_ref is missing: the committed fixture stands in")
def test_live_reference_foot_sliding_code_reproduces_the_fixture():
    import ref_delta_harness as dh
    from test_foot_sliding import FOOT_LEN
    mod = dh.load()
    for k, w, stair, ground, use_dv, want in _foot_cases(_delta()):
        if k % 5 == 0:
            r = dh.remove_foot_sliding(mod, FOOT_LEN, w["con"], w["initial"], w["marker_traj"], w["marker"], w["foot_center"] if stair else None, ground,
                                       w["desired_vel"] if use_dv else None)
            _close(r, want, 1e-13, f"foot sliding case {k}")


@pytest.mark.gpu
def test_cuda_equals_the_reference_foot_sliding_code(built_lib):
    """mmsto_remove_foot_sliding against the reference code's outputs directly, flat-ground cases (the batched API takes a constant height or
    an SRBTrack terrain; the wavy ground of the CPU cases is neither)."""
    import torch
    from ipcdnnwalk_b200.short_traj_opt import removeFootSliding_online
    from test_foot_sliding import FOOT_LEN
    n = 0
    for k, w, stair, ground, use_dv, want in _foot_cases(_delta()):
        if stair or ground(np.zeros(3)) != 0.0 or ground(np.ones(3)) != 0.0:
            continue
        mt = np.concatenate(w["marker_traj"], axis=1)[None]
        out = removeFootSliding_online(FOOT_LEN, w["con"][None], w["initial"][None], mt, w["marker"][None], w["desired_vel"][None] if use_dv else None, projectToGround=0.0)
        torch.cuda.synchronize()
        _close(out.cpu().numpy()[0], want, 1e-10, f"mmsto_remove_foot_sliding case {k}")
        n += 1
    assert n >= 5


@pytest.mark.gpu
def test_cuda_equals_the_reference_foot_sliding_code_on_the_playground(built_lib):
    """Both variants with projectToGround = the SRBTrack playground terrain on the device, against the reference code's outputs directly."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from ipcdnnwalk_b200.short_traj_opt import removeFootSliding_online
    from test_foot_sliding import FOOT_LEN
    from make_golden_refenv import playground_window
    g = _delta()
    env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), 1, variant="terrain", precision=64, terrain=TerrainModel.from_npz(os.path.join(helpers.GOLDEN, "terrain_playground.npz")))
    ws = [playground_window(j) for j in range(int(g["pg_count"]))]
    stack = lambda k: np.stack([w[k] for w in ws])                  # noqa: E731
    mt = np.stack([np.concatenate(w["marker_traj"], axis=1) for w in ws])
    fc = np.stack([np.stack(w["foot_center"]) for w in ws])
    for stair in (True, False):
        out = removeFootSliding_online(FOOT_LEN, stack("con"), stack("initial"), mt, stack("marker"), stack("desired_vel"), projectToGround=env, footCenterTraj=fc if stair else None)
        torch.cuda.synchronize()
        out = out.cpu().numpy()
        for j in range(len(ws)):
            _close(out[j], g[f"pg{j}_{'stair' if stair else 'plain'}_out"], 1e-10, f"playground {'stair' if stair else 'plain'} case {j}")
    env.close()
