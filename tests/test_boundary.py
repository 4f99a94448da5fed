"""The drop-in boundary: the reference-shaped `SRBEnv(mocap_path_list)`, the SB3 VecEnv seam and the product-side
loaders, exercised the way the reference's drivers use them.

  walk_SRBTrack2025_terrain_singlethreaded.py:97-120   push: env.impulse / env.force (6-vector) / impulse_lpos / gizmo
  :137-138  env = SRBEnv(mocap_path_list); env.impulse_interval = 1e9        :160-163  env.right_site_id, env.data.site_xpos[j + k]
  :190-201  env.mocap_pos_com / mocap_ori_com / feet_pos_list / feet_ori_list [env.current_frame + 1]
  :250-252  env.testing_mode = True; env.model_path           :386 env.reset(mocap_id=..., initial_pos_planar=...)
  :421-426  env.step / getFullState / pack_obs_extra          :717-724 restoreFullState / _get_obs      :746-765 contact_forces_list
  walk_SRBTrack2025_train_delta.py:94-100   make_vec_env(SRBEnv, n_envs=32, vec_env_cls=SubprocVecEnv)"""
import os

import numpy as np
import pytest

import helpers

DATA_ROOT = helpers.GOLDEN                      # holds gym_trackSRB/ref_contact/<clip>.motion.npz + .constraint.npz (raw clips)
CLIPS = ["ref_contact/aiming_show.motion.npz", "ref_contact/stand1.motion.npz"]
REF_XML = "/root/reference/gym_trackSRB/Characters/SRB5_playground.xml"


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_contact_file_mapping_follows_the_reference():
    from ipcdnnwalk_b200.srb_env import contact_path_for
    # env/testSRB_mujoco_original_viewer.py:769-798
    assert contact_path_for("motiondata_mujoco_refined/stand1.txt") == "ref_contact/stand1.bvh.constraint.npz"
    assert contact_path_for("motiondata_mujoco_refined/walk1_subject5.txt") == "ref_contact/walk1_subject5.bvh.constraint.npz"
    assert contact_path_for("motiondata_mujoco_refined_mirrored/run1_subject5_mirrored.txt") == "ref_contact_mirrored/run1_subject5_mirrored.bvh.constraint.npz"
    assert contact_path_for("ref_contact/mvae.motion.npz") == "ref_contact/mvae.constraint.npz"
    assert contact_path_for("ref_contact/crouchwalk_walk_poses.bvh.motion.npz") == "ref_contact/crouchwalk_walk_poses.bvh.constraint.npz"


def test_raw_clip_readers():
    from ipcdnnwalk_b200.srb_env import load_mocap_file, read_contact_npz
    root = os.path.join(DATA_ROOT, "gym_trackSRB")
    md = load_mocap_file(os.path.join(root, CLIPS[0]))
    left, right = read_contact_npz(os.path.join(root, "ref_contact/aiming_show.constraint.npz"))
    assert md.shape == (600, 41) and left.shape[0] >= 600 and set(np.unique(left)) <= {0.0, 1.0}
    ref_txt = "/root/reference/gym_trackSRB/motiondata_mujoco_refined/aiming_show.txt"
    if os.path.exists(ref_txt):                 # the text form the reference ships (a Python list literal)
        assert np.array_equal(load_mocap_file(ref_txt), md)


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference mount absent")
def test_playground_loader_equals_the_fixture_bit_for_bit():
    """Product loader (xml.etree + own PNG decoder) vs the oracle's parser (regex + PIL) recorded in tests/golden/terrain_playground.npz,
    and vs the constant table shipped in the package."""
    from ipcdnnwalk_b200 import TerrainModel
    T = TerrainModel.from_mjcf(REF_XML)
    for other in (TerrainModel.from_npz(os.path.join(helpers.GOLDEN, "terrain_playground.npz")), TerrainModel.builtin_playground()):
        assert T.plane_geom == other.plane_geom and np.array_equal(T.plane_size, other.plane_size)
        assert len(T.hfields) == len(other.hfields) == 2
        for a, b in zip(T.hfields, other.hfields):
            assert a["geom_id"] == b["geom_id"] and np.array_equal(a["pix"], b["pix"]) and np.array_equal(a["data"], b["data"])
            assert np.array_equal(a["size"], b["size"]) and np.array_equal(a["pos"], b["pos"])
        for k in ("box_pos", "box_size", "box_geom", "box_body", "body_id", "body_xpos", "body_snap"):
            assert np.array_equal(getattr(T, k), getattr(other, k)), k
    assert T.box_geom.shape == (324,) and list(T.body_names) == ["pyramid_1", "pyramid_2", "pyramid_3", "pyramid_4"]


def test_png_decoder_filters(tmp_path):
    """All five PNG filter types, 8-bit grey and RGB, against an encoder written here."""
    import struct
    import zlib
    from ipcdnnwalk_b200.terrain import read_png_gray8
    rng = np.random.default_rng(0)

    def encode(img, nch):
        h, w = img.shape[:2]
        rows = img.reshape(h, w * nch).astype(np.int32)
        raw = bytearray()
        prev = np.zeros(w * nch, dtype=np.int32)
        for r in range(h):
            f = r % 5
            line = rows[r]
            out = np.zeros_like(line)
            for i in range(w * nch):
                a = line[i - nch] if i >= nch else 0; c = prev[i - nch] if i >= nch else 0; up = prev[i]
                if f == 0: pr = 0
                elif f == 1: pr = a
                elif f == 2: pr = up
                elif f == 3: pr = (a + up) >> 1
                else:
                    pa, pb, pc = abs(up - c), abs(a - c), abs(a + up - 2 * c)
                    pr = a if (pa <= pb and pa <= pc) else (up if pb <= pc else c)
                out[i] = (line[i] - pr) & 255
            raw += bytes([f]) + bytes(out.astype(np.uint8).tolist())
            prev = line

        def chunk(tag, body):
            return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if nch == 1 else 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b"")
    g = rng.integers(0, 256, (13, 17)).astype(np.uint8)
    p = tmp_path / "g.png"; p.write_bytes(encode(g, 1))
    assert np.array_equal(read_png_gray8(str(p)), g)
    rgb = np.repeat(g[:, :, None], 3, axis=2)
    p = tmp_path / "c.png"; p.write_bytes(encode(rgb, 3))
    assert np.array_equal(read_png_gray8(str(p)), g)          # grey RGB -> the same grey


# ---------------------------------------------------------------------------------------------------------------- GPU
def _oracle(variant="test"):
    from oracle.srb_env import OracleSRBEnv
    o = OracleSRBEnv(helpers.fresh_clips(("aiming_show", "stand1")), variant)
    o.testing_mode = True
    return o


@pytest.mark.gpu
def test_srbenv_from_mocap_paths_replays_the_driver(built_lib):
    from ipcdnnwalk_b200 import SRBEnv
    from oracle import fullbody_map as fm
    env = SRBEnv(CLIPS, data_root=DATA_ROOT)                                  # :137
    env.impulse_interval = 1e9                                                # :138
    env.testing_mode = True                                                   # :250
    assert env.model_path == "gym_trackSRB/Characters/SRB5.xml" and abs(1 / env.model.opt.timestep - 120.0) < 0.01      # :252, :97
    obs, info = env.reset(mocap_id=0)                                         # :386 (flat env: no initial_pos_planar)
    o = _oracle()
    oo, _ = o.reset_explicit(0, 0, None, None)
    assert obs.shape == (150,) and obs.dtype == np.float32 and info == {} and np.abs(obs - oo).max() < 1e-5
    rng = np.random.default_rng(0)
    states = []
    for k in range(12):
        # show_SRB_mocap (:190-201): reference tables of the bound clip at current_frame + 1
        f = env.current_frame + 1
        assert np.abs(env.mocap_pos_com[f] - o.mocap_pos_com[f]).max() < 1e-9 and np.abs(env.mocap_ori_com[f] - o.mocap_ori_com[f]).max() < 1e-9
        assert np.abs(env.feet_pos_list[f] - o.feet_pos_list[f]).max() < 1e-9 and np.abs(np.asarray(env.feet_ori_list[f]) - o.feet_ori_list[f]).max() < 1e-9
        assert len(env.mocap_qpos) == len(o.mocap_qpos)                       # :765
        if k == 4:                                                            # push widget (:97-120): 0.2 s of a 6-vector wrench
            rate = 1 / env.model.opt.timestep
            env.impulse = 0.2 * rate
            force = np.array([-120.0, 0.0, 30.0]); lpos = np.array([-0.18, 0.13, 0.18]) * 0.5
            fvec = np.array([0, 0, 0, 0, 0, 0])                               # an INTEGER array, as the driver builds it (:117)
            fvec[0:3] = force; fvec[3:6] = np.cross(lpos, force)
            env.force = fvec; env.impulse_lpos = lpos; env.gizmo = ["RightElbow", (-0.27, 0, 0), False]
            o.impulse = 0.2 * rate; o.force = fvec
        a = helpers.tracking_action(o, rng, 0.05)
        obs, r, done, trunc, info = env.step(a)                               # :421
        states.append(env.getFullState())                                     # :422
        oo, rr, dd, _, ii = o.step(a)
        assert trunc is False and set(info["reward_info"]) == set(ii["reward_info"])
        assert np.abs(obs - oo).max() < 1e-5 and abs(r - rr) < 1e-5 * max(1, abs(rr)) and done == dd
        assert env.current_frame == o.current_frame and env.random_initial_frame == o.random_initial_frame
        assert (env.left_foot_in_contact, env.right_foot_in_contact) == (o.left_foot_in_contact, o.right_foot_in_contact)
        assert env.site_to_apply_force_on == o.site_to_apply_force_on
        assert np.abs(env.data.qpos - o.data.qpos).max() < 1e-6 and np.abs(env.data.qvel - o.data.qvel).max() < 1e-5
        for j in (env.left_site_id, env.right_site_id):                       # :160-163
            for kk in range(5):
                assert np.abs(env.data.site_xpos[j + kk] - o.data.site_xpos[j + kk]).max() < 1e-6
        assert (env.impulse > 0) == (o.impulse > 0) and env.frame_counter == o.frame_counter      # :674
        # contact forces of the last QP (:746-760)
        if env.left_foot_in_contact or env.right_foot_in_contact:
            assert len(env.contact_forces_list) == len(o.contact_forces_list)
            scale = max(1.0, max(np.abs(f).max() for f in o.contact_forces_list))
            for fa, fb in zip(env.contact_forces_list, o.contact_forces_list):
                assert np.abs(fa - fb).max() < 1e-4 * scale
        ex = env._vec.pack_obs_extra().cpu().numpy()[0]                       # ED.pack_obs_extra(env) (:426)
        ref = fm.pack_obs_extra(o)
        both = np.isnan(ref) & np.isnan(ex)
        assert np.abs(np.where(both, 0.0, ex - ref)).max() < 1e-5
    # replanning (:717-724): restore an earlier state, re-read the observation, step again -> same as the first time
    env.restoreFullState(states[-4])
    assert env.current_frame == o.current_frame - 3
    o2 = env._get_obs()
    assert o2.shape == (150,)
    env.close()


@pytest.mark.gpu
def test_srbenv_terrain_variant_surface(built_lib):
    """`from env.SRBEnv5_mocap_list_terrain_window import SRBEnv` (:12): playground model, reset(mocap_id, initial_pos_planar)."""
    from ipcdnnwalk_b200 import SRBEnv
    from oracle.terrain import Terrain
    from oracle.srb_env import OracleSRBEnv, lift_clip_to_terrain
    planar = (22.0, 24.0)
    env = SRBEnv(CLIPS[:1], variant="terrain", data_root=DATA_ROOT)            # no XML under the fixture root: the packaged table is used
    env.testing_mode = True
    assert env.model_path == "gym_trackSRB/Characters/SRB5_playground.xml"
    obs, _ = env.reset(mocap_id=0, initial_pos_planar=np.array(planar))
    T = Terrain.load(os.path.join(helpers.GOLDEN, "terrain_playground.npz"))
    o = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, planar)], "terrain", terrain=T)
    o.testing_mode = True
    oo, _ = o.reset_explicit(0, 0, None, None)
    assert np.abs(obs - oo).max() < 1e-5
    rng = np.random.default_rng(1)
    for k in range(6):
        a = helpers.tracking_action(o, rng, 0.05)
        obs, r, done, _, _ = env.step(a)
        oo, rr, dd, _, _ = o.step(a)
        assert np.abs(obs - oo).max() < 1e-5 and r == 0.0 and done == dd
        assert np.abs(env.data.qpos - o.data.qpos).max() < 1e-6
    env.close()


@pytest.mark.gpu
def test_sb3_vecenv_seam(built_lib):
    """What stable-baselines3 calls on the vec_env_cls of make_vec_env: numpy protocol, terminal observations, episode records,
    get_attr / set_attr / env_method / env_is_wrapped / seed."""
    import torch
    from ipcdnnwalk_b200 import SB3VecEnv, SRBVecEnv
    n = 32                                                                    # n_envs=32 (train_delta.py:94)
    clips = helpers.fresh_clips()
    rng = np.random.default_rng(3)
    plans = np.stack([helpers.random_plan(rng, clips) for _ in range(n)])
    plans[::2, 4] = -12.0                                                     # some environments fall and terminate soon
    venv = SB3VecEnv(clips, n, variant="training", precision=64, seed=11)
    twin = SRBVecEnv(clips, n, variant="training", precision=64, seed=11, auto_reset=True)
    twin.attach(term_obs=True)
    obs = venv.reset(plan=plans); twin.reset(plan=plans)
    assert isinstance(obs, np.ndarray) and obs.shape == (n, 150) and obs.dtype == np.float32 and venv.num_envs == n
    act = torch.zeros(n, 22, device="cuda")
    n_done = 0
    for k in range(40):
        twin.tracking_action(act, sigma=0.1, seed=4, step=k)
        a = act.cpu().numpy()
        venv.step_async(a)
        obs, rew, done, infos = venv.step_wait()
        o2, r2, d2, _ = twin.step(act)
        torch.cuda.synchronize()
        assert obs.dtype == np.float32 and rew.shape == (n,) and done.dtype == bool and len(infos) == n
        assert np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew, r2.cpu().numpy()) and np.array_equal(done, d2.cpu().numpy().astype(bool))
        for i in range(n):
            assert ("terminal_observation" in infos[i]) == bool(done[i]) and "reward_info" in infos[i]
            if done[i]:
                assert np.array_equal(infos[i]["terminal_observation"], twin.term_obs[i].cpu().numpy())
                assert infos[i]["episode"]["l"] >= 1 and np.isfinite(infos[i]["episode"]["r"])
                n_done += 1
    assert n_done >= 2
    assert venv.get_attr("current_frame", indices=[0, 5]) == [twin.named_state(0)["current_frame"], twin.named_state(5)["current_frame"]]
    assert venv.env_is_wrapped(object) == [False] * n and venv.seed() == [11] * n
    st = venv.env_method("getFullState", indices=[3])[0]
    venv.env_method("restoreFullState", st, indices=[3])
    venv.set_attr("impulse", 5, indices=[2])
    assert venv.get_attr("impulse", indices=2) == [5]
    with pytest.raises(AttributeError):
        venv.set_attr("no_such_attribute", 1)
    venv.close(); twin.close()
