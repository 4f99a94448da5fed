"""GPU parity of the device-side mocap -> SRB table builder (row a-13) against oracle/srbdata.py."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_clip_builder_matches_oracle_tables(built_lib):
    from ipcdnnwalk_b200.clip_builder import ClipBuilder
    g = np.load(os.path.join(helpers.GOLDEN, "srbdata_case.npz"))
    b = ClipBuilder()
    clip = b.build(g["mocap_raw"], g["left"], g["right"])
    assert clip["cycle_length"] == int(g["cycle_length"])
    n = g["mocap_qpos"].shape[0]
    assert clip["mocap_qpos"].shape == (n, 7)
    assert np.abs(clip["mocap_qpos"][:, :3] - g["mocap_qpos"][:, :3]).max() < 1e-11          # whole-body CoM
    # root quaternion (second cycle: same rotation, sign convention w >= 0)
    assert np.abs(clip["mocap_qpos"][:, 3:] - g["mocap_qpos"][:, 3:]).max() < 1e-10
    assert np.abs(clip["mocap_ori_com"] - g["mocap_ori_com"]).max() < 1e-10
    assert np.abs(clip["mocap_vel_com"] - g["mocap_vel_com"]).max() < 1e-8
    assert np.abs(clip["feet_pos_list"] - g["feet_pos_list"]).max() < 1e-11
    assert ((clip["feet_pos_list"][:, :, 2] == 0) == (g["feet_pos_list"][:, :, 2] == 0)).all()   # contact mask bit exact
    assert np.abs(clip["feet_yaw"] - g["feet_yaw"]).max() < 1e-10
    b.close()


def test_built_clip_drives_the_env(built_lib):
    """Tables from the builder go straight into the batched env and reproduce the oracle's reset observation."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv
    from ipcdnnwalk_b200.clip_builder import ClipBuilder
    from oracle.srb_env import OracleSRBEnv
    g = np.load(os.path.join(helpers.GOLDEN, "srbdata_case.npz"))
    b = ClipBuilder()
    clip = b.build(g["mocap_raw"], g["left"], g["right"])
    env = SRBVecEnv([clip], 2, variant="training", precision=64, auto_reset=False)
    plans = np.array([[0, 3, 0, 0, 0, np.nan, 0, 0, 0], [0, 60, 0, 0, 0, np.nan, 0, 0, 0]], dtype=float)
    obs = env.reset(plan=plans).cpu().numpy()
    n = g["mocap_qpos"].shape[0]
    yaw = g["feet_yaw"]
    fo = np.zeros((n, 2, 3, 3)); fo[:, :, 0, 0] = yaw[:, :, 0]; fo[:, :, 0, 1] = -yaw[:, :, 1]; fo[:, :, 1, 0] = yaw[:, :, 1]; fo[:, :, 1, 1] = yaw[:, :, 0]; fo[:, :, 2, 2] = 1
    oc = dict(mocap_dt=0.033333, cycle_length=int(g["cycle_length"]), mocap_qpos=g["mocap_qpos"], mocap_pos_com=g["mocap_qpos"][:, :3],
              mocap_ori_com=g["mocap_ori_com"], mocap_vel_com=g["mocap_vel_com"], feet_pos_list=g["feet_pos_list"], feet_ori_list=fo)
    for i in range(2):
        o = OracleSRBEnv([oc], "training")
        o.cycle_length = 600                      # the 90-frame test clip is shorter than the env's 512-frame episodes
        oo, _ = helpers.oracle_reset(o, plans[i])
        assert np.abs(obs[i] - oo).max() < 2e-6
    env.close(); b.close()
