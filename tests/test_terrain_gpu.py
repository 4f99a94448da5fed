"""GPU parity of the terrain variant (env/SRBEnv5_mocap_list_terrain_window.py) against the oracle."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
NPZ = os.path.join(helpers.GOLDEN, "terrain_playground.npz")


def test_terrain_height_ray_bit_exact_indices(built_lib):
    """Integer outputs (geom id, height-field cell, triangle) bit exact; heights to round-off."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from oracle.terrain import Terrain
    T = Terrain.load(NPZ)
    env = SRBVecEnv(helpers.fresh_clips(("aiming_show",)), 4, variant="terrain", precision=64, terrain=TerrainModel.from_npz(NPZ))
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-90, 90, (4000, 2)), rng.uniform(2, 38, (500, 2)), rng.uniform(-82, -2, (500, 1)) * [[1, -1]],
                          np.array([[0, 0], [20, 20], [2.0, 20], [38.0, 20.0], [-42, 42], [-2.0, 2.0], [-82, 82], [99.9, -99.9], [150.0, 0.0]])])
    h, hit = env.terrain_height(torch.as_tensor(pts))
    h = h.cpu().numpy(); hit = hit.cpu().numpy()
    n_hf = n_box = 0
    for i, (x, y) in enumerate(pts):
        ref_h, gid = T.height_ray(x, y)
        assert hit[i, 0] == gid, (x, y)
        if ref_h is None:
            continue
        assert abs(h[i] - ref_h) <= 1e-12
        for hf in T.hfields:
            if hf["geom_id"] == gid:
                c, r, tri, _ = T.hfield_cell(hf, x, y)
                assert (hit[i, 1], hit[i, 2], hit[i, 3]) == (c, r, tri)
                n_hf += 1
        n_box += gid >= T.boxes[0]["geom_id"]
    assert n_hf > 500 and n_box > 300
    env.close()


@pytest.mark.parametrize("planar", [(0.0, 0.0), (22.0, 24.0), (-42.0, 42.0), (42.0, -40.0)])
def test_terrain_rollout_matches_oracle(built_lib, planar):
    """flat ground, pyramid (edge snapping), hill height field, rough height field."""
    import torch
    from ipcdnnwalk_b200 import SRBVecEnv, TerrainModel
    from oracle.terrain import Terrain
    from oracle.srb_env import OracleSRBEnv, lift_clip_to_terrain
    T = Terrain.load(NPZ)
    base = helpers.fresh_clips(("aiming_show",))
    n = 3
    rng = np.random.default_rng(5)
    frames = [0, 40, 233]
    env = SRBVecEnv(base, n, variant="terrain", precision=64, auto_reset=False, terrain=TerrainModel.from_npz(NPZ),
                    initial_pos_planar=np.tile(np.array(planar), (n, 1)))
    plans = np.stack([np.concatenate([[0, f], rng.uniform(-.05, .05, 3), [1.0, 0, 0, 0]]) for f in frames])
    obs0 = env.reset(plan=plans).cpu().numpy()
    oracles = []
    for i in range(n):
        o = OracleSRBEnv([lift_clip_to_terrain(helpers.load_clip("aiming_show"), T, planar)], "terrain", terrain=T)
        oo, _ = helpers.oracle_reset(o, plans[i])
        assert np.abs(obs0[i] - oo).max() <= 2e-6, "reset obs"
        oracles.append(o)
    snapped = 0
    for k in range(25):
        acts = np.stack([helpers.tracking_action(o, rng, 0.05) for o in oracles])
        # force some touchdowns so that the pyramid edge snap is exercised
        if k % 5 == 2:
            acts[:, 18:22] = 0; acts[:, 18 + 2] = 1          # right foot only
        if k % 5 == 4:
            acts[:, 18:22] = 0; acts[:, 18 + 3] = 1          # both feet: left touches down
        obs, rew, done, _ = env.step(torch.as_tensor(acts, device="cuda"))
        torch.cuda.synchronize()
        for i, o in enumerate(oracles):
            was_swing = not o.left_foot_in_contact
            oo, rr, dd, _, _ = o.step(acts[i])
            if was_swing and o.left_foot_in_contact:
                # touchdown: on a pyramid the foot was projected to a layer edge, i.e. it now lies on the boundary of a
                # square centred on the pyramid
                c = np.abs(o.left_foot_contact_pos_global[:2] - np.array([20.0, 20.0]))
                if planar == (22.0, 24.0) and abs(max(c) * 10 - round(max(c) * 10)) < 1e-9:
                    snapped += 1
            st = helpers.cuda_state_vector(env.named_state(i)); ref = helpers.oracle_state_vector(o)
            assert (st[23:26] == ref[23:26]).all()
            err = np.abs(st[:23] - ref[:23]) / np.maximum(1, np.abs(ref[:23]))
            assert err.max() <= 1e-5, (k, i, err.argmax(), st[:23], ref[:23])
            assert (np.abs(obs[i].cpu().numpy() - oo) / np.maximum(1, np.abs(oo))).max() <= 1e-5
            assert float(rew[i]) == 0.0 and rr == 0 and bool(done[i]) == dd
    if planar == (22.0, 24.0):
        assert snapped >= 1, "the pyramid edge snap was never exercised"
    env.close()
