"""Shared test infrastructure of the AdaptiveSRB (gym_cdm2) path: golden tables, oracle <-> device state conversion."""
import json
import os

import numpy as np

import helpers

TABLES = os.path.join(helpers.GOLDEN, "cdm_walk_tables.npz")
ROLLOUT = os.path.join(helpers.GOLDEN, "cdm_rollout.npz")
TABLE_KEYS = ("cdm", "dcdm", "cdm_d", "fullbody", "iframe_tab", "con", "touch_down", "touch_off", "rtt_down", "rtt_off", "swing_dur",
              "feet_off", "nf", "leg_cols")


def load_tables(clip="walk"):
    """tests/golden/cdm_{walk,run}_tables.npz (tests/tools/make_golden_cdm.py) -> (tables dict, skeleton dict)."""
    d = np.load(os.path.join(helpers.GOLDEN, f"cdm_{clip}_tables.npz"), allow_pickle=False)
    tab = {k: d[k].copy() for k in TABLE_KEYS}
    skel = json.loads(str(d["skeleton_json"]))
    return tab, skel


def random_plan(rng, scale=1.0):
    """Reset draws with the reference's distributions (deepmimiccdm-v1.lua:1625-1647, noise scales of walk2-v1)."""
    p = np.zeros(16)
    p[0:4] = rng.uniform(-0.5, 0.5, 4) * scale
    p[4:7] = rng.uniform(-0.05, 0.05, 3) * scale
    p[7:13] = rng.uniform(-0.5, 0.5, 6) * scale
    p[13:16] = (rng.uniform(0, 1, 3) - 0.5) * np.array([0.05, 0.5, 0.05]) * scale
    return p


def oracle_reset(env, plan):
    return env.reset(dict(q=plan[0:4], pos=plan[4:7], dq=plan[7:13], vel3=plan[13:16]))


def oracle_state_vectors(env):
    """OracleCDMEnv -> (reals[61], ints[4], hist[8,8]) in the device layout (csrc/cdm_types.h)."""
    r = np.zeros(61)
    r[0:3] = env.p; r[3:7] = env.Q; r[7:10] = env.v; r[10:13] = env.wb
    r[13] = env.iframe; r[14] = env.global_time; r[15:19] = env.target_Y; r[19:22] = env.sim_prev[0]; r[22] = 0.0
    flags = 0
    for l, i in enumerate((1, 2)):
        li = env.leg[i]
        b = 23 + 19 * l
        r[b:b + 3] = li.contactPos; r[b + 3:b + 6] = li.contactPreFilter
        r[b + 6:b + 9] = [li.ball[k].x[0] for k in range(3)]
        r[b + 9:b + 12] = [li.ball[k].x[1] for k in range(3)]
        r[b + 12:b + 14] = li.filtQ.x
        r[b + 14:b + 18] = li.filtQorigin
        r[b + 18] = li.swingPhase
        if li.isSwing:
            flags |= 1 << l
    n = len(env.replay)
    hist = np.zeros((8, 8))
    for idx in range(max(0, n - 8), n):
        hist[idx & 7, :7] = env.replay[idx]
    ints = np.array([flags, n, 0, 0], dtype=np.int32)
    return r, ints, hist


def compare_state(dev, env, rtol, atol, what=""):
    """Device named/real state vs the oracle's; integer parts bit exact."""
    r, ints, hist = oracle_state_vectors(env)
    dr, di = dev["reals"], dev["ints"]
    assert int(di[0]) == int(ints[0]), f"{what}: swing flags differ {di[0]} vs {ints[0]}"
    assert int(di[1]) == int(ints[1]), f"{what}: replay length differs"
    sel = np.ones(61, bool)
    sel[22] = False                                   # running episode return is device bookkeeping
    for l in range(2):
        b = 23 + 19 * l
        if not (ints[0] >> l) & 1:
            sel[b + 9:b + 12] = True                  # particle velocity is carried through support by both
    err = np.abs(dr[sel] - r[sel]) - (atol + rtol * np.abs(r[sel]))
    assert (err <= 0).all(), f"{what}: state mismatch at fields {np.nonzero(sel)[0][err > 0]} dev {dr[sel][err > 0]} oracle {r[sel][err > 0]}"


def load_terrain():
    """tests/golden/cdm_terrain_heights.npz -> dict(heights [cm], size_x, size_z [cm], terrain_pos [m]) as CDMVecEnv(terrain=...) takes it."""
    d = np.load(os.path.join(helpers.GOLDEN, "cdm_terrain_heights.npz"))
    return dict(heights=d["heights"].copy(), size_x=float(d["size_x"]), size_z=float(d["size_z"]), terrain_pos=d["terrain_pos"].copy())


def load_terrain_oracle():
    from oracle import tl_terrain as tt
    t = load_terrain()
    return tt.OracleTerrain(t["heights"], t["size_x"], t["size_z"], True), t["terrain_pos"]
