"""Single-environment proxy with the reference's `SRBEnv` surface, over the batched CUDA step.

Reference: class SRBEnv(Env) -- env/SRBEnv5_mocap_list_window.py:122-176 (constructor, spaces), :208 step, :805 reset,
:1005-1053 getFullState / restoreFullState; attributes read by the drivers (walk_SRBTrack2025_terrain_singlethreaded.py:
98-120,137-138,163,190-201,386,401,421-426): data.qpos / data.qvel / data.site_xpos / data.site_xmat, current_frame,
random_initial_frame, left/right_foot_in_contact, site_to_apply_force_on, testing_mode, impulse*, model.opt.timestep.

This is a convenience for the per-environment drivers; training should use SRBVecEnv (one launch for all environments).
gymnasium is optional: if it is importable the class derives from gymnasium.Env and exposes Box spaces."""
import types

import numpy as np

from .srb_vec_env import SRBVecEnv, REWARD_INFO_KEYS

try:                                                     # pragma: no cover - not installed in the build image
    import gymnasium as _gym
    _Base = _gym.Env
except Exception:                                        # noqa: BLE001
    _gym = None
    _Base = object

_REL = np.array([[0.12, 0.06, 0.0], [0.12, -0.06, 0.0], [-0.12, 0.06, 0.0], [-0.12, -0.06, 0.0]])


class SRBEnv(_Base):
    """`clips`: list of reference-table dicts (what SRBEnv.loadMocap builds; see clip_builder.ClipBuilder)."""

    def __init__(self, clips, variant="test", precision=64, device=0, seed=0, terrain=None):
        self._vec = SRBVecEnv(clips, 1, variant=variant, precision=precision, auto_reset=False, device=device, seed=seed, terrain=terrain)
        self.testing_mode = False
        self.window_size = 5
        self.SRB_data_num = len(clips)
        self.mu = 1.2
        self.model = types.SimpleNamespace(opt=types.SimpleNamespace(timestep=0.0083332), nv=6)
        self.model_path = "gym_trackSRB/Characters/SRB5_playground.xml" if variant == "terrain" else "gym_trackSRB/Characters/SRB5.xml"
        self.left_site_id, self.right_site_id = 0, 5
        self.site_id_list = [0, 5]
        if _gym is not None:                             # pragma: no cover
            self.action_space = _gym.spaces.Box(low=-3.14, high=3.14, shape=(22,), dtype=np.float32)
            self.observation_space = _gym.spaces.Box(low=-np.inf, high=np.inf, shape=(150,), dtype=np.float32)
        self._episode = 0
        self._seed = seed
        self._refresh()

    # ---- gym API ---------------------------------------------------------------------------------
    def reset(self, seed=None, options=None, mocap_id=0, initial_pos_planar=None):
        if initial_pos_planar is not None:
            self._vec.set_planar_offsets(np.asarray(initial_pos_planar, dtype=np.float64).reshape(1, 2))
        if self.testing_mode:                            # chosen clip, frame 0, no noise (:826-842)
            plan = np.array([[mocap_id, 0, 0, 0, 0, np.nan, 0, 0, 0]], dtype=np.float64)
            obs = self._vec.reset(plan=plan)
        else:
            obs = self._vec.reset()
        self._refresh()
        return obs[0].cpu().numpy().copy(), {}

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 22)
        obs, rew, done, rinfo = self._vec.step(a)
        self._refresh()
        info = {"reward_info": dict(zip(REWARD_INFO_KEYS, map(float, rinfo[0].cpu().numpy())))}
        return obs[0].cpu().numpy().copy(), float(rew[0]), bool(done[0]), False, info

    def _get_obs(self):
        return self._vec.obs[0].cpu().numpy().copy()

    def close(self):
        self._vec.close()

    # ---- state mirrors -------------------------------------------------------------------------------
    def _refresh(self):
        s = self._vec.named_state(0)
        self.current_frame = s["current_frame"]
        self.random_initial_frame = s["random_initial_frame"]
        self.left_foot_in_contact = s["left_foot_in_contact"]
        self.right_foot_in_contact = s["right_foot_in_contact"]
        self.left_foot_contact_pos_global = s["left_foot_contact_pos_global"]
        self.right_foot_contact_pos_global = s["right_foot_contact_pos_global"]
        site_xpos = np.zeros((10, 3)); site_xmat = np.zeros((10, 9))
        for f, (pos, cs) in enumerate(((s["left_foot_contact_pos_global"], s["left_foot_yaw_cs"]), (s["right_foot_contact_pos_global"], s["right_foot_yaw_cs"]))):
            R = np.array([[cs[0], -cs[1], 0.0], [cs[1], cs[0], 0.0], [0.0, 0.0, 1.0]])
            site_xpos[5 * f] = pos
            site_xpos[5 * f + 1:5 * f + 5] = pos + _REL @ R.T
            site_xmat[5 * f] = R.reshape(9)
        self.data = types.SimpleNamespace(qpos=s["qpos"], qvel=s["qvel"], site_xpos=site_xpos, site_xmat=site_xmat)
        lc, rc = self.left_foot_in_contact, self.right_foot_in_contact
        self.site_to_apply_force_on = [i if lc else -1 for i in range(5)] + [5 + i if rc else -1 for i in range(5)]

    def getFullState(self):
        return self._vec.get_state(0)

    def restoreFullState(self, state):
        self._vec.set_state(0, state)
        self._refresh()
