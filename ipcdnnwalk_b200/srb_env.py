"""`SRBEnv` with the reference's constructor and attribute surface, over the batched CUDA step.

Reference: class SRBEnv(Env) -- env/SRBEnv5_mocap_list_window.py:122-176 (`SRBEnv(mocap_path_list)`, spaces), :182-206 (push schedule),
:208 step, :805 reset, :1005-1053 getFullState / restoreFullState, :1171-1197 loadMocap; the terrain variant
env/SRBEnv5_mocap_list_terrain_window.py (model gym_trackSRB/Characters/SRB5_playground.xml, reset(..., initial_pos_planar)).
Attributes the drivers read or write (walk_SRBTrack2025_terrain_singlethreaded.py:97-120,137-138,160-163,190-201,250-252,386,401,
421-426,674-704,717-724,746-765; gym_trackSRB/taesooLib/DeltaExtractor.py:1015-1034): data.qpos / qvel / site_xpos / site_xmat,
model.opt.timestep, model_path, current_frame, random_initial_frame, mocap_qpos / mocap_pos_com / mocap_ori_com / mocap_vel_com /
feet_pos_list / feet_ori_list, left/right_foot_in_contact, site_to_apply_force_on, contact_forces_list, right_site_id, impulse /
force / impulse_interval / impulse_duration / impulse_lpos / gizmo, testing_mode, getFullState / restoreFullState / _get_obs.

`mocap_path_list` holds the reference's clip paths (relative to `gym_trackSRB/`, e.g. 'motiondata_mujoco_refined/stand1.txt'); the
tables are built on the device by `clip_builder.ClipBuilder` from the clip file and its contact file (SRBDataProcess,
env/testSRB_mujoco_original_viewer.py:765-818).  A list of ready-made table dicts is accepted too.

This single-environment class serves the per-environment drivers; training uses SRBVecEnv / SB3VecEnv (one launch for all
environments).  gymnasium is optional: when importable the class derives from gymnasium.Env and exposes Box spaces."""
import ast
import math
import os
import types

import numpy as np

from .srb_vec_env import SRBVecEnv, REWARD_INFO_KEYS
from ._lib import DBG_DIM

try:                                                     # pragma: no cover - not installed in the build image
    import gymnasium as _gym
    _Base = _gym.Env
except Exception:                                        # noqa: BLE001
    _gym = None
    _Base = object

_REL = np.array([[0.12, 0.06, 0.0], [0.12, -0.06, 0.0], [-0.12, 0.06, 0.0], [-0.12, -0.06, 0.0]])
_MODEL_PATH = {"test": "gym_trackSRB/Characters/SRB5.xml", "training": "Characters/SRB5.xml", "terrain": "gym_trackSRB/Characters/SRB5_playground.xml"}
_STAND_CLIP = "motiondata_mujoco_refined/stand1.txt"     # ..._training.py:643-645: `is_stand` iff this very path


def contact_path_for(mocap_path):
    """SRBdata.__init__ (viewer.py:769-798): the contact file that belongs to a clip path (both relative to gym_trackSRB/)."""
    d, f = os.path.split(mocap_path)
    if d.startswith("motiondata_mujoco_refined") and f.endswith(".txt"):
        return os.path.join(d.replace("motiondata_mujoco_refined", "ref_contact"), f[:-4] + ".bvh.constraint.npz")
    return mocap_path[:-10] + "constraint.npz"


def load_mocap_file(path):
    """viewer.py:802-815: an .npz with 'motion' / 'arr_0', or a text file holding a Python list literal of 41-float rows."""
    if path.endswith("npz"):
        d = np.load(path)
        for k in ("motion", "arr_0"):
            if k in d:
                return np.asarray(d[k], dtype=np.float64)
        raise KeyError("Neither 'motion' nor 'arr_0' found in file.")
    with open(path, "r") as f:
        return np.array(ast.literal_eval(f.read()), dtype=np.float64)


def read_contact_npz(path):
    """viewer.py:750-762: per-frame foot contact = heel or toe flag."""
    d = np.load(path)
    left = np.logical_or(d["leftHeel"], d["leftToe"]).astype(np.float64)
    right = np.logical_or(d["rightHeel"], d["rightToe"]).astype(np.float64)
    return left, right


def build_clip_tables(mocap_path_list, data_root=None, device=0):
    """SRBEnv.loadMocap for a list of clip paths (relative to <data_root>/gym_trackSRB/): ClipBuilder tables ready for SRBVecEnv."""
    env = SRBEnv(mocap_path_list, data_root=data_root, device=device)
    for i in range(env.SRB_data_num):
        env.loadMocap(i)
    return env.SRB_data_list


class _Model:
    def __init__(self, timestep):
        self.opt = types.SimpleNamespace(timestep=timestep)
        self.nv = 6


class SRBEnv(_Base):
    def __init__(self, mocap_path_list, variant="test", precision=64, device=0, seed=0, terrain=None, data_root=None):
        """variant: "test" (env/SRBEnv5_mocap_list_window.py), "training" (..._training.py), "terrain" (..._terrain_window.py).
        data_root: directory that contains `gym_trackSRB/` (default: $IPCDNNWALK_DATA_ROOT, else the working directory, as the
        reference resolves its paths)."""
        self.variant = variant
        self.mocap_path_list = list(mocap_path_list)
        self.SRB_data_num = len(self.mocap_path_list)
        self.SRB_data_list = [None] * self.SRB_data_num
        self._data_root = data_root or os.environ.get("IPCDNNWALK_DATA_ROOT", os.getcwd())
        self._device, self._precision, self._seed = device, precision, seed
        self.model_path = _MODEL_PATH[variant]
        self.model = _Model(0.0083332)
        if variant == "terrain" and terrain is None:
            from .terrain import TerrainModel
            xml = os.path.join(self._data_root, self.model_path)
            terrain = TerrainModel.from_mjcf(xml) if os.path.exists(xml) else TerrainModel.builtin_playground()
        self._terrain = terrain
        self._vec = None
        self.minLogQ = 9
        self.window_size = 5
        self.mu = 1.2
        self.testing_mode = False
        self.cdm_id = 1
        self.left_site_id, self.right_site_id = 0, 5
        self.site_id_list = [0, 5]
        self.action = None
        # push schedule (:182-189; the training variant's defaults: ..._training.py:199-205)
        self._impulse_args = None
        if variant == "training":
            self.impulse_interval, self.force = 100000000, [1000 / math.sqrt(2) * 0.5, -1000 / math.sqrt(2) * 0.5, 0]
        else:
            self.impulse_interval, self.force = 1e9, [700, 700, -350]
        self.impulse_duration = 24
        self.impulse_lpos = None
        self.gizmo = None
        self.f = np.zeros(6)
        if _gym is not None:                             # pragma: no cover
            self.action_space = _gym.spaces.Box(low=-3.14, high=3.14, shape=(22,), dtype=np.float32)
            self.observation_space = _gym.spaces.Box(low=-np.inf, high=np.inf, shape=(150,), dtype=np.float32)
        self.current_frame = 0
        self.random_initial_frame = None
        self.left_foot_in_contact = False
        self.right_foot_in_contact = False
        self.left_foot_contact_pos_global = None
        self.right_foot_contact_pos_global = None
        self.site_to_apply_force_on = [-1, -1, -1, -1, -1, -1]
        self.contact_forces_list = []
        self.data = types.SimpleNamespace(qpos=np.zeros(7), qvel=np.zeros(6), site_xpos=np.zeros((10, 3)), site_xmat=np.zeros((10, 9)),
                                          xfrc_applied=np.zeros((2, 6)))
        self._clip = None

    # ---- clip tables (loadMocap :1171-1197) ---------------------------------------------------------------------------------
    def loadMocap(self, i):
        if self.SRB_data_list[i] is not None:
            return
        src = self.mocap_path_list[i]
        if isinstance(src, dict):                        # ready-made tables
            clip = dict(src)
        else:
            from .clip_builder import ClipBuilder
            root = os.path.join(self._data_root, "gym_trackSRB")
            md = load_mocap_file(os.path.join(root, src))
            left, right = read_contact_npz(os.path.join(root, contact_path_for(src)))
            b = ClipBuilder(device=self._device)
            clip = b.build(md, left, right, is_stand=(src == _STAND_CLIP or os.path.basename(src).startswith("stand1.")))
            b.close()
        clip.setdefault("mocap_dt", 0.033333)
        if "mocap_pos_com" not in clip:
            clip["mocap_pos_com"] = np.asarray(clip["mocap_qpos"])[:, :3]
        if "feet_ori_list" not in clip:                  # pure-yaw matrices from (cos, sin)
            yaw = np.asarray(clip["feet_yaw"]); n = yaw.shape[0]
            fo = np.zeros((n, 2, 3, 3))
            fo[:, :, 0, 0] = yaw[:, :, 0]; fo[:, :, 0, 1] = -yaw[:, :, 1]; fo[:, :, 1, 0] = yaw[:, :, 1]; fo[:, :, 1, 1] = yaw[:, :, 0]; fo[:, :, 2, 2] = 1.0
            clip["feet_ori_list"] = fo
        self.SRB_data_list[i] = clip

    def _ensure_device_env(self):
        if self._vec is None:
            for i in range(self.SRB_data_num):
                self.loadMocap(i)
            self._vec = SRBVecEnv(self.SRB_data_list, 1, variant=self.variant, precision=self._precision, auto_reset=False, device=self._device,
                                  seed=self._seed, terrain=self._terrain)
            self._vec.attach(dbg=True, contact_prob=(self.variant != "training"))
        return self._vec

    def _bind_clip(self, i):
        c = self.SRB_data_list[i]
        self._clip = i
        self.mocap_dt = c["mocap_dt"]; self.cycle_length = c["cycle_length"]
        self.mocap_qpos = c["mocap_qpos"]; self.mocap_pos_com = c["mocap_pos_com"]; self.mocap_ori_com = c["mocap_ori_com"]
        self.mocap_vel_com = c["mocap_vel_com"]; self.feet_pos_list = c["feet_pos_list"]; self.feet_ori_list = c["feet_ori_list"]

    # ---- push schedule: attributes are plain Python values as in the reference; they reach the device before a step ------------
    @property
    def impulse(self):
        return self._impulse

    @impulse.setter
    def impulse(self, v):
        self._impulse = v
        self._impulse_dirty = True

    _impulse, _impulse_dirty = 0, False

    def _push_impulse_settings(self):
        v = self._vec
        f = np.asarray([float(x) for x in (list(self.force) if not hasattr(self.force, "x") else (self.force.x, self.force.y, self.force.z))])
        args = (float(self.impulse_interval), int(self.impulse_duration), tuple(f))
        if args != self._impulse_args:
            v.set_impulse(int(min(float(self.impulse_interval), 2 ** 31 - 1)), int(self.impulse_duration), f)
            self._impulse_args = args
        if self._impulse_dirty:                          # `env.impulse = ...` from the driver: remaining pushes = the number of
            s = v.get_state(0)                           # sub-steps that still see impulse > 0 (a float count rounds up)
            s["ints"][5] = int(math.ceil(max(float(self._impulse), 0.0)))
            v.set_state(0, s)
            self._impulse_dirty = False

    # ---- gym API ------------------------------------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None, mocap_id=0, initial_pos_planar=None):
        v = self._ensure_device_env()
        if initial_pos_planar is not None:
            v.set_planar_offsets(np.asarray(initial_pos_planar, dtype=np.float64).reshape(1, 2))
        if self.testing_mode:                            # chosen clip, frame 0, no noise (:826-842)
            plan = np.array([[mocap_id, 0, 0, 0, 0, np.nan, 0, 0, 0]], dtype=np.float64)
            obs = v.reset(plan=plan)
        else:
            obs = v.reset()
        self._refresh()
        return obs[0].cpu().numpy().copy(), {}

    def step(self, action):
        v = self._ensure_device_env()
        self._push_impulse_settings()
        if v.contact_prob is not None:                   # Policy_prior_posterior_MoE_window._last_contact_prob side channel (:239)
            p = getattr(self, "contact_prob", None)
            row = [float("nan"), float("nan")] if (p is None or p[0] is None) else [float(p[0]), float(p[1])]
            v.contact_prob.copy_(v.torch.tensor([row], dtype=v.torch.float32))
        a = np.asarray(action, dtype=np.float32).reshape(1, 22)
        self.action = a[0]
        v.dbg.zero_()
        obs, rew, done, rinfo = v.step(a)
        self._refresh(after_step=True)
        info = {"reward_info": dict(zip(REWARD_INFO_KEYS, map(float, rinfo[0].cpu().numpy())))}
        return obs[0].cpu().numpy().copy(), float(rew[0]), bool(done[0]), False, info

    def _get_obs(self):
        return self._ensure_device_env().obs[0].cpu().numpy().copy()

    def close(self):
        if self._vec is not None:
            self._vec.close()
            self._vec = None

    # ---- state mirrors --------------------------------------------------------------------------------------------------------------
    def _refresh(self, after_step=False):
        s = self._vec.named_state(0)
        if s["clip"] != self._clip:
            self._bind_clip(s["clip"])
        self.current_frame = s["current_frame"]
        self.random_initial_frame = s["random_initial_frame"]
        self.left_foot_in_contact = s["left_foot_in_contact"]
        self.right_foot_in_contact = s["right_foot_in_contact"]
        self.left_foot_contact_pos_global = s["left_foot_contact_pos_global"]
        self.right_foot_contact_pos_global = s["right_foot_contact_pos_global"]
        self._impulse = s["impulse"]; self._impulse_dirty = False
        self.frame_counter = s["frame_counter"]
        site_xpos = np.zeros((10, 3)); site_xmat = np.zeros((10, 9))
        oris = []
        for f, (pos, cs) in enumerate(((s["left_foot_contact_pos_global"], s["left_foot_yaw_cs"]), (s["right_foot_contact_pos_global"], s["right_foot_yaw_cs"]))):
            R = np.array([[cs[0], -cs[1], 0.0], [cs[1], cs[0], 0.0], [0.0, 0.0, 1.0]])
            oris.append(R)
            site_xpos[5 * f] = pos
            site_xpos[5 * f + 1:5 * f + 5] = pos + _REL @ R.T
            site_xmat[5 * f] = R.reshape(9)
        self.left_foot_contact_ori, self.right_foot_contact_ori = oris
        self.data.qpos, self.data.qvel, self.data.site_xpos, self.data.site_xmat = s["qpos"], s["qvel"], site_xpos, site_xmat
        lc, rc = self.left_foot_in_contact, self.right_foot_in_contact
        self.site_to_apply_force_on = [i if lc else -1 for i in range(5)] + [5 + i if rc else -1 for i in range(5)]
        if after_step:
            self._contact_forces(s)

    def _contact_forces(self, s):
        """contact_forces_list of the last QP (:395-449): per active site f = sum_d lambda_d (R_proj b_d), in the order of
        site_to_apply_force_on, after the in-place 1e4 N clamp of the test env (Q3).  lambda comes from the step's debug record."""
        dbg = self._vec.dbg[0].cpu().numpy().reshape(4, DBG_DIM // 4)
        self.contact_forces_list = []
        if not (self.left_foot_in_contact or self.right_foot_in_contact):
            return
        lam = dbg[3, 24:64].reshape(10, 4)
        qw, qx, qy, qz = s["xquat_lagged"]               # the xmat of the last mj_step1
        x0, x1 = 1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy + qw * qz)
        n = math.hypot(x0, x1)
        fx, fy = (x0 / n, x1 / n) if n != 0 else (x0, x1)
        mu_b = 1.2 * self.mu
        B = np.array([[fx * bx - fy * by, fy * bx + fx * by, 1.0] for bx, by in ((mu_b, 0), (-mu_b, 0), (0, mu_b), (0, -mu_b))])
        forces = [(sid, lam[sid] @ B) for sid in self.site_to_apply_force_on if sid != -1]
        if self.variant == "test":
            sums = [np.zeros(3), np.zeros(3)]
            for sid, f in forces:
                sums[0 if sid < 5 else 1] = sums[0 if sid < 5 else 1] + f
            lens = [math.sqrt(float(v.dot(v))) for v in sums]
            forces = [(sid, f * (10000.0 / lens[1 if sid > 5 else 0]) if lens[1 if sid > 5 else 0] > 10000.0 else f) for sid, f in forces]
        self.contact_forces_list = [f for _, f in forces]

    def getFullState(self):
        return self._ensure_device_env().get_state(0)

    def restoreFullState(self, state):
        self._ensure_device_env().set_state(0, state)
        self._refresh()
