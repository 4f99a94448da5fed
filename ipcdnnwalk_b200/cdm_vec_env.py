"""Batched AdaptiveSRB environment (`deepmimiccdm-v1`, gym_cdm2): host-side mirror of the reference's gymnasium
interface over the CUDA step in libsrbtrack_b200.so (include/cdmenv_b200.h).

Reference interface mirrored here (paths relative to the reference root):
  gym_cdm2/luaEnvGymnasium.py:44-125     class LuaEnv(gymnasium.Env): observation Box(+-100, get_dim()[0]),
                                         action Box(+-1, get_dim()[1]); step -> (float32 state, reward, done,
                                         truncated, {}); reset -> (float32 state, {}); a terminal step resets the env
  gym_cdm2/deepmimiccdm-v1.lua:434,853-903   get_dim / step / reset Lua globals behind lua.F
  gym_cdm2/train_gym.py:165-176          the vectorised trainer loop (one process per env in the reference)

`CDMVecEnv` is the batched replacement (VecEnv protocol over device tensors); `LuaEnv` is the single-environment
proxy with the reference's class name and call signature.  torch is used only for device memory and streams.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import CDM_OBS_DIM, CDM_ACT_DIM, CDM_TAB_W, CDM_PLAN_DIM, CDM_DBG_DIM, CdmConfig, CdmSkeleton, check
from . import cdm_tables

AXIS_CODE = {"X": 0, "Y": 1, "Z": 2}
LEG_CHAIN = ("LeftHip", "LeftKnee", "LeftAnkle", "RightHip", "RightKnee", "RightAnkle")
TOE_LOCAL = (0.0, -0.01, 0.14)      # deepmimiccdm-v1.lua:984-985
HEEL_LOCAL = (0.0, -0.01, -0.14)


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    return C.c_void_p(t.data_ptr())


def skeleton_struct(skel):
    """The leg chains of a parsed .wrl skeleton (ipcdnnwalk_b200.vrml_skeleton.parse_wrl) as the C struct."""
    names = skel["names"]
    s = CdmSkeleton()
    for i in range(3):
        s.root_off[i] = float(skel["offset"][0][i])
        s.toe[i] = TOE_LOCAL[i]
        s.heel[i] = HEEL_LOCAL[i]
    for j, n in enumerate(LEG_CHAIN):
        b = names.index(n)
        exp_parent = 0 if j % 3 == 0 else names.index(LEG_CHAIN[j - 1])
        if skel["parent"][b] != exp_parent:
            raise ValueError(f"skeleton: {n} is not chained as Hips -> Hip -> Knee -> Ankle")
        for i in range(3):
            s.leg_off[j][i] = float(skel["offset"][b][i])
        ch = skel["channels"][b]
        s.leg_nchan[j] = len(ch)
        s.leg_axes[j] = sum(AXIS_CODE[c] << (2 * k) for k, c in enumerate(ch))
    return s


class CDMVecEnv:
    """N `deepmimiccdm-v1` environments stepped by one CUDA launch."""

    def __init__(self, tables, skel, num_envs, precision=64, lanes_per_env=0, auto_reset=True, device=0, seed=0, params=None, terrain=None):
        """`tables`: dict from cdm_tables.build_tables (or any table with the same keys); `skel`: parsed .wrl skeleton;
        `terrain`: dict(heights [rows, cols] in cm, size_x, size_z in cm, terrain_pos [3] in metres) selects the
        `walkterrain-v1` variant (deepmimiccdmterrain-v1.lua: 31-float observation)."""
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("CDMVecEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.num_envs = int(num_envs)
        self.precision = int(precision)
        cfg = CdmConfig(self.num_envs, self.precision, int(lanes_per_env), int(bool(auto_reset)), int(device), int(seed))
        self._h = C.c_void_p()
        check(self.lib.cdm_create(C.byref(cfg), C.byref(self._h)))
        self.terrain = None
        self.obs_dim = CDM_OBS_DIM
        if terrain is not None:
            from .tl_terrain import Terrain
            self.terrain = Terrain(terrain["heights"], float(terrain["size_x"]), float(terrain["size_z"]), True, device=device)
            tp = np.ascontiguousarray(terrain["terrain_pos"], dtype=np.float64)
            check(self.lib.cdm_set_terrain(self._h, self.terrain._h, _ptr(tp)))
            self.set_param("healthy_min", 0.7)              # spec/walkterrain-v1.lua: healthy_y_range = {0.7, 2.0}
            self.obs_dim = int(self.lib.cdm_obs_dim(self._h))
        for k, v in (params or {}).items():
            self.set_param(k, v)
        rows = np.ascontiguousarray(cdm_tables.pack_rows(tables), dtype=np.float64)
        assert rows.shape[1] == CDM_TAB_W
        self._skel = skeleton_struct(skel)
        check(self.lib.cdm_upload_table(self._h, rows.shape[0], int(tables["nf"]), _ptr(rows), C.byref(self._skel)))
        N = self.num_envs
        self.obs = torch.zeros(N, self.obs_dim, dtype=torch.float32, device=self.device)
        self.rew = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.term_obs = None
        self.reset_plan = None
        self.dbg = None
        self._actions = None
        self._pinned = None

    @classmethod
    def from_spec(cls, env_id, num_envs, tables=None, skel=None, motion_dir=None, gui=False, params=None, terrain=None, **kw):
        """One of the registered environments (gym_cdm2/__init__.py:9-11: `walk2-v1`, `walkterrain-v1`, `run2-v1`) with the
        parameter set its spec file selects (cdm_tables.SPECS).  Tables / skeleton are built from `motion_dir` (the reference's
        work/taesooLib/Resource/motion/MOB1) unless passed in; `gui=True` adds what the Lua main-loop viewer (test_cdm_deepmimic.lua, isMainloopInLua) changes --
        limited leg length, touchDownHeight -- the configuration the refTraj_*.dat roll-outs were recorded in; `terrain` as in the constructor (required for walkterrain-v1)."""
        spec = cdm_tables.SPECS[env_id]
        if skel is None or tables is None:
            import os
            from .vrml_skeleton import parse_wrl
            skel = skel or parse_wrl(os.path.join(motion_dir, "hanyang_lowdof_T.wrl"))
            tables = tables or cdm_tables.build_tables(cdm_tables.read_dof(os.path.join(motion_dir, spec["motion"])), skel, info=spec["info"])
        if spec["terrain"] and terrain is None:
            raise ValueError(env_id + " needs a terrain (heights, size_x, size_z, terrain_pos)")
        p = dict(spec["params"])
        if gui:
            p.update(spec["gui_params"])
        p.update(params or {})
        return cls(tables, skel, num_envs, params=p, terrain=terrain, **kw)

    # ---- parameters (spec-file overrides) ----------------------------------------------------------
    def set_param(self, name, value):
        check(self.lib.cdm_set_param(self._h, name.encode(), float(value)))

    def get_param(self, name):
        v = C.c_double()
        check(self.lib.cdm_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    def attach(self, term_obs=False, reset_plan=False, dbg=False):
        t, N = self.torch, self.num_envs
        if term_obs and self.term_obs is None:
            self.term_obs = t.zeros(N, self.obs_dim, dtype=t.float32, device=self.device)
        if reset_plan and self.reset_plan is None:
            self.reset_plan = t.zeros(N, CDM_PLAN_DIM, dtype=t.float64, device=self.device)
        if dbg and self.dbg is None:
            self.dbg = t.zeros(N, CDM_DBG_DIM, dtype=t.float64, device=self.device)
        check(self.lib.cdm_set_aux(self._h, _ptr(self.term_obs), _ptr(self.reset_plan), _ptr(self.dbg)))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---- VecEnv protocol ---------------------------------------------------------------------------
    def reset(self, env_idx=None, plan=None):
        """`plan` [count, 16] = explicit (already scaled) reset draws: quaternion noise 4, position noise 3, dq noise 6
        (angular, linear), extra linear-velocity noise 3; None = device RNG; zeros = noise-free start."""
        if env_idx is None:
            idx, count = None, self.num_envs
        else:
            idx = np.ascontiguousarray(np.asarray(env_idx, dtype=np.int32))
            count = idx.shape[0]
        if plan is not None:
            plan = np.ascontiguousarray(np.asarray(plan, dtype=np.float64).reshape(count, CDM_PLAN_DIM))
        check(self.lib.cdm_reset(self._h, _ptr(idx), count, _ptr(plan), _ptr(self.obs), self._stream()))
        return self.obs

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        a = self._actions
        t = self.torch
        if not isinstance(a, t.Tensor):
            a = t.as_tensor(np.asarray(a, dtype=np.float32), device=self.device)
        a = a.to(device=self.device, dtype=t.float32).contiguous()
        assert a.shape == (self.num_envs, CDM_ACT_DIM)
        check(self.lib.cdm_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.rew), _ptr(self.done), self._stream()))
        return self.obs, self.rew, self.done, None

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_host(self, actions=None):
        if actions is not None:
            self.pinned()["act"][...] = actions
        check(self.lib.cdm_step_host(self._h, None, None, None, None))
        p = self.pinned()
        return p["obs"], p["rew"], p["done"]

    def pinned(self):
        if self._pinned is None:
            ptrs = [C.c_void_p() for _ in range(4)]
            check(self.lib.cdm_pinned_buffers(self._h, *[C.byref(p) for p in ptrs]))
            N = self.num_envs

            def view(p, ctype, shape):
                n = int(np.prod(shape))
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).reshape(shape)
            self._pinned = dict(act=view(ptrs[0], C.c_float, (N, CDM_ACT_DIM)), obs=view(ptrs[1], C.c_float, (N, self.obs_dim)),
                                rew=view(ptrs[2], C.c_float, (N,)), done=view(ptrs[3], C.c_uint8, (N,)))
        return self._pinned

    def close(self):
        if self._h:
            self.lib.cdm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state access (RagdollSim:saveFullState / restoreFullState, deepmimiccdm-v1.lua:1730-1776) -----
    def state_dims(self):
        d = [C.c_int32() for _ in range(4)]
        check(self.lib.cdm_state_dims(*[C.byref(x) for x in d]))
        return tuple(x.value for x in d)

    def get_state(self, i):
        nr, ni, hs, hw = self.state_dims()
        reals = np.zeros(nr); ints = np.zeros(ni, dtype=np.int32); hist = np.zeros((hs, hw))
        check(self.lib.cdm_get_state(self._h, int(i), _ptr(reals), _ptr(ints), _ptr(hist)))
        return dict(reals=reals, ints=ints, hist=hist)

    def set_state(self, i, state):
        reals = np.ascontiguousarray(state["reals"], dtype=np.float64)
        ints = np.ascontiguousarray(state["ints"], dtype=np.int32)
        hist = np.ascontiguousarray(state["hist"], dtype=np.float64)
        check(self.lib.cdm_set_state(self._h, int(i), _ptr(reals), _ptr(ints), _ptr(hist)))

    def named_state(self, i):
        s = self.get_state(i)
        r, k = s["reals"], s["ints"]
        legs = []
        for l in range(2):
            b = 23 + 19 * l
            legs.append(dict(isSwing=bool(k[0] & (1 << l)), contactPos=r[b:b + 3].copy(), contactPreFilter=r[b + 3:b + 6].copy(),
                             contactFilter=r[b + 6:b + 9].copy(), particle_vel=r[b + 9:b + 12].copy(), yaw_filter=r[b + 12:b + 14].copy(),
                             lqrFilterQorigin=r[b + 14:b + 18].copy(), swingPhase=float(r[b + 18])))
        return dict(pos=r[0:3].copy(), quat=r[3:7].copy(), vel=r[7:10].copy(), angvel_body=r[10:13].copy(), iframe=float(r[13]),
                    globalTime=float(r[14]), target_Y=r[15:19].copy(), prev_com=r[19:22].copy(), episode_return=float(r[22]),
                    legs=legs, n_replay=int(k[1]), episodes=int(k[2]), steps=int(k[3]), replay=s["hist"])

    def episode_stats(self, clear=False):
        out = np.zeros(3)
        check(self.lib.cdm_episode_stats(self._h, _ptr(out), int(clear)))
        return out

    def launch_count(self):
        return int(self.lib.cdm_launch_count(self._h))


class LuaEnv:
    """Single-environment proxy with the reference's surface (gym_cdm2/luaEnvGymnasium.py:44-125): the Lua runtime is
    replaced by one lane group of the batched CUDA step.  `args` may carry `.tables` / `.skel`; otherwise they are built
    from `args.mot_file` / `args.skel_file`."""

    def __init__(self, args, render_mode=None, precision=64, device=0):
        from .vrml_skeleton import parse_wrl
        skel = getattr(args, "skel", None) or parse_wrl(args.skel_file)
        tables = getattr(args, "tables", None)
        env_id = getattr(args, "env_name", None) or "walk2-v1"       # gym_cdm2/test_gym.py: args.env_name selects spec/<env_name>.lua
        spec = cdm_tables.SPECS[env_id]
        if tables is None:
            tables = cdm_tables.build_tables(cdm_tables.read_dof(args.mot_file), skel, info=spec["info"])
        self.renderMode = render_mode
        self.args = args
        # driven from Python, isMainloopInLua is unset: always the training configuration (defaultRLsetting.lua:177)
        self._vec = CDMVecEnv.from_spec(env_id, 1, tables=tables, skel=skel, terrain=getattr(args, "terrain", None),
                                        precision=precision, auto_reset=False, device=device, seed=getattr(args, "seed", 0))
        self.observation_space_shape = (self._vec.obs_dim,)  # Box(-100, 100, (27,)), 31 on the terrain   :62-65
        self.action_space_shape = (CDM_ACT_DIM,)            # Box(-1, 1, (10,))
        self.actionSteps = 0
        self.step_counter = 0
        self.episodeTotalReward = 0

    def get_observation(self):
        return np.array([])

    def step(self, paction):
        a = np.asarray(paction, dtype=np.float32).reshape(1, CDM_ACT_DIM)
        obs, rew, done, _ = self._vec.step(a)
        state = obs[0].cpu().numpy().astype(np.float32)
        reward = float(rew[0])
        d = bool(done[0])
        self.step_counter += 1
        self.actionSteps += 1
        self.episodeTotalReward += reward
        if d:
            self.reset()
        return state, reward, d, False, {}

    def reset(self, seed=None, options=None):
        self.step_counter = 0
        self.actionSteps = 0
        self.episodeTotalReward = 0
        out = self._vec.reset()
        return out[0].cpu().numpy().astype(np.float32), {}

    def close(self):
        self._vec.close()
