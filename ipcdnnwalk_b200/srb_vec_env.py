"""Batched SRBTrack environment: host-side mirror of the reference's gymnasium interface over the
CUDA step in libsrbtrack_b200.so.

Reference interface mirrored here (paths relative to the reference repository root):
  env/SRBEnv5_mocap_list_window.py:122-176   class SRBEnv(Env): observation Box(150), action Box(22)
  env/SRBEnv5_mocap_list_window.py:208,805   step(action) -> (obs, reward, done, truncated, info), reset()
  walk_SRBTrack2025_train_delta.py:94-100    make_vec_env(SRBEnv, n_envs=32, vec_env_cls=SubprocVecEnv)

`SRBVecEnv` is the seam replacement for the SubprocVecEnv of per-process SRBEnv copies: the SB3
VecEnv protocol (num_envs / reset / step_async / step_wait / step / get_attr / env_method / close)
over device tensors.  torch is used only for device memory and streams.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import OBS_DIM, ACT_DIM, RINFO_DIM, PLAN_DIM, DBG_DIM, VARIANT_TEST, VARIANT_TRAINING, VARIANT_TERRAIN, SrbConfig, check

REWARD_INFO_KEYS = ("reward_survive", "reward_contact", "reward_end_effector", "reward_delta_pos",
                    "reward_delta_ori", "reward_pos", "reward_ori", "reward_total")


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    return C.c_void_p(t.data_ptr())


def clip_arrays(clip):
    """Normalise one clip-table dict (keys of SRBEnv.loadMocap, :1187-1197) to contiguous float64 arrays."""
    pos = np.ascontiguousarray(np.asarray(clip["mocap_pos_com"], dtype=np.float64))
    n = pos.shape[0]
    ori = np.ascontiguousarray(np.asarray(clip["mocap_ori_com"], dtype=np.float64).reshape(n, 9))
    vel = np.ascontiguousarray(np.asarray(clip["mocap_vel_com"], dtype=np.float64).reshape(n, 6))
    quat = np.ascontiguousarray(np.asarray(clip["mocap_qpos"], dtype=np.float64)[:, 3:7])
    feet = np.ascontiguousarray(np.asarray(clip["feet_pos_list"], dtype=np.float64).reshape(n, 6))
    if "feet_yaw" in clip:
        yaw = np.asarray(clip["feet_yaw"], dtype=np.float64).reshape(n, 4)
    else:   # feet_ori_list: pure-yaw 3x3 matrices -> (cos, sin) = first column
        fo = np.asarray(clip["feet_ori_list"], dtype=np.float64).reshape(n, 2, 3, 3)
        yaw = np.stack([fo[:, 0, 0, 0], fo[:, 0, 1, 0], fo[:, 1, 0, 0], fo[:, 1, 1, 0]], axis=1)
    yaw = np.ascontiguousarray(yaw)
    return n, pos, ori, vel, quat, feet, yaw


class SRBVecEnv:
    """N SRBEnv instances stepped by one CUDA launch."""

    def __init__(self, clips, num_envs, variant="test", precision=64, lanes_per_env=0, auto_reset=True,
                 device=0, seed=0, terrain=None, initial_pos_planar=None):
        """variant "terrain" (env/SRBEnv5_mocap_list_terrain_window.py) needs `terrain` (ipcdnnwalk_b200.terrain.TerrainModel);
        `clips` are the FLAT reference tables, `initial_pos_planar` [N,2] shifts each environment's reference."""
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("SRBVecEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.num_envs = int(num_envs)
        self.variant = {"test": VARIANT_TEST, "training": VARIANT_TRAINING, "terrain": VARIANT_TERRAIN}[variant] if isinstance(variant, str) else int(variant)
        self.precision = int(precision)
        self._seed = int(seed)
        cfg = SrbConfig(self.num_envs, self.variant, self.precision, int(lanes_per_env), int(bool(auto_reset)), int(device), int(seed))
        self._h = C.c_void_p()
        check(self.lib.srb_create(C.byref(cfg), C.byref(self._h)))
        self.clips = []
        for i, clip in enumerate(clips):
            n, pos, ori, vel, quat, feet, yaw = clip_arrays(clip)
            check(self.lib.srb_upload_clip(self._h, i, n, int(clip["cycle_length"]), int(bool(clip.get("is_stand", False))),
                                           _ptr(pos), _ptr(ori), _ptr(vel), _ptr(quat), _ptr(feet), _ptr(yaw)))
            self.clips.append(clip)
        N = self.num_envs
        self.terrain = terrain
        if self.variant == VARIANT_TERRAIN:
            if terrain is None:
                raise ValueError("variant 'terrain' needs a TerrainModel")
            hf, n_hf, bx, n_box, bd, n_body = terrain.c_tables()
            check(self.lib.srb_upload_terrain(self._h, _ptr(terrain.plane_size), terrain.plane_geom, n_hf, C.cast(hf, C.c_void_p),
                                              n_box, C.cast(bx, C.c_void_p), n_body, C.cast(bd, C.c_void_p)))
            if initial_pos_planar is not None:
                self.set_planar_offsets(initial_pos_planar)
        self.obs = torch.zeros(N, OBS_DIM, dtype=torch.float32, device=self.device)
        self.rew = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.rinfo = torch.zeros(N, RINFO_DIM, dtype=torch.float32, device=self.device)
        self.term_obs = None
        self.contact_prob = None
        self.reset_plan = None
        self.dbg = None
        self._actions = None

    # ---- side channels ---------------------------------------------------------------------------
    def attach(self, contact_prob=False, term_obs=False, reset_plan=False, dbg=False):
        t, N = self.torch, self.num_envs
        if contact_prob and self.contact_prob is None:
            self.contact_prob = t.full((N, 2), float("nan"), dtype=t.float32, device=self.device)
        if term_obs and self.term_obs is None:
            self.term_obs = t.zeros(N, OBS_DIM, dtype=t.float32, device=self.device)
        if reset_plan and self.reset_plan is None:
            self.reset_plan = t.zeros(N, PLAN_DIM, dtype=t.float64, device=self.device)
        if dbg and self.dbg is None:
            self.dbg = t.zeros(N, DBG_DIM, dtype=t.float64, device=self.device)
        check(self.lib.srb_set_aux(self._h, _ptr(self.contact_prob), _ptr(self.term_obs), _ptr(self.reset_plan), _ptr(self.dbg)))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---- VecEnv protocol -------------------------------------------------------------------------
    def reset(self, env_idx=None, plan=None):
        """Reset all (or the given) environments.  `plan` [count, 9] = explicit draws
        (clip, start frame, vel noise 3, noise quaternion wxyz 4); None = device RNG."""
        if env_idx is None:
            idx, count = None, self.num_envs
        else:
            idx = np.ascontiguousarray(np.asarray(env_idx, dtype=np.int32))
            count = idx.shape[0]
        if plan is not None:
            plan = np.ascontiguousarray(np.asarray(plan, dtype=np.float64).reshape(count, PLAN_DIM))
        check(self.lib.srb_reset(self._h, _ptr(idx), count, _ptr(plan), _ptr(self.obs), self._stream()))
        return self.obs

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        a = self._actions
        t = self.torch
        if not isinstance(a, t.Tensor):
            a = t.as_tensor(np.asarray(a, dtype=np.float32), device=self.device)
        a = a.to(device=self.device, dtype=t.float32).contiguous()
        assert a.shape == (self.num_envs, ACT_DIM)
        check(self.lib.srb_step(self._h, _ptr(a), _ptr(self.obs), _ptr(self.rew), _ptr(self.done), _ptr(self.rinfo), self._stream()))
        return self.obs, self.rew, self.done, self.rinfo

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_host(self, actions=None):
        """Step through host buffers (pinned staging inside the library): numpy in, numpy views out."""
        if actions is not None:
            self.pinned()["act"][...] = actions
        check(self.lib.srb_step_host(self._h, None, None, None, None, None))
        p = self.pinned()
        return p["obs"], p["rew"], p["done"], p["rinfo"]

    def set_host_mode(self, mode):
        """step_host data movement: 1 = zero copy through mapped pinned memory (default), 0 = staged copies, 2 / 3 = split forms
        (look-ahead windows by copy engine while the step kernel runs; see include/srbtrack_b200.h)."""
        check(self.lib.srb_set_host_mode(self._h, int(mode)))

    def host_timeline(self, enable=True):
        """Device timeline (ms) of the last mode-2 step_host: window gather / observation copy / step kernel / patch launch done."""
        out = (C.c_float * 5)()
        check(self.lib.srb_host_timeline(self._h, int(bool(enable)), out))
        return dict(zip(("t0", "window_gather", "obs_copy", "step_kernel", "patch"), list(out)))

    def pinned(self):
        if getattr(self, "_pinned", None) is None:
            ptrs = [C.c_void_p() for _ in range(5)]
            check(self.lib.srb_pinned_buffers(self._h, *[C.byref(p) for p in ptrs]))
            N = self.num_envs

            def view(p, ctype, shape):
                n = int(np.prod(shape))
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).reshape(shape)
            self._pinned = dict(act=view(ptrs[0], C.c_float, (N, ACT_DIM)), obs=view(ptrs[1], C.c_float, (N, OBS_DIM)),
                                rew=view(ptrs[2], C.c_float, (N,)), done=view(ptrs[3], C.c_uint8, (N,)),
                                rinfo=view(ptrs[4], C.c_float, (N, RINFO_DIM)))
        return self._pinned

    def infos(self):
        """list of {'reward_info': {...}} like SRBEnv.step (:490)."""
        ri = self.rinfo.cpu().numpy()
        return [{"reward_info": dict(zip(REWARD_INFO_KEYS, map(float, row)))} for row in ri]

    def get_attr(self, name, indices=None):
        return [self._attr(name, i) for i in self._indices(indices)]

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else indices

    def set_attr(self, name, value, indices=None):
        """VecEnv.set_attr for the attributes a driver writes on SRBEnv: per-environment `impulse` (remaining pushes) and the
        batch-wide push schedule / reward weights (they are parameters of the one kernel launch)."""
        if name == "impulse":
            for i in self._indices(indices):
                st = self.get_state(i)
                st["ints"][5] = int(np.ceil(max(float(value), 0.0)))
                self.set_state(i, st)
        elif name in ("force", "impulse_interval", "impulse_duration"):
            self._push = getattr(self, "_push", dict(force=[700, 700, -350], impulse_interval=1e9, impulse_duration=24))
            self._push[name] = value
            self.set_impulse(self._push["impulse_interval"], self._push["impulse_duration"], self._push["force"])
        elif name == "reward_weights":
            self.set_reward_weights(value)
        else:
            raise AttributeError(f"SRBVecEnv.set_attr: '{name}' is not a settable attribute of the batched environment")

    def env_method(self, name, *args, indices=None, **kw):
        return [getattr(self, "_m_" + name)(i, *args, **kw) for i in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def seed(self, seed=None):
        """The device reset stream is fixed at construction (srb_config.seed); returns it per environment like VecEnv.seed."""
        return [self._seed for _ in range(self.num_envs)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None

    def close(self):
        if self._h:
            self.lib.srb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state access (getFullState / restoreFullState, :1005-1053) --------------------------------
    def state_dims(self):
        d = [C.c_int32() for _ in range(4)]
        check(self.lib.srb_state_dims(*[C.byref(x) for x in d]))
        return tuple(x.value for x in d)

    def get_state(self, i):
        nr, ni, hs, hw = self.state_dims()
        reals = np.zeros(nr); ints = np.zeros(ni, dtype=np.int32); hist = np.zeros((hs, hw))
        check(self.lib.srb_get_state(self._h, int(i), _ptr(reals), _ptr(ints), _ptr(hist)))
        return dict(reals=reals, ints=ints, hist=hist)

    def set_state(self, i, state):
        reals = np.ascontiguousarray(state["reals"], dtype=np.float64)
        ints = np.ascontiguousarray(state["ints"], dtype=np.int32)
        hist = np.ascontiguousarray(state["hist"], dtype=np.float64)
        check(self.lib.srb_set_state(self._h, int(i), _ptr(reals), _ptr(ints), _ptr(hist)))

    def named_state(self, i):
        """State of env i under the reference's attribute names."""
        s = self.get_state(i)
        r, k = s["reals"], s["ints"]
        return dict(qpos=r[0:7].copy(), qvel=r[7:13].copy(), xquat_lagged=r[13:17].copy(),
                    left_foot_contact_pos_global=r[17:20].copy(), right_foot_contact_pos_global=r[20:23].copy(),
                    left_foot_yaw_cs=r[23:25].copy(), right_foot_yaw_cs=r[25:27].copy(),
                    contactFilter_yaw=r[27:29].copy(), sds_x=r[29:33].copy(), sds_v=r[33:37].copy(),
                    episode_return=float(r[37]), qp_dual_warm_start=r[38:44].copy(), initial_pos_planar=r[44:46].copy(),
                    left_foot_in_contact=bool(k[0] & 1), right_foot_in_contact=bool(k[0] & 2),
                    filter_initialised=bool(k[0] & 4), is_stand=bool(k[0] & 8),
                    current_frame=int(k[1]), random_initial_frame=int(k[2]), clip=int(k[3]),
                    frame_counter=int(k[4]), impulse=int(k[5]), episodes=int(k[6]), hist=s["hist"])

    def _attr(self, name, i):
        return self.named_state(i)[name]

    def _m_getFullState(self, i):
        return self.get_state(i)

    def _m_restoreFullState(self, i, state):
        self.set_state(i, state)

    # ---- misc ------------------------------------------------------------------------------------
    def set_reward_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.shape == (9,)
        check(self.lib.srb_set_reward_weights(self._h, _ptr(w)))

    def set_impulse(self, interval, duration, force):
        """impulse_interval / impulse_duration / force (:182-206); `force` has up to 6 components (world force, world torque:
        `f[0:len(self.force)] = self.force`)."""
        f = np.ascontiguousarray(np.asarray(force, dtype=np.float64).reshape(-1))
        check(self.lib.srb_set_impulse(self._h, int(min(interval, 2**31 - 1)), int(duration), _ptr(f), int(f.shape[0])))

    def set_planar_offsets(self, xy):
        """initial_pos_planar per environment ([N,2]); used from each environment's next reset on."""
        xy = np.ascontiguousarray(np.asarray(xy, dtype=np.float64).reshape(self.num_envs, 2))
        check(self.lib.srb_set_planar_offsets(self._h, _ptr(xy)))

    def terrain_height(self, xy):
        """Batched terrain_height_ray: xy float64 [n,2] (device) -> (height [n], hit [n,4] = geom id, hfield cell c, r, triangle)."""
        t = self.torch
        xy = xy.to(device=self.device, dtype=t.float64).contiguous()
        n = xy.shape[0]
        h = t.empty(n, dtype=t.float64, device=self.device); hit = t.empty(n, 4, dtype=t.int32, device=self.device)
        check(self.lib.srb_terrain_height(self._h, n, _ptr(xy), _ptr(h), _ptr(hit), self._stream()))
        return h, hit

    def obs_stats(self, acc):
        """acc (float64 [1+2*150], device) += (count, sum, sum of squares) of the current observations."""
        check(self.lib.srb_obs_stats(self._h, _ptr(self.obs), _ptr(acc), self._stream()))

    def pack_obs_extra(self, out=None):
        """Batched `DeltaExtractor.pack_obs_extra(env)` (DeltaExtractor.py:1015-1036): [N, 29] float64 on the device."""
        t = self.torch
        if out is None:
            out = t.empty(self.num_envs, 29, dtype=t.float64, device=self.device)
        check(self.lib.srb_pack_obs_extra(self._h, _ptr(out), self._stream()))
        return out

    def rollout_stats(self, acc):
        """acc (float64 [1 + 2*150 + 3], device): observation statistics accumulated into acc[:301] (obs_stats) and the episode
        counters (sum of returns, sum of lengths, count) copied into acc[301:304] -- the vector of the path's one all-reduce."""
        self.obs_stats(acc)
        check(self.lib.srb_episode_stats_device(self._h, C.c_void_p(acc.data_ptr() + 301 * 8), self._stream()))
        return acc

    def connect_ranks(self, dist, rank, world):
        """Set up the peer-memory exchange of `rollout_stats_allreduce` between the ranks of one node: every rank's exchange block is
        exported as a CUDA IPC handle, the handles are all-gathered with torch.distributed (plumbing) and opened.  Call once per env."""
        t = self.torch
        buf = (C.c_ubyte * 64)()
        check(self.lib.srb_xchg_create(self._h, int(rank), int(world), C.cast(buf, C.c_void_p)))
        mine = t.tensor(list(bytes(buf)), dtype=t.uint8, device=self.device)
        if world > 1:
            parts = [t.zeros(64, dtype=t.uint8, device=self.device) for _ in range(world)]
            dist.all_gather(parts, mine)
            allh = t.stack(parts).cpu().numpy().tobytes()
        else:
            allh = bytes(buf)
        self._xchg_handles = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
        check(self.lib.srb_xchg_connect(self._h, C.cast(self._xchg_handles, C.c_void_p)))
        if world > 1:
            dist.barrier()                                   # every rank has opened every block before the first collective

    def rollout_stats_allreduce(self, acc):
        """acc (float64 [304], device) = all-rank sum of this step's observation statistics and the episode counters, computed and
        exchanged by ONE kernel over NVLink peer memory (collective: call on every rank).  Overwrites acc."""
        check(self.lib.srb_rollout_stats_allreduce(self._h, _ptr(self.obs), _ptr(acc), self._stream()))
        return acc

    def xchg_status(self):
        e = C.c_int32(0)
        check(self.lib.srb_xchg_status(self._h, C.byref(e)))
        return int(e.value)

    def episode_stats(self, clear=False):
        out = np.zeros(3)
        check(self.lib.srb_episode_stats(self._h, _ptr(out), int(clear)))
        return out

    def tracking_action(self, out, sigma=0.1, seed=1234, step=0):
        """out (float32 [N,22], device) = reference-tracking action + N(0, sigma^2) noise (bench / roll-out driver)."""
        check(self.lib.srb_tracking_action(self._h, float(sigma), int(seed), int(step), _ptr(out), self._stream()))
        return out

    def launch_count(self):
        return int(self.lib.srb_launch_count(self._h))


class SB3VecEnv(SRBVecEnv):
    """The SubprocVecEnv seam of the reference's trainer -- `make_vec_env(SRBEnv, n_envs=32, vec_env_cls=SubprocVecEnv)`
    (walk_SRBTrack2025_train_delta.py:94-100), wrapped by VecNormalize (gym_cdm2/train_gym.py:98) -- as stable-baselines3 sees a
    VecEnv: numpy in / numpy out, `reset() -> obs`, `step_wait() -> (obs, rewards, dones, infos)` with
    infos[i]["terminal_observation"] on the step that ended an episode (the environment is reset inside the step, as SubprocVecEnv's
    workers do), "reward_info" as SRBEnv.step returns it, and "episode" = {"r", "l"} as the Monitor wrapper of make_vec_env adds.
    Data moves through the library's pinned host buffers (srb_step_host)."""

    def __init__(self, clips, num_envs, variant="training", precision=32, **kw):
        kw["auto_reset"] = True
        super().__init__(clips, num_envs, variant=variant, precision=precision, **kw)
        self.attach(term_obs=True)
        self._ep_ret = np.zeros(self.num_envs); self._ep_len = np.zeros(self.num_envs, dtype=np.int64)
        self.reward_info = True
        try:                                              # pragma: no cover - gymnasium is not in the build image
            import gymnasium as gym
            self.action_space = gym.spaces.Box(low=-3.14, high=3.14, shape=(ACT_DIM,), dtype=np.float32)
            self.observation_space = gym.spaces.Box(low=-np.inf, high=np.inf, shape=(OBS_DIM,), dtype=np.float32)
        except Exception:                                 # noqa: BLE001
            self.action_space = self.observation_space = None

    def reset(self, env_idx=None, plan=None):
        obs = super().reset(env_idx=env_idx, plan=plan)
        self.torch.cuda.synchronize(self.device)
        self._ep_ret[:] = 0; self._ep_len[:] = 0
        return obs.cpu().numpy()

    def step_wait(self):
        a = np.ascontiguousarray(np.asarray(self._actions, dtype=np.float32).reshape(self.num_envs, ACT_DIM))
        obs, rew, done, rinfo = self.step_host(a)
        obs, rew, done = obs.copy(), rew.copy(), done.astype(bool)
        self._ep_ret += rew; self._ep_len += 1
        infos = [{"TimeLimit.truncated": False} for _ in range(self.num_envs)]
        if self.reward_info:
            for i, row in enumerate(rinfo):
                infos[i]["reward_info"] = dict(zip(REWARD_INFO_KEYS, map(float, row)))
        if done.any():
            idx = np.nonzero(done)[0]
            term = self.term_obs[self.torch.as_tensor(idx, device=self.device)].cpu().numpy()
            for k, i in enumerate(idx):
                infos[i]["terminal_observation"] = term[k]
                infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}
            self._ep_ret[idx] = 0; self._ep_len[idx] = 0
        return obs, rew, done, infos
