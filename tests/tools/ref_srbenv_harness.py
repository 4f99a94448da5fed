"""Runs the REFERENCE'S OWN SRBEnv code -- env/SRBEnv5_mocap_list_window.py (test variant) and
env/SRBEnv5_mocap_list_window_training.py (training variant), imported unmodified from /root/reference -- in this container, with
the third-party back ends it cannot have here replaced by stated stand-ins:

  mujoco      not installed, not vendored, not version-pinned.  `FakeMujoco` below offers the dozen API entry points the environment
              calls (MjModel.from_xml_path, MjData, mj_name2id, mj_step1 / mj_step2 / mj_forward, mj_jac, mj_fullM, mj_applyFT) for the
              one-free-body model gym_trackSRB/Characters/SRB5.xml, implemented by oracle/mj_freebody.py (the restated MuJoCo semantics).
  clarabel    not installed.  `FakeClarabel.DefaultSolver` solves the QP it is handed EXACTLY (the problem is strictly convex, its optimum
              is unique, so any exact solver returns what an interior-point solver converges to): equality rows eliminated, NNLS on the rest.
  work.controlmodule.SDS   the reference's OWN class (work/controlmodule.py:911-1086), imported by tests/tools/ref_sds_harness.py; only
              libmainlib's matrixn / vectorn and two lua helpers under it are NumPy-backed stand-ins (test variant only; the training
              variant has no filter).
  gymnasium, transformations, work.settings / rendermodule, env.Policy_prior_posterior_MoE_window   import-time stubs (spaces.Box, Env,
              the `_last_contact_prob` side channel).
  SRBdata (humanoid2SRB needs MuJoCo's humanoid FK)   for the environments `SRBDataProcess` is answered by a holder of the clip tables
              of tests/golden/clip_*.npz, so loadMocap / __init__ fill SRB_data_list themselves; `reference_srbdata` runs the reference's
              own SRBdata.humanoid2SRB with the humanoid's mj_forward / mj_differentiatePos answered by oracle/srbdata.py's FK.

Everything else -- action decoding, contact logic and hysteresis, site placement, the reference's own get_site_jacobians /
apply_force_at_site / solve_qp_for_contact_forces problem assembly, force clamp, impulse schedule, history lists, get_obs_window,
_get_reward, _get_done, reset / set_state, getFullState -- is the reference's code, executed as it stands.  So a roll-out of this harness
pins oracle/srb_env.py's restatement of that code; what stays an assumption is exactly the MuJoCo / Clarabel semantics above.

Only usable where /root/reference is mounted: tests/tools/make_golden_refenv.py records roll-outs as tests/golden/ref_srbenv_rollouts.npz,
which travel to the GPU box."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
REF = "/root/reference"
SITES = ['left_foot_center', 'left_foot_lf', 'left_foot_rf', 'left_foot_lb', 'left_foot_rb',
         'right_foot_center', 'right_foot_lf', 'right_foot_rf', 'right_foot_lb', 'right_foot_rb']


def available():
    return os.path.isfile(os.path.join(REF, "env", "SRBEnv5_mocap_list_window.py"))


# ------------------------------------------------------------------------------------------------ mujoco stand-in
def _fake_mujoco():
    from oracle import mj_freebody as fb
    m = types.ModuleType("mujoco")

    class _Opt:
        pass

    class MjModel:
        def __init__(self):
            self._fb = fb.FreeBodyModel()
            self.nv = 6
            self.opt = _Opt()
            self.opt.timestep = self._fb.timestep
            self.geom_friction = np.array([self._fb.floor_friction.copy()])      # geom 0 = 'floor'
            self.site_bodyid = np.ones(10, dtype=int)
            self.site_rgba = np.zeros((10, 4))
            self._terrain = None; self._body_names = {0: "world", 1: "cdm"}; self._floor_geom = 0

        @staticmethod
        def from_xml_path(path):
            if path.endswith("humanoid_deepmimic_withpalm_Tpose.xml"):
                return HumanoidMjModel(path)
            mdl = MjModel()
            if path.endswith("SRB5_playground.xml"):
                # the terrain model: plane, two height fields, four pyramids of boxes -- geom / body numbering as MuJoCo compiles it
                # (world 0, cdm 1, pyramid_1..4 2..5; geoms: floor 0, height fields 1-2, the cdm box 3, boxes 4..), held by the
                # oracle's Terrain (tests/golden/terrain_playground.npz); mj_ray below answers from it
                from oracle.terrain import Terrain
                T = Terrain.load(os.path.join(ROOT, "tests", "golden", "terrain_playground.npz"))
                mdl._terrain = T
                ngeom = T.boxes[-1]["geom_id"] + 1
                mdl.geom_bodyid = np.zeros(ngeom, dtype=int); mdl.geom_size = np.zeros((ngeom, 3))
                mdl.geom_bodyid[3] = 1
                for b in T.boxes:
                    mdl.geom_bodyid[b["geom_id"]] = b["body_id"]; mdl.geom_size[b["geom_id"]] = b["size"]
                mdl.geom_friction = np.tile(mdl._fb.floor_friction, (ngeom, 1))
                mdl._body_names = {0: "world", 1: "cdm"}
                mdl._body_names.update({b["body_id"]: b["name"] for b in T.bodies})
                mdl._floor_geom = T.plane_geom
            else:
                assert path.endswith("SRB5.xml"), path
            return mdl

    class HumanoidMjModel:
        """the mocap humanoid of SRBdata.humanoid2SRB (testSRB_mujoco_original_viewer.py:871-1031): what mj_forward must deliver there is
        subtree_com[0], xmat of bodies, geom_xpos / geom_xmat of the foot geoms and site_xpos -- from the oracle's MJCF reader / FK"""

        def __init__(self, path):
            from oracle.srbdata import HumanoidModel
            self._h = HumanoidModel(path)
            self._humanoid = True
            self.nq = self._h.nq; self.nv = self._h.nq - 1
            self._geoms = [g[0] for b in self._h.bodies for g in b["geoms"]]
            self._sites = [st[0] for b in self._h.bodies for st in b["sites"]]

    class HumanoidMjData:
        def __init__(self, model):
            nb = len(model._h.bodies) + 1
            self.qpos = np.zeros(model.nq); self.qvel = np.zeros(model.nv)
            self.xpos = np.zeros((nb, 3)); self.xmat = np.zeros((nb, 9)); self.subtree_com = np.zeros((nb, 3))
            self.geom_xpos = np.zeros((len(model._geoms), 3)); self.geom_xmat = np.zeros((len(model._geoms), 9))
            self.site_xpos = np.zeros((len(model._sites), 3))

    class MjData:
        def __new__(cls, model):
            if getattr(model, "_humanoid", False):
                return HumanoidMjData(model)
            return object.__new__(cls)

        def __init__(self, model):
            d = fb.FreeBodyData(model._fb)
            self.qpos = d.qpos.copy(); self.qvel = d.qvel.copy()
            self.xpos = np.zeros((6, 3)); self.xmat = np.tile(np.eye(3).reshape(9), (6, 1))
            self.site_xpos = np.zeros((10, 3)); self.site_xmat = np.zeros((10, 9))
            self.qfrc_applied = np.zeros(6); self.xfrc_applied = np.zeros((6, 6))
            self.qfrc_bias = np.zeros(6); self.qfrc_passive = np.zeros(6); self.qacc = np.zeros(6)
            self.qM = np.zeros(21); self.time = 0.0
            self.subtree_com = np.zeros((6, 3))

    def _to(model, data):
        d = fb.FreeBodyData(model._fb)
        d.qpos = np.array(data.qpos, dtype=float).copy(); d.qvel = np.array(data.qvel, dtype=float).copy()
        d.xpos = data.xpos[1].copy(); d.xmat = data.xmat[1].copy()
        d.site_xpos = data.site_xpos.copy(); d.site_xmat = data.site_xmat.copy()
        d.qfrc_applied = data.qfrc_applied.copy(); d.xfrc_applied = data.xfrc_applied[1].copy()
        d.qfrc_bias = data.qfrc_bias.copy(); d.qfrc_passive = data.qfrc_passive.copy(); d.qacc = data.qacc.copy(); d.time = data.time
        return d

    def _back(d, data, kin):
        data.qpos = d.qpos.copy(); data.qvel = d.qvel.copy()
        if kin:
            data.xpos[1] = d.xpos; data.xmat[1] = d.xmat
            data.subtree_com[0] = d.xpos; data.subtree_com[1] = d.xpos      # one body, CoM at its origin
            data.site_xpos[:] = d.site_xpos; data.site_xmat[:] = d.site_xmat
        data.qfrc_bias[:] = d.qfrc_bias; data.qfrc_passive[:] = d.qfrc_passive; data.qacc[:] = d.qacc; data.time = d.time

    def mj_step1(model, data):
        d = _to(model, data); fb.mj_step1(model._fb, d); _back(d, data, True)

    def mj_step2(model, data):
        d = _to(model, data); fb.mj_step2(model._fb, d); _back(d, data, False)

    def mj_forward(model, data):
        if getattr(model, "_humanoid", False):
            fk = model._h.forward(np.asarray(data.qpos, dtype=float))
            data.xpos[1:] = fk["xpos"]; data.xmat[1:] = fk["xmat"].reshape(-1, 9)
            data.subtree_com[0] = fk["com"]
            for i, n in enumerate(model._geoms):
                data.geom_xpos[i] = fk["geoms"][n][0]; data.geom_xmat[i] = fk["geoms"][n][1].reshape(9)
            for i, n in enumerate(model._sites):
                data.site_xpos[i] = fk["sites"][n]
            return
        d = _to(model, data); fb.mj_forward(model._fb, d); _back(d, data, True)

    def mj_differentiatePos(model, qvel, dt, qpos1, qpos2):
        """mj_differentiatePos: free joint = position difference + quaternion difference as a rotation vector, hinges = angle difference"""
        from oracle.srbdata import differentiate_pos_free
        q1, q2 = np.asarray(qpos1, dtype=float), np.asarray(qpos2, dtype=float)
        lin, ang = differentiate_pos_free(dt, q1, q2)
        qvel[0:3] = lin; qvel[3:6] = ang
        qvel[6:] = (q2[7:] - q1[7:]) / dt

    def mj_jac(model, data, jacp, jacr, point, body):
        assert body == 1
        d = _to(model, data)
        if jacp is not None:
            jacp[:] = fb.mj_jacp(model._fb, d, np.asarray(point, dtype=float))
        if jacr is not None:
            jacr[:] = 0.0
            jacr[:, 3:6] = d.xmat.reshape(3, 3)

    def mj_fullM(model, dst, qM):
        dst[:] = np.diag(model._fb.fullM_diag())

    def mj_applyFT(model, data, force, torque, point, body, qfrc_target):
        assert body == 1
        d = _to(model, data)
        fb.mj_applyFT(model._fb, d, np.asarray(force, dtype=float).reshape(3), np.asarray(torque, dtype=float).reshape(3),
                      np.asarray(point, dtype=float), qfrc_target)

    class mjtObj:
        mjOBJ_BODY, mjOBJ_GEOM, mjOBJ_SITE = 1, 5, 6

    def mj_name2id(model, typ, name):
        if getattr(model, "_humanoid", False):
            if typ == mjtObj.mjOBJ_BODY:
                return model._h.body_id[name] + 1
            return (model._geoms if typ == mjtObj.mjOBJ_GEOM else model._sites).index(name)
        if typ == mjtObj.mjOBJ_BODY:
            return {v: k for k, v in model._body_names.items()}[name]
        if typ == mjtObj.mjOBJ_SITE:
            return SITES.index(name)
        if typ == mjtObj.mjOBJ_GEOM:
            return {"floor": model._floor_geom}[name]
        raise KeyError((typ, name))

    def mj_id2name(model, typ, idx):
        assert typ == mjtObj.mjOBJ_BODY
        return model._body_names.get(int(idx))               # None past the last body, as mujoco

    def mj_ray(model, data, pnt, vec, geomgroup=None, flg_static=1, bodyexclude=-1, geomid=None):
        """vertical ray of terrain_height_ray (testSRB_mujoco_original_viewer_terrain.py:741-752) over the group-0 geoms"""
        assert vec[0] == 0 and vec[1] == 0 and vec[2] == -1 and bodyexclude == 1
        h, gid = model._terrain.height_ray(float(pnt[0]), float(pnt[1]))
        geomid[0] = -1 if h is None else gid
        return -1.0 if h is None else float(pnt[2]) - h

    for k, v in dict(MjModel=MjModel, MjData=MjData, mj_step1=mj_step1, mj_step2=mj_step2, mj_forward=mj_forward, mj_jac=mj_jac,
                     mj_fullM=mj_fullM, mj_applyFT=mj_applyFT, mjtObj=mjtObj, mj_name2id=mj_name2id, mj_id2name=mj_id2name, mj_ray=mj_ray, mj_differentiatePos=mj_differentiatePos).items():
        setattr(m, k, v)
    return m


# ------------------------------------------------------------------------------------------------ clarabel stand-in
def _fake_clarabel():
    from scipy.optimize import nnls
    m = types.ModuleType("clarabel")

    class _Cone:
        def __init__(self, n):
            self.dim = n

    class ZeroConeT(_Cone):
        pass

    class NonnegativeConeT(_Cone):
        pass

    class DefaultSettings:
        verbose = False

    class _Sol:
        pass

    class DefaultSolver:
        """min 1/2 x'Px + q'x  s.t.  A x + s = b, s in (zero cone of n_eq rows) x (nonnegative cone): the reference hands over
        x = [qdd (n_eq), lambda], equality rows -[M, -C] x = -(-bias), inequality rows -lambda <= 0."""

        def __init__(self, P, q, A, b, cones, settings):
            P = np.asarray(P.todense(), dtype=float)
            self.P = P + P.T - np.diag(np.diag(P))                     # the reference passes the upper triangle
            self.q = np.asarray(q, dtype=float)
            self.A = np.asarray(A.todense(), dtype=float); self.b = np.asarray(b, dtype=float)
            assert isinstance(cones[0], ZeroConeT) and isinstance(cones[1], NonnegativeConeT)
            self.neq, self.nin = cones[0].dim, cones[1].dim

        def solve(self):
            neq, nin = self.neq, self.nin
            n = self.P.shape[0]
            nu = n - nin
            E, e = self.A[:neq], self.b[:neq]                            # E x = e
            Gi, gi = self.A[neq:], self.b[neq:]                          # Gi x <= gi  (= -lambda <= 0)
            assert np.array_equal(Gi[:, :nu], np.zeros((nin, nu))) and np.array_equal(Gi[:, nu:], -np.eye(nin)) and not gi.any()
            Eu, El = E[:, :nu], E[:, nu:]
            Eui = np.linalg.inv(Eu)
            # u = Eui (e - El lam)  ->  strictly convex QP in lam >= 0:  1/2 lam' Q lam + c' lam
            T = -Eui @ El; u0 = Eui @ e
            Puu, Pul, Pll = self.P[:nu, :nu], self.P[:nu, nu:], self.P[nu:, nu:]
            Q = T.T @ Puu @ T + T.T @ Pul + Pul.T @ T + Pll
            c = T.T @ (Puu @ u0 + self.q[:nu]) + Pul.T @ u0 + self.q[nu:]
            sol = _Sol()
            if nin:
                L = np.linalg.cholesky(Q).T                               # Q = L'L
                lam, _ = nnls(L, -np.linalg.solve(L.T, c), maxiter=100 * nin + 100)
                # polish on the final active set (nnls stops at its own tolerance)
                S = lam > 0
                if S.any():
                    lam = np.zeros(nin); lam[S] = np.linalg.solve(Q[np.ix_(S, S)], -c[S])
            else:
                lam = np.zeros(0)
            sol.x = np.concatenate([u0 + T @ lam, lam])
            return sol

    for k, v in dict(ZeroConeT=ZeroConeT, NonnegativeConeT=NonnegativeConeT, DefaultSettings=DefaultSettings, DefaultSolver=DefaultSolver).items():
        setattr(m, k, v)
    return m


# ------------------------------------------------------------------------------------------------ the rest of the import-time stubs
def _fake_control():
    """work.controlmodule as the environment sees it: the reference's OWN SDS / NonlinearController classes (work/controlmodule.py),
    imported by tests/tools/ref_sds_harness.py over NumPy-backed matrixn / vectorn objects (libmainlib cannot be built here)."""
    import ref_sds_harness as sh
    c = types.ModuleType("work.controlmodule")
    c.SDS = sh.load().SDS
    return c


def install():
    """Register the stand-ins in sys.modules and put the reference on sys.path.  Idempotent."""
    if "mujoco" in sys.modules and getattr(sys.modules["mujoco"], "_is_fake", False):
        return
    mj = _fake_mujoco(); mj._is_fake = True
    mj.__path__ = []
    sys.modules["mujoco"] = mj
    mv = types.ModuleType("mujoco.viewer"); mj.viewer = mv
    sys.modules["mujoco.viewer"] = mv
    sys.modules["clarabel"] = _fake_clarabel()
    tr = types.ModuleType("transformations")                   # third-party transformations.py: the two functions SRBdata.get_frame_data calls
    from oracle import srbdata as _sd

    def quaternion_matrix(q):
        M = np.eye(4); M[:3, :3] = _sd.quaternion_matrix3(q)
        return M

    tr.quaternion_matrix = quaternion_matrix
    tr.quaternion_from_matrix = lambda M: _sd.quaternion_from_matrix3(np.asarray(M)[:3, :3])
    import math
    _NEXT = [1, 2, 0, 1]
    _AXES = {"rxyz": (2, 1, 0, 1)}                                  # first axis, parity, repetition, frame

    def euler_from_matrix(matrix, axes="sxyz"):
        """transformations.py (Gohlke), the published algorithm, for the axes string the reference uses"""
        first, parity, rep, frame = _AXES[axes]
        i = first; j = _NEXT[i + parity]; k = _NEXT[i - parity + 1]
        M = np.array(matrix, dtype=np.float64)[:3, :3]
        cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
        if cy > np.finfo(float).eps * 4.0:
            ax = math.atan2(M[k, j], M[k, k]); ay = math.atan2(-M[k, i], cy); az = math.atan2(M[j, i], M[i, i])
        else:
            ax = math.atan2(-M[j, k], M[j, j]); ay = math.atan2(-M[k, i], cy); az = 0.0
        if parity:
            ax, ay, az = -ax, -ay, -az
        if frame:
            ax, az = az, ax
        return ax, ay, az

    def euler_matrix(ai, aj, ak, axes="sxyz"):
        first, parity, rep, frame = _AXES[axes]
        i = first; j = _NEXT[i + parity]; k = _NEXT[i - parity + 1]
        if frame:
            ai, ak = ak, ai
        if parity:
            ai, aj, ak = -ai, -aj, -ak
        si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
        ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
        cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
        M = np.identity(4)
        M[i, i] = cj * ck; M[i, j] = sj * sc - cs; M[i, k] = sj * cc + ss
        M[j, i] = cj * sk; M[j, j] = sj * ss + cc; M[j, k] = sj * cs - sc
        M[k, i] = -sj; M[k, j] = cj * si; M[k, k] = cj * ci
        return M

    tr.euler_from_matrix = euler_from_matrix; tr.euler_matrix = euler_matrix
    sys.modules["transformations"] = tr
    g = types.ModuleType("gymnasium"); sp = types.ModuleType("gymnasium.spaces")

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Env:
        def __init__(self, *a, **k):
            pass

    sp.Box = Box; g.spaces = sp; g.Env = Env
    sys.modules["gymnasium"] = g; sys.modules["gymnasium.spaces"] = sp
    w = types.ModuleType("work"); w.__path__ = []
    st = types.ModuleType("work.settings"); st.control = _fake_control()
    rm = types.ModuleType("work.rendermodule"); rm.output = lambda *a, **k: None
    w.settings = st; w.rendermodule = rm
    sys.modules["work"] = w; sys.modules["work.settings"] = st; sys.modules["work.rendermodule"] = rm; sys.modules["work.controlmodule"] = st.control
    for name in ("Policy_prior_posterior_MoE_window", "Policy_prior_posterior_MoE_window_training", "Policy_prior_posterior_MoE"):
        pm = types.ModuleType("env." + name); pm._last_contact_prob = [None, None]
        sys.modules["env." + name] = pm
    if REF not in sys.path:
        sys.path.insert(0, REF)


def reference_module(variant):
    """variant "test": imported as it stands.  variant "training": env/SRBEnv5_mocap_list_window_training.py does not parse as shipped
    (line 10 reads `from .testSRB_mujoco_original_viewer import * SRBDataProcess=SRBdata`, two statements fused on one line); the
    source is executed with that one line split into the two statements the sibling file has -- nothing else is touched."""
    install()
    import importlib
    if variant == "test":
        return importlib.import_module("env.SRBEnv5_mocap_list_window")
    if variant == "terrain":
        return importlib.import_module("env.SRBEnv5_mocap_list_terrain_window")
    name = "env.SRBEnv5_mocap_list_window_training"
    if name in sys.modules:
        return sys.modules[name]
    importlib.import_module("env.testSRB_mujoco_original_viewer")
    path = os.path.join(REF, "env", "SRBEnv5_mocap_list_window_training.py")
    src = open(path, encoding="utf-8").read()
    bad = "from .testSRB_mujoco_original_viewer import * SRBDataProcess=SRBdata"
    assert src.count(bad) == 1
    src = src.replace(bad, "from .testSRB_mujoco_original_viewer import *\nSRBDataProcess=SRBdata")
    mod = types.ModuleType(name); mod.__file__ = path; mod.__package__ = "env"
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


STAND_PATH = "motiondata_mujoco_refined/stand1.txt"        # the training variant recognises the stand clip by this name (:644)


def make_env(variant, clips, paths=None):
    """The reference's SRBEnv over the committed clip tables (helpers.fresh_clips()).  `SRBDataProcess` (= SRBdata, whose
    humanoid2SRB needs MuJoCo's humanoid FK) is answered by a holder of those tables, so loadMocap / __init__ run as they stand."""
    mod = reference_module(variant)
    paths = paths or [STAND_PATH if c.get("is_stand") else f"clip{i}.txt" for i, c in enumerate(clips)]
    by_path = dict(zip(paths, clips))

    class ClipTables:
        def __init__(self, mocap_path):
            self._c = by_path[mocap_path]
            self.dt = self._c["mocap_dt"]
            self.mocap_data = np.zeros((1, 2))                 # the terrain env adds initial_pos_planar to columns 0:2 before humanoid2SRB
            self._fill(self._c)

        def _fill(self, c):
            self.cycle_length = c["cycle_length"]; self.mocap_com_humanoid = c["mocap_qpos"]
            self.mocap_ori_com = c["mocap_ori_com"]; self.mocap_vel_com = c["mocap_vel_com"]
            self.feet_pos_list = c["feet_pos_list"]; self.feet_ori_list = c["feet_ori_list"]
            if "feet_pos_list_original" in c:
                self.feet_pos_list_original = c["feet_pos_list_original"]; self.mocap_height_com_terrain = c["mocap_height_com_terrain"]

        def humanoid2SRB(self, initial_frame=0, last_frame=None):
            if variant == "terrain":                           # tables lifted onto the terrain under the shifted clip (:1043-1065)
                from oracle.srb_env import lift_clip_to_terrain
                from oracle.terrain import Terrain
                T = Terrain.load(os.path.join(ROOT, "tests", "golden", "terrain_playground.npz"))
                self._fill(lift_clip_to_terrain(self._c, T, tuple(self.mocap_data[0, :2])))

    mod.SRBDataProcess = ClipTables
    cwd = os.getcwd()
    os.chdir(REF)                                          # relative model path 'gym_trackSRB/Characters/SRB5.xml'
    try:
        env = mod.SRBEnv(paths)
        for i in range(len(paths)):
            if hasattr(env, "loadMocap") and variant != "terrain":
                env.loadMocap(i)                           # test variant: lazy loading (:1171-1197); terrain: at reset, with the planar shift
    finally:
        os.chdir(cwd)
    return env, mod


def reset_explicit(env, mod, clip_id, frame, vel_noise=None, noise_quat=None, planar=None):
    """SRBEnv.reset (:805-845; terrain :528-571) with the random draws supplied: binds the clip as reset() does, then the reference's
    own set_state with np.random.uniform / random_quaternion_perturbation answering `vel_noise` / `noise_quat` (None = testing_mode).
    `planar` (terrain variant): initial_pos_planar, handed to the reference's loadMocap as reset(mocap_id, initial_pos_planar) does."""
    env.contactFilter = [None] * 4
    if planar is not None:
        env.loadMocap(clip_id, np.asarray(planar, dtype=float))
    d = env.SRB_data_list[clip_id]
    for k, v in d.items():                                    # SRBdata, mocap_dt, cycle_length, mocap_qpos, ... exactly the names reset() binds
        setattr(env, k, v)
    env.is_stand = env.mocap_path_list[clip_id] == STAND_PATH            # training variant's reset (:643-645); unused by the test variant
    env.random_initial_frame = int(frame); env.current_frame = 0
    env.testing_mode = vel_noise is None
    if vel_noise is None:
        env.set_state(int(frame))
    else:
        real_uniform, real_rqp = np.random.uniform, mod.random_quaternion_perturbation
        np.random.uniform = lambda low, high, size: np.asarray(vel_noise, dtype=float).copy()
        mod.random_quaternion_perturbation = lambda scale=5e-2: np.asarray(noise_quat, dtype=float).copy()
        try:
            env.set_state(int(frame))
        finally:
            np.random.uniform, mod.random_quaternion_perturbation = real_uniform, real_rqp
    env.testing_mode = False
    return env._get_obs()


def reference_srbdata(mocap_path):
    """The reference's own SRBdata(mocap_path).humanoid2SRB() (testSRB_mujoco_original_viewer.py:765-1031) on the shipped clip files:
    MuJoCo's humanoid FK / subtree CoM / mj_differentiatePos and transformations.py answered by oracle/srbdata.py's restatements, the
    reference's own frame stitching, yaw-aligned cycle continuation, velocity, foot-orientation and contact processing as it stands."""
    install()
    import importlib
    mod = importlib.import_module("env.testSRB_mujoco_original_viewer")
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        s = mod.SRBdata(mocap_path)
        s.humanoid2SRB()
    finally:
        os.chdir(cwd)
    return s
