"""ORACLE (test infrastructure only; nothing in the product imports this).

NumPy fp64 restatement of the AdaptiveSRB environment `deepmimiccdm-v1` as configured by gym_cdm2/spec/walk2-v1.lua in
TRAINING mode (no GUI: defaultRLsetting.lua:177-190 -> touchDownHeight 0, unlimitedLeglen true) -- SURVEY 8a rows
a-16 ... a-20.  Paths relative to the reference root:

  gym_cdm2/deepmimiccdm-v1.lua:441-517   startSwing / startSpprt / initContactFilter
                             :519-599   filterContact (mode 'sw': pre-filter + LQR particle + yaw SDS)
                             :602-754   calcContactPos (swing / support state machine, stride early touch-down)
                             :961-1012  getJointStateFromSimulator / getJointStateFromRefTree (markers)
                             :1014-1574 RagdollSim:step     :1578-1722 reset     :1779-1859 _get_obs
                             :1865-1903 getFilteredFootPos  :1907-1936 getDesiredFinalVel   :1946-2170 calcRefDif2
  gym_cdm2/RagdollSim.lua:128-176        replay buffer (pushSimPose / sampleSimPose), sampleTargetPose
  gym_cdm2/defaultRLsetting.lua:117-127,239-435,501-513   comparison frames, phase, rtt tables, ref tree, advanceTime
  gym_cdm2/module/QPservo_v3.lua:138-321 (Trbdl repacking), :689-737 (_calcDesiredAcceleration), :939-979
      (_buildProblem: 5-direction friction basis), :1058-1463 (_QPsolve), :776-938 (stepSimul), :1720-1866 (frameMove)
  src/QP_controller/eiquadprog.hpp (Goldfarb-Idnani) via quadprog.cpp:37-70: strictly convex QP with a unique
      optimum; here it is solved EXACTLY by eliminating (qdd, tau) through the equality constraints (M is diagonal
      for the single box) and running NNLS + active-set polish on the ridge problem
          min_{lam>=0} |A lam - y|^2 + (w_lam / w_ddq) |lam|^2,  A = M^-1 J^T V,  y = a_des + M^-1 b.
  work/taesooLib/PhysicsLib/rbdl/DynamicsSimulator_Trbdl_penalty.cpp:211-250 (integrateQuater), :701-763 (integrate:
      semi-implicit Euler, q <- normalize(q + qdot(q, w) dt)), :933-973 (mass matrix / bias), :1264-1484 (link data
      packing), :1910-1956 (getDQ / setDQ)
  work/taesooLib/Resource/scripts/control/SDRE.lua:37-146,470-620 (SDS, MassParticle3D), PhysicsLib/SDRE.cpp:100-245
      (continuous ARE -> LQR gain; closed form for the 2x2 spring-damper, see `lqr_gain`)

taesooLib / Lua cannot run here.  PARITY PARTLY PINNED against the reference's own recorded outputs:
  * gym_cdm2/spec/refTraj_walk2-v1.dat / refTraj_run2-v1.dat are roll-outs RECORDED BY THE REFERENCE SIMULATOR
    (deepmimiccdm-v1.lua:1095-1146: root pose, filtered feet, dq, foot yaws, raw contact targets, swing flags / phases
    at every RL step).  tests/test_oracle_cdm_refpins_cpu.py replays them through this file's own methods: the swing-foot
    contact filter + yaw filter + getFilteredFootPos (1572 swing rows, < 4e-6 m, 1e-8 rad), the swing / touch-down
    state machine against the restated rtt tables (3596 rows), and the contact-free sub-steps of the running clip
    (505 flight rows: gravity, Euler equations, quaternion integration, < 1e-7).
  * work/_LQR_CACHE_2.DAT holds one LQR gain computed by the reference (A, B, Q, R, K): a known-answer vector for
    `lqr_gain`.
  * trained_models/ppo/walk2-v1.pt, the policy the reference trained on this environment, walks in this restatement
    at the speed of the reference's own recorded roll-out (tests/test_cdm_policy.py).
  NOT pinned by reference outputs (no recorded intermediate exists): the contact QP and the reward; they are checked
  analytically in tests/test_oracle_cdm_cpu.py (KKT of the un-eliminated QP as _QPsolve states it, second solver) and
  by the committed golden roll-out.

Q15 (found through the recorded roll-outs): the shipped build sets NoLapack (src/CMakeLists.txt:41,
work/taesooLib/Samples/Common.cmake:3) => EXCLUDE_CLAPACK, and Physics.LQR then returns the gain cached in
work/_LQR_CACHE_2.DAT -- K = (9999.00005, 1095.39024412), computed for M=60, b=1e-3, k=1, Q=1e8 -- for EVERY 2x2
system (work/taesooLib/PhysicsLib/SDRE.cpp:136-154).  The contact filter (M=30, Q=10^7) therefore runs with that gain;
the recorded roll-outs reproduce only with it.  `filter_K` below; `lqr_gain` is the real ARE solution kept for the
known-answer test.

Stated simplifications (all below 1e-12 relative, i.e. far inside the 1e-5 gate): the JOINT_VALUE/JOINT_VELOCITY
round trips of QPsim:frameMove through an identity reference frame (refCoord, flat ground) are skipped; randomised
target-direction changes (`change_target_direction`) have no effect on the state in the reference because
`target_Y_UI` is never nil (:1066-1070) and are not modelled.
"""
import math
import numpy as np
from scipy.optimize import nnls
from . import tl_math as tl
from .tl_math import Y, QI

RL_STEP = 1.0 / 60.0
TIMESTEP_QP = 1.0 / 120.0
FRAME_RATE = 30.0
MASS = 60.0
INERTIA = np.array([9.683, 2.056, 10.209])          # defaultRLsetting.lua:20 (box principal moments, Y up)
GRAV = 9.8                                          # QPservo_v3.lua:1586
TOE_LOCAL = np.array([0.0, -0.01, 0.14])            # deepmimiccdm-v1.lua:984-985
HEEL_LOCAL = np.array([0.0, -0.01, -0.14])
PARAMS = dict(k_p=120.0, k_d=35.0, w_ddq=10000.0, w_tau=0.0001, w_lam=0.0001, acc_clamp=400.0, max_leg_len=1.10,
              max_foot_vel=3.0, logQ=7.0, stride_thr=0.45, max_contact_len=0.9, min_swing_phase=0.25, max_contact_y=1.1,
              healthy_min=0.8, healthy_max=2.0, fall_angle=math.radians(20), timing_mod_coef=0.4, facing_weight=1.5,
              ef_scale=1.0, limit_length=300.0, max_turning=0.5, filter_end_phase=0.4, inv_fric=1.0,
              filter_K=(9999.00005, 1095.39024412),
              unlimited_leglen=True, contact_thr=0.0, touch_down_height=0.0)        # training mode (defaultRLsetting.lua:183-187); contactThr: motionInfos.walk4 (:82)

# spec/run2-v1.lua over motionInfos.run4 (gym_cdm2/module/motionInfo.lua:133-183): overrides of PARAMS in training mode; defaults
# where both are silent: k_p_ID 340 (defaultRLsetting.lua:43), EFdif_scale 1/8 (deepmimiccdm-v1.lua:2108), maxContactLen 2.0 (:1246),
# minSwingPhase 0.4 (:615).  RUN2_GUI: what a rendered run adds (the spec's own unlimitedLeglen=false, touchDownHeight 0.7 of
# defaultRLsetting.lua:22) -- the configuration refTraj_run2-v1.dat was recorded in.
RUN2_PARAMS = dict(k_p=340.0, ef_scale=0.125, max_leg_len=1.5, contact_thr=0.1, stride_thr=None, max_contact_len=2.0, min_swing_phase=0.4,
                   healthy_min=0.8, healthy_max=1.1, facing_weight=10.0, max_turning=0.4)
RUN2_GUI = dict(unlimited_leglen=False, touch_down_height=0.7)


def lqr_gain(M, b, k, q, qd=0.0):
    """Physics.LQR for A=[[0,1],[-k/M,-b/M]], B=[0,1/M]^T, Q=diag(q,qd), R=1: closed-form stabilising ARE solution."""
    a0, a1, be = k / M, b / M, 1.0 / M
    p2 = (-a0 + math.sqrt(a0 * a0 + be * be * q)) / (be * be)
    p3 = (-a1 + math.sqrt(a1 * a1 + be * be * (2.0 * p2 + qd))) / (be * be)
    return be * p2, be * p3


class SDS:
    """SDRE.lua:470-560 with NonlinearController.oneStepRaw (D, C, G branch; the constant G = k/M is subtracted, :134)."""
    def __init__(self, M, b, k, q, dt):
        self.M, self.b, self.k, self.dt = M, b, k, dt
        self.K = lqr_gain(M, b, k, q)
        self.x = np.zeros(2); self.xd = np.zeros(2)

    def single_step(self):
        cf = self.K[0] * self.xd[0] + self.K[1] * self.xd[1] - (self.K[0] * self.x[0] + self.K[1] * self.x[1])
        dd = cf * (1.0 / self.M) - self.k / self.M - (self.b / self.M) * self.x[1]
        self.x[1] += dd * self.dt
        self.x[0] += self.x[1] * self.dt


class Leg:
    pass


def contact_basis(inv_fric=1.0):
    """QPservo:_buildProblem :956-968."""
    ang = math.atan(1.0 / inv_fric)
    X, Z = np.array([1.0, 0, 0]), np.array([0, 0, 1.0])
    return [Y.copy(), tl.qrot(tl.qaxis(X, -ang), Y), tl.qrot(tl.qaxis(Z, -ang), Y), tl.qrot(tl.qaxis(X, ang), Y),
            tl.qrot(tl.qaxis(Z, ang), Y)]


def solve_contact_qp(Q, p, a_des, bias, contacts, P):
    """_QPsolve with (qdd, tau) eliminated.  Returns qdd (body [dW, dV]) and lam [nc,5].  dq order: [w_body, v_body]."""
    Mdiag = np.array([INERTIA[0], INERTIA[1], INERTIA[2], MASS, MASS, MASS])
    if len(contacts) == 0:
        return -bias / Mdiag, np.zeros((0, 5))
    normals = contact_basis(P["inv_fric"])
    qi = tl.qinv(Q)
    cols = []
    for c in contacts:
        for n in normals:
            cols.append(np.concatenate([tl.qrot(qi, np.cross(c - p, n)), tl.qrot(qi, n)]))   # J^T V column (body frame)
    JtV = np.array(cols).T
    A = JtV / Mdiag[:, None]
    y = a_des + bias / Mdiag
    w = P["w_lam"] / P["w_ddq"]
    nl = A.shape[1]
    lam, _ = nnls(np.vstack([A, math.sqrt(w) * np.eye(nl)]), np.concatenate([y, np.zeros(nl)]), maxiter=200 * nl)
    S = lam > 0
    if S.any():                                          # polish on the support (removes the NNLS tolerance)
        As = A[:, S]
        ls = np.linalg.solve(As.T @ As + w * np.eye(As.shape[1]), As.T @ y)
        out = np.zeros_like(lam); out[S] = ls
        grad = A.T @ (A @ out - y) + w * out
        if (ls > 0).all() and (grad[~S] >= -1e-9 * (1 + np.abs(grad).max())).all():
            lam = out
    qdd = (JtV @ lam - bias) / Mdiag
    return qdd, lam.reshape(-1, 5)


TERRAIN_SAMPLES = ((0.0, 0.0, 0.5), (0.5, 0.0, 0.5), (-0.5, 0.0, 0.5), (1.0, 0.0, 0.5))     # deepmimiccdmterrain-v1.lua:3-8


class OracleCDMEnv:
    def __init__(self, tab, skel, params=None, terrain=None, terrain_pos=None):
        """terrain: oracle.tl_terrain.OracleTerrain (gym_cdm2 `walkterrain-v1`, deepmimiccdmterrain-v1.lua: 31-float observation,
        terrain-shifted targets, projectToGround everywhere); None = flat ground (`walk2-v1`)."""
        self.tab, self.skel = tab, skel
        self.P = dict(PARAMS)
        self.terrain, self.terrain_pos = terrain, (np.asarray(terrain_pos, float) if terrain is not None else None)
        if terrain is not None:
            self.P["healthy_min"] = 0.7                              # spec/walkterrain-v1.lua: healthy_y_range = {0.7, 2.0}
        if params:
            self.P.update(params)
        self.F = tab["cdm"].shape[0]
        self.nf = int(tab["nf"])
        names = skel["names"]
        self.iRA, self.iLA = names.index("RightAnkle"), names.index("LeftAnkle")
        self.K_filter = tuple(self.P["filter_K"])                # Q15: the cached gain, not lqr_gain(30, 1e-3, 1, 10^logQ)
        self.debug = {}

    def G(self, v):
        """RagdollSim:projectToGround / OBJloader.Terrain:getTerrainPos (deepmimiccdm-v1.lua:415-424,2171-2177): ground height under v."""
        if self.terrain is None:
            return 0.0
        return self.terrain.height((v[0] - self.terrain_pos[0]) * 100.0, (v[2] - self.terrain_pos[2]) * 100.0)[0] / 100.0 + self.terrain_pos[1]

    # ---------------------------------------------------------------- model (defaultRLsetting.lua)
    def _ti(self, arr, ileg):
        """integer / vector tables indexed with the real-valued frame: (int) truncation (Q8)."""
        return arr[ileg - 1][int(self.iframe)]

    def rtt_down(self, ileg): return int(self._ti(self.tab["rtt_down"], ileg))
    def rtt_off(self, ileg): return int(self._ti(self.tab["rtt_off"], ileg))

    def swing_phase(self, ileg):
        return tl.clamp_map(self.rtt_down(ileg), 0.0, float(self._ti(self.tab["swing_dur"], ileg)), 0.0, 1.0)

    def feet_offset(self, ileg):
        return np.array(self._ti(self.tab["feet_off"], ileg), dtype=float)

    def get_phase(self):
        """model:getPhase (defaultRLsetting.lua:239-283) -- mutates iframe when it runs off the stitched table."""
        tabf = self.tab["iframe_tab"]
        nf = self.nf
        if self.iframe > nf * 2:
            self.iframe -= nf
        if self.iframe >= len(tabf) - 1:
            self.iframe -= nf
        iframe = self.iframe
        ifl = int(math.floor(iframe))
        if tabf[ifl + 1] - tabf[ifl] < 0:
            iframe = tl.sop_map(iframe, ifl, ifl + 1, tabf[ifl], tabf[ifl + 1] + nf)
            if iframe > nf:
                iframe -= nf
        return tl.sample_row(tabf.reshape(-1, 1), iframe)[0] / nf

    def ref_tree(self, delta_frame=0.0):
        """model:setRefTree -> (CDM reference pose 7, ankle marker positions R toe, R heel, L toe, L heel)."""
        f = min(self.iframe + delta_frame, self.F - 1)
        cdm = tl.motiondof_sample(self.tab["cdm"], f)
        pose = tl.motiondof_sample(self.tab["fullbody"], f)
        from .bone_fk import bone_forward_kinematics
        rot, tr = bone_forward_kinematics(self.skel, pose)
        mk = []
        for b in (self.iRA, self.iLA):
            for lp in (TOE_LOCAL, HEEL_LOCAL):
                g = tl.qrot(rot[b], lp) + tr[b]
                g[1] = self.G(g)                               # projectToGround
                mk.append(g)
        return cdm, mk

    # ---------------------------------------------------------------- simulator access
    def get_dq(self):
        return np.concatenate([tl.qrot(self.Q, self.wb), self.v])          # [w_world, v_world]

    def sample_sim_pose(self, delta_frame):
        n = len(self.replay)
        sf = delta_frame * 2 + n - 1
        if sf > n - 1:
            sf = n - 1
        return tl.sample_row(np.array(self.replay), max(0, sf))

    # ---------------------------------------------------------------- contact filter
    def start_swing(self, li):
        li.isSwing = True
        li.swingPhase = 0.0
        li.prevContact = li.contactFilter.copy()
        li.contactPreFilter = li.prevContact.copy()
        for k in range(3):
            li.ball[k].x[:] = (li.prevContact[k], 0.0)
        li.filtQ.x[:] = 0.0
        li.contactFilterQ = QI.copy()

    def start_spprt(self, li, rotY):
        li.isSwing = False
        li.spprtPhase = 0.0
        li.rotY = rotY.copy()
        li.contactFilter = np.array([li.ball[k].x[0] for k in range(3)])
        li.dotContactFilter = np.zeros(3)
        li.filtQorigin = li.rotY.copy()
        li.contactFilterQ = QI.copy()

    def filter_contact_sw(self, li, rotY, gv):
        w = li.swingPhase if li.isSwing else self.P["filter_end_phase"]
        vel = li.contactPos - li.contactPreFilter
        alpha2 = tl.map_sin(w, 0.0, self.P["filter_end_phase"], 0.0, 1.0)
        vel = tl.smooth_clamp_vec3(vel, self.P["max_foot_vel"] * alpha2)
        li.contactPreFilter = li.contactPreFilter + vel
        v = gv.copy(); v[1] = 0.0
        for k in range(3):
            li.ball[k].K = self.K_filter
            li.ball[k].xd[:] = (li.contactPreFilter[k], v[k])
            li.ball[k].single_step(); li.ball[k].single_step()
        li.contactFilter = np.array([li.ball[k].x[0] for k in range(3)])
        li.dotContactFilter = np.array([li.ball[k].x[1] for k in range(3)])
        delta = tl.qalign(tl.qdifference(li.filtQorigin, rotY), QI)
        ang = tl.rotation_angle_about_axis(delta, Y)
        li.filtQ.K = self.K_filter
        li.filtQ.xd[:] = (ang, 0.0)
        li.filtQ.single_step(); li.filtQ.single_step()
        li.contactFilterQ = tl.qaxis(Y, li.filtQ.x[0])

    def update_swing_foot(self, li, contact, Q, p, rotY, gv):
        local = tl.qrot(tl.qinv(Q), contact - p)
        lim = self.P["max_leg_len"] + -0.01
        ln = math.sqrt(local.dot(local))
        if ln > lim:
            local = tl.vnormalize(local) * lim
        li.localContactPos = local
        li.contactPos = tl.qrot(Q, local) + p
        self.filter_contact_sw(li, rotY, gv)

    def calc_contact_pos(self, ileg, Q, p, gv, rotY, act2):
        li = self.leg[ileg]
        is_swing = li.isSwing
        if not is_swing:
            li.spprtPhase += RL_STEP / 0.73
            if self.rtt_off(ileg) <= 0:
                self.start_swing(li)
                self.update_swing_foot(li, li.contactPos, Q, p, rotY, gv)
        if is_swing:
            off = self.feet_offset(ileg)
            desired = p + tl.qrot(rotY, np.array([act2[0], 0.0, act2[1] + 0.0]) + off)
            desired[1] = self.G(desired)
            li.swingPhase += RL_STEP
            li.contactPos = desired.copy()
            self.update_swing_foot(li, li.contactPos, Q, p, rotY, gv)
            rt = self.rtt_down(ileg)
            com_height = p[1]
            if self.terrain is not None:                        # :686-692 height relative to the terrain between body and landing spot
                com_height -= tl.map_cos(self.swing_phase(ileg), 0.0, 1.0, self.G(p), self.G(li.contactPos))
            if rt <= 0:
                # :702 -- with a leg-length limit the foot lands only once the body is low enough (reference CDM height + contactThr)
                if self.P["unlimited_leglen"] or com_height < tl.sample_pose(self.tab["cdm"], self.iframe)[1] + self.P["contact_thr"]:
                    li.contactPos[1] = self.G(li.contactPos)    # projectToGround
                    self.start_spprt(li, rotY)
                else:
                    li.lateTouchDown = True
            elif com_height < self.P["touch_down_height"] and li.swingPhase > self.P["min_swing_phase"]:     # touchDownHeight: 0 in training
                li.earlyTouchDown = rt
            else:
                oli = self.leg[ileg % 2 + 1]
                st = self.P["stride_thr"]
                if st is not None:                              # `if st then` (:718): run4 has no strideThr
                    a = li.contactFilter.copy(); a[1] = 0.0
                    b = oli.contactFilter.copy(); b[1] = 0.0
                    d = a - (b + tl.qrot(rotY, np.array([0.0, 0.0, st])))
                    if (not oli.isSwing) and math.sqrt(d.dot(d)) > st and li.swingPhase > self.P["min_swing_phase"]:
                        li.earlyTouchDown = rt

    def filtered_foot_pos(self, gq, ilimb):
        li = self.leg[ilimb]
        pos = li.contactFilter.copy()
        ori = li.filtQorigin
        if li.isSwing:
            ori = tl.qmult(li.filtQorigin, li.contactFilterQ)
            oli = self.leg[ilimb % 2 + 1]
            if not oli.isSwing:
                side = tl.qrot(tl.rotation_y(gq), np.array([1.0, 0.0, 0.0]))
                depth = (pos - oli.contactFilter).dot(side)
                thr = 0.05
                weight = math.pow(math.sin(self.swing_phase(ilimb) * math.pi), 0.5)
                if li.isLeft:
                    if depth < thr:
                        pos = pos - side * ((depth - thr) * weight)
                elif -depth < thr:
                    pos = pos + side * ((-depth - thr) * weight)
        pos[1] = self.G(pos)
        return pos, ori

    def sim_markers(self, rotY):
        mk = []
        for ileg in (1, 2):
            pos, ori = self.filtered_foot_pos(rotY, ileg)
            for s in (1.0, -1.0):
                g = pos + tl.qrot(ori, np.array([0.0, 0.0, s * 0.12]))
                g[1] = self.G(g)
                mk.append(g)
        return self.p.copy(), mk

    # ---------------------------------------------------------------- reset
    def reset(self, noise=None):
        """noise = dict(q[4], pos[3], dq[6], vel3[3]) of the uniform draws of :1625-1647 (None -> zeros)."""
        z = dict(q=np.zeros(4), pos=np.zeros(3), dq=np.zeros(6), vel3=np.zeros(3))
        if noise:
            z.update({k: np.asarray(v, float) for k, v in noise.items()})
        self.change_count = 0
        self.target_Y = QI.copy()
        self.global_time = 0.0
        self.iframe = 0.0                                        # model.start (randomRestart is off)
        st = self.tab["cdm"][0].copy()
        vel = self.tab["dcdm"][0].copy()
        start_swing = [not bool(self.tab["con"][0][0]), not bool(self.tab["con"][1][0])]
        _, ref_mk = self.ref_tree()
        delta_root = np.array([st[0], 0.0, st[2]])
        st[0] = 0.0; st[2] = 0.0
        st[1] += self.G(np.zeros(3))                              # :1595-1596
        st[3:7] += z["q"]
        st[0:3] += z["pos"]
        st[3:7] = tl.qnormalize(st[3:7])
        self.p, self.Q = st[0:3].copy(), st[3:7].copy()
        dq = np.concatenate([tl.qrot(self.Q, vel[4:7]), tl.qrot(self.Q, vel[0:3])])      # dposeToDQ
        dq += z["dq"]
        dq[3:6] += z["vel3"]
        self.v = dq[3:6].copy()
        self.wb = tl.qrot(tl.qinv(self.Q), dq[0:3])                                  # setDQ
        self.replay = [st.copy()]
        rotY = tl.rotation_y(self.Q)
        self.leg = {}
        for i in (1, 2):
            li = Leg()
            li.rotY = rotY.copy()
            li.contactPos = ref_mk[2 * i - 2] * 0.5 + ref_mk[2 * i - 1] * 0.5
            li.contactPos[1] = 0.0
            li.contactPos = li.contactPos - delta_root
            li.prevContact = li.contactPos.copy()
            q = 10.0 ** self.P["logQ"]
            li.ball = [SDS(30.0, 0.001, 1.0, q, RL_STEP / 2) for _ in range(3)]
            for k in range(3):
                li.ball[k].x[:] = (li.contactPos[k], 0.0)
            li.contactPreFilter = li.contactPos.copy()
            li.contactFilter = li.contactPos.copy()
            li.dotContactFilter = np.zeros(3)
            li.filtQ = SDS(30.0, 0.001, 1.0, q, RL_STEP / 2)
            li.filtQorigin = li.rotY.copy()
            li.contactFilterQ = QI.copy()
            li.footOffset = np.array([-0.07, 0, 0]) if i == 1 else np.array([0.07, 0, 0])
            li.isLeft = (i == 2)
            li.earlyTouchDown = None
            li.lateTouchDown = None
            li.swingPhase = 0.0
            li.spprtPhase = 0.0
            li.isSwing = False
            if start_swing[i - 1]:
                self.start_swing(li)
            else:
                self.start_spprt(li, rotY)
            self.leg[i] = li
        self.sim_prev = self.sim_markers(rotY)
        return self.get_obs()

    # ---------------------------------------------------------------- observation
    def get_obs(self):
        """_get_obs (:1779-1859); with a terrain the version of deepmimiccdmterrain-v1.lua:25-103 (4 height samples appended)."""
        P = self.P
        rotY = tl.rotation_y(self.Q)
        dq = self.get_dq()
        ref_t = np.array([self.p[0], self.G(self.p), self.p[2]])       # getRefCoord: root projected to the ground
        offq = tl.qalign(tl.qmult(tl.qinv(rotY), self.Q), QI)
        iry = tl.qinv(rotY)
        o = np.zeros(27 if self.terrain is None else 27 + len(TERRAIN_SAMPLES))
        o[0] = self.p[1] - ref_t[1]
        o[1:5] = offq
        o[5:8] = tl.qrot(iry, dq[0:3])
        o[8:11] = tl.qrot(iry, dq[3:6])
        c = 11
        for i in (1, 2):
            li = self.leg[i]
            cpos = li.contactFilter.copy()
            if self.terrain is not None:
                cpos[1] = self.G(cpos)                            # terrain version: the filtered contact point is projected first
            o[c:c + 3] = tl.qrot(iry, cpos - ref_t)
            vv = tl.qrot(iry, li.dotContactFilter)
            o[c + 3], o[c + 4] = vv[0], vv[2]
            if li.isSwing:
                o[c + 5] = 0.0
            else:
                dqq = tl.qalign(tl.qmult(iry, li.filtQorigin), QI)
                o[c + 5] = tl.rotation_angle_about_axis(dqq, Y)
            c += 6
        phase = self.get_phase()
        o[c] = tl.map_cos(phase, 0.0, 0.25, 1.0, 0.0)
        o[c + 1] = tl.map_sin(phase, 0.0, 0.25, 0.0, 1.0)
        tv = tl.qrot(tl.qmult(iry, self.target_Y), np.array([0.0, 0.0, 1.0]))
        o[c + 2], o[c + 3] = tv[0], tv[2]
        if self.terrain is not None:
            for k, v in enumerate(TERRAIN_SAMPLES):
                pos = tl.qrot(rotY, np.array(v)) + ref_t
                o[c + 4 + k] = self.G(pos) - ref_t[1]
        return o

    # ---------------------------------------------------------------- step
    def desired_final_rot(self, rotY, v):
        """getDesiredFinalVel with target_Y_UI = identity (:1907-1936)."""
        v = max(0.0, v)
        front = tl.qrot(rotY, np.array([0.0, 0.0, 1.0]))
        tgt = np.array([0.0, 0.0, v])
        n = math.sqrt(tgt.dot(tgt))
        if n == 0.0:
            ang = 0.0                                            # degenerate setAxisRotation -> identity twist
        else:
            ang = math.atan2(tgt[0], tgt[2]) - math.atan2(front[0], front[2])
            ang = math.atan2(math.sin(ang), math.cos(ang))
        knots = [(0.0, math.radians(180)), (1.0, math.radians(180)), (2.5, math.radians(120)), (5.5, math.radians(50)),
                 (10.5, math.radians(30))]
        mx = knots[-1][1]
        for i in range(len(knots) - 1):
            t1 = knots[i + 1][0]
            if t1 >= v - 1e-10:
                mx = tl.sop_map(tl.sop_map(v, knots[i][0], t1, 0.0, 1.0), 0.0, 1.0, knots[i][1], knots[i + 1][1])
                break
        mx *= self.P["max_turning"]
        ang = min(max(ang, -mx), mx)
        return tl.qmult(rotY, tl.qaxis(Y, ang))

    def frame_move(self, contacts, theta_d, dtheta_d, daction, terrain_y=0.0):
        """QPsim:frameMove with niter = 2.  refCoord is a pure vertical translation by `terrain_y` (:1229); the QP and the
        integrator are translation invariant, so the only place the reference frame shows is the contact gate on the body height."""
        P = self.P
        dtheta_d = dtheta_d.copy()
        self.debug["qdd"] = []; self.debug["lam"] = []; self.debug["ades"] = []
        for _ in range(2):
            dtheta_d[0:3] += daction[0:3]
            dtheta_d[4:7] += daction[3:6]
            # _calcDesiredAcceleration
            q2 = tl.qalign(theta_d[3:7].copy(), self.Q)
            w_tw, v_tw = tl.twist(self.Q, self.p, q2, theta_d[0:3], 1.0)
            vb = tl.qrot(tl.qinv(self.Q), self.v)                     # JOINT_VELOCITY: body-frame linear velocity
            lin = P["k_p"] * v_tw + P["k_d"] * (dtheta_d[0:3] - vb)
            ang = P["k_p"] * w_tw + P["k_d"] * (dtheta_d[4:7] - self.wb)
            a_des = np.array([tl.smooth_clamp_sym(x, P["acc_clamp"]) for x in np.concatenate([ang, lin])])
            # mass matrix / bias in the body frame (calcMassMatrix3): b = [w x I w, R^T(m g) + m w x v_b]
            bias = np.concatenate([np.cross(self.wb, INERTIA * self.wb),
                                   tl.qrot(tl.qinv(self.Q), np.array([0.0, MASS * GRAV, 0.0])) + MASS * np.cross(self.wb, vb)])
            use = contacts if (self.p[1] - terrain_y > 0 and self.p[1] - terrain_y < P["max_contact_y"]) else []
            qdd, lam = solve_contact_qp(self.Q, self.p, a_des, bias, use, P)
            self.debug["qdd"].append(qdd.copy()); self.debug["lam"].append(lam.copy()); self.debug["ades"].append(a_des.copy())
            # stepKinematic: QDDot = [R dV + R (w x v_b), dW]; semi-implicit Euler
            lin_acc = tl.qrot(self.Q, qdd[3:6] + np.cross(self.wb, vb))
            h = TIMESTEP_QP
            self.v = self.v + h * lin_acc
            self.wb = self.wb + h * qdd[0:3]
            self.p = self.p + h * self.v
            w = self.wb
            q = self.Q                                               # qd = 0.5 * q (x) (0, w_body)
            qd = 0.5 * np.array([-q[1] * w[0] - q[2] * w[1] - q[3] * w[2],
                                 q[0] * w[0] + q[2] * w[2] - q[3] * w[1],
                                 q[0] * w[1] - q[1] * w[2] + q[3] * w[0],
                                 q[0] * w[2] + q[1] * w[1] - q[2] * w[0]])
            self.Q = tl.qnormalize(q + qd * h)

    def calc_ref_dif(self, rotY0):
        P = self.P
        pd_pose = np.concatenate([self.p, self.Q])
        delta_frame = FRAME_RATE * RL_STEP
        prev_delta = delta_frame + -2.0
        sample = self.sample_sim_pose(prev_delta)
        self.replay.append(pd_pose.copy())
        sim_prev = self.sim_prev
        cdm0, mk0 = self.ref_tree(0.0)
        ref_prev = (cdm0[0:3].copy(), mk0)
        cdm1, mk1 = self.ref_tree(delta_frame)
        self.ref_y = cdm1[1]                                         # loader_ref stays at iframe + 0.5 for is_healthy
        ref_pose = cdm1
        sim_rotY = tl.rotation_y(self.Q)
        sim_curr = self.sim_markers(sim_rotY)
        ref_curr = (cdm1[0:3].copy(), mk1)
        sample_ref = tl.sample_row(self.tab["cdm"], max(self.iframe + prev_delta, 0.0))
        prev_sim_ori = tl.rotation_y(sample[3:7])
        prev_ref_ori = tl.rotation_y(sample_ref[3:7])
        d_sim = tl.qmult(tl.qinv(prev_sim_ori), self.Q)
        d_ref = tl.qmult(tl.qinv(prev_ref_ori), ref_pose[3:7])
        pose_dif = tl.rotation_angle(tl.qmult(tl.qinv(d_sim), d_ref)) * 5
        v_sim = tl.qrot(tl.qinv(prev_sim_ori), self.p - sample[0:3])
        v_ref = tl.qrot(tl.qinv(prev_ref_ori), ref_pose[0:3] - sample_ref[0:3])
        dd = v_sim - v_ref
        pose_dif += math.sqrt(dd.dot(dd))
        rc_m = (prev_ref_ori, np.array([sample_ref[0], 0.0, sample_ref[2]]))
        rc_s = (prev_sim_ori, np.array([sample[0], self.G(sample), sample[2]]))

        def loc(rc, x): return tl.qrot(tl.qinv(rc[0]), x - rc[1])
        def locd(rc, x): return tl.qrot(tl.qinv(rc[0]), x)
        def dist(a, b): d = a - b; return math.sqrt(d.dot(d))
        def dist2(a, b): d = a - b; d = np.array([d[0], 0.0, d[2]]); return math.sqrt(d.dot(d))
        ef = 0.0
        for ileg in (1, 2):
            com_d = dist(loc(rc_m, ref_curr[0]), loc(rc_s, sim_curr[0]))
            if not self.leg[ileg].isSwing:
                ef += dist2(loc(rc_m, ref_curr[1][2 * ileg - 2]), loc(rc_s, sim_curr[1][2 * ileg - 2])) * 3
                ef += dist2(loc(rc_m, ref_curr[1][2 * ileg - 1]), loc(rc_s, sim_curr[1][2 * ileg - 1])) * 3
                ef += com_d * 3
            else:
                ef += com_d
        if not self.leg[1].isSwing and not self.leg[2].isSwing:
            ef += dist2(locd(rc_m, ref_curr[1][0] - ref_curr[1][2]), locd(rc_s, sim_curr[1][0] - sim_curr[1][2])) * 3
            ef += dist2(locd(rc_m, ref_curr[1][1] - ref_curr[1][3]), locd(rc_s, sim_curr[1][1] - sim_curr[1][3])) * 3
        pose_dif += ef * P["ef_scale"]
        pose_dif *= 10
        ref_rotY = tl.rotation_y(ref_pose[3:7])
        target_delta = tl.qmult(tl.qinv(sim_rotY), self.target_Y)
        v1 = tl.qrot(tl.qinv(sim_rotY), sim_curr[0] - sim_prev[0])
        v2 = tl.qrot(tl.qmult(target_delta, tl.qinv(ref_rotY)), ref_curr[0] - ref_prev[0])
        dxz = dist2(v1, v2)
        dy = abs(v1[1] - v2[1]) * 1.0
        com_lvdif = math.sqrt(dxz * dxz + dy * dy) / RL_STEP
        self.sim_prev = sim_curr
        return pose_dif / 10, com_lvdif, dist(ref_curr[0], ref_prev[0]) / RL_STEP

    def step(self, action):
        P = self.P
        action = np.asarray(action, dtype=float) * 0.2
        if self.iframe >= self.F:
            return self.get_obs(), 0.0, True
        Q, p = self.Q.copy(), self.p.copy()
        rotY = tl.rotation_y(Q)
        dq = self.get_dq()
        gv = dq[3:6]
        self.target_Y = self.desired_final_rot(rotY, gv.dot(tl.qrot(rotY, np.array([0.0, 0.0, 1.0]))))
        contact_pos = {}
        for i in (1, 2):
            self.calc_contact_pos(i, Q, p, gv, rotY, action[(i - 1) * 2:i * 2])
            pos, ori = self.filtered_foot_pos(Q, i)
            contact_pos[i] = [pos + tl.qrot(ori, np.array([0.0, 0.0, 0.12])), pos + tl.qrot(ori, np.array([0.0, 0.0, -0.12]))]
        mspc = FRAME_RATE * RL_STEP
        self.debug["late"] = bool(self.leg[1].lateTouchDown or self.leg[2].lateTouchDown)
        if self.leg[1].lateTouchDown or self.leg[2].lateTouchDown:
            for i in (1, 2):
                self.leg[i].lateTouchDown = None; self.leg[i].earlyTouchDown = None
            self.iframe += -mspc
        else:
            e1, e2 = self.leg[1].earlyTouchDown, self.leg[2].earlyTouchDown
            if e1 is not None and e2 is not None:
                if e1 < e2: e2 = None
                else: e1 = None
            etd = None
            for ileg, e in ((1, e1), (2, e2)):
                if e is not None:
                    if e > mspc:
                        e = mspc
                    else:
                        self.start_spprt(self.leg[ileg], rotY)          # enforceTouchDown (:767-774)
                        self.leg[ileg].contactPos[1] = self.G(self.leg[ileg].contactPos)
                    etd = e
                    break
            self.leg[1].earlyTouchDown = None; self.leg[2].earlyTouchDown = None
            if etd is not None:
                self.iframe += etd
        cdm_action = action[4:10]
        # sampleTargetPose (RagdollSim.lua:145-176)
        theta_d = tl.sample_pose(self.tab["cdm"], self.iframe)
        dtheta_d = tl.sample_row(self.tab["cdm_d"], self.iframe)
        s = self.sample_sim_pose(-2.0)
        Tq, Tp = tl.rotation_y(s[3:7]), s[0:3]
        r = tl.sample_row(self.tab["cdm"], max(self.iframe + -2.0, 0.0))
        Rq, Rp = tl.rotation_y(r[3:7]), r[0:3]
        # delta = T_ref * T^-1, projected to the ground plane; theta_d <- delta^-1 * theta_d
        dq_ = tl.qmult(Rq, tl.qinv(Tq))
        dp_ = Rp - tl.qrot(dq_, Tp)
        dp_[1] = 0.0
        dq_ = tl.rotation_y(dq_)
        iq = tl.qinv(dq_)
        theta_d[0:3] = tl.qrot(iq, theta_d[0:3] - dp_)
        theta_d[3:7] = tl.qmult(iq, theta_d[3:7])
        terrain_y = 0.0
        if self.terrain is not None:                              # :1205-1238
            terrain_y = self.G(theta_d[0:3])
            theta_d[1] += terrain_y
            nxt = self.G(p + gv * 0.5)
            dy = (nxt - self.G(p)) / 0.5
            if dy < 0:
                v_d = tl.qrot(Q, dtheta_d[0:3])
                v_d[1] += max(dy, -5.0)
                dtheta_d[0:3] = tl.qrot(tl.qinv(Q), v_d)
        contacts = []
        cmask = 0
        for i in (1, 2):
            if not self.leg[i].isSwing:
                for j in range(2):
                    c = contact_pos[i][j]
                    c[1] = self.G(c)
                    dd = c - p
                    if math.sqrt(dd.dot(dd)) < P["max_contact_len"]:
                        contacts.append(c.copy()); contacts.append(c.copy())       # pushed twice (Q7)
                        cmask |= 1 << (2 * (i - 1) + j)
        self.debug["cmask"] = cmask
        self.debug["contacts"] = [c.copy() for c in contacts]
        self.debug["theta_d"] = theta_d.copy()
        self.frame_move(contacts, theta_d, dtheta_d, cdm_action, terrain_y)
        pose_dif, com_lvdif, ref_comvel = self.calc_ref_dif(rotY)
        timing_mod = P["timing_mod_coef"] * (max(com_lvdif, 0.7) / max(ref_comvel, 0.7))
        # is_healthy (:1350-1398)
        min_y = min(P["healthy_min"], self.ref_y * 0.7)
        curr_y = self.p[1] - self.G(self.p)                       # :1361-1365 height above the terrain
        healthy = (min_y < curr_y) and (curr_y < P["healthy_max"])
        if tl.rotation_angle(tl.offset_q(self.Q)) > P["fall_angle"]:
            healthy = False
        target_delta = tl.qmult(tl.qinv(rotY), self.target_Y)
        delta_y = tl.rotation_angle_about_axis(target_delta, Y)
        facing = math.cos(delta_y) - 1.0
        reward = math.exp(-pose_dif * 2) * math.exp(-com_lvdif * 2) * math.exp(facing * P["facing_weight"])
        done = not healthy
        obs = self.get_obs()
        self.change_count += 1
        if self.global_time * FRAME_RATE > P["limit_length"]:
            done = True
        if np.isnan(obs).any():
            obs[:] = 0.0; reward = 0.0; done = True
        self.iframe += FRAME_RATE * RL_STEP * (1 + timing_mod)
        self.global_time += RL_STEP
        self.debug.update(pose_dif=pose_dif, com_lvdif=com_lvdif, ref_comvel=ref_comvel, timing_mod=timing_mod)
        return obs, reward, done

    # ---------------------------------------------------------------- state export (for loading into the CUDA env)
    def export_state(self):
        out = dict(p=self.p.copy(), Q=self.Q.copy(), v=self.v.copy(), wb=self.wb.copy(), iframe=self.iframe,
                   global_time=self.global_time, target_Y=self.target_Y.copy(), replay=np.array(self.replay[-5:]),
                   n_replay=len(self.replay), sim_prev_com=self.sim_prev[0].copy(), legs=[])
        for i in (1, 2):
            li = self.leg[i]
            out["legs"].append(dict(isSwing=li.isSwing, contactPos=li.contactPos.copy(), pre=li.contactPreFilter.copy(),
                                    filt=li.contactFilter.copy(), dfilt=li.dotContactFilter.copy(),
                                    ball=np.array([li.ball[k].x.copy() for k in range(3)]), filtQ=li.filtQ.x.copy(),
                                    filtQorigin=li.filtQorigin.copy(), contactFilterQ=li.contactFilterQ.copy(),
                                    swingPhase=li.swingPhase, spprtPhase=li.spprtPhase))
        return out
