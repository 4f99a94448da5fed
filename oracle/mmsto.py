"""ORACLE (test infrastructure only).  CPU restatement of the momentum-mapped space-time optimiser (MMSTO) of
gym_cdm2 -- SURVEY 8f rank 1 -- in the ONLINE configuration that walk_SRBTrack2025.py:529-587 runs every frame:

  gym_cdm2/shortTermOptimizer.lua:164-265      addDefaultTerms_euler (ddx / dx / summed-angle quadratic terms)
  gym_cdm2/shortTermOptimizer.lua:385-818      customGradientFunction: velocity half-space terms (:466-503), momentum
                                               term with the Gaussian window and the cached mean Jacobian (:650-711),
                                               Gaussian-filtered marker term (:713-743), contact-averaged marker term
                                               (:745-784), CoM term (:786-809)
  gym_cdm2/shortTermOptimizer.lua:826-860      initial guess: frames 0,1 from the input motion, clamped to x_original +- 25 deg
  BaseLib/motion/IK_sdls/NodeWrap.h:102-140    createNodes: one slide / hinge node per channel (root = slides XYZ + hinges YZX)
  BaseLib/motion/IK_sdls/NodeWrap.cpp:496-573  setQ / setDQ in the Euler-root mode (every dof is a 1-dof joint)
  BaseLib/motion/IK_sdls/Node.cpp:59-97,196-238,386-400,429-470   ComputeS / ComputeDS / hinge + slide locals / _updateGrad_S_JT
  BaseLib/motion/IK_sdls/NodeWrap.cpp:107-160  calcMomentumJacobianTranspose
  BaseLib/motion/IK_sdls/NodeWrap.cpp:325-400,1005-1036  calcMomentumCOM / calcCOM / calcCOMjacobianTranspose / updateGrad_S_JT
  BaseLib/math/OperatorStitch.cpp:262-288      addSquaredWeighted (coefficients times sqrt(weight))
  BaseLib/math/Filter.cpp:257-277              GetGaussFilter
  BaseLib/motion/IK_lbfgs/lbfgs.cpp:114-119,246-640,646-738   libLBFGS defaults, main loop, backtracking (regular Wolfe) line search
  BaseLib/math/optimize.h:426-470              LBFGS_METHOD: linesearch = LBFGS_LINESEARCH_BACKTRACKING, epsilon from the caller

PARITY: the Euler-root tree (frames, CoM, momentum, every Jacobian, updateGrad_S_JT) and the L-BFGS loop are PINNED to the reference's
own compiled code (oracle/_ref: IK_sdls and libLBFGS built from /root/reference; tests/test_refpins.py -- libLBFGS driving this
file's objective reproduces status code, iteration / evaluation counts exactly and the iterates at 1e-12).  UNPINNED: the Lua wrapper
`subRoutines/OptimizerLBFGS` that shortTermOptimizer.lua:1 requires is NOT in the reference tree.  It is restated from its call site
(`LBFGS_opta(nf*ndof, 1e-9, ndof*2, 0)`, `h:addWeighted`, `h.customGradientFunction`, `h:solve`) with these stated
assumptions: (W1) the objective is the sum of the addWeighted squared terms and of customGradientFunction, the gradient
likewise; (W2) the third / fourth constructor arguments are the numbers of leading / trailing variables that are held at
their initial values (the Lua comments call frames 0 and 1 "constants"), L-BFGS runs on the remaining variables only;
(W3) the second argument is lbfgs_parameter_t::epsilon, everything else is the libLBFGS default (m = 6, unlimited
iterations, 40 line-search trials); (W4) tree joint numbers equal the dof indices (bones are stored depth first).
Analytic checks in addition (tests/test_oracle_mmsto_cpu.py): finite differences of every term, momentum Jacobian
against calcMomentumCOM, convergence on consistent targets.

Layout: q = [root xyz, root Euler YZX, joint angles in tree order] (ndof = 59 for hanyang_lowdof_T), x = nf stacked q."""
import math

import numpy as np

from . import bone_fk as fk

INV_DT = 30.0


def gauss_filter(kernelsize):
    """Filter::GetGaussFilter."""
    if kernelsize % 2 == 0:
        kernelsize = kernelsize // 2 * 2 + 1
    k = kernelsize // 2
    gamma = 1.0 / kernelsize
    a = np.array([math.exp(gamma * -1 * i * i) for i in range(-k, k + 1)])
    return a / a.sum()


class EulerTree:
    """LoaderToTree(loader, {}, {}, useEulerRoot=true, useFixedRootPos=false)."""

    def __init__(self, skel):
        self.skel = skel
        nb = len(skel["parent"])
        self.nb = nb
        self.dof_bone, self.dof_kind, self.dof_axis, self.dof_first = [], [], [], []
        self.bone_dofs = []
        for b in range(nb):
            ch = [("S", c) for c in "XYZ"] + [("H", c) for c in "YZX"] if b == 0 else [("H", c) for c in skel["channels"][b]]
            s = len(self.dof_bone)
            for k, (kind, c) in enumerate(ch):
                self.dof_bone.append(b); self.dof_kind.append(kind); self.dof_axis.append(fk.AXIS[c].copy()); self.dof_first.append(k == 0)
            self.bone_dofs.append(list(range(s, len(self.dof_bone))))
        self.ndof = len(self.dof_bone)
        self.dof_parent = []
        for j in range(self.ndof):
            b = self.dof_bone[j]
            if not self.dof_first[j]:
                self.dof_parent.append(j - 1)
            elif b == 0:
                self.dof_parent.append(-1)
            else:
                self.dof_parent.append(self.bone_dofs[skel["parent"][b]][-1])
        self.mass = np.array(skel["mass"], dtype=float)
        self.total_mass = float(self.mass.sum())

    # --- Node::ComputeS over the whole tree (Tree::Compute)
    def set_q(self, q):
        n = self.ndof
        self.q = np.array(q, dtype=float)
        self.grot = np.zeros((n, 4)); self.gtr = np.zeros((n, 3)); self.W = np.zeros((n, 3))
        self.lrot = np.zeros((n, 4)); self.ltr = np.zeros((n, 3))
        for j in range(n):
            off = np.array(self.skel["offset"][self.dof_bone[j]], dtype=float) if self.dof_first[j] else np.zeros(3)
            if self.dof_kind[j] == "S":
                lq = np.array([1.0, 0, 0, 0]); lt = off + self.dof_axis[j] * self.q[j]
            else:
                lq = fk.qaxis(self.dof_axis[j], self.q[j]); lt = off
            self.lrot[j], self.ltr[j] = lq, lt
            p = self.dof_parent[j]
            if p < 0:
                self.grot[j], self.gtr[j] = lq, lt
            else:
                self.grot[j] = fk.qmult(self.grot[p], lq)
                self.gtr[j] = fk.vrotate(self.grot[p], lt) + self.gtr[p]
            self.W[j] = fk.vrotate(self.grot[j], self.dof_axis[j])          # HingeJoint::ComputeW
        last = [d[-1] for d in self.bone_dofs]
        self.rot = self.grot[last].copy(); self.tr = self.gtr[last].copy()

    # --- setDQ + Tree::ComputeDS
    def set_dq(self, dq):
        n = self.ndof
        lw = np.zeros((n, 3)); lv = np.zeros((n, 3))
        for j in range(n):
            rel_w = self.dof_axis[j] * dq[j] if self.dof_kind[j] == "H" else np.zeros(3)
            rel_v = self.dof_axis[j] * dq[j] if self.dof_kind[j] == "S" else np.zeros(3)
            p = self.dof_parent[j]
            if p < 0:
                lw[j], lv[j] = rel_w, rel_v
            else:
                t_rel = fk.qinv(self.lrot[j])
                v1 = np.cross(lw[p], self.ltr[j])
                lv[j] = fk.vrotate(t_rel, v1) + fk.vrotate(t_rel, lv[p]) + rel_v
                lw[j] = fk.vrotate(t_rel, lw[p]) + rel_w
        last = [d[-1] for d in self.bone_dofs]
        self.bw = lw[last].copy(); self.bv = lv[last].copy()

    def global_pos(self, bone, lpos):
        return fk.vrotate(self.rot[bone], np.asarray(lpos, dtype=float)) + self.tr[bone]

    def update_grad_S_JT(self, g, deltaS, bone, lpos):
        """g += J^T deltaS for the point lpos of `bone` (walks the node chain to the root)."""
        target = self.global_pos(bone, lpos)
        j = self.bone_dofs[bone][-1]
        while j >= 0:
            if self.dof_kind[j] == "H":
                g[j] += deltaS.dot(np.cross(self.W[j], target - self.gtr[j]))
            else:
                g[j] += deltaS.dot(self.W[j])
            j = self.dof_parent[j]

    def jacobian_T_at(self, bone, lpos):
        JT = np.zeros((self.ndof, 3))
        target = self.global_pos(bone, lpos)
        j = self.bone_dofs[bone][-1]
        while j >= 0:
            JT[j] = np.cross(self.W[j], target - self.gtr[j]) if self.dof_kind[j] == "H" else self.W[j]
            j = self.dof_parent[j]
        return JT

    def rot_jacobian_T(self, bone):
        JT = np.zeros((self.ndof, 3))
        j = self.bone_dofs[bone][-1]
        while j >= 0:
            if self.dof_kind[j] == "H":
                JT[j] = self.W[j]
            j = self.dof_parent[j]
        return JT

    def calc_com(self):
        return fk.calc_com(self.skel, self.rot, self.tr)

    def calc_momentum_com(self):
        return fk.calc_momentum_com(self.skel, self.rot, self.tr, self.bw, self.bv)[1]

    def calc_com_jacobian_T(self):
        JT = np.zeros((self.ndof, 3))
        for b in range(self.nb):
            JT += self.jacobian_T_at(b, self.skel["com"][b]) * self.mass[b]
        return JT * (1.0 / self.total_mass)

    def calc_momentum_jacobian_T(self):
        """NodeWrap.cpp:107-160, literally: body twist Jacobian -> Inertia * se3 -> dAd(inv(body) * T(COM))."""
        JT = np.zeros((self.ndof, 6))
        com = self.calc_com()
        for b in range(self.nb):
            j1 = self.jacobian_T_at(b, np.zeros(3))
            j2 = self.rot_jacobian_T(b)
            Rinv = fk.qinv(self.rot[b])
            mass = self.mass[b]
            c = np.array(self.skel["com"][b], dtype=float)
            I = np.array(self.skel["inertia"][b]).reshape(3, 3)
            r = mass * c
            # T = inv(body) * translate(COM): rotation R^-1, translation R^-1 (COM - t)
            p = fk.vrotate(Rinv, com - self.tr[b])
            for j in range(self.ndof):
                w = fk.vrotate(Rinv, j2[j]); v = fk.vrotate(Rinv, j1[j])
                M = I.dot(w) + np.cross(r, v)
                F = np.cross(w, r) + mass * v
                # dAd(T) (Liegroup.h:207-218): [R^T, R^T skew(-p); 0, R^T] with R = R_body^-1
                JT[j, :3] += fk.vrotate(self.rot[b], M - np.cross(p, F))
                JT[j, 3:] += fk.vrotate(self.rot[b], F)
        return JT


class ShortTrajOpt:
    """ShortTrajOpt with onlineOpt = true.  All inputs are what walk_SRBTrack2025.py:529-587 sets before solveEndEffectors."""

    DEFAULT_WEIGHTS = dict(weightDDX=10.0, weightDX=100.0, weightAngleOriginal=100000.0, weightEE=100000.0, weightEEcon=0.0,
                           weightVelHalfSpace=100000.0, weightMomentum=1000.0, weightCOM=100000.0)

    def __init__(self, skel, nf, markers, **weights):
        self.tree = EulerTree(skel)
        self.trees = [EulerTree(skel) for _ in range(nf)]
        self.nf = nf
        self.ndof = self.tree.ndof
        self.markers = markers                      # list of dict(bone=, lpos=, weight=)
        self.dof_weights = np.ones(self.ndof)
        self.clamp_thr = 25.0
        for k, v in self.DEFAULT_WEIGHTS.items():
            setattr(self, k, float(weights.get(k, v)))
        self.x_original = np.zeros((nf, self.ndof)); self.dx = np.zeros((nf - 1, self.ndof)); self.ddx = np.zeros((nf - 2, self.ndof))
        self.com = None; self.momentum = None; self.con = None
        self.gpos_target = None; self.hs_vel = {}
        self.mcache = None
        self.n_eval = 0

    # --- addDefaultTerms_euler
    def default_terms(self, x, grad):
        nf, nd = self.nf, self.ndof
        X = x.reshape(nf, nd); G = grad.reshape(nf, nd)
        o = 0.0
        inv, invs = INV_DT, INV_DT * INV_DT
        if self.weightDDX != 0:
            for i in range(1, nf - 1):
                s = np.sqrt(self.weightDDX * self.dof_weights)
                e = s * invs * X[i - 1] + (s * (-2 * invs)) * X[i] + s * invs * X[i + 1] + s * (-self.ddx[i - 1])
                o += float((e * e).sum())
                G[i - 1] += 2 * e * (s * invs); G[i] += 2 * e * (s * (-2 * invs)); G[i + 1] += 2 * e * (s * invs)
        if self.weightDX != 0:
            for i in range(1, nf - 1):
                s = np.sqrt(self.weightDX * self.dof_weights)
                e = (s * -inv) * X[i - 1] + (s * inv) * X[i] + s * (-self.dx[i - 1])
                o += float((e * e).sum())
                G[i - 1] += 2 * e * (s * -inv); G[i] += 2 * e * (s * inv)
        if self.weightAngleOriginal != 0:
            s = np.sqrt(self.weightAngleOriginal * self.dof_weights)
            e = s * X.sum(axis=0) + s * (-self.x_original.sum(axis=0))
            o += float((e * e).sum())
            G += (2 * e * s)[None, :]
        return o

    # --- customGradientFunction
    def custom_terms(self, x, grad):
        nf, nd = self.nf, self.ndof
        X = x.reshape(nf, nd); G = grad.reshape(nf, nd)
        trees = self.trees
        o = 0.0
        last_added = 1
        for i in range(nf):
            trees[i].set_q(X[i])
            if i > last_added:
                info = self.hs_vel.get(i)
                if info is not None and self.weightVelHalfSpace != 0:
                    w = self.weightVelHalfSpace
                    for speed, normal, bone, lpos in info:
                        normal = np.asarray(normal, dtype=float)
                        gvel = (trees[i].global_pos(bone, lpos) - trees[i - 1].global_pos(bone, lpos)) * INV_DT
                        d = (normal * speed - gvel).dot(normal)
                        if d > 0:
                            o += d * d * w
                            trees[i].update_grad_S_JT(G[i], -(2.0 * w * INV_DT * d) * normal, bone, lpos)
                            trees[i - 1].update_grad_S_JT(G[i - 1], (2.0 * w * INV_DT * d) * normal, bone, lpos)
                last_added = i
        if self.momentum is not None and self.weightMomentum > 0:
            wm = self.weightMomentum
            filt = gauss_filter(nf)
            if self.mcache is None:
                mc = np.zeros((nd, 6))
                for i in range(nf - 1):
                    mc += trees[i].calc_momentum_jacobian_T()
                self.mcache = mc * (1.0 / (nf - 1))
            for i in range(nf - 1):
                trees[i].set_dq((X[i + 1] - X[i]) * INV_DT)
                r = trees[i].calc_momentum_com()
                dS = self.momentum[i] - r
                o += float(dS.dot(dS)) * wm * filt[i]
                gf = self.mcache.dot((-2.0 * wm * filt[i] * INV_DT) * dS)
                G[i + 1] += gf; G[i] -= gf
        filt = gauss_filter(nf)
        if self.weightEE > 0:
            for m, mk in enumerate(self.markers):
                w = self.weightEE * mk["weight"]
                dS = np.zeros(3)
                for i in range(nf):
                    dS += (self.gpos_target[i, 3 * m:3 * m + 3] - trees[i].global_pos(mk["bone"], mk["lpos"])) * filt[i]
                o += float(dS.dot(dS)) * w
                for i in range(nf):
                    trees[i].update_grad_S_JT(G[i], -(2.0 * w * filt[i]) * dS, mk["bone"], mk["lpos"])
        if self.con is not None and self.weightEEcon > 0:
            for m, mk in enumerate(self.markers):
                w = self.weightEEcon * mk["weight"]
                dS = np.zeros(3); c = 0
                for i in range(nf):
                    if self.con[m][i]:
                        c += 1
                        dS += self.gpos_target[i, 3 * m:3 * m + 3] - trees[i].global_pos(mk["bone"], mk["lpos"])
                if c > 0:
                    dS *= 1 / c
                    o += float(dS.dot(dS)) * w
                    for i in range(nf):
                        if self.con[m][i]:
                            trees[i].update_grad_S_JT(G[i], -(2.0 * w / c) * dS, mk["bone"], mk["lpos"])
        if self.com is not None and self.weightCOM > 0:
            wc = self.weightCOM
            fs = max(nf // 2, 2)
            dS = np.zeros(3)
            for i in range(fs):
                dS += self.com[i] - trees[i].calc_com()
            dS *= 1.0 / fs
            o += float(dS.dot(dS)) * wc
            for i in range(fs):
                G[i] += trees[i].calc_com_jacobian_T().dot((-2.0 * wc / fs) * dS)
        return o

    def evaluate(self, x):
        """-> (objective, gradient) over all nf * ndof variables (W1)."""
        self.n_eval += 1
        g = np.zeros(self.nf * self.ndof)
        o = self.default_terms(x, g)
        o += self.custom_terms(x, g)
        return o, g

    def initial_guess(self, euler_mot):
        """shortTermOptimizer.lua:826-838: x_original with frames 0, 1 replaced by the (clamped) input motion."""
        init = self.x_original.copy()
        thr = math.radians(self.clamp_thr)
        for f in (0, 1):
            init[f] = np.clip(euler_mot[f], self.x_original[f] - thr, self.x_original[f] + thr)
        return init

    def solve(self, euler_mot, epsilon=1e-9, max_iterations=0):
        """solveEndEffectors_euler -> (x [nf, ndof], info)."""
        self.mcache = None
        self.n_eval = 0
        x0 = self.initial_guess(np.asarray(euler_mot, dtype=float)).ravel()
        nfix = 2 * self.ndof                                              # W2

        def ev(z):
            x = x0.copy(); x[nfix:] = z
            o, g = self.evaluate(x)
            return o, g[nfix:]
        z, fx, ret, k = lbfgs(x0[nfix:].copy(), ev, epsilon=epsilon, max_iterations=max_iterations)
        x = x0.copy(); x[nfix:] = z
        return x.reshape(self.nf, self.ndof), dict(fx=fx, ret=ret, iterations=k, evaluations=self.n_eval)


LBFGS_SUCCESS, LBFGS_ALREADY_MINIMIZED = 0, 2
LBFGSERR_MINIMUMSTEP, LBFGSERR_MAXIMUMSTEP, LBFGSERR_MAXIMUMLINESEARCH, LBFGSERR_MAXIMUMITERATION, LBFGSERR_INCREASEGRADIENT = -1000, -999, -998, -997, -994


def _linesearch_backtracking(x, f, g, s, stp, xp, evaluate, ftol=1e-4, wolfe=0.9, min_step=1e-20, max_step=1e20, max_linesearch=40):
    """lbfgs.cpp:646-738 with linesearch == LBFGS_LINESEARCH_BACKTRACKING_WOLFE."""
    count = 0
    dginit = float(g.dot(s))
    if 0 < dginit:
        return LBFGSERR_INCREASEGRADIENT, x, f, g, stp
    finit = f
    dgtest = ftol * dginit
    while True:
        x = xp + s * stp
        f, g = evaluate(x)
        count += 1
        if f > finit + stp * dgtest:
            width = 0.5
        else:
            dg = float(g.dot(s))
            if dg < wolfe * dginit:
                width = 2.1
            else:
                return count, x, f, g, stp
        if stp < min_step:
            return LBFGSERR_MINIMUMSTEP, x, f, g, stp
        if stp > max_step:
            return LBFGSERR_MAXIMUMSTEP, x, f, g, stp
        if max_linesearch <= count:
            return LBFGSERR_MAXIMUMLINESEARCH, x, f, g, stp
        stp *= width


def lbfgs(x, evaluate, epsilon=1e-5, m=6, max_iterations=0):
    """lbfgs.cpp:246-640 without the orthant-wise and `past` branches -> (x, fx, ret, iterations)."""
    n = len(x)
    S = np.zeros((m, n)); Y = np.zeros((m, n)); YS = np.zeros(m); alpha = np.zeros(m)
    fx, g = evaluate(x)
    d = -g
    xnorm = max(math.sqrt(float(x.dot(x))), 1.0)
    gnorm = math.sqrt(float(g.dot(g)))
    if gnorm / xnorm <= epsilon:
        return x, fx, LBFGS_ALREADY_MINIMIZED, 0
    step = 1.0 / math.sqrt(float(d.dot(d)))
    k = 1; end = 0
    while True:
        xp = x.copy(); gp = g.copy()
        ls, x, fx, g, step = _linesearch_backtracking(x, fx, g, d, step, xp, evaluate)
        if ls < 0:
            return xp, fx, ls, k
        xnorm = max(math.sqrt(float(x.dot(x))), 1.0)
        gnorm = math.sqrt(float(g.dot(g)))
        if gnorm / xnorm <= epsilon:
            return x, fx, LBFGS_SUCCESS, k
        if max_iterations != 0 and max_iterations < k + 1:
            return x, fx, LBFGSERR_MAXIMUMITERATION, k
        S[end] = x - xp; Y[end] = g - gp
        ys = float(Y[end].dot(S[end])); yy = float(Y[end].dot(Y[end]))
        YS[end] = ys
        bound = m if m <= k else k
        k += 1
        end = (end + 1) % m
        d = -g
        j = end
        for _ in range(bound):
            j = (j + m - 1) % m
            alpha[j] = float(S[j].dot(d)) / YS[j]
            d = d + Y[j] * (-alpha[j])
        d = d * (ys / yy)
        for _ in range(bound):
            beta = float(Y[j].dot(d)) / YS[j]
            d = d + S[j] * (alpha[j] - beta)
            j = (j + 1) % m
        step = 1.0


# ---------------------------------------------------------------------------------------------------------------------
# removeCOMsliding_online (gym_trackSRB/taesooLib/DeltaExtractor.py:421-512): the step right after the MMSTO solve in the
# online loop (walk_SRBTrack2025.py:617, walk_SRBTrack2025_terrain_singlethreaded.py:610).  SURVEY 8f rank 2.
#   QuadraticFunctionHardCon: BaseLib/math/OperatorStitch.cpp:531-620 (H = 2 sum c c^T, R = -2 sum v c, KKT matrix),
#   Resource/scripts/module.lua:2108-2124,2165-2207 (addWeighted, con, solve = math.LUsolve of the KKT system).
# PARITY PINNED to the reference's own removeCOMsliding_online run from /root/reference (tests/tools/ref_delta_harness.py, tests/test_refenv_pins.py:
# 8e-16, flat and with projectToGround), besides the KKT conditions and constraint residuals in tests/test_oracle_mmsto_cpu.py.

def hardcon_solve(nvar, sq_terms, con_terms):
    """sq_terms: (weight, [(coef, index)...], const); con_terms: ([(coef, index)...], const) -> x[nvar]."""
    ncon = len(con_terms)
    K = np.zeros((nvar + ncon, nvar + ncon)); d = np.zeros(nvar + ncon)
    for w, ci, const in sq_terms:
        s = math.sqrt(w)
        for a, (ca, ia) in enumerate(ci):
            K[ia, ia] += 2.0 * (s * ca) ** 2
            for cb, ib in ci[a + 1:]:
                K[ia, ib] += 2.0 * (s * ca) * (s * cb); K[ib, ia] = K[ia, ib]
            d[ia] -= 2.0 * (s * const) * (s * ca)
    for r, (ci, const) in enumerate(con_terms):
        for ca, ia in ci:
            K[nvar + r, ia] = ca; K[ia, nvar + r] = ca
        d[nvar + r] = -const
    return np.linalg.solve(K, d)[:nvar]


def remove_com_sliding_online(skel, com_srb, desired_com, output, fix_y_also=True, ground_h=None):
    """output: poses [nvar, 7 + #angles] (root xyz, quaternion wxyz, angles), modified copy returned with (currCOM, optCOM).
    ground_h[nvar]: terrain height under each current CoM (projectToGround), None = no projection (flat script)."""
    output = np.array(output, dtype=float)
    nvar = output.shape[0]
    curr = np.zeros((nvar, 3))
    for i in range(nvar):
        rot, tr = fk.bone_forward_kinematics(skel, output[i])
        curr[i] = fk.calc_com(skel, rot, tr)
    opt = np.zeros((nvar, 3))
    for dim in range(3):
        sq = []
        for v in range(1, nvar - 1):
            sq.append((1.5, [(-1.0, v - 1), (2.0, v), (-1.0, v + 1)], -1 * (-1 * com_srb[v - 1][dim] + 2 * com_srb[v][dim] - com_srb[v + 1][dim])))
        con = [([(1.0, 0)], -curr[0][dim]), ([(1.0, 1)], -curr[1][dim]),
               ([(1.0, i) for i in range(1, nvar)], -1 * float(np.sum(desired_com[0:nvar - 1, dim])))]
        opt[:, dim] = hardcon_solve(nvar, sq, con)
    if not fix_y_also:
        opt[:, 1] = curr[:, 1]
    for i in range(nvar):
        delta = opt[i] - curr[i]
        delta[1] = 0
        axis = np.cross(delta, np.array([0.0, 1.0, 0.0]))
        ln = math.sqrt(float(axis.dot(axis)))
        axis = axis * (0.0 if ln == 0 else 1.0 / ln)
        cy = curr[i, 1]
        if ground_h is not None:
            cy -= ground_h[i]
        R = fk.qaxis(axis, -math.sqrt(float(delta.dot(delta))) * cy * 0.5)
        # (plusCOM * R * minusCOM) * root: rotation R q, translation R (t - curr) + opt
        q = output[i, 3:7].copy(); t = output[i, 0:3].copy()
        output[i, 3:7] = fk.qmult(R, q)
        output[i, 0:3] = fk.vrotate(R, t - curr[i]) + opt[i]
    return output, curr, opt
